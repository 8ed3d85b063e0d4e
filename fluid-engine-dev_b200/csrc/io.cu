// io.cu — the particle frame files the reference's examples write: saveParticleAsPos / saveParticleAsXyz,
// src/examples/hybrid_liquid_sim/main.cpp:29-63, fed straight from the device-resident particle set.
//
//   .pos  jet::serialize(const uint8_t*, size_t, …) (src/jet/serialization.cpp:25-38): a FlatBuffers `FlatData { data:[ubyte] }`
//         table around the raw Vector3D array.  With one field the builder's output is a fixed 24-byte frame —
//         root offset 12 | 2 pad | vtable {6, 8, 4} | table {soffset 6, uoffset 4} | u32 length — followed by the payload and
//         zero padding to a multiple of 4 (checked byte for byte against the unmodified reference, tests/test_oracle_io.py).
//   .xyz  `file << x << ' ' << y << ' ' << z << std::endl`: operator<< on doubles with the default format == "%g".
//
// Host code only (nvcc compiles it as C++): positions come through jetcu_hybrid_get_particles in the original particle order.
#include <stdio.h>
#include <string.h>

#include <new>
#include <string>
#include <atomic>
#include <thread>
#include <vector>

#include "common.cuh"

extern "C" {

uint64_t jetcu_pos_frame_size(uint64_t payload_bytes) { return 24 + ((payload_bytes + 3) & ~(uint64_t)3); }

int jetcu_pos_frame_header(uint64_t payload_bytes, uint8_t header[24]) {
    // FlatBuffers vectors carry a 32-bit length (the reference's builder additionally asserts buffers < 2 GiB)
    if (!header || payload_bytes > 0xffffffffull - 24) return JETCU_ERR_INVALID_ARG;
    static const uint8_t frame[20] = {0x0c, 0, 0, 0, 0, 0, 0x06, 0, 0x08, 0, 0x04, 0, 0x06, 0, 0, 0, 0x04, 0, 0, 0};
    memcpy(header, frame, 20);
    const uint32_t len = (uint32_t)payload_bytes;
    header[20] = (uint8_t)(len & 0xff); header[21] = (uint8_t)((len >> 8) & 0xff);
    header[22] = (uint8_t)((len >> 16) & 0xff); header[23] = (uint8_t)((len >> 24) & 0xff);
    return JETCU_OK;
}

// one "x y z\n" line per particle into `out` (capacity `cap`); *out_size = bytes needed (call with cap 0 to size)
int jetcu_format_xyz(const double* positions, uint64_t n, char* out, uint64_t cap, uint64_t* out_size) {
    if (!out_size || (n && !positions)) return JETCU_ERR_INVALID_ARG;
    uint64_t at = 0;
    char line[128];
    for (uint64_t i = 0; i < n; ++i) {
        const int len = snprintf(line, sizeof(line), "%g %g %g\n", positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]);
        if (out && at + (uint64_t)len <= cap) memcpy(out + at, line, (size_t)len);
        at += (uint64_t)len;
    }
    *out_size = at;
    return (out && at > cap) ? JETCU_ERR_INVALID_ARG : JETCU_OK;
}

int jetcu_write_pos_file(const char* path, const double* positions, uint64_t n) {
    if (!path || (n && !positions)) return JETCU_ERR_INVALID_ARG;
    uint8_t header[24];
    const uint64_t bytes = n * 24;
    if (jetcu_pos_frame_header(bytes, header) != JETCU_OK) return JETCU_ERR_INVALID_ARG;
    FILE* f = fopen(path, "wb");
    if (!f) return JETCU_ERR_INVALID_ARG;
    bool ok = fwrite(header, 1, 24, f) == 24;
    if (ok && bytes) ok = fwrite(positions, 1, (size_t)bytes, f) == (size_t)bytes;   // 24 n is a multiple of 4: no padding
    ok = (fclose(f) == 0) && ok;
    return ok ? JETCU_OK : JETCU_ERR_INVALID_ARG;
}

static int write_xyz_file(const char* path, const double* positions, uint64_t n) {
    FILE* f = fopen(path, "w");
    if (!f) return JETCU_ERR_INVALID_ARG;
    // formatting is the cost (~100 ns per number): blocks of particles are formatted by a few threads, written in order
    const uint64_t block = 1u << 18;
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    if (nt > 32) nt = 32;
    bool ok = true;
    for (uint64_t base = 0; base < n && ok; base += block * nt) {
        std::vector<std::string> text(nt);
        std::vector<std::thread> pool;
        std::atomic<bool> failed(false);   // a worker ran out of memory: no exception may leave a thread (std::terminate)
        bool spawn_failed = false;
        for (unsigned t = 0; t < nt && !spawn_failed; ++t) {
            const uint64_t b0 = base + (uint64_t)t * block;
            if (b0 >= n) break;
            const uint64_t cnt = (n - b0 < block) ? n - b0 : block;
            try {
                pool.emplace_back([&text, &failed, t, positions, b0, cnt]() {
                    try {
                        std::string& s = text[t];
                        s.resize((size_t)cnt * 48);
                        uint64_t need = 0;
                        if (jetcu_format_xyz(positions + 3 * b0, cnt, &s[0], s.size(), &need) != JETCU_OK) {
                            s.resize((size_t)need);
                            jetcu_format_xyz(positions + 3 * b0, cnt, &s[0], s.size(), &need);
                        }
                        s.resize((size_t)need);
                    } catch (...) {
                        failed = true;
                    }
                });
            } catch (...) {   // thread creation failed: join what runs, then report
                spawn_failed = true;
            }
        }
        for (auto& th : pool) th.join();
        if (spawn_failed || failed) { fclose(f); return JETCU_ERR_INVALID_ARG; }
        for (unsigned t = 0; t < nt && ok; ++t)
            if (!text[t].empty()) ok = fwrite(text[t].data(), 1, text[t].size(), f) == text[t].size();
    }
    ok = (fclose(f) == 0) && ok;
    return ok ? JETCU_OK : JETCU_ERR_INVALID_ARG;
}

// no C++ exception may cross the C boundary: allocation (text blocks, the host copy of the positions) and thread creation
// failures become status codes
int jetcu_write_xyz_file(const char* path, const double* positions, uint64_t n) {
    if (!path || (n && !positions)) return JETCU_ERR_INVALID_ARG;
    try {
        return write_xyz_file(path, positions, n);
    } catch (const std::bad_alloc&) {
        return JETCU_ERR_OOM;
    } catch (...) {
        return JETCU_ERR_INVALID_ARG;
    }
}

// saveParticleAsPos / saveParticleAsXyz for the device-resident particle set (original particle order)
int jetcu_hybrid_save_positions(jetcu_hybrid* h, const char* path, int format) {
    if (!h || !path || (format != JETCU_FRAME_POS && format != JETCU_FRAME_XYZ)) return JETCU_ERR_INVALID_ARG;
    try {
        const uint64_t n = jetcu_hybrid_num_particles(h);
        std::vector<double> pos((size_t)n * 3);
        if (n) {
            const int rc = jetcu_hybrid_get_particles(h, pos.data(), nullptr);
            if (rc != JETCU_OK) return rc;
        }
        return format == JETCU_FRAME_POS ? jetcu_write_pos_file(path, pos.data(), n) : jetcu_write_xyz_file(path, pos.data(), n);
    } catch (const std::bad_alloc&) {
        return JETCU_ERR_OOM;
    } catch (...) {
        return JETCU_ERR_INVALID_ARG;
    }
}

}  // extern "C"
