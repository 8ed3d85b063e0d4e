// p2p_ops.cuh — addressing and element moves of the peer-memory halo exchange (comm.cu, JETCU_P2P_HALO=1) as plain
// __host__ __device__ code: where in which mailbox a rank's boundary planes go, which flag announces them, where the halo planes
// come from.  comm.cu's kernels add the parallel loops, the fences and the bounded waits; tests/hostcheck runs this very source
// for a chain of ranks in host memory and checks that every halo plane ends up holding the right neighbour's plane.
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifndef __CUDACC__
struct uint4 { unsigned x, y, z, w; };
#endif

struct P2PFlags {
    unsigned long long data[2];  // last message number published by the rank below [0] / above [1] — written by the PEER
    unsigned long long seq;      // exchanges completed on this rank
    unsigned long long err;      // sticky: a wait timed out
    unsigned int pushed, pulled; // CTA completion counters of the running kernel
};
#define P2P_HEADER 256

// One exchange as seen by one rank.  Mailbox layout: [P2P_HEADER | from-below slot 0, slot 1 | from-above slot 0, slot 1].
struct P2PExchange {
    P2PFlags* me;
    // push: my lowest owned planes -> the lower neighbour's "from above" side, my highest -> the upper neighbour's "from below" side
    const uint4* src_dn; uint4* box_dn; unsigned long long* flag_dn;
    const uint4* src_up; uint4* box_up; unsigned long long* flag_up;
    // pull: my own mailbox -> my halo planes below / above the slab
    const uint4* in_dn; uint4* dst_dn;
    const uint4* in_up; uint4* dst_up;
    size_t n16, slot16;          // 16-byte elements per message / per slot
};

// `box` = this rank's mailbox, `peer_lo` / `peer_hi` = the mailboxes of rank - 1 / rank + 1 as mapped into this process,
// `p` = base of the globally indexed array, planes [k0, k1) owned, `depth` planes exchanged each way where a neighbour exists.
__host__ __device__ inline P2PExchange p2p_exchange(char* box, char* peer_lo, char* peer_hi, size_t slot_bytes, char* p,
                                                    size_t plane_bytes, int k0, int k1, int depth, bool down, bool up) {
    P2PExchange e;
    const size_t bytes = (size_t)depth * plane_bytes;
    e.n16 = bytes / 16; e.slot16 = slot_bytes / 16;
    e.me = reinterpret_cast<P2PFlags*>(box);
    const size_t side0 = P2P_HEADER, side1 = P2P_HEADER + 2 * slot_bytes;   // written by the rank below / above
    e.src_dn = down ? reinterpret_cast<const uint4*>(p + (size_t)k0 * plane_bytes) : nullptr;
    e.box_dn = down ? reinterpret_cast<uint4*>(peer_lo + side1) : nullptr;                        // I am ABOVE my lower neighbour
    e.flag_dn = down ? &reinterpret_cast<P2PFlags*>(peer_lo)->data[1] : nullptr;
    e.src_up = up ? reinterpret_cast<const uint4*>(p + (size_t)(k1 - depth) * plane_bytes) : nullptr;
    e.box_up = up ? reinterpret_cast<uint4*>(peer_hi + side0) : nullptr;                          // I am BELOW my upper neighbour
    e.flag_up = up ? &reinterpret_cast<P2PFlags*>(peer_hi)->data[0] : nullptr;
    e.in_dn = down ? reinterpret_cast<const uint4*>(box + side0) : nullptr;
    e.dst_dn = down ? reinterpret_cast<uint4*>(p + (size_t)(k0 - depth) * plane_bytes) : nullptr;
    e.in_up = up ? reinterpret_cast<const uint4*>(box + side1) : nullptr;
    e.dst_up = up ? reinterpret_cast<uint4*>(p + (size_t)k1 * plane_bytes) : nullptr;
    return e;
}

// message `seq` lives in slot seq & 1
__host__ __device__ inline size_t p2p_slot_offset(const P2PExchange& e, unsigned long long seq) { return (size_t)(seq & 1ull) * e.slot16; }
__host__ __device__ inline bool p2p_arrived(const P2PExchange& e, unsigned long long data_dn, unsigned long long data_up,
                                            unsigned long long seq) {
    return (!e.dst_dn || data_dn >= seq) && (!e.dst_up || data_up >= seq);
}
