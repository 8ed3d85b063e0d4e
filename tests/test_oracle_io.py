"""The particle frame files of the reference's examples (saveParticleAsPos / saveParticleAsXyz,
src/examples/hybrid_liquid_sim/main.cpp:29-63): the product's host-side writers (csrc/io.cu — no GPU involved) against the
UNMODIFIED reference's jet::serialize and ostream formatting (oracle/_ref), byte for byte, plus the frame's known answer so the
check still runs where the reference library is absent."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle_api import REF_PATH, have_ref
from jetcu_loader import load_package

load_package()
from fluid_engine_dev_b200 import LIB_PATH  # noqa: E402

pytestmark = pytest.mark.skipif(not os.path.exists(LIB_PATH), reason="libjetcu.so not built")


@pytest.fixture(scope="module")
def lib():
    l = C.CDLL(LIB_PATH)
    l.jetcu_pos_frame_size.restype = C.c_uint64
    l.jetcu_pos_frame_size.argtypes = [C.c_uint64]
    return l


@pytest.fixture(scope="module")
def ref():
    if not have_ref():
        pytest.skip("oracle/_ref/libjetref.so not built")
    return C.CDLL(REF_PATH)


def positions(n, seed=0):
    rng = np.random.default_rng(seed)
    p = rng.uniform(0.0, 1.0, (n, 3))
    if n >= 8:   # exponents, integers, negative zero, tiny and huge values: everything "%g" switches format on
        p[0] = (0.0, 1.0, -0.0); p[1] = (1e-5, 9.9999995e-5, 123456.5); p[2] = (1234567.0, 1e21, -2.5e-310)
        p[3] = (0.1, 1.0 / 3.0, 100000.0); p[4] = (999999.5, 0.000123456789, 5e-324)
    return np.ascontiguousarray(p)


def test_pos_frame_known_answer(lib):
    """root offset 12, vtable {6, 8, 4} at 6, table {6, 4} at 12, u32 length at 20, payload at 24, padded to 4 bytes."""
    hdr = (C.c_uint8 * 24)()
    assert lib.jetcu_pos_frame_header(C.c_uint64(48), hdr) == 0
    assert bytes(hdr).hex() == "0c0000000000060008000400060000000400000030000000"
    assert lib.jetcu_pos_frame_size(48) == 72 and lib.jetcu_pos_frame_size(5) == 32 and lib.jetcu_pos_frame_size(0) == 24
    assert lib.jetcu_pos_frame_header(C.c_uint64(1 << 32), hdr) != 0          # FlatBuffers vectors have 32-bit lengths


@pytest.mark.parametrize("n", [0, 1, 2, 7, 1000])
def test_pos_file_equals_reference_serialize(lib, ref, n, tmp_path):
    p = positions(n)
    path = tmp_path / "frame_000000.pos"
    assert lib.jetcu_write_pos_file(C.c_char_p(os.fsencode(str(path))), p.ctypes.data_as(C.POINTER(C.c_double)), C.c_uint64(n)) == 0
    got = path.read_bytes()
    out = (C.c_uint8 * (24 * n + 64))()
    size = C.c_uint64(0)
    assert ref.jeto_serialize_bytes(p.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_uint64(24 * n), out, C.c_uint64(len(out)), C.byref(size)) == 0
    assert got == bytes(out[:size.value])
    assert len(got) == lib.jetcu_pos_frame_size(24 * n)


@pytest.mark.parametrize("size", [0, 1, 2, 3, 4, 5, 7, 8, 9, 16, 1001])
def test_frame_header_and_padding_for_any_payload(lib, ref, size):
    data = np.arange(1, size + 1, dtype=np.uint8)
    out = (C.c_uint8 * (size + 64))()
    n = C.c_uint64(0)
    buf = data.ctypes.data_as(C.POINTER(C.c_uint8)) if size else None
    assert ref.jeto_serialize_bytes(buf, C.c_uint64(size), out, C.c_uint64(len(out)), C.byref(n)) == 0
    want = bytes(out[:n.value])
    hdr = (C.c_uint8 * 24)()
    assert lib.jetcu_pos_frame_header(C.c_uint64(size), hdr) == 0
    total = lib.jetcu_pos_frame_size(size)
    assert bytes(hdr) + data.tobytes() + b"\0" * (total - 24 - size) == want


@pytest.mark.parametrize("n", [0, 1, 8, 5000])
def test_xyz_equals_reference_stream(lib, ref, n, tmp_path):
    p = positions(n, seed=n)
    pp = p.ctypes.data_as(C.POINTER(C.c_double))
    cap = 64 * n + 16
    want = C.create_string_buffer(cap)
    size = C.c_uint64(0)
    assert ref.jeto_format_xyz(pp, C.c_uint64(n), want, C.c_uint64(cap), C.byref(size)) == 0
    need = C.c_uint64(0)
    assert lib.jetcu_format_xyz(pp, C.c_uint64(n), None, C.c_uint64(0), C.byref(need)) == 0 and need.value == size.value
    got = C.create_string_buffer(cap)
    assert lib.jetcu_format_xyz(pp, C.c_uint64(n), got, C.c_uint64(cap), C.byref(need)) == 0
    assert got.raw[:need.value] == want.raw[:size.value]
    path = tmp_path / "frame_000000.xyz"
    assert lib.jetcu_write_xyz_file(C.c_char_p(os.fsencode(str(path))), pp, C.c_uint64(n)) == 0
    assert path.read_bytes() == want.raw[:size.value]


def test_xyz_many_blocks(lib, tmp_path):
    """More particles than one formatting block per thread: blocks land in the file in order."""
    n = (1 << 18) * 3 + 17
    p = positions(n, seed=5)
    path = tmp_path / "big.xyz"
    assert lib.jetcu_write_xyz_file(C.c_char_p(os.fsencode(str(path))), p.ctypes.data_as(C.POINTER(C.c_double)), C.c_uint64(n)) == 0
    back = np.loadtxt(path)
    assert back.shape == (n, 3)
    assert np.allclose(back, p, rtol=1e-5, atol=1e-300)
