"""The addressing of the peer-memory halo exchange (csrc/p2p_ops.cuh: which mailbox side and slot a rank's boundary planes go to,
which flag announces them, where the halo planes are read from) executed for a chain of ranks in host memory through
tests/hostcheck — the same source comm.cu's kernels use.  Complements tests/test_p2p_protocol_model.py (the protocol under every
interleaving): here the pointer arithmetic, there the synchronisation."""
import ctypes as C

import pytest

from test_hostcheck_collider import _load_hostcheck


@pytest.fixture(scope="module")
def hc():
    lib = _load_hostcheck()
    if lib is None:
        pytest.skip("nvcc not available to build tests/hostcheck")
    lib.hc_p2p_chain.restype = C.c_longlong
    lib.hc_p2p_chain.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_uint64, C.c_int]
    return lib


@pytest.mark.parametrize("reverse", [0, 1])
@pytest.mark.parametrize("world,T,plane_bytes,depth", [(2, 4, 64, 1), (3, 8, 256, 2), (4, 8, 1024, 4), (8, 4, 48, 1), (8, 16, 512, 8),
                                                       (2, 32, 4096, 16)])
def test_halo_planes_arrive_where_they_belong(hc, world, T, plane_bytes, depth, reverse):
    """5 rounds (slots alternate, message numbers advance): every halo plane holds the neighbour's plane of THIS round, owned planes
    and everything beyond the halos are untouched."""
    slot = max(depth * plane_bytes, 64)
    assert hc.hc_p2p_chain(world, T, plane_bytes, depth, 5, slot, reverse) == 0


def test_slot_larger_than_message(hc):
    assert hc.hc_p2p_chain(4, 8, 128, 2, 4, 1 << 16, 0) == 0
