"""The C restatement (oracle/libjetport.so) against the committed golden vectors produced by the unmodified
reference.  This is the check that still runs where /root/reference (and oracle/_ref) are absent."""
import pytest

import golden_cases as gc
from oracle_api import APIC, FLIP, PIC, have_port, load_port

pytestmark = pytest.mark.skipif(not have_port(), reason="oracle/libjetport.so not built (make -C oracle port)")


@pytest.fixture(scope="module")
def port():
    o = load_port()
    o.set_num_threads(0)
    return o


def test_linear_ops(port):
    gc.linear_ops(port)


def test_mgpcg_poisson(port):
    gc.mgpcg_poisson(port, exact_bits=True)


@pytest.mark.parametrize("n", [16, 32])
def test_pressure_column_box(port, n):
    gc.pressure_column_box(port, n, exact_bits=True)


def test_bin_points(port):
    gc.bin_points(port)


@pytest.mark.parametrize("mode", [PIC, FLIP, APIC])
def test_hybrid(port, mode):
    gc.hybrid(port.hybrid, mode, exact_bits=True)


@pytest.mark.parametrize("name", sorted(gc.COLLIDER_CASES))
def test_hybrid_with_collider(port, name):
    gc.collider(port.hybrid, name, exact_bits=True)


def test_smoke(port):
    gc.smoke(port.smoke, exact_bits=True)


@pytest.mark.parametrize("advection", ["cubic", "linear"])
def test_smoke_example1(port, advection):
    gc.smoke_example1(port.smoke, exact_bits=True, advection=advection)


def test_emission(port):
    gc.emission(port.hybrid)

