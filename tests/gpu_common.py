"""Shared fixtures for the -m gpu tests: the jetcu context (product, through the C ABI) and the checker oracle."""
import os

import pytest

from oracle_api import have_port, have_ref, load_port, load_ref


def jetcu_or_fail():
    """The CUDA library must load on a GPU box — a missing extension is a hard failure, never a skip."""
    from fluid_engine_dev_b200 import Jetcu
    return Jetcu(0)


def checker():
    """Prefer the unmodified reference when its prebuilt shim travelled with the repo; else the C port."""
    if have_ref() and os.environ.get("JETCU_TEST_ORACLE", "") != "port":
        o = load_ref()
    elif have_port():
        o = load_port()
    else:
        pytest.fail("no oracle library built: run `make -C oracle port` (and `make -C oracle ref` where possible)")
    o.set_num_threads(0)
    return o
