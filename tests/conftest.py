import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running CPU check")


def pytest_collection_modifyitems(config, items):
    """A hung kernel or a stuck collective must fail ONE test, not the whole run: every test gets a hard limit when the
    pytest-timeout plugin is present (it is in this image); tests that set their own limit keep it."""
    if not config.pluginmanager.hasplugin("timeout"):
        return
    for item in items:
        if item.get_closest_marker("timeout") is None:
            # method="thread": a test stuck inside a native call (a CUDA sync that never returns) cannot be interrupted by
            # a signal handler; the watchdog thread ends the run instead of letting it hang for good
            item.add_marker(pytest.mark.timeout(900, method="thread"))
