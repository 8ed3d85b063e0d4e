"""fluid-engine-dev_b200 — B200-native (sm_100a CUDA) hybrid-liquid step behind Jet's interfaces.

The directory name is not a Python identifier; load it with `load_package()` from the repo-root helper
`jetcu_loader.py`, which registers it as module `fluid_engine_dev_b200`.
"""
from .jetcu import *  # noqa: F401,F403
from .jetcu import Jetcu, JetcuError, LIB_PATH  # noqa: F401
