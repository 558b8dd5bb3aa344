"""rover_slam_b200: B200-native SuperPoint + LightGlue front end of Rover-SLAM.

The product is the C-ABI library `librover_fe.so` (include/rover_fe.h) built from csrc/ for sm_100a.
This package is only the thin ctypes binding used by the tests, bench.py and __graft_entry__.py; it
fails loudly when the CUDA library is missing -- there is no CPU fallback.
"""
from .api import FrontEnd, RoverFeError, lib_path, load_library, exported_symbols  # noqa: F401
