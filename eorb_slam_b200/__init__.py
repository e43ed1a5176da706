"""eorb_slam_b200 — B200-native (sm_100a) front-end hot path of EORB-SLAM.

Only what the hot path needs lives here: `csrc/` (hand-written CUDA kernels + the C-ABI library
`libeorb_b200.so`), `shim/` (C++ classes with the reference's ORBextractor / ORBmatcher / EvImConverter
signatures forwarding to the C ABI) and a thin ctypes host mirror (`api.py`) used by tests and bench.py.
There is NO CPU fallback: importing `eorb_slam_b200.api` fails loudly if the CUDA library is missing.
"""
__version__ = "0.1.0"
