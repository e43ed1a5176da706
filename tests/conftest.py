import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu() -> bool:
    try:
        from eorb_slam_b200 import api
        return api.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # a gpu-marked test that lands on a box without a device is an error in the product path, not a skip:
    # the library has no CPU fallback.  Only skip when the user did not ask for gpu tests explicitly.
    if _has_gpu():
        return
    markexpr = config.getoption("-m") or ""
    if "gpu" in markexpr and "not gpu" not in markexpr:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def host_model():
    """tests/host_model/_hm.so: the product's shared host/device sources compiled as plain C++"""
    import ctypes as C
    d = os.path.join(ROOT, "tests", "host_model")
    so = os.path.join(d, "_hm.so")
    srcs = [os.path.join(d, "host_model.cc"), os.path.join(ROOT, "eorb_slam_b200", "csrc", "octree_core.cuh"),
            os.path.join(ROOT, "eorb_slam_b200", "csrc", "eorb_math.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                               "-o", so, srcs[0]])
    L = C.CDLL(so)
    L.hm_octree.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    L.hm_fast_max_arc_min.argtypes = [C.c_int, C.c_void_p]
    L.hm_fast_atan2.argtypes = [C.c_float, C.c_float]; L.hm_fast_atan2.restype = C.c_float
    L.hm_brief_offset.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    L.hm_resize_px.argtypes = [C.c_int] * 8
    L.hm_hamming.argtypes = [C.c_void_p, C.c_void_p]
    return L
