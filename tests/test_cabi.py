"""CPU suite: the C-ABI library loads, exports every symbol include/eorb_b200.h declares, and fails loudly
(no CPU fallback) when there is no CUDA device.  No compute calls are made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "eorb_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(eorb_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from eorb_slam_b200 import api
    L = C.CDLL(api.LIB_PATH)
    names = _declared()
    assert len(names) >= 40
    for n in names:
        assert hasattr(L, n), "missing export " + n
    assert set(api.EXPORTED) == set(names), set(api.EXPORTED) ^ set(names)
    assert api.lib.eorb_version() >= 100


def test_struct_layouts_match_reference_types():
    from eorb_slam_b200 import synth
    assert synth.KEYPOINT_DTYPE.itemsize == 28     # cv::KeyPoint
    assert synth.EVENT_DTYPE.itemsize == 24        # EORB_SLAM::EventData (double, float, float, bool + padding)
    assert synth.EVENT_DTYPE.fields["x"][1] == 8 and synth.EVENT_DTYPE.fields["p"][1] == 16
    assert synth.MATCH_DTYPE.itemsize == 16


def test_host_helpers_need_no_gpu():
    from eorb_slam_b200 import api
    a = np.zeros(32, np.uint8); b = np.full(32, 255, np.uint8)
    assert api.ORBmatcher.DescriptorDistance(a, b) == 256
    assert api.ORBmatcher.DescriptorDistance(a, a) == 0
    a1 = np.array([10, 50, 100, 359, 200, 45], np.float32); a2 = np.zeros(6, np.float32)
    m = np.array([0, 1, 2, 3, 4, -1], np.int32)
    import oracle_lib as O
    n_ref, m_ref = O.rotation_filter(a1, a2, m)
    n, out = api.rotation_filter(a1, a2, m)
    assert n == n_ref and np.array_equal(out, m_ref)
    rng = np.random.default_rng(3)
    for _ in range(50):
        n1 = int(rng.integers(1, 400)); n2 = int(rng.integers(1, 400))
        a1 = rng.uniform(0, 360, n1).astype(np.float32); a2 = rng.uniform(0, 360, n2).astype(np.float32)
        if _ % 2:   # a dominant rotation
            a2[:] = 0; a1 = (rng.normal(40, 12, n1) % 360).astype(np.float32)
        m = np.where(rng.random(n1) < 0.7, rng.integers(0, n2, n1), -1).astype(np.int32)
        n_ref, m_ref = O.rotation_filter(a1, a2, m)
        n, out = api.rotation_filter(a1, a2, m)
        assert n == n_ref and np.array_equal(out, m_ref)


def test_no_cpu_fallback_without_device():
    from eorb_slam_b200 import api
    if api.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(api.EorbError):
        api.ORBextractor(api.ORBxParams())
    with pytest.raises(api.EorbError):
        api.ORBmatcher(0.7)
    with pytest.raises(api.EorbError):
        api.EvImConverter()


def test_constructor_tables_need_no_gpu():
    """eorb_orb_params_tables = ORBextractor::ORBextractor's host arithmetic (ORBextractor.cc:420-489) without a device or a handle;
    equals the oracle's (= the reference's, tests/test_ref_pin.py) tables"""
    import oracle_lib as O
    from eorb_slam_b200 import api
    for (nf, sf, nl, edge, w, h) in [(1000, 1.2, 8, 19, 752, 480), (400, 1.0, 1, 9, 240, 180), (2500, 1.26, 6, -1, 863, 517), (1, 1.5, 2, 25, 300, 200)]:
        p = api._OrbParams(nf, sf, nl, 20, 7, edge, w, h)
        n = C.c_int(0); e = C.c_int(0)
        arrs = [np.zeros(nl, np.float32) for _ in range(4)]; fpl = np.zeros(nl, np.int32)
        rc = api.lib.eorb_orb_params_tables(C.byref(p), C.byref(n), C.byref(e), *[a.ctypes.data_as(C.c_void_p) for a in arrs],
                                            fpl.ctypes.data_as(C.c_void_p))
        assert rc == 0 and n.value == nl
        orc = O.OrbOracle(nf, sf, nl, 20, 7, edge, w, h)
        assert e.value == orc.edge and list(fpl) == list(orc.features_per_level())
        for a, b in zip(arrs, orc.scale_factors()):
            assert a.tobytes() == b.tobytes()
