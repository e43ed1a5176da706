"""CPU suite: the oracle against the committed golden vectors (generated from cv2 4.13.0 and an independent
cv2-assisted restatement by tests/golden/make_golden.py).  The reference ships no tests of its own."""
import ast
import os
import zlib

import numpy as np
import pytest

import oracle_lib as O
from eorb_slam_b200 import synth


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def test_primitives_match_cv2_goldens(golden_dir):
    g = np.load(os.path.join(golden_dir, "prims.npz"))
    small = g["small"]
    assert np.array_equal(O.resize_linear(small, 81, 109), g["small_resize_81x109"])
    assert np.array_equal(O.resize_linear(small, 50, 77), g["small_resize_50x77"])
    assert np.array_equal(O.border_reflect101(small, 19), g["small_border19"])
    assert np.array_equal(O.gauss5(small), g["small_blur"])
    tile = g["fast_tile"]
    for t in (0, 7, 20):
        xs, ys, sc = O.fast(tile, t, True)
        assert np.array_equal(np.stack([xs, ys, sc], 1).reshape(-1, 3), g["fast_tile_t%d" % t])
    got = np.array([O.fast_atan2(float(y), float(x)) for y, x in zip(g["atan_y"], g["atan_x"])], np.float32)
    assert np.array_equal(got.view(np.uint32), g["atan_out"].view(np.uint32))
    m = O.hamming_best2(g["bf_q"], g["bf_db"], th=50, ratio=0.7)
    assert np.array_equal(np.stack([m["best_dist"], m["best_idx"], m["second_dist"]], 1), g["bf_out"])


def test_descriptor_distance_is_popcount():
    rng = np.random.default_rng(7)
    a = rng.integers(0, 256, (200, 32), dtype=np.uint8); b = rng.integers(0, 256, (200, 32), dtype=np.uint8)
    for i in range(200):
        assert O.descriptor_distance(a[i], b[i]) == int(np.unpackbits(a[i] ^ b[i]).sum())
    assert O.descriptor_distance(a[0], a[0]) == 0
    assert O.descriptor_distance(np.zeros(32, np.uint8), np.full(32, 255, np.uint8)) == 256


ORB_CASES = ["cfg1_seed0", "cfg1_seed1_stereo", "cfg1_flat", "mvsec_346x260", "ethz_240x180_e9", "ev_single_level"]


@pytest.mark.parametrize("name", ORB_CASES)
def test_orb_pipeline_matches_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "orb_%s.npz" % name))
    fkw = ast.literal_eval(str(g["frame_kw"])); okw = ast.literal_eval(str(g["orb_kw"]))
    img = synth.make_frame(**fkw)
    assert crc(img) == int(g["frame_crc"][0]), "synthetic generator drifted"
    orc = O.OrbOracle(okw["nfeatures"], okw["scale_factor"], okw["nlevels"], okw["ini_th"], okw["min_th"], okw["edge"], fkw["w"], fkw["h"])
    ret, kps, desc = orc.extract(img, tuple(g["lapping"]), True)
    assert ret == int(g["ret"][0])
    assert list(orc.features_per_level()) == list(g["quota"])
    assert [crc(orc.level(l)) for l in range(okw["nlevels"])] == [int(c) for c in g["pyr_crc"]]
    assert [len(orc.candidates(l)[0]) for l in range(okw["nlevels"])] == list(g["ncand"])
    assert orc.fallback_cells() == int(g["nfallback"][0])
    assert kps.tobytes() == g["kps"].tobytes()
    assert np.array_equal(desc, g["desc"])
    # keypoints-only overload returns the same keypoints
    ret2, kps2, _ = orc.extract(img, tuple(g["lapping"]), False)
    assert ret2 == ret and kps2.tobytes() == kps.tobytes()


def test_orb_ctor_tables():
    orc = O.OrbOracle(1000, 1.2, 8, 20, 7, 19, 752, 480)
    assert list(orc.features_per_level()) == [217, 181, 151, 126, 105, 87, 73, 60]
    assert list(orc.umax()) == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert [orc.level_size(l) for l in range(0)] == []
    s, inv, s2, inv2 = orc.scale_factors()
    assert s[0] == 1.0 and abs(s[7] - 1.2 ** 7) < 1e-5
    # adaptive edge threshold: 19*W/752 forced odd (ORBextractor.cc:481-485)
    assert O.OrbOracle(400, 1.2, 4, 10, 0, -1, 240, 180).edge == 5
    assert O.OrbOracle(400, 1.2, 4, 10, 0, -1, 752, 480).edge == 19
    assert O.OrbOracle(400, 1.0, 1, 0, 0, 9, 240, 180).features_per_level()[0] == 400


def test_orb_edge_cases():
    orc = O.OrbOracle()
    ret, kps, desc = orc.extract(np.zeros((480, 752), np.uint8))
    assert ret == 0 and len(kps) == 0
    ret, kps, desc = orc.extract(np.zeros((0, 0), np.uint8))
    assert ret == -1
    # stereo lapping {0,0}: forward order, return value = number of keypoints outside the lapping area
    img = synth.make_frame(9)
    r_mono, k_mono, d_mono = orc.extract(img, (0, 1000))
    r_st, k_st, d_st = orc.extract(img, (0, 0))
    assert r_mono == 0 and r_st == len(k_st) == len(k_mono)
    assert k_st.tobytes() == k_mono[::-1].tobytes() and np.array_equal(d_st, d_mono[::-1])
    assert len(k_mono) >= 1000 and len(k_mono) <= 1000 + 3 * 8


def test_events_match_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "events_2000.npz"))
    ev = synth.make_events(2000, seed=int(g["seed"][0]), w=240, h=180)
    img, (mn, mx), u8 = O.ev_accumulate(ev, 240, 180, 1.0, mode=1, normalize=True)
    assert np.array_equal(img, g["gauss"]) and np.array_equal(u8, g["gauss_u8"])
    assert float(np.abs(img - g["py_gauss"]).max()) <= 2e-6 * float(img.max())
    imgn, _, _ = O.ev_accumulate(ev, 240, 180, 1.0, mode=0, normalize=True)
    assert np.array_equal(imgn, g["nearest"])
    img3, _, _ = O.ev_accumulate(ev, 240, 180, 1.0, mode=2, Tcw=g["Trot"], depth=1.0, K=g["K"])
    assert np.array_equal(img3, g["se3"])
    img4, _, _ = O.ev_accumulate(ev, 240, 180, 1.0, mode=3, K=g["K"], se2=g["se2_params"])
    assert np.array_equal(img4, g["se2"])
    # mass conservation of the splat: an interior event deposits sum_ij exp(..)/(2 pi s^2) ~ 1
    one = np.zeros(1, synth.EVENT_DTYPE); one["x"] = 100.3; one["y"] = 90.7; one["ts"] = 1.0
    im1, _, _ = O.ev_accumulate(one, 240, 180, 1.0, mode=1)
    assert abs(float(im1.sum()) - 1.0) < 2e-2 and np.count_nonzero(im1) == 49


def test_rotation_filter_quirk():
    # factor = 1/30: rot in [0,360) only reaches bins 0..12
    a1 = np.array([10, 50, 100, 359, 200, 45], np.float32); a2 = np.array([0, 0, 0, 0, 0, 0], np.float32)
    m = np.array([0, 1, 2, 3, 4, -1], np.int32)
    n, out = O.rotation_filter(a1, a2, m)
    assert n == int((out >= 0).sum()) and out[5] == -1
