"""The parity pin: the CPU oracle against the REFERENCE'S OWN CODE.

tests/golden/ref_*.npz hold the outputs of oracle/_ref/libref.so — /root/reference/src/ORBextractor.cc,
src/Event/EventConversion.cc and ORBmatcher::DescriptorDistance compiled unmodified behind the header-only stand-ins of
oracle/ref_mock/ (recipe `make -C oracle ref`; generator tests/golden/make_ref_golden.py).  The oracle must reproduce them
byte for byte.  Where libref itself is present (the authoring container; or its prebuilt .so on the GPU box) it is also run
live, on the committed cases (fixtures not stale) and on fresh random ones."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle_lib as O
import ref_cases as RC
from eorb_slam_b200 import synth


# ------------------------------------------------------------------------------------------------ oracle vs committed fixtures
@pytest.mark.parametrize("name", RC.ORB_NAMES)
def test_oracle_equals_reference_golden(name):
    g, fkw, okw, img, lap = RC.orb_golden(name)
    orc = O.OrbOracle(*RC.orb_args(okw, fkw))
    ret, kps, desc = orc.extract(img, lap, True)
    assert ret == int(g["ret"][0])
    assert kps.tobytes() == g["kps"].tobytes(), "keypoints (coordinates, size, angle bits, response, octave) and their order"
    rows = g["desc_rows_defined"]          # all rows when the margin is >= 19 (the reference reads out of bounds below that)
    assert np.array_equal(desc[rows], g["desc"][rows])
    assert orc.edge == int(g["edge"][0]) and list(orc.features_per_level()) == list(g["features_per_level"])
    for a, b in zip(orc.scale_factors(), (g["scale"], g["inv_scale"], g["sigma2"], g["inv_sigma2"])):
        assert a.tobytes() == b.tobytes()
    assert np.array_equal(orc.umax(), g["umax"])
    for l in range(okw["nlevels"]):
        assert orc.level_size(l) == (int(g["level_w"][l]), int(g["level_h"][l]))
        assert RC.sha(orc.level(l)) == str(g["level_sha"][l]), "pyramid level %d" % l
        if str(g["blur_sha"][l]):
            assert RC.sha(orc.blurred(l)) == str(g["blur_sha"][l]), "blurred level %d" % l
    assert sum(len(orc.candidates(l)[0]) for l in range(okw["nlevels"])) == int(g["candidates"][0])
    rk, kk, _ = orc.extract(img, lap, False)
    assert rk == ret and kk.tobytes() == kps.tobytes()


def test_oracle_equals_reference_fuzz():
    """200 random configurations (sizes, 1..9 levels, scale factors, thresholds, margins incl. the adaptive one, 1..2500 features,
    lapping areas, textured / flat frames), 78 646 keypoints: same return value, keypoint bytes and descriptor bytes"""
    bad = []
    for i, c, ret, n, ksha, dsha in RC.fuzz_cases():
        r, k, d = O.OrbOracle(*RC.fuzz_args(c)).extract(RC.fuzz_frame(c), (c["lap0"], c["lap1"]), bool(c["want"]))
        if not (r == ret and len(k) == n and RC.sha(k) == ksha and (not c["want"] or RC.sha(d) == dsha)):
            bad.append((i, c))
    assert not bad, bad[:3]


def test_oracle_octree_equals_reference():
    """DistributeOctTree alone on 300 candidate sets built to tie (lattices, blobs, few distinct responses)"""
    for i, w, h, N, x, y, resp, osha, on in RC.octree_cases():
        out = O.distribute_octtree(x, y, resp, 16, 16 + w, 16, 16 + h, N)
        assert len(out) == on and RC.sha(out) == osha, ("octree case", i, w, h, N, len(x))


def test_oracle_secondary_api_and_distance_equal_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_secondary.npz"))
    img = synth.make_frame(int(g["frame_seed"][0]))
    orc = O.OrbOracle()
    assert np.array_equal(orc.tracked_desc(img, g["sel"]), g["tracked_desc"])
    moved = np.roll(img, (2, 3), axis=(0, 1))
    assert orc.assign_level_by_best_desc(g["tracked_desc"], moved, g["sel"]).tobytes() == g["assigned"].tobytes()
    q, db, D = g["q"], g["db"], g["dist"].astype(np.int32)
    assert np.array_equal(np.unpackbits(q[:, None, :] ^ db[None, :, :], axis=2).sum(2), D)     # the bit-hack == a popcount
    out = O.hamming_best2(q, db, 256, 2.0)
    assert np.array_equal(out["best_dist"], D.min(1)) and np.array_equal(out["best_idx"], D.argmin(1))


def test_oracle_event_frames_equal_reference():
    for i, s, ev, kw, g in RC.event_cases():
        f, mm, u8 = O.ev_accumulate(ev, s["w"], s["h"], normalize=True, **kw)
        assert RC.sha(f) == str(g["f_sha"][i]), ("float frame, bit for bit", i, s)
        if "f%d" % i in g.files:
            assert f.tobytes() == g["f%d" % i].tobytes()
        if str(g["u_sha"][i]):
            assert u8 is not None and np.array_equal(u8, g["u%d" % i]), ("normalised u8 frame", i)
        else:
            assert u8 is None
    # motion-compensation Jacobian
    evj = synth.make_events(6000, 21)
    for glob in (0, 1):
        jo = np.zeros(6)
        O.lib().orc_ev_mci_jac(O._p(evj), len(evj), 240, 180, 1.0, O._p(np.ascontiguousarray(g["jac_Rt"])), 1.5,
                               O._p(np.ascontiguousarray(g["jac_K"])), 0, glob, O._p(jo))
        jr = g["jac%d" % glob]
        assert np.abs(jo - jr).max() <= 1e-9 * max(1.0, float(np.abs(jr).max()))


def test_reference_report_is_committed(golden_dir):
    rep = json.load(open(os.path.join(golden_dir, "ref_report.json")))
    assert rep["fuzz"]["oracle_equals_libref"] and rep["octree"]["oracle_equals_libref"]
    assert set(rep["cases"]) == set(RC.ORB_NAMES)


# ------------------------------------------------------------------------------------------------ libref live (when present)
def _ref():
    import ref_lib as R
    if not R.available():
        pytest.skip("oracle/_ref/libref.so absent and no reference tree to build it from")
    return R


def test_libref_reproduces_committed_goldens():
    R = _ref()
    for name in RC.ORB_NAMES:
        g, fkw, okw, img, lap = RC.orb_golden(name)
        ret, kps, desc = R.RefOrb(*RC.orb_args(okw, fkw)).extract(img, lap, True)
        rows = g["desc_rows_defined"]
        assert ret == int(g["ret"][0]) and kps.tobytes() == g["kps"].tobytes() and np.array_equal(desc[rows], g["desc"][rows]), name
    for i, c, ret, n, ksha, dsha in list(RC.fuzz_cases())[::10]:
        r, k, d = R.RefOrb(*RC.fuzz_args(c)).extract(RC.fuzz_frame(c), (c["lap0"], c["lap1"]), bool(c["want"]))
        assert r == ret and RC.sha(k) == ksha and (not c["want"] or RC.sha(d) == dsha), (i, c)


def test_libref_vs_oracle_fresh_cases():
    """cases that are NOT in the fixtures: new seeds every frame shape of BASELINE.json's configs"""
    R = _ref()
    for seed, (w, h, args) in enumerate([(752, 480, (1000, 1.2, 8, 20, 7, 19)), (346, 260, (1000, 1.2, 8, 20, 7, 19)),
                                         (240, 180, (500, 1.2, 4, 10, 0, 19)), (640, 480, (1500, 1.2, 8, 20, 7, 19)),
                                         (1241, 376, (2000, 1.2, 8, 20, 7, 19)), (752, 480, (1000, 1.2, 8, 20, 7, -1))]):
        for k in range(3):
            img = synth.make_frame(1000 + 17 * seed + k, w, h)
            a = args + (w, h)
            rr, rk, rd = R.RefOrb(*a).extract(img, (0, 1000) if k else (200, 400), True)
            orr, ok, od = O.OrbOracle(*a).extract(img, (0, 1000) if k else (200, 400), True)
            assert rr == orr and rk.tobytes() == ok.tobytes() and np.array_equal(rd, od), (w, h, k)


def test_libref_events_vs_oracle_fresh_cases():
    R = _ref()
    K = np.array([226.38, 226.15, 173.65, 133.73], np.float32)
    T = np.eye(4, dtype=np.float32)
    c, s = np.cos(0.07), np.sin(0.07)
    T[:3, :3] = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], np.float32); T[:3, 3] = [0.01, 0.0, -0.01]
    for seed in (101, 102):
        ev = synth.make_events(20000, seed, 346, 260)
        for kw in (dict(mode=1), dict(mode=1, pol=True), dict(mode=2, Tcw=T, depth=1.2, K=K), dict(mode=3, se2=[0.02, -1.0, 0.5], K=K)):
            rf, ru = R.ev_accumulate(ev, 346, 260, 1.0, normalize=True, **kw)
            of, _, ou = O.ev_accumulate(ev, 346, 260, 1.0, normalize=True, **kw)
            assert rf.tobytes() == of.tobytes() and np.array_equal(ru, ou), kw


# ------------------------------------------------------------------------------------------------ tracking-thread matchers
def test_oracle_matchers_equal_reference_golden():
    """SearchByProjection (last frame: mono, forward / backward level windows, rectified-stereo column test; local map: mono and stereo;
    relocalisation), SearchForInitialization and SearchByBoW: the oracle reproduces every match array the reference's own function
    bodies produced (tests/golden/ref_guided.npz, 96 cases)"""
    g = RC.guided_golden()
    ncase = 0
    for key, kind, a, kw in RC.guided_cases():
        r = RC.guided_run(O, kind, a, kw)
        assert r[0] == int(g[key + "_n"][0]) and np.array_equal(r[1], g[key]), key
        if kind == "init":
            assert r[2].tobytes() == g[key + "_prev"].tobytes(), key
        ncase += 1
    assert ncase == sum(k.endswith("_n") for k in g.files)


def test_libref_matchers_vs_oracle_fresh_cases():
    R = _ref()
    g = RC.guided_golden()
    for j, (key, kind, a, kw) in enumerate(RC.guided_cases()):
        if j % 5 == 0:                                   # the fixtures are not stale
            r = RC.guided_run(R, kind, a, kw)
            assert r[0] == int(g[key + "_n"][0]) and np.array_equal(r[1], g[key]), key
    rng = np.random.default_rng(77)
    for seed in (901, 902, 903):                         # cases that are not in the fixtures
        c = synth.make_projection_case(700, 650, seed, zero_obs_frac=0.2)
        ur = RC._stereo_u_right(rng, c["kps2"])
        a = (c["x3Dc"], c["valid1"], c["obs1"], c["kps1"], c["descMP"], c["kps2"], c["desc2"], c["bounds"], c["K"], c["scale_factors"])
        for mode in (0, 1, 2):
            kw = dict(th=15.0, check_ori=True, level_mode=mode, mbf=40.0, u_right2=ur)
            o = O.search_by_projection_ex(*a, **kw); r = R.search_by_projection(*a, **kw)
            assert o[0] == r[0] and np.array_equal(o[1], r[1]), (seed, mode)
        cb = RC.bow_case(700, 650, seed, levelsup=2)
        o = O.search_by_bow(*cb, 0.7, True); r = R.search_by_bow(*cb, 0.7, True)
        assert o[0] == r[0] and np.array_equal(o[1], r[1]), seed
        # keyframe-keyframe form (ORBmatcher.cc:833-990): the second side needs good map points too, strict TH_LOW, result indexed by side 1
        k1, d1, v1, fv1, k2, d2, fv2 = cb
        v2 = (np.random.default_rng(seed).random(len(k2)) < 0.75).astype(np.uint8)
        for ratio, ori in ((0.7, True), (0.9, False)):
            o = O.search_by_bow_kf(k1, d1, v1, fv1, k2, d2, v2, fv2, ratio, ori); r = R.search_by_bow_kf(k1, d1, v1, fv1, k2, d2, v2, fv2, ratio, ori)
            assert o[0] == r[0] and np.array_equal(o[1], r[1]) and o[0] > 0, (seed, ratio, ori)


# ------------------------------------------------------------------------------------------------ keyframe-side searches
def test_oracle_keyframe_searches_equal_reference_golden():
    """SearchByProjection(KeyFrame*, Scw, ...) (both overloads), Fuse (both overloads) and SearchBySim3: the host part of
    tests/kf_cases.py + the oracle's matching core reproduce what the reference's own function bodies produced
    (tests/golden/ref_guided_kf.npz, 30 cases: match tables, the ordered Add / Replace events of Fuse, nmatches)"""
    import kf_cases as KC
    g = KC.kf_golden()
    ncase = 0
    for key, kind, ov, c in KC.kf_cases():
        r = KC.run_composed(O.search_windows, kind, ov, c)
        assert r[0] == int(g[key + "_n"][0]) and r[1].shape == g[key].shape and np.array_equal(r[1], g[key]), key
        assert r[0] > 50, key                                    # the cases are not vacuous
        ncase += 1
    assert ncase == sum(k.endswith("_n") for k in g.files)


def test_libref_keyframe_searches_reproduce_golden():
    import kf_cases as KC
    R = _ref()
    g = KC.kf_golden()
    for key, kind, ov, c in KC.kf_cases():
        r = KC.run_ref(R, kind, ov, c)
        assert r[0] == int(g[key + "_n"][0]) and np.array_equal(r[1], g[key]), key


def test_oracle_search_for_triangulation_equals_reference_golden():
    """ORBmatcher::SearchForTriangulation (:975-1214, pinhole keyframes): the oracle reproduces the pairs the reference's own function body and
    Pinhole::epipolarConstrain produced (tests/golden/ref_triangulation.npz, 40 cases: epipole at infinity / inside the image, right-image
    columns, bOnlyStereo, bCoarse, rotation check off); libref, where it exists, still reproduces the file, F12 and epipole included"""
    import kf_cases as KC
    g = KC.tri_golden()
    ncase = 0
    import ref_lib
    R = ref_lib if ref_lib.available() else None
    for key, a, kw in KC.tri_cases():
        o = KC.tri_compose(O.search_for_triangulation, a, kw, g[key + "_F"], g[key + "_ep"])
        assert o[0] == int(g[key + "_n"][0]) and np.array_equal(o[1], g[key]), key
        if R is not None and ncase % 4 == 0:
            r = R.search_for_triangulation(*a, **kw)
            assert r[0] == o[0] and np.array_equal(r[1], g[key]) and r[2].tobytes() == g[key + "_F"].tobytes() and r[3].tobytes() == g[key + "_ep"].tobytes(), key
        ncase += 1
    assert ncase == sum(k.endswith("_n") for k in g.files) == 40
    # nothing to match: empty sides, no common node, every feature holding a map point
    key, a, kw = next(iter(KC.tri_cases(1)))
    k1, d1, h1, u1, fv1, k2, d2, h2, u2, fv2, K, t1, t2, sc, sg = a
    F, ep = g[key + "_F"], g[key + "_ep"]
    f1 = O.triangulation_flags(h1, None, False); f2 = O.triangulation_flags(h2, None, False)
    n, m = O.search_for_triangulation(k1, d1, np.zeros_like(f1), fv1, k2, d2, f2, fv2, F, ep, sc, sg)
    assert n == 0 and (m == -1).all()
    far = (fv2[0] + np.uint32(1 << 20), fv2[1], fv2[2])
    n, m = O.search_for_triangulation(k1, d1, f1, fv1, k2, d2, f2, far, F, ep, sc, sg)
    assert n == 0 and (m == -1).all()


def test_oracle_search_windows_edge_cases():
    """no queries / no keypoints / every point gated out / every slot held"""
    import kf_cases as KC
    key, kind, ov, c = next(iter(KC.kf_cases(1)))
    q, _ = KC.host_windows(c["pt8"], c["level1"], np.zeros(len(c["level1"]), bool), c["bounds"], c["K"], c["scale_factors"], 6, 0)
    n, bi, bd, m2 = O.search_windows(q[:0], None, c["descMP"][:0], c["kps2"], c["desc2"], None, None, c["bounds"])
    assert n == 0 and len(bi) == 0 and (m2 == -1).all()
    n, bi, bd, m2 = O.search_windows(q, None, c["descMP"], c["kps2"][:0], c["desc2"][:0], None, None, c["bounds"])
    assert n == 0 and (bi == -1).all() and (bd == 256).all()
    q0 = q.copy(); q0["r"] = -1
    n, bi, bd, _ = O.search_windows(q0, None, c["descMP"], c["kps2"], c["desc2"], None, None, c["bounds"])
    assert n == 0 and (bi == -1).all() and (bd == 256).all()
    held = np.ones(len(c["kps2"]), np.uint8)
    for blocking in (False, True):
        n, bi, bd, _ = O.search_windows(q, None, c["descMP"], c["kps2"], c["desc2"], held, None, c["bounds"], blocking=blocking, th_high=100)
        assert n == 0 and (bi == -1).all() and (bd == 256).all()
    n_nb, bi_nb, _, _ = O.search_windows(q, None, c["descMP"], c["kps2"], c["desc2"], None, None, c["bounds"], blocking=False, th_high=100)
    n_b, bi_b, _, m2 = O.search_windows(q, None, c["descMP"], c["kps2"], c["desc2"], None, None, c["bounds"], blocking=True, th_high=100)
    assert n_b <= n_nb and len(set(bi_b[bi_b >= 0])) == n_b and n_b == int((m2 >= 0).sum())     # blocking: one point per keypoint


# ------------------------------------------------------------------------------------------------ KannalaBrandt8 motion compensation
KB8_K = (226.38018519795807, 226.15002947047415, 173.6470807871759, 133.73271487507847)          # Examples/Event/EvMVSEC.yaml:53-63
KB8_D = (-0.048031442223833355, 0.011330957517194437, -0.055378166304281135, 0.021500973881459395)


def kb8_cases():
    T = np.eye(4, dtype=np.float32)
    c, s = np.cos(0.07), np.sin(0.07)
    T[:3, :3] = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], np.float32); T[:3, 3] = [0.01, 0.0, -0.01]
    for seed in (201, 202):
        ev = synth.make_events(20000, seed, 346, 260)
        yield "se3_%d" % seed, ev, dict(mode=2, Tcw=T, depth=1.2, K=KB8_K, kb8=KB8_D)
        yield "se2_%d" % seed, ev, dict(mode=3, se2=[0.02, -1.0, 0.5], K=KB8_K, kb8=KB8_D)
        yield "se2s_%d" % seed, ev, dict(mode=3, se2=[0.02, -1.0, 0.5, 0.9], K=KB8_K, kb8=KB8_D)


def test_oracle_kb8_motion_compensation_equals_reference(golden_dir):
    """ev2mci_gg_f through the reference's own KannalaBrandt8::unproject / project (src/CameraModels/KannalaBrandt8.cpp:86-129, 163-190, cut
    into libref): float frames bit for bit (tests/golden/ref_events_kb8.npz, written by this test's generator below when libref exists)"""
    path = os.path.join(golden_dir, "ref_events_kb8.npz")
    import ref_lib as R
    if R.available() and os.environ.get("EORB_WRITE_GOLDEN") == "1":
        out = {}
        for key, ev, kw in kb8_cases():
            f, u8 = R.ev_accumulate(ev, 346, 260, 1.0, normalize=True, **kw)
            out[key + "_sha"] = np.array(RC.sha(f)); out[key + "_u8sha"] = np.array(RC.sha(u8)); out[key + "_max"] = np.array([f.max()], np.float32)
        np.savez_compressed(path, **out)
    g = np.load(path)
    for key, ev, kw in kb8_cases():
        f, _, u8 = O.ev_accumulate(ev, 346, 260, 1.0, normalize=True, **kw)
        assert RC.sha(f) == str(g[key + "_sha"]) and RC.sha(u8) == str(g[key + "_u8sha"]), key
        pin, _, _ = O.ev_accumulate(ev, 346, 260, 1.0, **{k: v for k, v in kw.items() if k != "kb8"})
        assert np.abs(f - pin).max() > 0.5, "the fisheye model must differ from the pinhole one on this camera"
        if R.available():
            rf, ru = R.ev_accumulate(ev, 346, 260, 1.0, normalize=True, **kw)
            assert rf.tobytes() == f.tobytes() and np.array_equal(ru, u8), key


# ------------------------------------------------------------------------------------------------ ELK_Tracker post-LK bookkeeping
def _lk_refine_case(seed, n=700, w=240, h=180):
    rng = np.random.default_rng(seed)
    ref = np.zeros(n, O.KEYPOINT_DTYPE)
    ref["x"] = rng.uniform(0, w, n).astype(np.float32); ref["y"] = rng.uniform(0, h, n).astype(np.float32)
    ref["size"] = 31; ref["angle"] = rng.uniform(0, 360, n).astype(np.float32); ref["response"] = rng.integers(1, 200, n)
    ref["octave"] = rng.integers(0, 3, n); ref["class_id"] = rng.integers(-1, 50, n)
    cur = np.stack([ref["x"], ref["y"]], 1) + rng.normal(0, 6, (n, 2)).astype(np.float32)
    edge = rng.integers(0, n, 40)   # points exactly on and just past the image edges (isInImage :99-102: 0 <= x < W)
    cur[edge[:10], 0] = 0.0; cur[edge[10:20], 0] = np.float32(w); cur[edge[20:30], 1] = np.nextafter(np.float32(h), np.float32(0)); cur[edge[30:], 1] = -0.0
    status = (rng.uniform(0, 1, n) < 0.8).astype(np.uint8)
    status[rng.integers(0, n, 5)] = 2   # the test is status == 1, not != 0
    return cur.astype(np.float32), status, ref


def test_oracle_lk_refine_equals_reference():
    """oracle == the reference's own refineTrackedPts / refineFirstOctaveLevel bodies, byte for byte, fresh and carried-over vectors"""
    R = _ref()
    for seed in range(6):
        cur, st, ref = _lk_refine_case(seed)
        for init in (False, True):
            a = R.lk_refine(cur, st, ref, 240, 180, init)
            b = O.lk_refine(cur, st, ref, 240, 180, init)
            assert a[0] == b[0] and a[1].tobytes() == b[1].tobytes() and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
            assert a[4].tobytes() == b[4].tobytes()
            # second frame with the vectors of the first (the event trackers keep them across frames)
            cur2, st2, _ = _lk_refine_case(100 + seed)
            a2 = R.lk_refine(cur2, st2, ref, 240, 180, init, a[2], a[3])
            b2 = O.lk_refine(cur2, st2, ref, 240, 180, init, b[2], b[3])
            assert a2[0] == b2[0] and a2[1].tobytes() == b2[1].tobytes() and np.array_equal(a2[2], b2[2]) and np.array_equal(a2[3], b2[3])
            assert a2[4].tobytes() == b2[4].tobytes()
