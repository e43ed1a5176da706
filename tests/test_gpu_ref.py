"""GPU parity against the REFERENCE-PINNED fixtures: the CUDA path (through the C ABI) must reproduce what the reference's
own ORBextractor.cc / EventConversion.cc / DescriptorDistance produced (tests/golden/ref_*.npz, written from
oracle/_ref/libref.so by tests/golden/make_ref_golden.py).  Integer work byte for byte; event frames within 1e-4 of peak."""
import os

import numpy as np
import pytest

import ref_cases as RC
from eorb_slam_b200 import synth

pytestmark = pytest.mark.gpu
REL_TOL = 1e-4
LOUD_LIMITS = ("shared-memory budget exceeded", "is empty", "edge threshold")


def _api():
    from eorb_slam_b200 import api
    return api


def _mk(api, a):
    return api.ORBextractor(api.ORBxParams(a[0], a[1], a[2], a[3], a[4], a[5], (a[6], a[7])))


@pytest.mark.parametrize("name", RC.ORB_NAMES)
def test_cuda_orb_equals_reference_golden(name):
    api = _api()
    g, fkw, okw, img, lap = RC.orb_golden(name)
    ex = _mk(api, RC.orb_args(okw, fkw))
    ret, kps, desc = ex(img, None, lap, True)
    assert ret == int(g["ret"][0])
    assert kps.tobytes() == g["kps"].tobytes(), "keypoint bytes and order"
    rows = g["desc_rows_defined"]
    assert np.array_equal(desc[rows], g["desc"][rows]), "descriptors"
    assert ex.edge_threshold() == int(g["edge"][0]) and list(ex.features_per_level()) == list(g["features_per_level"])
    assert ex.GetScaleFactors().tobytes() == g["scale"].tobytes() and ex.GetInverseScaleFactors().tobytes() == g["inv_scale"].tobytes()
    assert ex.GetScaleSigmaSquares().tobytes() == g["sigma2"].tobytes() and ex.GetInverseScaleSigmaSquares().tobytes() == g["inv_sigma2"].tobytes()
    ncand = 0
    for l in range(okw["nlevels"]):
        assert ex.level_size(l) == (int(g["level_w"][l]), int(g["level_h"][l]))
        assert RC.sha(ex.pyramid_level(l)) == str(g["level_sha"][l]), "pyramid level %d" % l
        if str(g["blur_sha"][l]):
            assert RC.sha(ex.debug_blurred(l)) == str(g["blur_sha"][l]), "blurred level %d" % l
        ncand += len(ex.debug_candidates(l)[0])
    assert ncand == int(g["candidates"][0]), "FAST candidates before the octree"
    r2, k2, d2 = ex(img, None, lap, False)
    assert r2 == ret and d2 is None and k2.tobytes() == kps.tobytes()


def test_cuda_orb_equals_reference_fuzz():
    api = _api()
    bad, loud, done = [], 0, 0
    for i, c, ret, n, ksha, dsha in RC.fuzz_cases():
        try:
            r, k, d = _mk(api, RC.fuzz_args(c))(RC.fuzz_frame(c), None, (c["lap0"], c["lap1"]), bool(c["want"]))
        except Exception as e:    # documented capacity limits fail loudly: not a mismatch, but counted
            if any(m in repr(e) for m in LOUD_LIMITS):
                loud += 1
                continue
            raise
        done += 1
        if not (r == ret and len(k) == n and RC.sha(k) == ksha and (not c["want"] or RC.sha(d) == dsha)):
            bad.append((i, c, len(k), n))
    assert not bad, bad[:3]
    assert done >= 180, "only %d of 200 cases ran (%d loud capacity errors)" % (done, loud)


def test_cuda_batch_equals_reference_golden():
    """the batched entry point (configs[2]'s path) on a stack holding the golden frames, each frame checked against its fixture"""
    api = _api()
    names = ["cfg1_seed0", "cfg1_seed1_stereo", "cfg1_flat"]
    gs = [RC.orb_golden(n) for n in names]
    frames = np.stack([x[3] for x in gs] * 3)
    ex = api.ORBextractor(api.ORBxParams(), 0, len(frames))
    for lap_i, lap in enumerate([(0, 1000), gs[1][4]]):
        kps, desc, n, mono = ex.extract_batch(frames, lap, True)
        for f in range(len(frames)):
            g = gs[f % 3][0]
            if tuple(int(v) for v in g["lapping"]) != tuple(lap):
                continue
            assert mono[f] == int(g["ret"][0]) and n[f] == len(g["kps"])
            assert kps[f][:n[f]].tobytes() == g["kps"].tobytes() and np.array_equal(desc[f][:n[f]], g["desc"])


def test_cuda_secondary_api_and_distance_equal_reference(golden_dir):
    api = _api()
    g = np.load(os.path.join(golden_dir, "ref_secondary.npz"))
    img = synth.make_frame(int(g["frame_seed"][0]))
    ex = api.ORBextractor(api.ORBxParams())
    assert np.array_equal(ex.ComputeTrackedKPtsDesc(img, g["sel"]), g["tracked_desc"])
    moved = np.roll(img, (2, 3), axis=(0, 1))
    assert ex.AssignKPtLevelByBestDesc(g["tracked_desc"], moved, g["sel"]).tobytes() == g["assigned"].tobytes()
    q, db, D = g["q"], g["db"], g["dist"].astype(np.int32)
    assert all(api.ORBmatcher.DescriptorDistance(q[i], db[j]) == D[i, j] for i in range(0, 64, 5) for j in range(0, 512, 31))
    m = api.ORBmatcher(2.0)
    m.set_db(db)
    out = m.search(q, th=256)
    assert np.array_equal(out["best_dist"], D.min(1)) and np.array_equal(out["best_idx"], D.argmin(1))
    second = np.sort(D, axis=1)[:, 1]
    assert np.array_equal(out["second_dist"], second)


def test_cuda_event_frames_equal_reference():
    api = _api()
    cv = api.EvImConverter(0, 1, 100000, 346, 260)
    checked = 0
    for i, s, ev, kw, g in RC.event_cases():
        if "f%d" % i not in g.files:
            continue
        ref = g["f%d" % i]
        cam = tuple(kw["K"]) if kw["K"] is not None else None
        if s["mode"] == 0:
            img = cv.ev2im(ev, s["w"], s["h"], kw["pol"], False)
        elif s["mode"] == 1:
            img = cv.ev2im_gauss(ev, s["w"], s["h"], s["sigma"], kw["pol"], False)
        elif s["mode"] == 2:
            img = cv.ev2mci_gg_f(ev, cam, kw["Tcw"], kw["depth"], s["w"], s["h"], s["sigma"], kw["pol"], False)
        else:
            img = cv.ev2mci_gg_f_2d(ev, cam, kw["se2"], s["w"], s["h"], s["sigma"], kw["pol"], False)
        peak = float(np.abs(ref).max())
        assert float(np.abs(img - ref).max()) <= REL_TOL * max(peak, 1e-12), (i, s)
        if not kw["pol"] and "u%d" % i in g.files:     # pol = false: the running extremes equal the final ones
            if s["mode"] == 1:
                u8 = cv.ev2im_gauss(ev, s["w"], s["h"], s["sigma"], False, True)
            elif s["mode"] == 2:
                u8 = cv.ev2mci_gg_f(ev, cam, kw["Tcw"], kw["depth"], s["w"], s["h"], s["sigma"], False, True)
            elif s["mode"] == 3:
                u8 = cv.ev2mci_gg_f_2d(ev, cam, kw["se2"], s["w"], s["h"], s["sigma"], False, True)
            else:
                u8 = cv.ev2im(ev, s["w"], s["h"], False, True)
            assert int(np.abs(u8.astype(int) - g["u%d" % i].astype(int)).max()) <= 1, (i, s)
        checked += 1
    assert checked >= 4


def _guided_cuda(api, gm, kind, a, kw):
    """one case of tests/ref_cases.py::guided_cases through the C ABI (host-pointer entry points)"""
    if kind == "proj":
        gm.mbCheckOrientation = kw["check_ori"]
        if kw["level_mode"] == 0 and kw["u_right2"] is None:
            return gm.SearchByProjection(*a, th=kw["th"])
        return gm.SearchByProjectionStereo(*a, th=kw["th"], level_mode=kw["level_mode"], mbf=kw["mbf"], u_right2=kw["u_right2"])
    if kind == "reloc":
        gm.mbCheckOrientation = kw["check_ori"]
        return gm.SearchByProjectionReloc(*a, th=kw["th"], ORBdist=kw["orb_dist"])
    if kind == "map":
        gm.mfNNratio = kw["nnratio"]
        pts, xr, dmp, k2, d2, held, ur, b, sf = a
        if xr is None:
            return gm.SearchByProjectionMapPoints(pts, dmp, k2, d2, held, b, sf, th=kw["th"], bFarPoints=kw["far_points"], thFarPoints=kw["th_far"])
        return gm.SearchByProjectionMapPointsStereo(pts, xr, dmp, k2, d2, held, ur, b, sf, th=kw["th"], bFarPoints=kw["far_points"],
                                                    thFarPoints=kw["th_far"])
    if kind == "init":
        gm.mfNNratio = kw["nnratio"]; gm.mbCheckOrientation = kw["check_ori"]
        return gm.SearchForInitialization(*a, windowSize=kw["window_size"])
    if kind == "bow":
        gm.mfNNratio = kw["nnratio"]; gm.mbCheckOrientation = kw["check_ori"]
        return gm.SearchByBoW(*a)
    raise KeyError(kind)


def test_cuda_matchers_equal_reference_golden():
    """every tracking-thread matcher on the device against the match arrays the REFERENCE'S OWN function bodies produced
    (tests/golden/ref_guided.npz: last-frame search mono / forward / backward / rectified stereo, relocalisation search, local-map
    search mono / stereo, SearchForInitialization, SearchByBoW; 96 cases), bit for bit"""
    api = _api()
    g = RC.guided_golden()
    gm = api.GuidedMatcher()
    bad = []
    for key, kind, a, kw in RC.guided_cases():
        r = _guided_cuda(api, gm, kind, a, kw)
        ok = r[0] == int(g[key + "_n"][0]) and np.array_equal(r[1], g[key])
        if ok and kind == "init":
            ok = np.asarray(r[2], np.float32).tobytes() == g[key + "_prev"].tobytes()
        if not ok:
            bad.append((key, r[0], int(g[key + "_n"][0]), int((r[1] != g[key]).sum())))
    assert not bad, bad[:5]


def test_cuda_keyframe_searches_equal_reference_golden():
    """the keyframe-side searches of local mapping / loop closing -- SearchByProjection(KeyFrame*, Scw, ...) (:480, :595), Fuse (:1407,
    :1619), SearchBySim3 (:1743) -- with their matching core on the device (eorb_guided_search_windows) against what the REFERENCE'S OWN
    function bodies produced (tests/golden/ref_guided_kf.npz, 30 cases), bit for bit"""
    import kf_cases as KC
    api = _api()
    g = KC.kf_golden()
    gm = api.GuidedMatcher()
    bad = []
    for key, kind, ov, c in KC.kf_cases():
        r = KC.run_composed(gm.SearchWindows, kind, ov, c)
        if not (r[0] == int(g[key + "_n"][0]) and r[1].shape == g[key].shape and np.array_equal(r[1], g[key])):
            bad.append((key, r[0], int(g[key + "_n"][0])))
    assert not bad, bad[:5]


def test_cuda_search_windows_equals_oracle_raw_outputs():
    """every output array of the matching core (best_idx, best_dist, match2, nmatches) against the oracle: blocking and not, with the
    reprojection gate and right-image columns, held slots, distorted-camera bounds, whole-image windows (candidate-buffer growth), empty sides"""
    import kf_cases as KC
    import oracle_lib as O
    api = _api()
    gm = api.GuidedMatcher()
    rng = np.random.default_rng(5)
    ncmp = 0
    for key, kind, ov, c in KC.kf_cases(3):
        if kind != "fuse" or ov != 0:
            continue
        n1 = len(c["level1"])
        q, ur = KC.host_windows(c["pt8"], c["level1"], np.zeros(n1, bool), c["bounds"], c["K"], c["scale_factors"], 8.0, 0, mbf=40.0)
        held = (rng.random(len(c["kps2"])) < 0.3).astype(np.uint8)
        ur2 = np.where(rng.random(len(c["kps2"])) < 0.5, c["kps2"]["x"] - 5, -1).astype(np.float32)
        qm = KC._qmin(c["bounds"])
        variants = [dict(blocking=False, th_high=50), dict(blocking=True, th_high=50), dict(blocking=False, th_high=100, inv_level_sigma2=c["inv_sigma2"]),
                    dict(blocking=True, th_high=255, inv_level_sigma2=c["inv_sigma2"]), dict(blocking=True, th_high=0)]
        for kw in variants:
            for hd in (None, held):
                for u2 in (None, ur2):
                    a = (q, ur if u2 is not None else None, c["descMP"], c["kps2"], c["desc2"], hd, u2, c["bounds"])
                    o = O.search_windows(*a, query_min_xy=qm, **kw); r = gm.SearchWindows(*a, query_min_xy=qm, **kw)
                    assert o[0] == r[0] and all(np.array_equal(x, y) for x, y in zip(o[1:], r[1:])), (key, kw, hd is None, u2 is None)
                    ncmp += 1
        big = q.copy(); big["r"] = np.where(big["r"] > 0, 2000.0, -1).astype(np.float32)          # every keypoint of the level range is a candidate
        for blocking in (False, True):
            o = O.search_windows(big, None, c["descMP"], c["kps2"], c["desc2"], held, None, c["bounds"], blocking=blocking, th_high=100)
            r = gm.SearchWindows(big, None, c["descMP"], c["kps2"], c["desc2"], held, None, c["bounds"], blocking=blocking, th_high=100)
            assert o[0] == r[0] and all(np.array_equal(x, y) for x, y in zip(o[1:], r[1:])), (key, "whole image", blocking)
        r = gm.SearchWindows(q[:0], None, c["descMP"][:0], c["kps2"], c["desc2"], None, None, c["bounds"])
        assert r[0] == 0 and len(r[1]) == 0
        r = gm.SearchWindows(q, None, c["descMP"], c["kps2"][:0], c["desc2"][:0], None, None, c["bounds"])
        assert r[0] == 0 and (r[1] == -1).all() and (r[2] == 256).all()
    assert ncmp >= 40


def test_cuda_search_for_triangulation_equals_reference_golden():
    """ORBmatcher::SearchForTriangulation (:975-1214) on the device against the pairs the REFERENCE'S OWN function body + Pinhole::epipolarConstrain
    produced (tests/golden/ref_triangulation.npz, 40 cases), bit for bit; plus the empty / disjoint inputs"""
    import kf_cases as KC
    import oracle_lib as O
    api = _api()
    g = KC.tri_golden()
    bad = []
    for key, a, kw in KC.tri_cases():
        gm = api.GuidedMatcher(0, 0.6, kw["check_ori"])
        r = KC.tri_compose(lambda *x: gm.SearchForTriangulation(*x[:-2], bCoarse=x[-2]), a, kw, g[key + "_F"], g[key + "_ep"])
        if not (r[0] == int(g[key + "_n"][0]) and np.array_equal(r[1], g[key])):
            bad.append((key, r[0], int(g[key + "_n"][0]), int((r[1] != g[key]).sum())))
    assert not bad, bad[:5]
    key, a, kw = next(iter(KC.tri_cases(1)))
    k1, d1, h1, u1, fv1, k2, d2, h2, u2, fv2, K, t1, t2, sc, sg = a
    F, ep = g[key + "_F"], g[key + "_ep"]
    f1 = O.triangulation_flags(h1, None, False); f2 = O.triangulation_flags(h2, None, False)
    gm = api.GuidedMatcher()
    n, m = gm.SearchForTriangulation(k1, d1, np.zeros_like(f1), fv1, k2, d2, f2, fv2, F, ep, sc, sg)
    assert n == 0 and (m == -1).all()
    far = (fv2[0] + np.uint32(1 << 20), fv2[1], fv2[2])
    n, m = gm.SearchForTriangulation(k1, d1, f1, fv1, k2, d2, f2, far, F, ep, sc, sg)
    assert n == 0 and (m == -1).all()
    n, m = gm.SearchForTriangulation(k1[:0], d1[:0], f1[:0], (fv1[0][:0], np.zeros(1, np.int32), fv1[2][:0]), k2, d2, f2, fv2, F, ep, sc, sg)
    assert n == 0 and len(m) == 0
    # the device-resident entry point on the same case == the host entry point
    import torch
    from eorb_slam_b200.synth import KEYPOINT_DTYPE
    dev = lambda a_: torch.from_numpy(np.ascontiguousarray(a_).view(np.uint8).reshape(-1).copy()).cuda()
    tk1, td1, tf1, tk2, td2, tf2 = dev(np.ascontiguousarray(k1, KEYPOINT_DTYPE)), dev(d1), dev(f1), dev(np.ascontiguousarray(k2, KEYPOINT_DTYPE)), dev(d2), dev(f2)
    t1 = [dev(np.ascontiguousarray(fv1[0], np.uint32)), dev(np.ascontiguousarray(fv1[1], np.int32)), dev(np.ascontiguousarray(fv1[2], np.uint32))]
    t2 = [dev(np.ascontiguousarray(fv2[0], np.uint32)), dev(np.ascontiguousarray(fv2[1], np.int32)), dev(np.ascontiguousarray(fv2[2], np.uint32))]
    d_m = torch.zeros(len(k1), dtype=torch.int32, device="cuda")
    nh, mh = gm.SearchForTriangulation(k1, d1, f1, fv1, k2, d2, f2, fv2, F, ep, sc, sg)
    nd = gm.SearchForTriangulation_device(tk1.data_ptr(), td1.data_ptr(), tf1.data_ptr(), len(k1), [t.data_ptr() for t in t1], len(fv1[0]), int(fv1[1][-1]),
                                          tk2.data_ptr(), td2.data_ptr(), tf2.data_ptr(), len(k2), [t.data_ptr() for t in t2], len(fv2[0]), F, ep, sc, sg,
                                          d_m.data_ptr())
    assert nd == nh and nh > 0 and np.array_equal(d_m.cpu().numpy(), mh)

