"""GPU parity of the guided-matching row: Frame grid, GetFeaturesInArea and
ORBmatcher::SearchForInitialization through the C ABI against the CPU oracle — every index, count and float bit-exact."""
import numpy as np
import pytest

import oracle_lib as O
from eorb_slam_b200 import synth

pytestmark = pytest.mark.gpu


def _api():
    from eorb_slam_b200 import api
    return api


@pytest.mark.parametrize("n,seed,bounds", [(0, 1, None), (1, 2, None), (700, 3, None), (4096, 4, None), (5003, 5, (-12.5, -9.25, 771.0, 493.5))])
def test_frame_grid_and_features_in_area(n, seed, bounds):
    api = _api()
    _, _, k2, _, b = synth.make_keypoint_frame_pair(max(n, 2), max(n, 2), seed)
    k2 = k2[:n]
    if bounds is not None:
        b = np.array(bounds, np.float32)
    fg = api.FrameGrid(k2, b)
    cs, ci = fg.AssignFeaturesToGrid()
    ecs, eci = O.frame_grid(k2, b)
    assert np.array_equal(cs, ecs) and np.array_equal(ci, eci)
    rng = np.random.default_rng(seed)
    nq = 300
    q = np.zeros(nq, api.AREA_QUERY_DTYPE)
    q["x"] = rng.uniform(-60, 820, nq); q["y"] = rng.uniform(-60, 540, nq)
    q["r"] = rng.choice([5.0, 15.0, 100.0, 900.0], nq)
    lv = [(0, -1), (0, 0), (2, 4), (1, -1), (3, 2), (-1, -1)]
    for i in range(nq):
        q["min_level"][i], q["max_level"][i] = lv[i % len(lv)]
    got = fg.GetFeaturesInAreaBatch(q)
    for i in range(nq):
        exp = O.features_in_area(k2, b, ecs, eci, float(q["x"][i]), float(q["y"][i]), float(q["r"][i]), int(q["min_level"][i]), int(q["max_level"][i]))
        assert np.array_equal(got[i], exp), i


@pytest.mark.parametrize("n1,n2,seed,window,ratio,ori", [
    (500, 520, 3, 100, 0.9, True), (500, 520, 4, 30, 0.7, True), (500, 520, 5, 100, 0.9, False), (300, 1, 6, 100, 0.9, True),
    (5000, 5000, 7, 100, 0.9, True),        # the initialisation extractor's 5 x nFeatures keypoints (Tracking.cc:127-128)
    (3000, 3000, 8, 2000, 0.9, True),       # window covers the whole image: > 32 candidates everywhere, buffer growth + slow path
    (64, 3000, 9, 15, 0.95, True), (1, 1, 10, 100, 0.9, True)])
def test_search_for_initialization(n1, n2, seed, window, ratio, ori):
    api = _api()
    k1, d1, k2, d2, b = synth.make_keypoint_frame_pair(max(n1, 2), max(n2, 2), seed)
    k1, d1, k2, d2 = k1[:n1], d1[:n1], k2[:n2], d2[:n2]
    prev = np.stack([k1["x"], k1["y"]], 1)
    gm = api.GuidedMatcher(0, ratio, ori)
    n, m12, p = gm.SearchForInitialization(k1, d1, k2, d2, b, prev, window)
    en, em12, ep = O.search_for_initialization(k1, d1, k2, d2, b, prev, window, ratio, ori)
    assert n == en and np.array_equal(m12, em12) and p.tobytes() == ep.tobytes()
    # second round with the updated vbPrevMatched, same handle (state must not leak between calls)
    n_b, m12_b, p_b = gm.SearchForInitialization(k1, d1, k2, d2, b, p, window)
    en_b, em12_b, ep_b = O.search_for_initialization(k1, d1, k2, d2, b, ep, window, ratio, ori)
    assert n_b == en_b and np.array_equal(m12_b, em12_b) and p_b.tobytes() == ep_b.tobytes()


def test_search_for_initialization_duplicate_descriptors_take_over():
    """several frame-1 keypoints with IDENTICAL descriptors and positions compete for one frame-2 keypoint: the first claims it,
    the later ones are filtered by vMatchedDistance <= dist (ORBmatcher.cc:755) and fall back to their second candidate"""
    api = _api()
    rng = np.random.default_rng(11)
    n = 200
    k1 = np.zeros(n, synth.KEYPOINT_DTYPE); k2 = np.zeros(n, synth.KEYPOINT_DTYPE)
    k1["x"] = 100 + (np.arange(n) % 10) * 3; k1["y"] = 100 + (np.arange(n) // 10) * 3
    k2["x"] = k1["x"] + 1; k2["y"] = k1["y"] - 1
    k1["angle"] = 10; k2["angle"] = 15
    base = rng.integers(0, 256, (4, 32), dtype=np.uint8)
    d1 = base[np.arange(n) % 4].copy(); d2 = base[(np.arange(n) // 3) % 4].copy()
    d2[:, 1] ^= 3                             # distance 2 to the twin (0 < 0 * ratio would never pass the ratio test)
    d2[::7, 0] ^= 1                           # a few one-bit variations -> distance ties and near ties everywhere
    b = np.array([0, 0, 752, 480], np.float32)
    prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
    for ratio in (0.9, 1.5):                  # 1.5: the ratio test passes even for equal best / second distances
        gm = api.GuidedMatcher(0, ratio, True)
        n_, m12, p = gm.SearchForInitialization(k1, d1, k2, d2, b, prev, 100)
        en, em12, ep = O.search_for_initialization(k1, d1, k2, d2, b, prev, 100, ratio, True)
        assert n_ == en and np.array_equal(m12, em12) and p.tobytes() == ep.tobytes()
    assert en > 0


def test_extract_two_frames_then_guided_search_without_leaving_hbm():
    """frame pair -> ORB extraction on the device -> SearchForInitialization on the device outputs (level-0 keypoints,
    100-px window), checked against oracle extraction + oracle search"""
    import torch
    api = _api()
    img1 = synth.make_frame(21)
    img2 = np.roll(img1, (3, -5), axis=(0, 1))
    p = api.ORBxParams(5000, 1.2, 8, 20, 7, 19, (752, 480))      # the initialisation extractor: 5 x nFeatures
    ex = api.ORBextractor(p, 0, 2)
    cap = ex.cap
    st = torch.cuda.current_stream().cuda_stream
    ex.set_stream(st)
    d_frames = torch.from_numpy(np.stack([img1, img2])).cuda()
    d_kps = torch.zeros(2 * cap * 28, dtype=torch.uint8, device="cuda"); d_desc = torch.zeros(2 * cap * 32, dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(2, dtype=torch.int32, device="cuda"); d_mono = torch.zeros(2, dtype=torch.int32, device="cuda")
    ex.extract_batch_raw(d_frames.data_ptr(), 2, 752, 480, 752, 752 * 480, (0, 0), True, d_kps.data_ptr(), d_desc.data_ptr(), cap,
                         d_n.data_ptr(), d_mono.data_ptr(), device=True)
    torch.cuda.synchronize()
    n1, n2 = (int(v) for v in d_n.cpu().numpy())
    kps = d_kps.cpu().numpy().view(synth.KEYPOINT_DTYPE).reshape(2, cap)
    desc = d_desc.cpu().numpy().reshape(2, cap, 32)
    d_prev = torch.from_numpy(np.stack([kps[0]["x"][:n1], kps[0]["y"][:n1]], 1).astype(np.float32).copy()).cuda()
    d_m12 = torch.zeros(n1, dtype=torch.int32, device="cuda")
    gm = api.GuidedMatcher(0, 0.9, True)
    gm.set_stream(st)
    b = np.array([0, 0, 752, 480], np.float32)
    nm = gm.SearchForInitialization_device(d_kps.data_ptr(), d_desc.data_ptr(), n1, d_kps.data_ptr() + cap * 28, d_desc.data_ptr() + cap * 32,
                                           n2, b, d_prev.data_ptr(), d_m12.data_ptr(), 100)
    torch.cuda.synchronize()
    orc = O.OrbOracle(5000, 1.2, 8, 20, 7, 19, 752, 480)
    _, ok1, od1 = orc.extract(img1, (0, 0), True); _, ok2, od2 = orc.extract(img2, (0, 0), True)
    assert ok1.tobytes() == kps[0][:n1].tobytes() and ok2.tobytes() == kps[1][:n2].tobytes()
    en, em12, ep = O.search_for_initialization(ok1, od1, ok2, od2, b, np.stack([ok1["x"], ok1["y"]], 1), 100, 0.9, True)
    assert nm == en and np.array_equal(d_m12.cpu().numpy(), em12) and d_prev.cpu().numpy().tobytes() == ep.tobytes()
    assert en > 100          # a shifted copy of a textured frame matches well
    gm.set_stream(None); ex.set_stream(None)


@pytest.mark.parametrize("n1,n2,seed,th,ori,zero_obs", [
    (500, 520, 21, 15.0, True, 0.05), (500, 520, 22, 7.0, True, 0.0), (500, 520, 23, 15.0, False, 0.3), (500, 520, 24, 40.0, True, 0.5),
    (1009, 1009, 25, 15.0, True, 0.02),      # tracking-rate call: two frames of the 1000-feature extractor (th = 15, Tracking.cc)
    (3000, 3000, 26, 400.0, True, 0.1),      # whole-image windows: long candidate lists, blocked heads -> the slow path
    (64, 2000, 27, 15.0, True, 0.0), (0, 10, 28, 15.0, True, 0.0), (300, 1, 29, 15.0, True, 0.0)])
def test_search_by_projection(n1, n2, seed, th, ori, zero_obs):
    """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono=True): match table and nmatches bit-exact"""
    api = _api()
    c = synth.make_projection_case(max(n1, 2), max(n2, 2), seed, zero_obs_frac=zero_obs)
    for k in ("x3Dc", "valid1", "obs1", "kps1", "descMP"):
        c[k] = c[k][:n1]
    c["kps2"], c["desc2"] = c["kps2"][:n2], c["desc2"][:n2]
    gm = api.GuidedMatcher(0, 0.9, ori)
    for rep in range(2):       # twice on one handle: no state may leak between calls
        n, mc = gm.SearchByProjection(c["x3Dc"], c["valid1"], c["obs1"], c["kps1"], c["descMP"], c["kps2"], c["desc2"], c["bounds"], c["K"],
                                      c["scale_factors"], th)
        en, emc = O.search_by_projection(c["x3Dc"], c["valid1"], c["obs1"], c["kps1"], c["descMP"], c["kps2"], c["desc2"], c["bounds"], c["K"],
                                         c["scale_factors"], th, ori)
        assert n == en and np.array_equal(mc, emc)
    if n1 >= 500:
        assert (emc >= 0).sum() > 50


def test_search_by_projection_contested_slots():
    """many last-frame points project onto the same few current-frame keypoints: first come first served for points with
    observations, overwrite (and double counting, as in the reference) for points without"""
    api = _api()
    rng = np.random.default_rng(31)
    n1, n2 = 400, 40
    c = synth.make_projection_case(n1, n2, 31)
    K = c["K"]
    tgt = rng.integers(0, n2, n1)
    z = rng.uniform(1, 5, n1).astype(np.float32)
    u = c["kps2"]["x"][tgt] + rng.normal(0, 1, n1).astype(np.float32); v = c["kps2"]["y"][tgt] + rng.normal(0, 1, n1).astype(np.float32)
    c["x3Dc"] = np.stack([(u - K[2]) / K[0] * z, (v - K[3]) / K[1] * z, z], 1).astype(np.float32)
    c["valid1"][:] = 1
    c["kps1"]["octave"] = c["kps2"]["octave"][tgt]
    c["descMP"] = c["desc2"][tgt].copy()
    c["descMP"][:, 0] ^= rng.integers(0, 4, n1).astype(np.uint8)
    for zero in (0.0, 0.5, 1.0):
        c["obs1"] = np.where(rng.random(n1) < zero, 0, 3).astype(np.int32)
        n, mc = api.GuidedMatcher(0, 0.9, True).SearchByProjection(c["x3Dc"], c["valid1"], c["obs1"], c["kps1"], c["descMP"], c["kps2"], c["desc2"],
                                                                 c["bounds"], K, c["scale_factors"], 15.0)
        en, emc = O.search_by_projection(c["x3Dc"], c["valid1"], c["obs1"], c["kps1"], c["descMP"], c["kps2"], c["desc2"], c["bounds"], K,
                                         c["scale_factors"], 15.0, True)
        assert n == en and np.array_equal(mc, emc)


@pytest.mark.parametrize("n1,n2,seed,th,far,ratio,zero_obs,held", [
    (600, 520, 31, 1.0, False, 0.8, 0.05, True), (600, 520, 32, 3.0, True, 0.8, 0.0, True), (600, 520, 33, 5.0, False, 0.6, 0.4, False),
    (2000, 1009, 34, 3.0, False, 0.8, 0.02, True),      # Tracking::SearchLocalPoints shape: a local map against one 1000-feature frame
    (3000, 3000, 35, 60.0, False, 0.9, 0.1, True),      # very wide windows: long candidate lists, blocked heads -> the slow path
    (5000, 5000, 36, 15.0, True, 0.8, 0.05, True), (0, 10, 37, 1.0, False, 0.8, 0.0, True), (300, 1, 38, 1.0, False, 0.8, 0.0, False)])
def test_search_by_projection_map_points(n1, n2, seed, th, far, ratio, zero_obs, held):
    """ORBmatcher::SearchByProjection(F, vpMapPoints, th, bFarPoints, thFarPoints), monocular: match table and nmatches bit-exact"""
    api = _api()
    c = synth.make_local_map_case(max(n1, 2), max(n2, 2), seed, zero_obs_frac=zero_obs)
    c["pts"], c["descMP"] = c["pts"][:n1], c["descMP"][:n1]
    c["kps2"], c["desc2"], c["held2"] = c["kps2"][:n2], c["desc2"][:n2], (c["held2"][:n2] if held else None)
    gm = api.GuidedMatcher(0, ratio, True)
    for rep in range(2):
        n, mc = gm.SearchByProjectionMapPoints(c["pts"], c["descMP"], c["kps2"], c["desc2"], c["held2"], c["bounds"], c["scale_factors"], th, far, 20.0)
        en, emc = O.search_by_projection_map_points(c["pts"], c["descMP"], c["kps2"], c["desc2"], c["held2"], c["bounds"], c["scale_factors"], th, far,
                                                    20.0, ratio)
        assert n == en and np.array_equal(mc, emc)
    if n1 >= 600:
        assert (emc >= 0).sum() > 40


def test_search_by_projection_map_points_crowded():
    """thousands of near-identical map points over a few hundred keypoints: every head runs dry (the full-list path), slots
    fill up first come first served, points without observations are overwritten and counted again"""
    api = _api()
    rng = np.random.default_rng(41)
    n1, n2 = 3000, 400
    c = synth.make_local_map_case(n1, n2, 41)
    base = rng.integers(0, 256, 32).astype(np.uint8)
    def noisy(n, bits):
        d = np.tile(base, (n, 1))
        for _ in range(bits):
            d[np.arange(n), rng.integers(0, 32, n)] ^= (1 << rng.integers(0, 8, n)).astype(np.uint8)
        return d
    c["descMP"], c["desc2"] = noisy(n1, 12), noisy(n2, 12)
    c["kps2"]["x"] = rng.uniform(300, 380, n2).astype(np.float32); c["kps2"]["y"] = rng.uniform(200, 280, n2).astype(np.float32)
    c["kps2"]["octave"] = rng.integers(0, 3, n2)
    c["pts"]["proj_x"] = rng.uniform(300, 380, n1).astype(np.float32); c["pts"]["proj_y"] = rng.uniform(200, 280, n1).astype(np.float32)
    c["pts"]["scale_level"] = rng.integers(0, 4, n1)
    c["pts"]["in_view"] = 1; c["pts"]["bad"] = 0
    for zero, ratio in ((0.0, 0.8), (0.3, 0.95), (1.0, 0.8)):
        c["pts"]["observations"] = np.where(rng.random(n1) < zero, 0, 2).astype(np.int32)
        n, mc = api.GuidedMatcher(0, ratio, True).SearchByProjectionMapPoints(c["pts"], c["descMP"], c["kps2"], c["desc2"], c["held2"], c["bounds"],
                                                                              c["scale_factors"], 10.0)
        en, emc = O.search_by_projection_map_points(c["pts"], c["descMP"], c["kps2"], c["desc2"], c["held2"], c["bounds"], c["scale_factors"], 10.0,
                                                    False, 0.0, ratio)
        assert n == en and np.array_equal(mc, emc)
        assert en > 100


def _bow_case(api, n1, n2, seed, k, L, levelsup, valid_frac=0.8, max_flips=24):
    """keyframe / frame features, their FeatureVectors through the device vocabulary transform (the reference's flow:
    ComputeBoW on both, then SearchByBoW), and the keyframe's map-point flags"""
    k1, d1, k2, d2, _ = synth.make_keypoint_frame_pair(max(n1, 2), max(n2, 2), seed, max_flips=max_flips)
    k1, d1, k2, d2 = k1[:n1], d1[:n1], k2[:n2], d2[:n2]
    v = api.ORBVocabulary(synth.make_vocabulary(k, L, seed))
    def fv(d):
        if len(d) == 0:
            return (np.zeros(0, np.uint32), np.zeros(1, np.int32), np.zeros(0, np.uint32))
        t = v.transform(d, levelsup)
        return (t["fv_nodes"], t["fv_start"], t["fv_feats"])
    valid = (np.random.default_rng(seed).random(n1) < valid_frac).astype(np.uint8)
    return k1, d1, valid, fv(d1), k2, d2, fv(d2)


@pytest.mark.parametrize("n1,n2,seed,k,L,levelsup,ratio,ori", [
    (500, 520, 41, 6, 3, 2, 0.7, True), (500, 520, 42, 6, 3, 1, 0.9, True), (500, 520, 43, 6, 3, 0, 0.7, False),
    (1009, 1009, 44, 10, 4, 2, 0.7, True),       # Tracking::TrackReferenceKeyFrame shape: ORBmatcher(0.7, true), ~100 nodes
    (5000, 5000, 45, 10, 3, 2, 0.75, True),      # 10 nodes of ~500 features: long per-node lists (memory-resident loop)
    (600, 600, 61, 10, 2, 1, 0.7, True),         # 10 nodes of ~60 frame features: two register slots per lane
    (1009, 1009, 62, 10, 3, 2, 0.7, True),       # 10 nodes of ~100: four slots (the shape of smoke())
    (2000, 2000, 63, 10, 2, 1, 0.8, True),       # 10 nodes of ~200: eight slots
    (800, 800, 46, 5, 2, 4, 0.8, True),          # levelsup >= L: everything under the root, one ordered list
    (0, 10, 47, 6, 3, 2, 0.7, True), (300, 1, 48, 6, 3, 2, 0.7, True), (1, 300, 49, 6, 3, 2, 0.7, True)])
def test_search_by_bow(n1, n2, seed, k, L, levelsup, ratio, ori):
    """ORBmatcher::SearchByBoW(pKF, F, vpMapPointMatches), monocular: match table and nmatches bit-exact"""
    api = _api()
    c = _bow_case(api, n1, n2, seed, k, L, levelsup)
    gm = api.GuidedMatcher(0, ratio, ori)
    for rep in range(2):
        n, mf = gm.SearchByBoW(*c)
        en, emf = O.search_by_bow(*c, ratio, ori)
        assert n == en and np.array_equal(mf, emf)
    if n1 >= 500 and levelsup >= 2:
        assert en > 50


@pytest.mark.parametrize("n1,n2,seed,k,L,levelsup,ratio,ori", [
    (1009, 1009, 71, 10, 4, 2, 0.7, True), (1009, 1009, 72, 10, 4, 2, 0.9, False), (5000, 5000, 73, 10, 3, 2, 0.75, True),
    (2000, 2000, 74, 10, 2, 1, 0.8, True), (800, 800, 75, 5, 2, 4, 0.8, True), (0, 10, 76, 6, 3, 2, 0.7, True), (300, 1, 77, 6, 3, 2, 0.7, True)])
def test_search_by_bow_keyframe_keyframe(n1, n2, seed, k, L, levelsup, ratio, ori):
    """ORBmatcher::SearchByBoW(pKF1, pKF2, vpMatches12) (ORBmatcher.cc:833-990): second-side validity, strict TH_LOW, result indexed by the
    first keyframe; bit-exact against the oracle (which equals the reference's own body, tests/test_ref_pin.py)"""
    api = _api()
    k1, d1, v1, fv1, k2, d2, fv2 = _bow_case(api, n1, n2, seed, k, L, levelsup)
    v2 = (np.random.default_rng(seed).random(len(k2)) < 0.7).astype(np.uint8)
    gm = api.GuidedMatcher(0, ratio, ori)
    for rep in range(2):
        n, m12 = gm.SearchByBoW_KF(k1, d1, v1, fv1, k2, d2, v2, fv2)
        en, em12 = O.search_by_bow_kf(k1, d1, v1, fv1, k2, d2, v2, fv2, ratio, ori)
        assert n == en and np.array_equal(m12, em12)
    if n1 >= 500 and levelsup >= 2:
        assert en > 30
    # a distance of exactly TH_LOW = 50 is accepted by the keyframe-frame form (<=) and rejected here (<)
    if n1 >= 1009 and ori:
        q = d2[:1].copy()
        bits = np.unpackbits(q[0]); bits[:50] ^= 1
        kk1 = k1[:1]; dd1 = np.packbits(bits)[None, :]
        one = (np.array([0], np.uint32), np.array([0, 1], np.int32), np.array([0], np.uint32))
        gm2 = api.GuidedMatcher(0, 0.99, False)
        assert gm2.SearchByBoW(kk1, dd1, np.ones(1, np.uint8), one, k2[:1], q, one)[0] == 1
        assert gm2.SearchByBoW_KF(kk1, dd1, np.ones(1, np.uint8), one, k2[:1], q, np.ones(1, np.uint8), one)[0] == 0
        assert O.search_by_bow_kf(kk1, dd1, np.ones(1, np.uint8), one, k2[:1], q, np.ones(1, np.uint8), one, 0.99, False)[0] == 0


def test_search_by_bow_rejects_malformed_feature_vectors():
    api = _api()
    c = list(_bow_case(api, 200, 200, 51, 6, 3, 2))
    nodes, start, feats = c[3]
    bad = feats.copy(); bad[0] = 10**6
    with pytest.raises(RuntimeError):
        api.GuidedMatcher(0, 0.7, True).SearchByBoW(c[0], c[1], c[2], (nodes, start, bad), c[4], c[5], c[6])
    with pytest.raises(RuntimeError):
        api.GuidedMatcher(0, 0.7, True).SearchByBoW(c[0], c[1], c[2], (nodes[::-1].copy(), start, feats), c[4], c[5], c[6])


def test_tracking_chain_on_the_device_extract_bow_search_by_bow_and_local_points():
    """Tracking::TrackReferenceKeyFrame / SearchLocalPoints with the frame data resident in HBM: keyframe and frame are extracted
    on the device, their descriptors go through the vocabulary transform where they lie, SearchByBoW and the local-map
    SearchByProjection run on the device arrays; every result equals oracle extraction + oracle transform + oracle search"""
    import torch
    api = _api()
    img1 = synth.make_frame(31)
    img2 = np.roll(img1, (2, -3), axis=(0, 1))
    p = api.ORBxParams(1000, 1.2, 8, 20, 7, 19, (752, 480))
    ex = api.ORBextractor(p, 0, 2)
    cap = ex.cap
    st = torch.cuda.current_stream().cuda_stream
    ex.set_stream(st)
    d_frames = torch.from_numpy(np.stack([img1, img2])).cuda()
    d_kps = torch.zeros(2 * cap * 28, dtype=torch.uint8, device="cuda"); d_desc = torch.zeros(2 * cap * 32, dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(2, dtype=torch.int32, device="cuda"); d_mono = torch.zeros(2, dtype=torch.int32, device="cuda")
    ex.extract_batch_raw(d_frames.data_ptr(), 2, 752, 480, 752, 752 * 480, (0, 0), True, d_kps.data_ptr(), d_desc.data_ptr(), cap,
                         d_n.data_ptr(), d_mono.data_ptr(), device=True)
    torch.cuda.synchronize()
    n1, n2 = (int(v) for v in d_n.cpu().numpy())
    voc = synth.make_vocabulary(10, 4, 7)
    v = api.ORBVocabulary(voc, 0); v.set_stream(st)
    t1 = v.transform_device(d_desc.data_ptr(), n1, 2); t2 = v.transform_device(d_desc.data_ptr() + cap * 32, n2, 2)
    def up(t):
        a = [torch.from_numpy(np.ascontiguousarray(t[k]).view(np.int32).copy()).cuda() for k in ("fv_nodes", "fv_start", "fv_feats")]
        return a, tuple(x.data_ptr() for x in a)
    keep1, fv1 = up(t1); keep2, fv2 = up(t2)
    rng = np.random.default_rng(5)
    valid = (rng.random(n1) < 0.85).astype(np.uint8)
    d_valid = torch.from_numpy(valid).cuda()
    d_mf = torch.zeros(n2, dtype=torch.int32, device="cuda")
    gm = api.GuidedMatcher(0, 0.7, True); gm.set_stream(st)
    nb = gm.SearchByBoW_device(d_kps.data_ptr(), d_desc.data_ptr(), d_valid.data_ptr(), n1, fv1, len(t1["fv_nodes"]), d_kps.data_ptr() + cap * 28,
                               d_desc.data_ptr() + cap * 32, n2, fv2, len(t2["fv_nodes"]), d_mf.data_ptr())
    torch.cuda.synchronize()
    orc = O.OrbOracle(1000, 1.2, 8, 20, 7, 19, 752, 480)
    _, ok1, od1 = orc.extract(img1, (0, 0), True); _, ok2, od2 = orc.extract(img2, (0, 0), True)
    vo = O.VocabOracle(voc)
    e1, e2 = vo.transform(od1, 2), vo.transform(od2, 2)
    for k in ("fv_nodes", "fv_start", "fv_feats"):
        assert np.array_equal(t1[k], e1[k]) and np.array_equal(t2[k], e2[k])
    en, emf = O.search_by_bow(ok1, od1, valid, (e1["fv_nodes"], e1["fv_start"], e1["fv_feats"]), ok2, od2,
                              (e2["fv_nodes"], e2["fv_start"], e2["fv_feats"]), 0.7, True)
    assert nb == en and np.array_equal(d_mf.cpu().numpy(), emf)
    assert en > 100
    # local map = the keyframe's features as map points predicted at their frame-2 positions
    pts = np.zeros(n1, synth.TRACK_POINT_DTYPE)
    pts["proj_x"] = ok1["x"] - 3.0 + rng.normal(0, 0.7, n1).astype(np.float32); pts["proj_y"] = ok1["y"] + 2.0 + rng.normal(0, 0.7, n1).astype(np.float32)
    pts["view_cos"] = rng.uniform(0.99, 1.0, n1).astype(np.float32); pts["depth"] = 5.0
    pts["scale_level"] = ok1["octave"]; pts["observations"] = rng.integers(0, 4, n1); pts["in_view"] = valid
    held = (emf >= 0).astype(np.uint8)                      # slots TrackReferenceKeyFrame filled are not searched again
    sf = np.ones(8, np.float32)
    for i in range(1, 8):
        sf[i] = np.float32(np.float64(sf[i - 1]) * np.float64(np.float32(1.2)))
    b = np.array([0, 0, 752, 480], np.float32)
    d_pts = torch.from_numpy(pts.view(np.uint8).copy()).cuda(); d_held = torch.from_numpy(held).cuda()
    d_mc = torch.zeros(n2, dtype=torch.int32, device="cuda")
    gl = api.GuidedMatcher(0, 0.8, True); gl.set_stream(st)
    nl = gl.SearchByProjectionMapPoints_device(d_pts.data_ptr(), d_desc.data_ptr(), n1, d_kps.data_ptr() + cap * 28, d_desc.data_ptr() + cap * 32,
                                               d_held.data_ptr(), n2, b, sf, d_mc.data_ptr(), 3.0)
    torch.cuda.synchronize()
    eln, elmc = O.search_by_projection_map_points(pts, od1, ok2, od2, held, b, sf, 3.0, False, 0.0, 0.8)
    assert nl == eln and np.array_equal(d_mc.cpu().numpy(), elmc)
    assert not np.any((elmc >= 0) & (held != 0)) and eln > 20
    for h in (gm, gl, ex, v):
        h.set_stream(None)
