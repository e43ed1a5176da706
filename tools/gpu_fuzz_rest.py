"""Random-parameter fuzz of the event-frame, LK, guided-matching and bag-of-words rows against the oracle (GPU box).
usage: gpu_fuzz_rest.py [n_cases] [seed]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
from eorb_slam_b200 import api, synth

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0

def report(what, cfg, detail=""):
    global bad
    bad += 1
    print("MISMATCH", what, cfg, detail)

# ---- event frames
for case in range(n_cases):
    w = int(rng.integers(64, 400)); h = int(rng.integers(48, 300)); n = int(rng.choice([1, 7, 300, 2000, 9000]))
    sigma = float(rng.choice([0.5, 1.0, 1.0, 1.5, 2.0])); mode = int(rng.integers(0, 4)); pol = bool(rng.integers(0, 2))
    cfg = dict(w=w, h=h, n=n, sigma=sigma, mode=mode, pol=pol)
    try:
        ev = synth.make_events(n, int(rng.integers(0, 10**6)), w=w, h=h)
        ev["x"] += rng.uniform(-6, 6, n).astype(np.float32) * (rng.random(n) < 0.05)      # some events near / outside the border
        cv = api.EvImConverter(0, 1, max(n, 1), w, h)
        K = (w * 0.65, w * 0.65, w / 2.0, h / 2.0)
        dt = float(ev["ts"][-1] - ev["ts"][0]) if n > 1 else 0.0
        T = synth.rotation_tcw(rng.uniform(-2, 2, 3) * dt, rng.uniform(-0.02, 0.02, 3))
        se2 = rng.uniform(-0.05, 0.05, 3).astype(np.float32)
        if mode == 0:
            got = cv.ev2im(ev, w, h, pol, False); ref, _, _ = O.ev_accumulate(ev, w, h, 1.0, mode=0, pol=pol)
        elif mode == 1:
            got = cv.ev2im_gauss(ev, w, h, sigma, pol, False); ref, _, _ = O.ev_accumulate(ev, w, h, sigma, mode=1, pol=pol)
        elif mode == 2:
            got = cv.ev2mci_gg_f(ev, K, T, 1.0, w, h, sigma, pol, False); ref, _, _ = O.ev_accumulate(ev, w, h, sigma, mode=2, Tcw=T, depth=1.0, K=K, pol=pol)
        else:
            got = cv.ev2mci_gg_f_2d(ev, K, se2, w, h, sigma, pol, False); ref, _, _ = O.ev_accumulate(ev, w, h, sigma, mode=3, K=K, se2=se2, pol=pol)
        peak = max(float(np.abs(ref).max()), 1e-12)
        err = float(np.abs(got - ref).max()) / peak
        if not err <= 1e-4:
            report("events", cfg, "rel err %.3g" % err)
    except Exception as e:
        report("events", cfg, repr(e)[:200])

# ---- LK
for case in range(n_cases // 2):
    w = int(rng.integers(40, 400)); h = int(rng.integers(40, 300)); win = int(rng.choice([5, 11, 15, 21, 23, 31])); lvl = int(rng.integers(0, 4))
    npts = int(rng.choice([1, 17, 200])); it = int(rng.choice([3, 10, 30])); eps = float(rng.choice([0.01, 0.03, 0.3]))
    cfg = dict(w=w, h=h, win=win, lvl=lvl, npts=npts, it=it, eps=eps)
    try:
        a = synth.make_frame(int(rng.integers(0, 10**6)), w, h, nrect=60, noise=3)
        b = np.roll(a, (int(rng.integers(-3, 4)), int(rng.integers(-3, 4))), axis=(0, 1))
        pts = np.stack([rng.uniform(-2, w + 2, npts), rng.uniform(-2, h + 2, npts)], 1).astype(np.float32)
        tr = api.ELK_Tracker(win, lvl, it, eps, 0, (w, h), npts)
        tr.setRefImage(a, pts)
        p, s, e = tr.trackCurrImage(b)
        ep, es, ee, _ = O.lk_track(a, b, pts, None, win, lvl, it, eps)
        if not (np.array_equal(s, es) and p[s > 0].tobytes() == ep[es > 0].tobytes()):
            report("lk", cfg, "status equal %s" % np.array_equal(s, es))
    except Exception as e:
        report("lk", cfg, repr(e)[:200])

# ---- guided matching
for case in range(n_cases // 2):
    n1 = int(rng.choice([1, 30, 500, 2000])); n2 = int(rng.choice([1, 30, 500, 2000])); win = int(rng.choice([5, 30, 100, 1000]))
    ratio = float(rng.choice([0.6, 0.9, 1.2])); ori = bool(rng.integers(0, 2)); w = int(rng.integers(100, 1500)); h = int(rng.integers(80, 900))
    th = float(rng.choice([3.0, 15.0, 60.0]))
    cfg = dict(n1=n1, n2=n2, win=win, ratio=ratio, ori=ori, w=w, h=h, th=th)
    try:
        c = synth.make_projection_case(max(n1, 2), max(n2, 2), int(rng.integers(0, 10**6)), w=w, h=h, K=(w * 0.6, w * 0.6, w / 2.0, h / 2.0),
                                       zero_obs_frac=float(rng.choice([0.0, 0.1, 0.6])))
        for k in ("x3Dc", "valid1", "obs1", "kps1", "descMP"): c[k] = c[k][:n1]
        c["kps2"], c["desc2"] = c["kps2"][:n2], c["desc2"][:n2]
        gm = api.GuidedMatcher(0, ratio, ori)
        prev = np.stack([c["kps1"]["x"], c["kps1"]["y"]], 1)
        a = gm.SearchForInitialization(c["kps1"], c["descMP"], c["kps2"], c["desc2"], c["bounds"], prev, win)
        b = O.search_for_initialization(c["kps1"], c["descMP"], c["kps2"], c["desc2"], c["bounds"], prev, win, ratio, ori)
        if not (a[0] == b[0] and np.array_equal(a[1], b[1]) and a[2].tobytes() == b[2].tobytes()):
            report("search_init", cfg)
        a = gm.SearchByProjection(c["x3Dc"], c["valid1"], c["obs1"], c["kps1"], c["descMP"], c["kps2"], c["desc2"], c["bounds"], c["K"], c["scale_factors"], th)
        b = O.search_by_projection(c["x3Dc"], c["valid1"], c["obs1"], c["kps1"], c["descMP"], c["kps2"], c["desc2"], c["bounds"], c["K"], c["scale_factors"], th, ori)
        if not (a[0] == b[0] and np.array_equal(a[1], b[1])):
            report("search_proj", cfg)
        lc = synth.make_local_map_case(max(n1, 2), max(n2, 2), int(rng.integers(0, 10**6)), w=w, h=h, zero_obs_frac=float(rng.choice([0.0, 0.1, 0.6])),
                                       held_frac=float(rng.choice([0.0, 0.2, 0.9])))
        lc["pts"], lc["descMP"] = lc["pts"][:n1], lc["descMP"][:n1]
        lc["kps2"], lc["desc2"], lc["held2"] = lc["kps2"][:n2], lc["desc2"][:n2], lc["held2"][:n2]
        far = bool(rng.integers(0, 2)); thl = float(rng.choice([1.0, 3.0, 15.0, 80.0]))
        a = gm.SearchByProjectionMapPoints(lc["pts"], lc["descMP"], lc["kps2"], lc["desc2"], lc["held2"], lc["bounds"], lc["scale_factors"], thl, far, 15.0)
        b = O.search_by_projection_map_points(lc["pts"], lc["descMP"], lc["kps2"], lc["desc2"], lc["held2"], lc["bounds"], lc["scale_factors"], thl, far,
                                              15.0, ratio)
        if not (a[0] == b[0] and np.array_equal(a[1], b[1])):
            report("search_map_points", dict(cfg, thl=thl, far=far))
    except Exception as e:
        report("guided", cfg, repr(e)[:200])

# ---- SearchByBoW on FeatureVectors from the device transform
for case in range(n_cases // 4):
    k = int(rng.integers(2, 12)); L = int(rng.integers(1, 4)); lup = int(rng.integers(0, 5))
    n1 = int(rng.choice([1, 30, 500, 1500])); n2 = int(rng.choice([1, 30, 500, 1500]))
    ratio = float(rng.choice([0.6, 0.75, 0.9, 1.1])); ori = bool(rng.integers(0, 2)); flips = int(rng.choice([4, 24, 64]))
    cfg = dict(k=k, L=L, lup=lup, n1=n1, n2=n2, ratio=ratio, ori=ori, flips=flips)
    try:
        k1, d1, k2, d2, _ = synth.make_keypoint_frame_pair(max(n1, 2), max(n2, 2), int(rng.integers(0, 10**6)), max_flips=flips)
        k1, d1, k2, d2 = k1[:n1], d1[:n1], k2[:n2], d2[:n2]
        v = api.ORBVocabulary(synth.make_vocabulary(k, L, int(rng.integers(0, 10**6))))
        t1, t2 = v.transform(d1, lup), v.transform(d2, lup)
        fv1 = (t1["fv_nodes"], t1["fv_start"], t1["fv_feats"]); fv2 = (t2["fv_nodes"], t2["fv_start"], t2["fv_feats"])
        valid = (rng.random(n1) < float(rng.choice([0.3, 0.8, 1.0]))).astype(np.uint8)
        a = api.GuidedMatcher(0, ratio, ori).SearchByBoW(k1, d1, valid, fv1, k2, d2, fv2)
        b = O.search_by_bow(k1, d1, valid, fv1, k2, d2, fv2, ratio, ori)
        if not (a[0] == b[0] and np.array_equal(a[1], b[1])):
            report("search_by_bow", cfg)
    except Exception as e:
        report("search_by_bow", cfg, repr(e)[:200])

# ---- bag of words
for case in range(n_cases // 2):
    k = int(rng.integers(2, 21)); L = int(rng.integers(1, 5)); n = int(rng.choice([1, 10, 500, 3000])); lup = int(rng.integers(0, 7))
    sc = int(rng.integers(0, 6)); wt = int(rng.integers(0, 4))
    cfg = dict(k=k, L=L, n=n, lup=lup, scoring=sc, weighting=wt)
    try:
        voc = synth.make_vocabulary(k, L, int(rng.integers(0, 10**6)), sc, wt)
        feats = synth.make_vocabulary_features(voc, n, int(rng.integers(0, 10**6)))
        got = api.ORBVocabulary(voc).transform(feats, lup); exp = O.VocabOracle(voc).transform(feats, lup)
        if not (all(np.array_equal(got[q], exp[q]) for q in ("word_id", "node_id", "bow_ids", "fv_nodes", "fv_start", "fv_feats")) and
                got["bow_vals"].tobytes() == exp["bow_vals"].tobytes()):
            report("bow", cfg)
    except Exception as e:
        report("bow", cfg, repr(e)[:200])
print("fuzz done: %d mismatches" % bad)
sys.exit(1 if bad else 0)
