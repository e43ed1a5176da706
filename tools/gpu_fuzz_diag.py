"""Stage-by-stage diff of one ORB configuration (GPU vs oracle): levels, candidates, per-level selections, final keypoints."""
import os, sys, ast
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
from eorb_slam_b200 import api, synth

def diag(seed_img, w, h, nlev, sf, nfeat, ini, mn, edge, lap, want, kind, nrect, noise):
    img = synth.make_frame(seed_img, w, h, nrect=nrect, noise=noise, kind=kind)
    ex = api.ORBextractor(api.ORBxParams(nfeat, sf, nlev, ini, mn, edge, (w, h)))
    r1, k1, d1 = ex(img, None, lap, want)
    orc = O.OrbOracle(nfeat, sf, nlev, ini, mn, edge, w, h)
    r2, k2, d2 = orc.extract(img, lap, want)
    print("ret", r1, r2, "n", len(k1), len(k2), "edge", ex.edge if hasattr(ex, "edge") else None, orc.edge)
    for l in range(nlev):
        try:
            gl = ex.pyramid_level(l); ol = orc.level(l)
            same_lvl = gl.shape == ol.shape and np.array_equal(gl, ol)
        except Exception as e:
            same_lvl = "exc %r" % e
        gx, gy, gs = ex.debug_candidates(l); ox, oy, os_ = orc.candidates(l)
        same_c = len(gx) == len(ox) and np.array_equal(gx, ox) and np.array_equal(gy, oy) and np.array_equal(gs, os_)
        kx, ky, ks, ka = ex.debug_level_kps(l); qx, qy, qs, qa = orc.level_kps(l)
        same_k = len(kx) == len(qx) and np.array_equal(kx, qx) and np.array_equal(ky, qy) and np.array_equal(ks, qs)
        same_a = same_k and ka.tobytes() == qa.tobytes()
        print("level", l, ex.level_size(l), "pyr", same_lvl, "cand", same_c, (len(gx), len(ox)), "sel", same_k, (len(kx), len(qx)), "angle", same_a)
    if len(k1) == len(k2):
        diff = [f for f in k1.dtype.names if not np.array_equal(k1[f], k2[f])]
        print("differing keypoint fields", diff)
        if diff:
            i = int(np.flatnonzero(k1[diff[0]] != k2[diff[0]])[0]); print("first", i, k1[i], k2[i])
        if want and d1 is not None and d2 is not None:
            bad = np.flatnonzero((d1 != d2).any(1)); print("differing descriptors", len(bad), bad[:5], [(int(k1[j]["octave"]), float(k1[j]["x"]), float(k1[j]["y"])) for j in bad[:5]])

rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
target = [int(x) for x in sys.argv[1].split(",")]
for case in range(max(target) + 1):
    w = int(rng.integers(64, 900)); h = int(rng.integers(48, 620))
    nlev = int(rng.integers(1, 10)); sf = float(np.float32(rng.choice([1.0, 1.1, 1.2, 1.26, 1.5, 2.0]) if nlev > 1 else 1.0))
    if nlev > 1 and sf == 1.0: sf = 1.2
    nfeat = int(rng.choice([1, 50, 400, 1000, 2500]))
    ini = int(rng.integers(0, 40)); mn = int(rng.integers(0, ini + 1))
    edge = int(rng.choice([9, 15, 19, 25, -1]))
    lap = [(0, 1000), (0, 0), (100, 300)][int(rng.integers(0, 3))]
    want = bool(rng.integers(0, 2))
    kind = ["textured", "flat", "textured"][int(rng.integers(0, 3))]
    si = int(rng.integers(0, 10**6)); nrect = int(rng.integers(5, 500)); noise = int(rng.integers(0, 12))
    if case in target:
        print("==== case", case, dict(w=w, h=h, nlev=nlev, sf=sf, nfeat=nfeat, ini=ini, mn=mn, edge=edge, lap=lap, want=want, kind=kind))
        diag(si, w, h, nlev, sf, nfeat, ini, mn, edge, lap, want, kind, nrect, noise)
