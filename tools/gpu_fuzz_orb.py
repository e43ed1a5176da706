"""Random-parameter fuzz of the ORB path against the oracle (run on a GPU box): sizes, levels, scale factors, thresholds,
margins, feature counts, lapping areas, frame kinds.  usage: gpu_fuzz_orb.py [n_cases] [seed]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
from eorb_slam_b200 import api, synth

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
for case in range(n_cases):
    w = int(rng.integers(64, 900)); h = int(rng.integers(48, 620))
    nlev = int(rng.integers(1, 10)); sf = float(np.float32(rng.choice([1.0, 1.1, 1.2, 1.26, 1.5, 2.0]) if nlev > 1 else 1.0))
    if nlev > 1 and sf == 1.0: sf = 1.2
    nfeat = int(rng.choice([1, 50, 400, 1000, 2500]))
    ini = int(rng.integers(0, 40)); mn = int(rng.integers(0, ini + 1))
    edge = int(rng.choice([9, 15, 19, 25, -1]))
    lap = [(0, 1000), (0, 0), (100, 300)][int(rng.integers(0, 3))]
    want = bool(rng.integers(0, 2))
    kind = ["textured", "flat", "textured"][int(rng.integers(0, 3))]
    img = synth.make_frame(int(rng.integers(0, 10**6)), w, h, nrect=int(rng.integers(5, 500)), noise=int(rng.integers(0, 12)), kind=kind)
    try:
        ex = api.ORBextractor(api.ORBxParams(nfeat, sf, nlev, ini, mn, edge, (w, h)))
        r1, k1, d1 = ex(img, None, lap, want)
        orc = O.OrbOracle(nfeat, sf, nlev, ini, mn, edge, w, h)
        r2, k2, d2 = orc.extract(img, lap, want)
        ok = r1 == r2 and k1.tobytes() == k2.tobytes() and (not want or (d1 is None and d2 is None) or np.array_equal(d1, d2))
    except Exception as e:
        msg = repr(e)
        # documented limits fail loudly and are not mismatches: octree node capacity (per-level quota above ~2300), a pyramid
        # level that rounds to zero pixels (cv::resize would assert in the reference), margins below 3
        if "shared-memory budget exceeded" in msg or "is empty" in msg or "edge threshold" in msg:
            limits = globals().get("limits", 0) + 1
            continue
        ok = False
        print("EXC", msg[:200])
    if not ok:
        bad += 1
        print("MISMATCH case", case, dict(w=w, h=h, nlev=nlev, sf=sf, nfeat=nfeat, ini=ini, mn=mn, edge=edge, lap=lap, want=want, kind=kind),
              "n", len(k1) if "k1" in dir() else None, len(k2) if "k2" in dir() else None)
print("fuzz done: %d cases, %d mismatches, %d loud limit errors" % (n_cases, bad, globals().get("limits", 0)))
sys.exit(1 if bad else 0)
