#!/usr/bin/env python3
"""Random-case fuzz of the keyframe-side matching core (eorb_guided_search_windows) and SearchForTriangulation: CUDA vs the oracle on every
output array.  Sizes from 0 to a few thousand, windows from a few pixels to the whole image (candidate-buffer growth, heads longer than 32 with
every entry blocked), crowded duplicates, held slots, thresholds 0..255, distorted-camera bounds, level tables of 1..8 levels."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
from eorb_slam_b200 import api, synth

ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
gm = api.GuidedMatcher()
bad = 0
tot = [0, 0]
for it in range(ncases):
    n1 = int(rng.choice([0, 1, 7, 50, 400, 1500, 3000])); n2 = int(rng.choice([0, 1, 9, 64, 500, 1200, 4000]))
    nl = int(rng.integers(1, 9))
    b = np.array([0, 0, 752, 480], np.float32) if rng.random() < 0.5 else np.array([-20.5 * rng.random(), -9.3, 764.2, 489.6 + rng.random()], np.float32)
    k2 = np.zeros(n2, synth.KEYPOINT_DTYPE)
    crowd = rng.random() < 0.3
    k2["x"] = (rng.normal(300, 15, n2) if crowd else rng.uniform(b[0] - 5, b[2] + 5, n2)).astype(np.float32)
    k2["y"] = (rng.normal(200, 15, n2) if crowd else rng.uniform(b[1] - 5, b[3] + 5, n2)).astype(np.float32)
    k2["octave"] = rng.integers(0, nl, n2); k2["angle"] = rng.uniform(0, 360, n2).astype(np.float32)
    base = rng.integers(0, 256, (max(1, n2 // 20 if crowd else n2), 32), dtype=np.uint8)
    d2 = base[rng.integers(0, len(base), n2)].copy() if n2 else np.zeros((0, 32), np.uint8)
    if n2:
        flips = rng.integers(0, 256, (n2, 3)); d2[np.arange(n2)[:, None], flips // 8] ^= (1 << (flips % 8)).astype(np.uint8)
    dm = (d2[rng.integers(0, n2, n1)].copy() if n2 and n1 else rng.integers(0, 256, (n1, 32), dtype=np.uint8))
    if n1 and n2 and rng.random() < 0.5:
        fl = rng.integers(0, 256, (n1, 6)); dm[np.arange(n1)[:, None], fl // 8] ^= (1 << (fl % 8)).astype(np.uint8)
    q = np.zeros(n1, api.AREA_QUERY_DTYPE)
    q["x"] = (rng.normal(300, 20, n1) if crowd else rng.uniform(b[0] - 30, b[2] + 30, n1)).astype(np.float32)
    q["y"] = (rng.normal(200, 20, n1) if crowd else rng.uniform(b[1] - 30, b[3] + 30, n1)).astype(np.float32)
    q["r"] = rng.choice([-1.0, 2.0, 8.0, 30.0, 150.0, 2000.0], n1, p=[0.1, 0.2, 0.3, 0.2, 0.15, 0.05]).astype(np.float32)
    lv = rng.integers(0, nl, n1)
    q["min_level"] = lv - 1; q["max_level"] = lv
    if rng.random() < 0.2:
        q["min_level"] = -1; q["max_level"] = -1                 # no level test at all
    held = (rng.random(n2) < rng.choice([0.0, 0.3, 0.95])).astype(np.uint8) if rng.random() < 0.6 else None
    inv = (1.0 / (1.2 ** np.arange(nl)) ** 2).astype(np.float32) if rng.random() < 0.4 else None
    ur = (q["x"] - rng.random(n1).astype(np.float32) * 10).astype(np.float32) if inv is not None and rng.random() < 0.5 else None
    ur2 = np.where(rng.random(n2) < 0.5, k2["x"] - 5, -1).astype(np.float32) if ur is not None else None
    qm = np.array([int(b[0]), int(b[1])], np.float32) if rng.random() < 0.7 else None
    for blocking in (False, True):
        kw = dict(query_min_xy=qm, inv_level_sigma2=inv, blocking=blocking, th_high=int(rng.choice([0, 30, 50, 100, 255])))
        a = (q, ur, dm, k2, d2, held, ur2, b)
        o = O.search_windows(*a, **kw); r = gm.SearchWindows(*a, **kw)
        tot[int(blocking)] += int(o[0])
        ok = o[0] == r[0] and all(np.array_equal(x, y) for x, y in zip(o[1:], r[1:]))
        if not ok:
            bad += 1
            print("MISMATCH case", it, "n1", n1, "n2", n2, "blocking", blocking, kw["th_high"], "crowd", crowd, "nm", o[0], r[0],
                  [int((x != y).sum()) for x, y in zip(o[1:], r[1:])])
print("search_windows fuzz: %d cases x 2, %d mismatches; %d non-blocking and %d blocking matches compared" % (ncases, bad, tot[0], tot[1]))
sys.exit(1 if bad else 0)
