#!/usr/bin/env python3
"""Where the LK call time goes: fixed overhead (1 point) vs 400 / 4000 points."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from eorb_slam_b200 import api, synth
import oracle_lib as O
per, w, h = 2000, 240, 180
ev = synth.make_events(per * 2, seed=7, w=w, h=h)
i0 = O.normalize_minmax_u8(O.ev_accumulate(ev[:per], w, h, 1.0, mode=1)[0])
i1 = O.normalize_minmax_u8(O.ev_accumulate(ev[per // 2:per + per // 2], w, h, 1.0, mode=1)[0])
_, kps, _ = O.OrbOracle(400, 1.0, 1, 0, 0, 9, w, h).extract(i0, (0, 1000), False)
pts = np.stack([kps["x"], kps["y"]], 1).astype(np.float32)
for n in (1, 50, 400, 4000):
    p = np.tile(pts, (n // len(pts) + 1, 1))[:n]
    tr = api.ELK_Tracker(23, 1, 10, 0.03, 0, (w, h), n)
    tr.setRefImage(i0, p)
    for _ in range(10): tr.trackCurrImage(i1)
    t0 = time.perf_counter()
    for _ in range(100): tr.trackCurrImage(i1)
    print("n=%5d  %.1f us per call" % (n, (time.perf_counter() - t0) * 1e4))
