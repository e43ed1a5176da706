#!/usr/bin/env python3
"""Per-stage device time of ONE 752x480 frame through eorb_orb_extract (stage events around every kernel group), next to the
whole host call with the CUDA graph on."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from eorb_slam_b200 import api, synth

img = synth.make_frame(0)
ex = api.ORBextractor(api.ORBxParams())
for _ in range(20):
    ex(img)
ts = []
for i in range(200):
    t0 = time.perf_counter(); ex(img); ts.append(time.perf_counter() - t0)
print("host call (graph): median %.1f us  min %.1f" % (np.median(ts) * 1e6, min(ts) * 1e6))
ex.stage_timing(True)
for _ in range(5):
    ex(img)
ex.stage_times()
n = 100
for _ in range(n):
    ex(img)
ms, launches = ex.stage_times()
print("per-stage device time of one frame (us):", {k: round(v / n * 1e3, 1) for k, v in ms.items()}, "sum %.1f" % (sum(ms.values()) / n * 1e3))
