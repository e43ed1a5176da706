"""Latency of ONE ORBextractor::operator() call (config 1: one 752x480 frame, host image in, keypoints + descriptors out)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from eorb_slam_b200 import api, synth
imgs = [synth.make_frame(i) for i in range(8)]
ex = api.ORBextractor(api.ORBxParams())
for i in range(20): ex(imgs[i % 8])
ts = []
for i in range(300):
    t0 = time.perf_counter(); ret, k, d = ex(imgs[i % 8]); ts.append(time.perf_counter() - t0)
ts = np.array(ts) * 1e3
print("single frame operator(): median %.3f ms  p10 %.3f  p90 %.3f  (n=%d keypoints)" % (np.median(ts), np.percentile(ts, 10), np.percentile(ts, 90), len(k)))
ex2 = api.ORBextractor(api.ORBxParams(400, 1.0, 1, 0, 0, 9, (240, 180)))
ev = [np.ascontiguousarray(imgs[i][:180, :240]) for i in range(8)]
for i in range(20): ex2(ev[i % 8], None, (0, 1000), False)
ts = []
for i in range(300):
    t0 = time.perf_counter(); ex2(ev[i % 8], None, (0, 1000), False); ts.append(time.perf_counter() - t0)
ts = np.array(ts) * 1e3
print("event L1 extractor (240x180, 1 level, keypoints only): median %.3f ms" % np.median(ts))
ex.stage_timing(True)
for i in range(50): ex(imgs[i % 8])
ms, launches = ex.stage_times()
print("per-stage us per frame (single-frame calls):", {k: round(v / 50 * 1000, 1) for k, v in ms.items()})
