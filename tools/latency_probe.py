#!/usr/bin/env python3
"""Single-frame call latency of the host API (the call Tracking makes once per frame), GPU vs the CPU oracle."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from eorb_slam_b200 import api, synth
import oracle_lib as O

img = synth.make_frame(0)
ex = api.ORBextractor(api.ORBxParams())
for _ in range(20):
    ex(img)
ts = []
for i in range(300):
    t0 = time.perf_counter(); ex(img); ts.append(time.perf_counter() - t0)
ts = np.array(ts) * 1e6
print("gpu single-frame call: median %.1f us  p10 %.1f  p90 %.1f  min %.1f" % (np.median(ts), np.percentile(ts, 10), np.percentile(ts, 90), ts.min()))
orc = O.OrbOracle()
orc.extract(img)
t0 = time.perf_counter()
for _ in range(20):
    orc.extract(img)
print("cpu oracle single frame: %.1f us" % ((time.perf_counter() - t0) / 20 * 1e6))
