#!/usr/bin/env python3
"""Where the single-frame host call spends its time: pageable vs pinned input, with / without descriptors, graph on / off."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from eorb_slam_b200 import api, synth

img = synth.make_frame(0)
pin = torch.from_numpy(img.copy()).pin_memory().numpy()

def med(fn, n=200):
    for _ in range(20): fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return np.median(ts) * 1e6

ex = api.ORBextractor(api.ORBxParams())
print("graph=%s" % os.environ.get("EORB_ORB_GRAPH", "1"))
print("pageable, desc   : %.1f us" % med(lambda: ex(img)))
print("pinned,   desc   : %.1f us" % med(lambda: ex(pin)))
print("pinned,   no desc: %.1f us" % med(lambda: ex(pin, None, (0, 1000), False)))
t = torch.zeros(1, device="cuda")
print("reference points: empty kernel + sync %.1f us, 361 KB pageable H2D + sync %.1f us, pinned %.1f us" % (
    med(lambda: (t.add_(1), torch.cuda.synchronize())), med(lambda: (torch.from_numpy(img).cuda(), torch.cuda.synchronize())),
    med(lambda: (torch.from_numpy(pin).cuda(non_blocking=True), torch.cuda.synchronize()))))
