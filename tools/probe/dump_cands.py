"""Dump the FAST candidates of every level of one synthetic frame (kernel-tuning A/B: run with EORB_B200_LIB=...)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from eorb_slam_b200 import api, synth
img = synth.make_frame(0)
ex = api.ORBextractor(api.ORBxParams())
ret, kps, desc = ex(img)
out = {}
for l in range(8):
    xs, ys, sc = ex.debug_candidates(l)
    out["x%d" % l] = xs; out["y%d" % l] = ys; out["s%d" % l] = sc
out["kps"] = kps
np.savez(sys.argv[1], **out)
print(sys.argv[1], [len(out["x%d" % l]) for l in range(8)], len(kps))
