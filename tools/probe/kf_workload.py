"""runs the keyframe-side matching core (Fuse / SearchByProjection(KeyFrame*) shapes) and SearchForTriangulation a few times -- the command
profiled for profiles/r02w_launches_kf_matchers.csv (ncu --metrics gpu__time_duration.sum)"""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
import kf_cases as KC
from eorb_slam_b200 import api, synth

c = synth.make_local_map_case(3000, 1009, 47)
q = np.zeros(3000, api.AREA_QUERY_DTYPE)
lv = np.clip(c["pts"]["scale_level"], 0, len(c["scale_factors"]) - 1)
q["x"] = c["pts"]["proj_x"]; q["y"] = c["pts"]["proj_y"]; q["r"] = np.float32(3.0) * c["scale_factors"][lv]
q["min_level"] = lv - 1; q["max_level"] = lv
inv = (np.float32(1.0) / (c["scale_factors"] ** 2)).astype(np.float32)
gm = api.GuidedMatcher(0, 0.8, True)
a = (q, None, c["descMP"], c["kps2"], c["desc2"], None, None, c["bounds"])
for _ in range(3):
    r0 = gm.SearchWindows(*a, inv_level_sigma2=inv, blocking=False, th_high=50)
    r1 = gm.SearchWindows(*a, blocking=True, th_high=50)
key, ta, kw = next(iter(KC.tri_cases(1)))
g = KC.tri_golden()
for _ in range(3):
    r2 = KC.tri_compose(lambda *x: gm.SearchForTriangulation(*x[:-2], bCoarse=x[-2]), ta, kw, g[key + "_F"], g[key + "_ep"])
print(r0[0], r1[0], r2[0])
