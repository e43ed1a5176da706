"""host-call latency of SearchByBoW / SearchByProjection(map points) (probe, not a bench)"""
import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests"))
import numpy as np
from eorb_slam_b200 import api, synth
import oracle_lib as O
rng = np.random.default_rng(4)
voc = synth.make_vocabulary_regular(10, 6, 3)
leaves = np.flatnonzero(voc["is_leaf"])
feats = voc["desc"][rng.choice(leaves, 1009)].copy()
feats[:, :4] ^= rng.integers(0, 256, (1009, 4), dtype=np.uint8)
v = api.ORBVocabulary(voc, 0)
kk1, _, kk2, _, _ = synth.make_keypoint_frame_pair(1009, 1009, 17)
df = feats.copy()
fl = rng.integers(0, 256, (len(df), 2))
df[np.arange(len(df)), fl[:, 0] % 28 + 4] ^= (1 << (fl[:, 1] % 8)).astype(np.uint8)
t1 = v.transform(feats, 4); t2 = v.transform(df, 4)
fv1 = (t1["fv_nodes"], t1["fv_start"], t1["fv_feats"]); fv2 = (t2["fv_nodes"], t2["fv_start"], t2["fv_feats"])
valid = np.ones(len(feats), np.uint8)
gm = api.GuidedMatcher(0, 0.7, False)
for _ in range(5):
    r = gm.SearchByBoW(kk1, feats, valid, fv1, kk2, df, fv2)
t0 = time.perf_counter()
for _ in range(200):
    r = gm.SearchByBoW(kk1, feats, valid, fv1, kk2, df, fv2)
print("SearchByBoW host call: %.1f us, %d matches" % ((time.perf_counter() - t0) / 200 * 1e6, r[0]))
e = O.search_by_bow(kk1, feats, valid, fv1, kk2, df, fv2, 0.7, False)
print("bit exact:", r[0] == e[0] and np.array_equal(r[1], e[1]))
lc = synth.make_local_map_case(3000, 1009, 43)
a = (lc["pts"], lc["descMP"], lc["kps2"], lc["desc2"], lc["held2"], lc["bounds"], lc["scale_factors"])
gl = api.GuidedMatcher(0, 0.8, True)
for _ in range(5):
    r = gl.SearchByProjectionMapPoints(*a, 3.0, False, 0.0)
t0 = time.perf_counter()
for _ in range(200):
    r = gl.SearchByProjectionMapPoints(*a, 3.0, False, 0.0)
print("SearchByProjection(map points) host call: %.1f us, %d matches" % ((time.perf_counter() - t0) / 200 * 1e6, r[0]))
e = O.search_by_projection_map_points(*a, 3.0, False, 0.0, 0.8)
print("bit exact:", r[0] == e[0] and np.array_equal(r[1], e[1]))
import torch
st = torch.cuda.current_stream().cuda_stream
gl.set_stream(st)
d = {k: torch.from_numpy(np.ascontiguousarray(v).view(np.uint8).reshape(-1).copy()).cuda() for k, v in
     dict(pts=lc["pts"], dmp=lc["descMP"], k2=lc["kps2"], d2=lc["desc2"], held=lc["held2"]).items()}
d_mc = torch.zeros(len(lc["kps2"]), dtype=torch.int32, device="cuda")
n1, n2 = len(lc["pts"]), len(lc["kps2"])
for _ in range(5):
    nm = gl.SearchByProjectionMapPoints_device(d["pts"].data_ptr(), d["dmp"].data_ptr(), n1, d["k2"].data_ptr(), d["d2"].data_ptr(), d["held"].data_ptr(), n2,
                                               lc["bounds"], lc["scale_factors"], d_mc.data_ptr(), 3.0)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    nm = gl.SearchByProjectionMapPoints_device(d["pts"].data_ptr(), d["dmp"].data_ptr(), n1, d["k2"].data_ptr(), d["d2"].data_ptr(), d["held"].data_ptr(), n2,
                                               lc["bounds"], lc["scale_factors"], d_mc.data_ptr(), 3.0)
torch.cuda.synchronize()
print("SearchByProjection(map points) device-resident call: %.1f us, %d matches, equal %s" %
      ((time.perf_counter() - t0) / 200 * 1e6, nm, np.array_equal(d_mc.cpu().numpy(), e[1])))
gl.set_stream(None)
