"""runs the non-ORB rows of bench.py once each (events, Hamming, guided matching, bag of words) — the command profiled for
profiles/r01m_ncu_full_other_rows.md (ncu -k regex picks the kernels); prints the extras' JSON"""
import json, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from eorb_slam_b200 import api
which = sys.argv[1:] or ["events", "hamming", "guided", "bow"]
out = {}
if "events" in which:
    out["events"] = bench.bench_events(api, torch, 0, 3, 3)
if "hamming" in which:
    out["hamming"] = bench.bench_hamming(api, torch, 0, 1, 1, 1, 0, None)
if "guided" in which:
    out["guided"] = bench.bench_guided(api, torch, 0, 1, 1)
if "bow" in which:
    out["bow"] = bench.bench_bow(api, torch, 0, 1, 1)

if "matchers" in which:   # every guided matcher + the vocabulary transform twice (ncu: -k regex:"guided_|frame_grid|search_by_bow|bow_")
    import numpy as np
    import oracle_lib as O
    from eorb_slam_b200 import synth
    c = synth.make_projection_case(5000, 5000, 40)
    prev = np.stack([c["kps1"]["x"], c["kps1"]["y"]], 1)
    lc = synth.make_local_map_case(3000, 1009, 43)
    pc = synth.make_projection_case(1009, 1009, 41)
    voc = synth.make_vocabulary_regular(10, 6, 3)
    rng = np.random.default_rng(4)
    leaves = np.flatnonzero(voc["is_leaf"])
    feats = voc["desc"][rng.choice(leaves, 1009)].copy()
    feats[:, :4] ^= rng.integers(0, 256, (1009, 4), dtype=np.uint8)
    v = api.ORBVocabulary(voc, 0)
    gm = api.GuidedMatcher(0, 0.9, True)
    for _ in range(2):
        r1 = gm.SearchForInitialization(c["kps1"], c["descMP"], c["kps2"], c["desc2"], c["bounds"], prev, 100)
        r2 = gm.SearchByProjection(pc["x3Dc"], pc["valid1"], pc["obs1"], pc["kps1"], pc["descMP"], pc["kps2"], pc["desc2"], pc["bounds"], pc["K"],
                                   pc["scale_factors"], 15.0)
        r3 = gm.SearchByProjectionMapPoints(lc["pts"], lc["descMP"], lc["kps2"], lc["desc2"], lc["held2"], lc["bounds"], lc["scale_factors"], 3.0)
        t = v.transform(feats, 4)
        fv = (t["fv_nodes"], t["fv_start"], t["fv_feats"])
        r4 = gm.SearchByBoW(pc["kps1"], feats, np.ones(1009, np.uint8), fv, pc["kps2"], feats, fv)
    out["matchers"] = [int(r1[0]), int(r2[0]), int(r3[0]), int(r4[0])]
    print(json.dumps(out["matchers"]))
print(json.dumps(out)[:3000])
