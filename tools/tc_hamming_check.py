#!/usr/bin/env python3
"""Tensor-core Hamming engine against the POPC engine and the oracle: parity on awkward shapes, then timing."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from eorb_slam_b200 import api, synth
import oracle_lib as O

def run(nq, ndb, seed, dup=False):
    db = synth.make_descriptor_db(ndb, seed)
    q, _ = synth.make_queries(db, nq, seed + 1)
    if dup and ndb > 10:
        db[ndb // 2] = db[3]; db[ndb - 1] = db[3]
    m = api.ORBmatcher(0.7)
    m.set_db(db)
    m.set_engine(m.HAMMING_POPC); a = m.search(q)
    m.set_engine(m.HAMMING_TENSOR); b = m.search(q)
    assert m.last_engine() == 1
    ok = all(np.array_equal(a[k], b[k]) for k in a.dtype.names)
    if not ok:
        bad = np.nonzero((a["best_idx"] != b["best_idx"]) | (a["best_dist"] != b["best_dist"]) | (a["second_dist"] != b["second_dist"]))[0]
        print("MISMATCH nq=%d ndb=%d: %d queries differ; first %s popc %s tensor %s" % (nq, ndb, len(bad), bad[:5], a[bad[:3]], b[bad[:3]]))
    else:
        print("ok nq=%d ndb=%d" % (nq, ndb))
    return ok

allok = True
for nq, ndb in [(128, 256), (128, 512), (1, 256), (100, 300), (129, 1000), (300, 70001), (2000, 200000)]:
    allok &= run(nq, ndb, nq + ndb, dup=True)
if allok and "--time" in sys.argv:
    ndb = 1 << 22
    db = synth.make_descriptor_db(ndb, 5); q, _ = synth.make_queries(db, 2000, 6)
    m = api.ORBmatcher(0.7); m.set_db(db)
    for eng in (0, 1):
        m.set_engine(eng); m.search(q)
        t0 = time.perf_counter()
        for _ in range(3): r = m.search(q)
        dt = (time.perf_counter() - t0) / 3
        print("engine %d: %.2f ms per 2000 x %d search = %.1f Gmatch/s" % (eng, dt * 1e3, ndb, 2000 * ndb / dt / 1e9))
print("ALL OK" if allok else "FAILED")
