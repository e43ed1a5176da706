"""tensor Hamming engine: time of one 2000 x 4 Mi search with parts of the pipeline switched off (EORB_HT_PROBE bit 1: no epilogue
work, 2: no expansion, 4: no MMA)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from eorb_slam_b200 import api, synth
ndb = 1 << 22
db = synth.make_descriptor_db(ndb, 5); q, _ = synth.make_queries(db, 2000, 6)
m = api.ORBmatcher(0.7); m.set_db(db); m.set_engine(1)
for _ in range(2): m.search(q)
t0 = time.perf_counter()
for _ in range(5): m.search(q)
print("probe %s: %.3f ms" % (os.environ.get("EORB_HT_PROBE", "0"), (time.perf_counter() - t0) / 5 * 1e3))
