import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from eorb_slam_b200 import api, synth
ndb = 1 << 22
db = synth.make_descriptor_db(ndb, 5); q, _ = synth.make_queries(db, 2000, 6)
m = api.ORBmatcher(0.7); m.set_db(db); m.set_engine(1)
for _ in range(3): m.search(q)
