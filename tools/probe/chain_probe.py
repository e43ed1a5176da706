import sys, json; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch, bench
from eorb_slam_b200 import api
print(json.dumps(bench.bench_chain(api, torch, 0, 5, 3), indent=1))
