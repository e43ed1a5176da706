#!/bin/bash
# A/B: shuffle-based block scan also for the 128-thread launch-set octree blocks (rebuilds on the box)
cd /root/repo
for thr in 128 127; do
  sed -i "s/    if (NTC > [0-9]*) {/    if (NTC > $thr) {/" eorb_slam_b200/csrc/octree_core.cuh
  python -c "from eorb_slam_b200 import build as b; b.build_lib()" > /dev/null 2>&1
  echo "shuffle scan when NTC > $thr"
  python bench.py --steps 3 --warmup 3 --no-extras --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['stages']['octree']['ms_per_frame']*1e3,4), round(d['stages']['index']['ms_per_frame']*1e3,4))"
done
