#!/bin/bash
# A/B of orient_desc_kernel build parameters: rebuild on the box with the given defines, run the stage timing
cd /root/repo
for cfg in "4 11" "8 5" "2 22" "4 12"; do
  set -- $cfg
  sed -i "s/^#define EORB_KP_GROUP .*/#define EORB_KP_GROUP $1/; s/^#define EORB_OD_MINB .*/#define EORB_OD_MINB $2/" eorb_slam_b200/csrc/orb_kernels.cu
  python -c "from eorb_slam_b200 import build as b; b.build_lib()" > /dev/null 2>&1
  echo "KP_GROUP=$1 MINB=$2"
  python bench.py --steps 3 --warmup 3 --no-extras --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['stages']['orient_desc']['ms_per_frame']*1e3,4))"
done
