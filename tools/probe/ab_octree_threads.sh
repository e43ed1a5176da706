#!/bin/bash
# A/B: threads per block of the launch-set octree kernel (rebuilds on the box)
cd /root/repo
cp eorb_slam_b200/csrc/orb_kernels.cu /tmp/orb_kernels.cu.bak
for nt in 128 64 96 192 256; do
  cp /tmp/orb_kernels.cu.bak eorb_slam_b200/csrc/orb_kernels.cu
  sed -i "s/octree_kernel<128><<<grd, 128,/octree_kernel<$nt><<<grd, $nt,/; s/cudaFuncSetAttribute(octree_kernel<128>,/cudaFuncSetAttribute(octree_kernel<$nt>,/" eorb_slam_b200/csrc/orb_kernels.cu
  python -c "from eorb_slam_b200 import build as b; b.build_lib()" > /dev/null 2>&1
  echo -n "octree threads $nt: "
  python bench.py --steps 3 --warmup 3 --no-extras --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['stages']['octree']['ms_per_frame']*1e3,4))"
done
cp /tmp/orb_kernels.cu.bak eorb_slam_b200/csrc/orb_kernels.cu
