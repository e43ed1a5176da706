#!/bin/bash
# pyramid tile height sweep (EORB_PYR_TH, no rebuild): per-stage time of the default launch set
cd /root/repo
for th in 96 112 144 160 192; do
  echo -n "EORB_PYR_TH=$th  "
  EORB_PYR_TH=$th python bench.py --steps 3 --warmup 3 --no-extras --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['stages']['pyramid']['ms_per_frame']*1e3,4))"
done
