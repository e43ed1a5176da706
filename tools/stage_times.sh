#!/bin/bash
# quick A/B helper: ORB parity tests, then the per-stage times of the default launch set (no extras, no CPU baseline)
python -m pytest tests/test_gpu_orb.py tests/test_gpu_ref.py -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 --no-extras --no-cpu > gpurun_out/stage_bench.json 2> gpurun_out/stage_bench.err || tail -c 800 gpurun_out/stage_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/stage_bench.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["e2e"]["value"]), {k:round(v["ms_per_frame"]*1e3,4) for k,v in d["stages"].items()})
PY
