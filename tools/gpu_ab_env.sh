#!/bin/bash
# A/B of one environment switch on the GPU box: GPU parity suite once, then the N=1 bench value with VAR=0 / VAR=1
# (three alternating passes each, so clock drift shows).   usage: gpu_ab_env.sh <tag> <VAR>
mkdir -p gpurun_out
TAG=${1:-ab}; VAR=${2:-EAGCN_BWD_TICKETS}
timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee gpurun_out/${TAG}_tests.log | tail -12
for i in 1 2 3; do
  for v in 1 0; do
    env $VAR=$v timeout -s KILL 300 python bench.py --steps 200 --warmup 20 --no-cpu --no-extra --no-global-bn \
      > gpurun_out/${TAG}_${v}_$i.json 2> gpurun_out/${TAG}_${v}_$i.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_${v}_$i.json").read().strip().splitlines()[-1])
    print("$VAR=$v pass $i:", round(d["value"]), d["unit"], d["ms_per_step"], "ms  launches", d.get("gpu_launches"))
except Exception as e:
    print("$VAR=$v pass $i: failed", e)
PY
  done
done
