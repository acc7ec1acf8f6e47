#!/bin/bash
# A/B of one environment switch on the other BASELINE configurations.   usage: gpu_ab_cfg_env.sh <tag> <VAR> "<cfg1 cfg2>"
mkdir -p gpurun_out
TAG=$1; VAR=$2; CFGS=$3
for i in 1 2; do
  for c in $CFGS; do
    for v in 1 0; do
      env $VAR=$v timeout -s KILL 300 python bench.py --config $c --steps 100 --warmup 10 --no-cpu --no-extra --no-global-bn \
        > gpurun_out/${TAG}_${c}_${v}_$i.json 2> gpurun_out/${TAG}_${c}_${v}_$i.err
      python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_${c}_${v}_$i.json").read().strip().splitlines()[-1])
    print("$c $VAR=$v pass $i:", round(d["value"]), d["unit"], round(d["ms_per_step"], 4), "ms")
except Exception as e:
    print("$c $VAR=$v pass $i: failed", e)
PY
    done
  done
done
