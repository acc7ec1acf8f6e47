#!/bin/bash
# A/B of one bench.py flag on the other BASELINE configurations.   usage: gpu_ab_cfg.sh <tag> "<flag>" "<v1 v2>" "<cfg1 cfg2>"
mkdir -p gpurun_out
TAG=$1; FLAG=$2; VALS=$3; CFGS=$4
for i in 1 2; do
  for c in $CFGS; do
    for v in $VALS; do
      timeout -s KILL 300 python bench.py --config $c --steps 100 --warmup 10 --no-cpu --no-extra --no-global-bn $FLAG $v \
        > gpurun_out/${TAG}_${c}_${v}_$i.json 2> gpurun_out/${TAG}_${c}_${v}_$i.err
      python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_${c}_${v}_$i.json").read().strip().splitlines()[-1])
    print("$c $FLAG $v pass $i:", round(d["value"]), d["unit"], round(d["ms_per_step"], 4), "ms")
except Exception as e:
    print("$c $FLAG $v pass $i: failed", e)
PY
    done
  done
done
