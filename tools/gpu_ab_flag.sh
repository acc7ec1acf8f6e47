#!/bin/bash
# GPU-box A/B of one bench.py flag: [pytest -m gpu on the files given in $3], then the N=1 bench value for each value of
# the flag, alternating, three passes.   usage: gpu_ab_flag.sh <tag> "<flag>" "<v1 v2 ...>" ["pytest args"]
mkdir -p gpurun_out
TAG=$1; FLAG=$2; VALS=$3; PYT=${4:-tests}
timeout -s KILL 900 python -m pytest $PYT -m gpu -q -x 2>&1 | tail -40 | tee gpurun_out/${TAG}_tests.log | tail -15
for i in 1 2 3; do
  for v in $VALS; do
    timeout -s KILL 300 python bench.py --steps 200 --warmup 20 --no-cpu --no-extra --no-global-bn $FLAG $v \
      > gpurun_out/${TAG}_${v}_$i.json 2> gpurun_out/${TAG}_${v}_$i.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_${v}_$i.json").read().strip().splitlines()[-1])
    print("$FLAG $v pass $i:", round(d["value"]), d["unit"], round(d["ms_per_step"], 4), "ms  launches", d.get("gpu_launches"))
except Exception as e:
    print("$FLAG $v pass $i: failed", e)
PY
  done
done
