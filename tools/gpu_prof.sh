#!/bin/bash
# GPU-box visit for profiles: in-graph CUPTI trace (no PDL: clean per-kernel times) + ncu full capture of selected kernels.
mkdir -p gpurun_out
TAG=${1:-p}
KREGEX=${2:-layer_fwd_fused}
SKIP=${3:-6}
COUNT=${4:-2}
timeout -s KILL 300 python bench.py --trace --no-cpu --no-extra --no-pdl > gpurun_out/${TAG}_trace_nopdl.json 2> gpurun_out/${TAG}_trace_nopdl.err
timeout -s KILL 300 python bench.py --trace --no-cpu --no-extra > gpurun_out/${TAG}_trace.json 2> gpurun_out/${TAG}_trace.err
timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k "regex:${KREGEX}" -s ${SKIP} -c ${COUNT} \
    -o gpurun_out/${TAG}_prof -f python bench.py --profile-only --steps 2 --warmup 3 --nbatches 2 --no-extra > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out | grep ${TAG}
