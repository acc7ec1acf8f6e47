#!/bin/bash
# Diagnostic GPU visit: in-graph kernel trace, GEMM pipeline trace, ncu full capture of selected kernels.
mkdir -p gpurun_out
TAG=${1:-diag}
KREGEX=${2:-agg_bwd_tile|agg_fwd_tile|stat_reduce}
timeout -s KILL 300 python -m pytest tests -m gpu -q -x -k "gemm or agree or layer_vs_oracle" 2>&1 | tail -8 | tee gpurun_out/${TAG}_tests.log
timeout -s KILL 300 python bench.py --trace --no-cpu > gpurun_out/${TAG}_trace.json 2> gpurun_out/${TAG}_trace.err
tail -c 300 gpurun_out/${TAG}_trace.err
timeout -s KILL 200 python bench.py --gemm-trace --no-cpu --nbatches 2 > gpurun_out/${TAG}_gemm_trace.json 2> gpurun_out/${TAG}_gemm_trace.err
tail -c 300 gpurun_out/${TAG}_gemm_trace.err
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k "regex:${KREGEX}" -s 12 -c 8 \
    -o gpurun_out/${TAG}_prof -f python bench.py --profile-only --steps 2 --warmup 3 --nbatches 2 > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | grep ${TAG}
