#!/bin/bash
# Short GPU-box visit: full GPU parity suite + bench + in-graph trace (+ optional extra bench args as $2, ncu regex as $3).
mkdir -p gpurun_out
TAG=${1:-quick}
XARGS=${2:-}
KREGEX=${3:-}
timeout -s KILL 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 | tee gpurun_out/${TAG}_tests.log
timeout -s KILL 400 python bench.py --steps 100 --warmup 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 2500 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout -s KILL 300 python bench.py --trace --no-cpu > gpurun_out/${TAG}_trace.json 2> gpurun_out/${TAG}_trace.err
timeout -s KILL 200 python bench.py --gemm-trace --no-cpu --nbatches 2 > gpurun_out/${TAG}_gemm_trace.json 2> gpurun_out/${TAG}_gemm_trace.err
if [ -n "$XARGS" ]; then
timeout -s KILL 300 python bench.py --steps 100 --warmup 10 --no-cpu $XARGS > gpurun_out/${TAG}_bench_alt.json 2> gpurun_out/${TAG}_bench_alt.err
timeout -s KILL 300 python bench.py --trace --no-cpu $XARGS > gpurun_out/${TAG}_trace_alt.json 2> gpurun_out/${TAG}_trace_alt.err
fi
if [ -n "$KREGEX" ]; then
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k "regex:${KREGEX}" -s 8 -c 6 \
    -o gpurun_out/${TAG}_prof -f python bench.py --profile-only --steps 2 --warmup 3 --nbatches 2 > gpurun_out/${TAG}_ncu_full.log 2>&1
fi
