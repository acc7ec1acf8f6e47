#!/bin/bash
# Short GPU-box visit: full GPU parity suite + bench (+ optional in-graph trace).
mkdir -p gpurun_out
TAG=${1:-quick}
timeout -s KILL 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 | tee gpurun_out/${TAG}_tests.log
timeout -s KILL 400 python bench.py --steps 100 --warmup 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 2500 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout -s KILL 300 python bench.py --trace --no-cpu > gpurun_out/${TAG}_trace.json 2> gpurun_out/${TAG}_trace.err
timeout -s KILL 200 python bench.py --gemm-trace --no-cpu --nbatches 2 > gpurun_out/${TAG}_gemm_trace.json 2> gpurun_out/${TAG}_gemm_trace.err
