#!/bin/bash
# Standard GPU-box cycle: GPU parity suite, bench line, in-graph trace without PDL (clean per-kernel times).
mkdir -p gpurun_out
TAG=${1:-c}
XARGS=${2:-}
timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -80 | tee gpurun_out/${TAG}_tests.log | tail -15
timeout -s KILL 500 python bench.py --steps 100 --warmup 10 $XARGS > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
head -c 300 gpurun_out/${TAG}_bench.json; echo; tail -2 gpurun_out/${TAG}_bench.err | cut -c1-300
timeout -s KILL 300 python bench.py --trace --no-cpu --no-extra --no-pdl > gpurun_out/${TAG}_trace_nopdl.json 2> gpurun_out/${TAG}_trace_nopdl.err
