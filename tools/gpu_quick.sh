#!/bin/bash
# Short GPU-box visit: full GPU parity suite + bench with both aggregation engines.
mkdir -p gpurun_out
TAG=${1:-quick}
timeout -s KILL 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 | tee gpurun_out/${TAG}_tests.log
timeout -s KILL 400 python bench.py --steps 100 --warmup 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 2500 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout -s KILL 300 python bench.py --steps 100 --warmup 10 --agg generic --no-cpu > gpurun_out/${TAG}_bench_generic.json 2> gpurun_out/${TAG}_bench_generic.err
tail -c 1200 gpurun_out/${TAG}_bench_generic.json; tail -3 gpurun_out/${TAG}_bench_generic.err
