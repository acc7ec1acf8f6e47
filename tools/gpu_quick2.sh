#!/bin/bash
# Short GPU-box visit: GPU parity suite + one bench line (+ optional extra bench args as $2), everything under hard timeouts.
mkdir -p gpurun_out
TAG=${1:-q}
XARGS=${2:-}
timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -150 | tee gpurun_out/${TAG}_tests.log
timeout -s KILL 400 python bench.py --steps 100 --warmup 10 $XARGS > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 3000 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
