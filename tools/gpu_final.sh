#!/bin/bash
# last call of a round on a nearly spent budget: the GPU suite and one short bench line, nothing else
mkdir -p gpurun_out
TAG=${1:-fin}
timeout -s KILL 200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/${TAG}_tests.log
timeout -s KILL 120 python bench.py --steps 100 --warmup 10 --no-cpu --no-extra --no-global-bn > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
head -c 260 gpurun_out/${TAG}_bench.json; echo
