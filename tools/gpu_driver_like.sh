#!/bin/bash
# What the driver does at round end, on one box: smoke(), the GPU suite, both bench arms with the driver's flags.
mkdir -p gpurun_out
TAG=${1:-drv}
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.log
timeout -s KILL 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/${TAG}_tests.log
timeout -s KILL 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err
head -c 300 gpurun_out/${TAG}_ref.json; echo
timeout -s KILL 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
head -c 300 gpurun_out/${TAG}_bench.json; echo
