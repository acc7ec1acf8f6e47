#!/bin/bash
# GPU-box visit: the synthetic N x K x B sweep (BASELINE.json configs[4]) -> gpurun_out/r02_sweep.json
mkdir -p gpurun_out
timeout -s KILL 1200 python bench.py --config sweep > gpurun_out/r02_sweep_summary.json 2> gpurun_out/r02_sweep.err
tail -c 600 gpurun_out/r02_sweep_summary.json; tail -5 gpurun_out/r02_sweep.err | cut -c1-300
