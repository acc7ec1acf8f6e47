#!/bin/bash
# GPU-box visit: the full GPU parity suite (not -x: every failure is listed), output kept under gpurun_out/.
mkdir -p gpurun_out
TAG=${1:-t}
timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -60 | tee gpurun_out/${TAG}_tests.log
