#!/bin/bash
# A-in-TMEM bring-up: GEMM tests first (short timeout: a pipeline bug hangs rather than fails), then suite + A/B.
mkdir -p gpurun_out
TAG=${1:-atm}
timeout -s KILL 240 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | tail -30 | tee gpurun_out/${TAG}_gemm.log | tail -12
if ! grep -q " passed" gpurun_out/${TAG}_gemm.log || grep -q "failed" gpurun_out/${TAG}_gemm.log; then echo "GEMM tests not green: stop"; exit 1; fi
bash tools/gpu_ab_flag.sh $TAG "--tc-a-tmem" "3 0 1 2" tests
