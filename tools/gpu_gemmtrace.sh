#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-gt}
for m in 3 0; do
  timeout -s KILL 200 python bench.py --gemm-trace --no-cpu --no-extra --tc-a-tmem $m > gpurun_out/${TAG}_gemmtrace_$m.json 2> gpurun_out/${TAG}_gemmtrace_$m.err
  timeout -s KILL 300 python bench.py --trace --no-cpu --no-extra --no-pdl --tc-a-tmem $m > gpurun_out/${TAG}_trace_nopdl_$m.json 2> gpurun_out/${TAG}_trace_nopdl_$m.err
done
python tools/show_gemm_trace.py gpurun_out/${TAG}_gemmtrace_3.json gpurun_out/${TAG}_gemmtrace_0.json
