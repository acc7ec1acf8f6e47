#!/bin/bash
# One GPU-box visit: GEMM engine tests, full GPU parity suite, bench, ncu launch list + full capture.
# Every python process runs under its own hard timeout (a hung kernel must not hang the box).
mkdir -p gpurun_out
TAG=${1:-run}
timeout -s KILL 200 python -m pytest tests/test_gpu_gemm.py -q -s -k "gemm_nt or unaligned" 2>&1 | grep -v "^$" | tail -40 | tee gpurun_out/${TAG}_gemm.log
timeout -s KILL 200 python -m pytest tests/test_gpu_gemm.py -q -s -k "gemm_tn" 2>&1 | grep -v "^$" | tail -40 | tee gpurun_out/${TAG}_gemm_tn.log
if grep -q "failed\|Killed\|rror" gpurun_out/${TAG}_gemm.log; then ENG=ffma;
elif grep -q "failed\|Killed\|rror" gpurun_out/${TAG}_gemm_tn.log; then ENG=tcgen05-nt; else ENG=tcgen05; fi
echo "engine for the rest of this visit: $ENG" | tee -a gpurun_out/${TAG}_gemm.log
EAGCN_GEMM=$ENG timeout -s KILL 500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_gemm.py 2>&1 | tail -30 | tee gpurun_out/${TAG}_tests.log
timeout -s KILL 500 python bench.py --steps 100 --warmup 10 --gemm $ENG > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 1500 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout -s KILL 300 python bench.py --steps 100 --warmup 10 --gemm $ENG --layers-only --no-cpu 2>/dev/null | tail -1 | tee gpurun_out/${TAG}_layers_only.json
# launch list of 2 steady-state eager steps (3 warm-up steps skipped by kernel count is fragile -> profile all 5, aggregate offline)
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --profile-only --steps 2 --warmup 3 --nbatches 2 --gemm $ENG > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k "regex:gemm_tc|gemm_simt|agg_bwd|agg_fwd|pack_fill|bn_" -s 60 -c 16 \
    -o gpurun_out/${TAG}_prof -f python bench.py --profile-only --steps 2 --warmup 3 --nbatches 2 --gemm $ENG > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | tail -12
