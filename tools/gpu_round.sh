#!/bin/bash
# Full GPU-box visit of a round: parity suite, both bench arms, in-graph traces, ncu launch list + full capture.
# Every python process runs under its own hard timeout (a hung kernel must not hang the box).
mkdir -p gpurun_out
TAG=${1:-run}
timeout -s KILL 600 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/${TAG}_tests.log
timeout -s KILL 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err
tail -c 600 gpurun_out/${TAG}_ref.json
timeout -s KILL 500 python bench.py --steps 100 --warmup 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 1500 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout -s KILL 300 python bench.py --steps 100 --warmup 10 --layers-only --no-cpu 2>/dev/null | tail -1 > gpurun_out/${TAG}_layers_only.json
# in-graph per-kernel durations: with PDL (as shipped; kernels overlap) and without (clean per-kernel times)
timeout -s KILL 300 python bench.py --trace --no-cpu > gpurun_out/${TAG}_trace.json 2> gpurun_out/${TAG}_trace.err
timeout -s KILL 300 python bench.py --trace --no-cpu --no-pdl > gpurun_out/${TAG}_trace_nopdl.json 2> gpurun_out/${TAG}_trace_nopdl.err
timeout -s KILL 200 python bench.py --gemm-trace --no-cpu --nbatches 2 > gpurun_out/${TAG}_gemm_trace.json 2> gpurun_out/${TAG}_gemm_trace.err
# ncu: launch list of eager steps (3 warm-up + 2 profiled; aggregate offline), then a full capture of the large kernels
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --profile-only --steps 2 --warmup 3 --nbatches 2 > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k "regex:gemm_tc|agg_bwd|agg_fwd|pack_fill|bn_stat_apply|bn_bwd_partial|mm_tile|bn_act" -s 60 -c 30 \
    -o gpurun_out/${TAG}_prof -f python bench.py --profile-only --steps 2 --warmup 3 --nbatches 2 > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | grep ${TAG}
