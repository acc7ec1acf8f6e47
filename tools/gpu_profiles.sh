#!/bin/bash
# GPU-box visit for the committed profiles of a round: bench line, reference arm, in-graph traces, ncu launch list + full
# capture of one eager step's kernels, and the per-kernel table of the headline workload at a GPU-filling batch.
mkdir -p gpurun_out
TAG=${1:-r02}
timeout -s KILL 500 python bench.py --steps 100 --warmup 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
head -c 200 gpurun_out/${TAG}_bench.json; echo
timeout -s KILL 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err
head -c 200 gpurun_out/${TAG}_ref.json; echo
timeout -s KILL 300 python bench.py --trace --no-cpu --no-extra > gpurun_out/${TAG}_trace.json 2> gpurun_out/${TAG}_trace.err
timeout -s KILL 300 python bench.py --trace --no-cpu --no-extra --no-pdl > gpurun_out/${TAG}_trace_nopdl.json 2> gpurun_out/${TAG}_trace_nopdl.err
timeout -s KILL 300 python bench.py --batch 4096 --nbatches 2 --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/${TAG}_bench_b4096.json 2> gpurun_out/${TAG}_bench_b4096.err
head -c 200 gpurun_out/${TAG}_bench_b4096.json; echo
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --profile-only --steps 2 --warmup 3 --nbatches 2 --no-extra > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k "regex:layer_fwd_fused|gemm_tc|agg_bwd|pack_fill|bn_stat_apply|bn_bwd_partial|mm_tile|bn_act" -s 52 -c 34 \
    -o gpurun_out/${TAG}_prof -f python bench.py --profile-only --steps 2 --warmup 3 --nbatches 2 --no-extra > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out | grep ${TAG}_ | head -20
