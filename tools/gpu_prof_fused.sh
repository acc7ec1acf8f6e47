mkdir -p gpurun_out
timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k "regex:layer_fwd_fused" -s 6 -c 2 -o gpurun_out/r02d_prof -f python bench.py --profile-only --steps 2 --warmup 3 --nbatches 2 --no-extra > gpurun_out/r02d_ncu_full.log 2>&1
tail -2 gpurun_out/r02d_ncu_full.log
