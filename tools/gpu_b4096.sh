mkdir -p gpurun_out
timeout -s KILL 300 python bench.py --batch 4096 --nbatches 2 --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r02f_bench_b4096.json 2> gpurun_out/r02f_bench_b4096.err
head -c 300 gpurun_out/r02f_bench_b4096.json; tail -3 gpurun_out/r02f_bench_b4096.err | cut -c1-300
