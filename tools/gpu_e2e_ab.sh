#!/bin/bash
# A/B of the e2e input pipeline: one vs two concurrent zero-copy gathers
mkdir -p gpurun_out
for P in 1 2; do
timeout -s KILL 300 python bench.py --steps 60 --warmup 10 --no-cpu --no-extra --e2e-pack-streams $P > gpurun_out/e2e_ab_p$P.json 2> gpurun_out/e2e_ab_p$P.err
python tools/show_bench.py gpurun_out/e2e_ab_p$P.json | head -2
done
