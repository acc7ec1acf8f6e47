#!/bin/bash
# 2-GPU visit: data-parallel parity test + 2-rank bench (both arms).
# torchrun is started in its OWN process group and the whole group is killed at the limit: `timeout` alone kills only
# torchrun, and ranks hung in a collective then keep the box (and the GPU budget) until gpurun's own limit.
mkdir -p gpurun_out
TAG=${1:-dp2}
run_group() {  # seconds, command...
  local limit=$1; shift
  setsid "$@" &
  local pid=$!
  ( sleep "$limit"; kill -KILL -- -"$pid" 2>/dev/null ) &
  local killer=$!
  wait "$pid"; local rc=$?
  kill "$killer" 2>/dev/null
  return $rc
}
timeout -s KILL 400 python -m pytest tests/test_gpu_dp.py -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/${TAG}_tests.log
run_group 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 1500 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
run_group 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err
tail -c 800 gpurun_out/${TAG}_ref.json; tail -3 gpurun_out/${TAG}_ref.err
