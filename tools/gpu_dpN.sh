#!/bin/bash
# N-GPU visit: data-parallel parity test + N-rank bench with the in-graph (overlapped) and the eager gradient all-reduce.
# torchrun is started in its OWN process group and the whole group is killed at the limit: `timeout` alone kills only
# torchrun, and ranks hung in a collective then keep the box (and the GPU budget) until gpurun's own limit.
mkdir -p gpurun_out
TAG=${1:-dp}
N=${2:-2}
STEPS=${3:-100}
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
if [ "$N" = "2" ]; then
timeout -s KILL 300 python -m pytest tests/test_gpu_dp.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/${TAG}_tests.log
fi
for MODE in ${4:-eager}; do
run_group 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
    bench.py --gpus $N --steps $STEPS --warmup 10 --allreduce $MODE --no-extra > gpurun_out/${TAG}_bench_${MODE}.json 2> gpurun_out/${TAG}_bench_${MODE}.err
echo "== $MODE rc=$?"; head -c 250 gpurun_out/${TAG}_bench_${MODE}.json; echo; grep -v "Warning\|warn\|^$\|run_backward" gpurun_out/${TAG}_bench_${MODE}.err | tail -5 | cut -c1-300
done
