#!/bin/bash
mkdir -p gpurun_out
run_group() { local limit=$1; shift; setsid "$@" & local pid=$!; ( sleep "$limit"; kill -KILL -- -"$pid" 2>/dev/null ) & local k=$!; wait "$pid"; local rc=$?; kill "$k" 2>/dev/null; return $rc; }
echo "=== default env"; run_group 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 tools/experiments/nccl_probe.py 2>&1 | grep -v "^$\|Warning\|warn" | tail -8
echo "=== ASYNC_ERROR_HANDLING=0 + graph"; TORCH_NCCL_ASYNC_ERROR_HANDLING=0 run_group 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 tools/experiments/nccl_probe.py graph 2>&1 | grep -v "^$\|Warning\|warn" | tail -12
echo "=== default env + graph"; run_group 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29603 tools/experiments/nccl_probe.py graph 2>&1 | grep -v "^$\|Warning\|warn" | tail -12
