"""2-rank probe (diagnostic): cost of the eager gradient all-reduce variants and whether an NCCL collective can be captured
into a CUDA graph on this stack.  torchrun --nproc-per-node 2 tools/experiments/nccl_probe.py [graph]"""
import os, sys, time
import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
flat = torch.ones(493904, device=dev)
work = torch.cuda.Stream(); comm = torch.cuda.Stream()
torch.cuda.set_stream(work)


def bench(name, fn, n=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    host = (time.perf_counter() - t0) / n * 1e6
    if rank == 0:
        print(f"{name:40s} device {e0.elapsed_time(e1) / n * 1e3:8.1f} us/iter   host {host:8.1f} us/iter", flush=True)


def same_sum():
    dist.all_reduce(flat); flat.mul_(0.5)


def same_avg():
    dist.all_reduce(flat, op=dist.ReduceOp.AVG)


def comm_avg():
    comm.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(comm):
        dist.all_reduce(flat, op=dist.ReduceOp.AVG)
    torch.cuda.current_stream().wait_stream(comm)


def comm_avg_view():
    comm.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(comm):
        dist.all_reduce(flat[:400000], op=dist.ReduceOp.AVG)
    torch.cuda.current_stream().wait_stream(comm)


if rank == 0:
    print("env TORCH_NCCL_ASYNC_ERROR_HANDLING =", os.environ.get("TORCH_NCCL_ASYNC_ERROR_HANDLING"), flush=True)
bench("same stream, SUM + mul", same_sum)
bench("same stream, AVG", same_avg)
bench("comm stream, AVG", comm_avg)
bench("comm stream, AVG, view", comm_avg_view)

if len(sys.argv) > 1 and sys.argv[1] == "graph":
    x = torch.ones(1 << 20, device=dev)
    g = torch.cuda.CUDAGraph()
    if rank == 0:
        print("capturing ...", flush=True)
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        x.mul_(1.0001)
        comm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(comm):
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)
        torch.cuda.current_stream().wait_stream(comm)
        x.add_(1.0)
    if rank == 0:
        print("captured; replaying ...", flush=True)
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    if rank == 0:
        print("replayed ok", flush=True)
    bench("graph replay (mul + AVG on comm + add)", g.replay)
dist.destroy_process_group()
