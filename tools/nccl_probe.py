"""Time the half-step's collective in isolation (eager / graph, in-place / out-of-place)."""
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
per, ld = 128, 70
full = torch.zeros(per * world, ld, dtype=torch.float64, device="cuda")
local_view = full[rank * per:(rank + 1) * per]
local = torch.ones(per, ld, dtype=torch.float64, device="cuda")


def timed(fn, n=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


res = {}
res["eager_inplace"] = timed(lambda: dist.all_gather_into_tensor(full, local_view))
res["eager_outofplace"] = timed(lambda: dist.all_gather_into_tensor(full, local))
for name, src in (("graph_inplace", local_view), ("graph_outofplace", local)):
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            dist.all_gather_into_tensor(full, src)
    torch.cuda.current_stream().wait_stream(s)
    res[name] = timed(g.replay)
# a graph with 2 collectives separated by small kernels (like one ensemble step)
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    with torch.cuda.graph(g, stream=s):
        for _ in range(2):
            local_view.add_(1.0)
            dist.all_gather_into_tensor(full, local_view)
            full.mul_(0.5)
torch.cuda.current_stream().wait_stream(s)
res["graph_step_like_x2"] = timed(g.replay)
if rank == 0:
    print("world", world, {k: round(v, 1) for k, v in res.items()}, "us per call", flush=True)
torch.cuda.synchronize()
os._exit(0)
