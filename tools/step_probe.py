"""Device time per sharded ensemble step under torchrun: graph / eager / graph without the
collective (diagnostic)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
import naima_b200 as nb
import bench_workloads as wl
from naima_b200 import parallel

W = 256 * world
xt, gt = wl.c3_tables(wl.c3_device_flux)
data = nb.validate_data_table([xt, gt])
plan = nb.LikelihoodPlan(wl.c3_model, wl.c3_prior, data, 4)
p0 = wl.walkers(wl.C3_PTRUE, W)
nsteps = 60


def run(ens, label):
    ens.set_state(p0)
    ens.load_draws(nsteps + 10)
    ens.run_loaded(10)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ens.run_loaded(nsteps)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1e3 / nsteps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("%-28s %8.1f us/step" % (label, t.item()), flush=True)


run(parallel.ShardedDeviceEnsemble(plan, W, seed=1, transport="fused"),
    "graph, replicated state (mc)")
run(parallel.ShardedDeviceEnsemble(plan, W, seed=1, transport="fused", multicast=False),
    "graph, replicated state (p2p)")
run(parallel.ShardedDeviceEnsemble(plan, W, seed=1, transport="p2p"), "graph + multicast stores")
run(parallel.ShardedDeviceEnsemble(plan, W, seed=1, transport="p2p", multicast=False),
    "graph + peer stores")
run(parallel.ShardedDeviceEnsemble(plan, W, seed=1, transport="nccl"), "graph + all-gather")
run(parallel.ShardedDeviceEnsemble(plan, W, seed=1, transport="nccl", use_graph=False),
    "eager + all-gather")
single = nb.DeviceEnsemble(plan, 256, seed=1)
single.set_state(p0[:256])
single.load_draws(nsteps + 10)
single.run_loaded(10)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
single.run_loaded(nsteps)
e1.record()
torch.cuda.synchronize()
if rank == 0:
    print("%-28s %8.1f us/step" % ("single-GPU 256 walkers", e0.elapsed_time(e1) * 1e3 / nsteps),
          flush=True)
torch.cuda.synchronize()
dist.barrier()
os._exit(0)
