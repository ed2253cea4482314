"""Run under torchrun with N >= 2 GPUs: the sharded device ensemble and the sharded
PlanSampler must reproduce the single-GPU chain BITWISE (every walker is evaluated by
exactly one rank with identical code; accept decisions are replicated).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \\
        --master-port 29511 tests/multi/check_sharded.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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

W, nsteps, seed = 64 * world, 12, 5
xt, gt = wl.c3_tables(wl.c3_device_flux)
data = nb.validate_data_table([xt, gt])
plan = nb.LikelihoodPlan(wl.c3_model, wl.c3_prior, data, 4)
p0 = wl.walkers(wl.C3_PTRUE, W)

ref = nb.DeviceEnsemble(plan, W, seed=seed)  # single GPU, no communication
ref.set_state(p0)
rchain, rlp, rrows = ref.run(nsteps)
racc = ref.acceptance_counts.copy()

variants = (("nccl", False), ("fused", False), ("fused", True))
only = os.environ.get("NB_CHECK_TRANSPORTS")  # e.g. "nccl,fused" to shorten a many-GPU run
if only:
    variants = tuple(v for v in variants if v[0] in only.split(","))
want2 = None
for transport, mc in variants:
    sh = parallel.ShardedDeviceEnsemble(plan, W, seed=seed, transport=transport, multicast=mc)
    sh.set_state(p0)
    chain, lp, rows = sh.run(nsteps)
    assert sh.transport == transport
    assert sh.transport != "nccl" or sh.collectives >= 2, sh.collectives
    assert np.array_equal(chain, rchain), np.abs(chain - rchain).max()
    assert np.array_equal(lp, rlp)
    assert np.array_equal(rows, rrows)
    assert np.array_equal(sh.acceptance_counts, racc)
    # a second block on the same ensemble (buffers, flags and generation counter carry on)
    chain2, _, _ = sh.run(5)
    if want2 is None:
        want2, _, _ = ref.run(5)
    assert np.array_equal(chain2, want2)
    if rank == 0:
        print("transport %s (multicast %s): chain bitwise equal"
              % (sh.transport, getattr(sh, "uses_multicast", False)), flush=True)

ps1 = nb.PlanSampler(W, 4, plan, seed=seed, sharded=False)
ps1.run_mcmc(p0, nsteps)
ps = nb.PlanSampler(W, 4, plan, seed=seed)
assert ps.sharded
ps.run_mcmc(p0, nsteps)
assert np.array_equal(ps.get_chain(), ps1.get_chain())
assert np.array_equal(ps.get_log_prob(), ps1.get_log_prob())
if rank == 0:  # the blob records are read back on rank 0 only
    b, b1 = ps.get_blobs()[-1, 3], ps1.get_blobs()[-1, 3]
    assert np.array_equal(b[0].value, b1[0].value) and np.array_equal(b[1].value, b1[1].value)
else:
    assert ps.get_blobs() is None
# several steps per CUDA graph, with steps left over that do not fill a graph
ps4 = nb.PlanSampler(W, 4, plan, seed=seed, block=8, steps_per_graph=4)
ps4.run_mcmc(p0, nsteps - 2)
ps4.run_mcmc(None, 2)
assert np.array_equal(ps4.get_chain(), ps1.get_chain())
assert np.array_equal(ps4.get_log_prob(), ps1.get_log_prob())
if rank == 0:
    print("sharded PlanSampler, 4 steps per graph: chain bitwise equal", flush=True)
# all ranks hold the same chain
t = torch.from_numpy(ps.get_chain().copy()).cuda()
full = [torch.empty_like(t) for _ in range(world)]
dist.all_gather(full, t)
assert all(torch.equal(full[0], f) for f in full)
dist.barrier()
if rank == 0:
    print("sharded == single-GPU chain (bitwise) on %d ranks: OK" % world, flush=True)
# graphs with captured NCCL collectives are alive: leave without tearing the group down
torch.cuda.synchronize()
os._exit(0)
