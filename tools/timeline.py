"""Per-kernel device times of one likelihood half-step on the C3 workload, measured with
CUDA events between launches while the GPU is kept busy by a blocker (so host launch
latency does not leak into the numbers).  Diagnostic; prints averages in microseconds."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import naima_b200 as nb
from naima_b200 import engine as eng
import bench_workloads as wl
from naima_b200._lib import check, lib

W = 256
xt, gt = wl.c3_tables(wl.c3_device_flux)
data = nb.validate_data_table([xt, gt])
plan = nb.LikelihoodPlan(wl.c3_model, wl.c3_prior, data, 4, use_graph=False)
p0 = wl.walkers(wl.C3_PTRUE, W)
ex = plan.executable(W // 2)
ex.pars.copy_(eng.to_dev(p0[: W // 2]))
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
L = lib()


def stages():
    return plan.stages(ex)


for cold in (True, False):
    tot = {}
    reps = 30
    for _ in range(reps + 3):
        for _ in range(6):
            flush_buf.zero_()  # blocker (and the L2 flush when cold)
        if not cold:
            for name, fn in stages():
                fn()  # warm L2 behind the blocker
        evs = [torch.cuda.Event(enable_timing=True)]
        evs[0].record()
        names = []
        for name, fn in stages():
            fn()
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            evs.append(e)
            names.append(name)
        torch.cuda.synchronize()
        if _ >= 3:
            for k, name in enumerate(names):
                tot[name] = tot.get(name, 0.0) + evs[k].elapsed_time(evs[k + 1]) * 1e3
    print("cold L2" if cold else "warm L2", {k: round(v / reps, 2) for k, v in tot.items()},
          "sum", round(sum(tot.values()) / reps, 2))
