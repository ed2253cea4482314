"""Host-side time breakdown of PlanSampler on the C3 workload (diagnostic)."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import naima_b200 as nb
import bench_workloads as wl

W = 256
xt, gt = wl.c3_tables(wl.c3_device_flux)
data = nb.validate_data_table([xt, gt])
plan = nb.LikelihoodPlan(wl.c3_model, wl.c3_prior, data, 4)
p0 = wl.walkers(wl.C3_PTRUE, W)
sampler = nb.PlanSampler(W, 4, plan, seed=1)
state = sampler.run_mcmc(p0, 20)
torch.cuda.synchronize()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
t0 = time.perf_counter()
state = sampler.run_mcmc(state, n)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("plain: us/step", 1e6 * dt / n)
pr = cProfile.Profile()
pr.enable()
state = sampler.run_mcmc(state, n)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
