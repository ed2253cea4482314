"""Round-2 diagnostics (GPU box): C5 NaN hunt, host profile of PlanSampler."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench_workloads as wl
import naima_b200 as nb

what = sys.argv[1]
if what == "c5nan":
    wk = wl.WORKLOADS["C5"]
    data = nb.validate_data_table(wk.tables())
    plan = nb.LikelihoodPlan(wk.model, wk.prior, data, wk.P)
    rng = np.random.default_rng(1)
    P = np.column_stack([rng.uniform(20, 70, 8192), rng.uniform(-1.5, 5.5, 8192)])
    lnp, flux, blobs = plan(P)
    bad = np.isnan(lnp)
    print("NaN lnprob:", bad.sum(), "of", len(P))
    for p, f in list(zip(P[bad], flux[bad]))[:12]:
        print(p, "flux nan:", np.isnan(f).sum(), "inf:", np.isinf(f).sum(), f[:3], f[-3:])
    ens = nb.DeviceEnsemble(plan, 512, seed=wl.SEED)
    ens.set_state(wk.walkers(512))
    try:
        ens.run(30)
        print("device run: no NaN")
    except ValueError as e:
        print("device run:", e)
        lp = ens.chain_lp[:30].cpu().numpy()
        t, w = np.argwhere(np.isnan(lp))[0]
        print("first NaN at step", t, "walker", w, "state", ens.chain[t, w].cpu().numpy())
elif what == "profile":
    name = sys.argv[2] if len(sys.argv) > 2 else "C3"
    wk = wl.WORKLOADS[name]
    W = wk.walkers_per_gpu
    data = nb.validate_data_table(wk.tables())
    plan = nb.LikelihoodPlan(wk.model, wk.prior, data, wk.P)
    p0 = wk.walkers(W, spread=0.02)
    flush_buf = torch.empty(160 << 20, dtype=torch.uint8, device="cuda")
    sampler = nb.PlanSampler(W, wk.P, plan, seed=1)
    sampler._device().before_step = flush_buf.zero_
    state = sampler.run_mcmc(p0, 40)
    torch.cuda.synchronize()
    for n in (20, 200):
        t0 = time.perf_counter()
        state = sampler.run_mcmc(state, n)
        torch.cuda.synchronize()
        print("plain: %d steps, us/step %.1f" % (n, 1e6 * (time.perf_counter() - t0) / n))
    pr = cProfile.Profile()
    pr.enable()
    state = sampler.run_mcmc(state, 200)
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("tottime").print_stats(22)
