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
elif what == "nanstep":
    name, W, nst = sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    wk = wl.WORKLOADS[name]
    data = nb.validate_data_table(wk.tables())
    plan = nb.LikelihoodPlan(wk.model, wk.prior, data, wk.P)
    ens = nb.DeviceEnsemble(plan, W, seed=wl.SEED, use_graph=False)
    ens.set_state(wk.walkers(W))
    ens.load_draws(nst)
    ex = ens.ex
    found = False
    for t in range(nst):
        for split in range(2):
            plan._enqueue(ex, mv=ens._stretch(split))
            torch.cuda.synchronize()
            lnp = ex.lnp.cpu().numpy()
            bad = np.flatnonzero(np.isnan(lnp))
            if bad.size:
                w = int(bad[0])
                q = ex.pars[w].cpu().numpy()
                row = ex.row[w].cpu().numpy()
                print("NaN at step", t, "split", split, "proposal row", w, "q =", repr(q))
                print("prior", ex.prior[w].item(), "flux nan", np.isnan(row[:plan.N_E]).sum(),
                      "flux inf", np.isinf(row[:plan.N_E]).sum())
                print("flux", row[:plan.N_E])
                out = ex.outs[0][w].cpu().numpy()
                print("contract out nan", np.isnan(out).sum(), out[:6])
                p = ex.preps[plan.comps[0]["prep"]]
                xn, ds = p.xn[w].cpu().numpy(), p.ds1[w].cpu().numpy()
                print("xn nan/inf", np.isnan(xn).sum(), np.isinf(xn).sum(), xn[:4], xn[-4:])
                print("ds1 nan/inf", np.isnan(ds).sum(), np.isinf(ds).sum(), ds[:4], ds[-4:])
                print("pm", ex.pm.cpu().numpy().reshape(-1)[w * 8:(w + 1) * 8])
                for k, o_ in enumerate(ex.outs):
                    ov = o_[w].cpu().numpy()
                    print("comp", k, plan.comps[k]["kind"], "nan", np.isnan(ov).sum(), "inf",
                          np.isinf(ov).sum(), "first nan idx", np.flatnonzero(np.isnan(ov))[:5])
                coords = ens.coords.cpu().numpy()
                print("ensemble range:", coords.min(axis=0), coords.max(axis=0))
                l2, f2, _ = plan(q[None, :])
                print("same proposal through plan():", l2, "flux nan", np.isnan(f2).sum())
                found = True
                break
        if found:
            break
    print("done, found =", found)
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
