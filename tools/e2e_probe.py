"""Wall time per next(gen) of the public-API loop (bench.py's e2e leg), to find stalls."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench_workloads as wl  # noqa: E402
import naima_b200 as nb  # noqa: E402

wk = wl.WORKLOADS["C3"]
W, steps = 256, 200
data = nb.validate_data_table(wk.tables())
plan = nb.LikelihoodPlan(wk.model, wk.prior, data, wk.P)
p0 = wk.walkers(W)
flush_buf = torch.empty(160 << 20, dtype=torch.uint8, device="cuda")
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    sampler = nb.PlanSampler(W, wk.P, plan, seed=wl.SEED, block=32)
    sampler._device().before_step = flush_buf.zero_
    state = sampler.run_mcmc(p0, 10)
    torch.cuda.synchronize()
    gen = sampler.sample(state, iterations=steps, store=True)
    ts = [time.perf_counter()]
    for k in range(steps):
        next(gen)
        ts.append(time.perf_counter())
    torch.cuda.synchronize()
    tend = time.perf_counter()
    gen.close()
    dt = np.diff(ts) * 1e3
    big = np.argsort(dt)[-4:][::-1]
    print("rep %d: %.4f ms/step; per-call ms: median %.4f, sum of calls > 1 ms: %.2f ms; largest: %s; "
          "final sync %.2f ms" % (rep, (tend - ts[0]) * 1e3 / steps, np.median(dt), dt[dt > 1].sum(),
                                   ["%d: %.2f" % (i, dt[i]) for i in big], (tend - ts[-1]) * 1e3))
