"""Gantt data of one one-launch likelihood evaluation (nb_program_launch's trace): runs the
C3 half-ensemble evaluation a few times and prints, per work-item kind, when its CTAs started
and ended relative to the first CTA, their duration, and how many were resident over time."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench_workloads as wl  # noqa: E402
import naima_b200 as nb  # noqa: E402
from naima_b200 import engine as eng  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
W = int(sys.argv[2]) if len(sys.argv) > 2 else 128
wk = wl.WORKLOADS[name]
data = nb.validate_data_table(wk.tables())
plan = nb.LikelihoodPlan(wk.model, wk.prior, data, wk.P)
plan.use_graph = False
ex = plan.executable(W)
ex.graph = None
ex.trace = eng.zeros(4 * 8192, dtype=torch.int64)
p0 = wk.walkers(W)
ex.pars.copy_(eng.to_dev(p0))
flush = torch.empty(160 << 20, dtype=torch.uint8, device="cuda")
for rep in range(4):
    if len(sys.argv) > 3:
        flush.zero_()
    plan._enqueue(ex)
    torch.cuda.synchronize()
t = ex.trace.cpu().numpy().reshape(-1, 4)
t = t[t[:, 1] > 0]
t0 = t[:, 0].min()
names = {0: "set-up", 1: "synchrotron", 2: "table", 3: "combine", 4: "pdist"}
print("%d items, kernel span %.1f us" % (len(t), (t[:, 1].max() - t0) / 1e3))
for key in np.unique(t[:, 2]):
    m = t[:, 2] == key
    s, e = (t[m, 0] - t0) / 1e3, (t[m, 1] - t0) / 1e3
    print("%-12s #%d  n=%4d  start %6.1f..%6.1f  end %6.1f..%6.1f  duration med %5.1f max %5.1f"
          % (names[int(key) // 16], int(key) % 16, m.sum(), s.min(), s.max(), e.min(), e.max(),
             np.median(e - s), (e - s).max()))
edges = np.arange(0, (t[:, 1].max() - t0) / 1e3 + 2, 2.0)
print("resident CTAs every 2 us:", [int(((t[:, 0] - t0) / 1e3 <= x).sum() - ((t[:, 1] - t0) / 1e3 <= x).sum())
                                    for x in edges])
print("SMs used:", len(np.unique(t[:, 3])))
