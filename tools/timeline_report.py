"""Reads the per-rank half-step time stamps bench.py --timeline wrote and prints where a
sharded half-step's time goes (median over the timed half-steps, microseconds).

columns per row h: nb_stretch.timeline's stamps (include/naima_b200.h), last column: number of
half-steps run."""
import sys

import numpy as np


def main(prefix, world, last):
    tls = [np.load("%s%d.npy" % (prefix, r)) for r in range(world)]
    gen = int(tls[0][0, -1])
    cap = tls[0].shape[0]
    hs = np.arange(max(gen - last, 1), gen - 1)  # rows are 2 t + split  # half-steps whose successor also exists
    out = {}
    for r, t in enumerate(tls):
        a, b = t[hs % cap].astype(float), t[(hs + 1) % cap].astype(float)
        seg = {
            "wait": a[:, 1] - a[:, 0],
            "evaluate": a[:, 2] - a[:, 1],
            "accept": a[:, 4] - a[:, 2],
            "release": np.where(a[:, 3] > 0, a[:, 4] - a[:, 3], 0.0),
            "to_next_wait": b[:, 0] - a[:, 4],
            "half_step": b[:, 0] - a[:, 0],
        }
        if t.shape[1] > 12 and a[:, 5].any():  # per-kernel stamps, relative to the set-up's start
            for name, c0, c1 in (("synchrotron", 5, 6), ("set-up", 0, 7), ("energy blobs", 8, 9),
                                 ("contraction", 10, 11), ("accept", 2, 4)):
                if a[:, c1].any():
                    seg["%s starts at" % name] = a[:, c0] - a[:, 0]
                    seg["%s ends at" % name] = a[:, c1] - a[:, 0]
        for par in (0, 1):
            m = (hs % 2) == par
            for k in [k for k in seg if "[" not in k and k not in ("wait", "release", "accept")]:
                seg["%s[split %d]" % (k, par)] = seg[k][m]
        out[r] = {k: (float(np.median(v)) / 1e3, float(np.percentile(v, 90)) / 1e3)
                  for k, v in seg.items()}
        print("rank %d  (median / p90 us over %d half-steps)" % (r, len(hs)))
        for k, (m, p) in out[r].items():
            print("   %-44s %7.2f %7.2f" % (k, m, p))
    if world == 2 and tls[0][:, 3].any():
        # one-way flag latency, clock offset removed NTP-style: the flag rank A stored at
        # a[4] lets rank B leave its wait at b[1] >= a[4] + latency + offset(B - A)
        A, B = tls[0], tls[1]
        d01 = (B[(hs + 1) % cap, 1] - A[hs % cap, 4]).astype(float)
        d10 = (A[(hs + 1) % cap, 1] - B[hs % cap, 4]).astype(float)
        print("min over half-steps of (peer leaves wait - flag stored): 0->1 %.2f us, 1->0 %.2f us;"
              " their mean bounds the one-way latency: %.2f us"
              % (d01.min() / 1e3, d10.min() / 1e3, (d01.min() + d10.min()) / 2e3))
        # who was the later rank, and by how much the earlier one waited for it
        print("median of max(wait) over the two ranks: %.2f us"
              % (np.median(np.maximum(A[hs % cap, 1] - A[hs % cap, 0],
                                      B[hs % cap, 1] - B[hs % cap, 0])) / 1e3))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 300)
