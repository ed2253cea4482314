# -*- coding: utf-8 -*-
"""Walker sharding across GPUs (one process per GPU, torch.distributed).

The reference's only parallelism is a multiprocessing.Pool over walkers
(core.py:446-457).  Here the proposals of a half-ensemble are independent units:
rank r evaluates rows [r*per, (r+1)*per) of the proposal matrix and ONE
all-gather per half-step exchanges the log-probabilities (and the model-flux
blobs every rank needs to keep its replica of the ensemble state).  Positions
stay replicated without communication because every rank runs the same
proposal/accept random stream.  Because each walker is evaluated by exactly one
rank with identical code, chains are bitwise identical for any world size.

Backends: NCCL on device tensors (product), gloo on CPU tensors (host-logic
tests with a stand-in evaluator; the product path has no CPU compute).
"""
import ctypes

import numpy as np

from .sampler import DeviceEnsemble, EnsembleSampler

__all__ = ["shard_bounds", "ShardedSampler", "ShardedDeviceEnsemble"]


def _dist():
    import torch.distributed as dist

    return dist


def world_info(group=None):
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_bounds(n, world):
    """Equal-size padded shards: per = ceil(n/world); rank r owns [r*per, min((r+1)*per, n))."""
    per = -(-n // world)
    return per, [(min(r * per, n), min((r + 1) * per, n)) for r in range(world)]


class ShardedSampler(EnsembleSampler):
    """Host-driven stretch-move sampler whose half-ensemble evaluation is sharded.

    evaluator(q[per, P]) -> (lnp[per], flux[per, N_E] | None): a LikelihoodPlan
    wrapper in the product, any callable in the CPU tests."""

    def __init__(self, nwalkers, ndim, evaluator, seed=0, group=None, **kw):
        if seed is None:
            raise ValueError("ShardedSampler needs an explicit seed shared by all ranks")
        self.group = group
        self.rank, self.world = world_info(group)
        self._eval = evaluator
        self._plan = evaluator if hasattr(evaluator, "blobs_for") else None
        super().__init__(nwalkers, ndim, self._sharded_log_prob, vectorize=True, seed=seed, **kw)
        self.collectives = 0

    def _sharded_log_prob(self, q):
        import torch

        n = q.shape[0]
        per, bounds = shard_bounds(n, self.world)
        lo, hi = bounds[self.rank]
        mine = np.empty((per, q.shape[1]))
        mine[: hi - lo] = q[lo:hi]
        if hi - lo < per:  # pad with a valid row so the batch size is fixed
            mine[hi - lo:] = q[hi - 1] if hi > lo else q[0]
        res = self._eval(mine)
        lnp, flux = np.asarray(res[0], dtype=float), res[1]
        nf = 0 if flux is None else flux.shape[1]
        pack = np.empty((per, 1 + nf))
        pack[:, 0] = lnp
        if nf:
            pack[:, 1:] = flux
        if self.world > 1:
            dist = _dist()
            dev = "cuda" if dist.get_backend(self.group) == "nccl" else "cpu"
            local = torch.from_numpy(pack).to(dev)
            full = torch.empty((self.world * per, 1 + nf), dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(full, local, group=self.group)
            self.collectives += 1
            full = full.cpu().numpy()
        else:
            full = pack
        out = np.concatenate([full[r * per: r * per + (b - a)] for r, (a, b) in enumerate(bounds)])
        lnp_all = out[:, 0].copy()
        if not nf:
            return lnp_all
        flux_all = out[:, 1:]
        if self._plan is not None:
            from .sampler import BlobBatch

            return lnp_all, BlobBatch(self._plan, flux_all, [])
        return lnp_all, [(f,) for f in flux_all]


class ShardedDeviceEnsemble(DeviceEnsemble):
    """Device-resident ensemble step with the half-ensemble's proposals sharded over ranks.

    Per half-step, on every rank:
      set-up kernel (+ the proposals [lo, lo + per) of the half, computed in place)
      -> radiative components -> combine, writing packed records
         [blob record | lnprob | proposal] straight into this rank's slice of the gather buffer
      -> ONE in-place all-gather of the slices (NCCL over NVLink)
      -> accept step + chain append for all proposals (replicated, identical on every rank).
    The whole ensemble step (both halves, both collectives) is one CUDA-graph replay."""

    def __init__(self, plan, nwalkers, a=2.0, seed=0, store_blobs=True, group=None,
                 use_graph=True):
        from . import engine as eng

        if seed is None:
            raise ValueError("ShardedDeviceEnsemble needs an explicit seed shared by all ranks")
        self.group = group
        self.rank, self.world = world_info(group)
        super().__init__(plan, nwalkers, a=a, seed=seed, store_blobs=True, use_graph=use_graph)
        self.per, self.bounds = shard_bounds(self.Ns, self.world)
        if self.per * self.world != self.Ns:
            raise ValueError("the half-ensemble (%d) must divide evenly over %d ranks"
                             % (self.Ns, self.world))
        self.ld = plan.pack_width()
        self.pack_full = eng.zeros(self.Ns, self.ld)
        lo = self.rank * self.per
        self.pack_local = self.pack_full[lo:lo + self.per]
        self.ex = plan.executable(self.per, pack=self.pack_local)
        self.collectives = 0
        self.kernel_launches_per_step = 2 * (plan.launches_per_eval + 1)

    def set_state(self, coords, log_prob=None, rows=None):
        """Evaluate the initial ensemble sharded (unless given), then replicate."""
        from . import engine as eng

        coords = np.ascontiguousarray(coords, dtype=float)
        if log_prob is None or rows is None:
            W = coords.shape[0]
            per, bounds = shard_bounds(W, self.world)
            if per * self.world != W:
                raise ValueError("walkers must divide evenly over ranks")
            lo, hi = bounds[self.rank]
            lnp, rws = self.plan.eval_rows(coords[lo:hi])
            pack = np.concatenate([lnp[:, None], rws], axis=1)
            if self.world > 1:
                dist = _dist()
                local = eng.to_dev(pack)
                full = eng.zeros(W, pack.shape[1])
                dist.all_gather_into_tensor(full, local, group=self.group)
                pack = full.cpu().numpy()
            log_prob, rows = pack[:, 0].copy(), np.ascontiguousarray(pack[:, 1:])
        super().set_state(coords, log_prob, rows)

    def _enqueue_step(self):
        from . import engine as eng
        from ._lib import check, lib

        L = lib()
        for split in range(2):
            mv = self._stretch(split)
            mv.i0, mv.pars_ld = self.rank * self.per, self.ld
            self.plan._enqueue(self.ex, mv=mv, fuse_update=False)
            if self.world > 1:
                _dist().all_gather_into_tensor(self.pack_full, self.pack_local, group=self.group)
                self.collectives += 1
            check(L.nb_stretch_update_packed(ctypes.byref(self._stretch(split)),
                                             eng.ptr(self.pack_full), self.ld, eng.stream()),
                  "nb_stretch_update_packed")
