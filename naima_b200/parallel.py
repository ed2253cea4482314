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

    transport "fused" (default when symmetric memory can be set up): the ensemble state
    (coords, lp, blob records) and the chain live in symmetric memory on every rank; the
    combine kernel's accept step writes every rank's copy of the walkers it decides (one
    multimem.st per element through the NVSwitch when `multicast`, else one store per peer)
    and its last CTA raises per-rank flags that the next half-step's kernels wait on.  The
    sharded step has exactly the launches of the single-GPU step and no collective.

    transport "nccl": per half-step every rank's combine kernel writes packed records
    [blob record | lnprob | proposal] into its slice of a gather buffer, ONE in-place
    all-gather (NCCL over NVLink) exchanges the slices and a replicated accept kernel runs
    identically on every rank.  The north_star's "single all-gather per step" design; kept
    as the fallback and as the cross-check of the fused transport.

    Either way the whole ensemble step is one CUDA-graph replay and the chain is bitwise
    identical to the single-GPU chain (tests/multi/check_sharded.py)."""

    def __init__(self, plan, nwalkers, a=2.0, seed=0, store_blobs=True, group=None,
                 use_graph=True, transport="auto", multicast=True, timeline=False):
        self.multicast = multicast
        from . import engine as eng

        if seed is None:
            raise ValueError("ShardedDeviceEnsemble needs an explicit seed shared by all ranks")
        if transport not in ("auto", "fused", "nccl"):
            raise ValueError("transport must be 'auto', 'fused' or 'nccl'")
        self.group = group
        self.rank, self.world = world_info(group)
        super().__init__(plan, nwalkers, a=a, seed=seed, store_blobs=store_blobs,
                         use_graph=use_graph, timeline=timeline)
        self.per, self.bounds = shard_bounds(self.Ns, self.world)
        if self.per * self.world != self.Ns:
            raise ValueError("the half-ensemble (%d) must divide evenly over %d ranks"
                             % (self.Ns, self.world))
        self.ld = plan.pack_width()
        self.collectives = 0
        self.transport = "nccl"
        lo = self.rank * self.per
        if self.world > 1 and transport in ("auto", "fused"):
            try:
                self._setup_fused()
                self.transport = "fused"
            except Exception as e:  # no symmetric memory on this build / topology
                if transport == "fused":
                    raise
                import warnings

                warnings.warn("naima_b200: replicated-state transport unavailable (%r); using "
                              "the NCCL all-gather" % (e,))
                for name in ("peers_fused", "_flags_local"):
                    self.__dict__.pop(name, None)
                super().__init__(plan, nwalkers, a=a, seed=seed, store_blobs=store_blobs,
                                 use_graph=use_graph)
        if self.transport == "fused":
            self.ex = plan.executable(self.per)
            self.kernel_launches_per_step = 2 * plan.launches_per_eval
            return
        self.pack_full = eng.zeros(self.Ns, self.ld)
        self.pack_local = self.pack_full[lo:lo + self.per]
        self.ex = plan.executable(self.per, pack=self.pack_local)
        self.kernel_launches_per_step = 2 * (plan.launches_per_eval + 1)

    def _setup_fused(self):
        """Replicated state in symmetric memory: every rank holds coords / lp / blob records /
        acceptance counts (arena 0) and the chain (arena 1, per block of steps); the combine
        kernel's accept step writes all copies."""
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        from . import engine as eng
        from ._lib import NB_MAX_PEERS, nb_peers

        if self.world > NB_MAX_PEERS:
            raise ValueError("more ranks than NB_MAX_PEERS")
        self._group = self.group if self.group is not None else dist.group.WORLD
        dev = eng.device()
        W, P, nb = self.W, self.P, max(self.nb, 1)
        total = W * P + W + W * nb
        arena = symm.empty(total, dtype=torch.float64, device=dev)
        arena.zero_()
        flags = symm.empty(NB_MAX_PEERS, dtype=torch.int64, device=dev)
        flags.zero_()
        h_arena = symm.rendezvous(arena, self._group)
        h_flags = symm.rendezvous(flags, self._group)
        self._sym = [arena, flags, h_arena, h_flags]
        o = 0
        self.coords = arena[o:o + W * P].view(W, P); o += W * P
        self.lp = arena[o:o + W]; o += W
        self.blobs = arena[o:o + W * nb].view(W, nb); o += W * nb
        # self.n_acc stays a local tensor: each rank counts the acceptances it decided
        self.gen = eng.zeros(1, dtype=torch.int64)
        self.ticket = eng.zeros(1, dtype=torch.int32)
        pr = nb_peers()
        pr.world, pr.rank, pr.i0, pr.ld = self.world, self.rank, self.rank * self.per, 0
        pr.arena_local[0] = arena.data_ptr()
        pr.arena_bytes[0] = 8 * total
        mc = int(getattr(h_arena, "multicast_ptr", 0) or 0) if self.multicast else 0
        pr.arena_mc[0] = mc or None
        mcf = int(getattr(h_flags, "multicast_ptr", 0) or 0) if self.multicast else 0
        pr.mc_flags = mcf or None
        for r in range(self.world):
            pr.arena_peer[0][r] = int(h_arena.buffer_ptrs[r])
            pr.flags[r] = int(h_flags.buffer_ptrs[r])
        pr.gen, pr.ticket = self.gen.data_ptr(), self.ticket.data_ptr()
        self.peers_fused = pr
        self._flags_local = flags
        self.uses_multicast = bool(mc)
        torch.cuda.synchronize()
        dist.barrier(group=self.group)

    def _alloc_chain(self, n):
        if getattr(self, "transport", "") != "fused" and not hasattr(self, "peers_fused"):
            return super()._alloc_chain(n)
        import torch
        import torch.distributed._symmetric_memory as symm

        from . import engine as eng

        W, P, nb = self.W, self.P, self.nb
        total = n * W * (P + 1 + nb)
        arena = symm.empty(total, dtype=torch.float64, device=eng.device())
        arena.zero_()
        h = symm.rendezvous(arena, self._group)
        self._sym_chain = (arena, h)
        pr = self.peers_fused
        pr.arena_local[1] = arena.data_ptr()
        pr.arena_bytes[1] = 8 * total
        mc = int(getattr(h, "multicast_ptr", 0) or 0) if self.multicast else 0
        pr.arena_mc[1] = mc or None
        for r in range(self.world):
            pr.arena_peer[1][r] = int(h.buffer_ptrs[r])
        o = n * W * P
        chain = arena[:o].view(n, W, P)
        chain_lp = arena[o:o + n * W].view(n, W)
        chain_blobs = arena[o + n * W:].view(n, W, nb) if nb else None
        self._sync_ranks()
        return chain, chain_lp, chain_blobs

    def _sync_ranks(self):
        import torch

        if self.world > 1:
            torch.cuda.synchronize()
            _dist().barrier(group=self.group)

    def _wait_pushes(self):
        if getattr(self, "transport", "") == "fused" and hasattr(self, "s_idx"):
            from . import engine as eng
            from ._lib import check, lib

            check(lib().nb_peer_wait(ctypes.byref(self._stretch(0)), eng.stream()),
                  "nb_peer_wait")

    @property
    def acceptance_counts(self):
        if getattr(self, "transport", "") != "fused":
            return self.n_acc.cpu().numpy()
        tot = self.n_acc.clone()  # per-rank counts of the proposals this rank decided
        _dist().all_reduce(tot, group=self.group)
        return tot.cpu().numpy()

    def _stretch(self, split):
        mv = super()._stretch(split)
        if getattr(self, "transport", "") == "fused":
            mv.i0 = self.rank * self.per
            mv.wait_flags = self._flags_local.data_ptr()
            mv.wait_gen = self.gen.data_ptr()
            mv.wait_world = self.world
        return mv

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
        self._sync_ranks()  # no rank is still stepping (and pushing into our copy)
        super().set_state(coords, log_prob, rows)
        self._sync_ranks()  # every copy is in place before anybody steps

    def _enqueue_step(self):
        from . import engine as eng
        from ._lib import check, lib

        L = lib()
        for split in range(2):
            if self.transport == "fused":
                # same launches as the single-GPU step: the accept step inside the combine
                # kernel writes every rank's copy of the state and raises the flags
                self.plan._enqueue(self.ex, mv=self._stretch(split), fuse_update=True,
                                   peers=self.peers_fused)
                continue
            mv = self._stretch(split)
            mv.i0, mv.pars_ld = self.rank * self.per, self.ld
            self.plan._enqueue(self.ex, mv=mv, fuse_update=False)
            if self.world > 1:
                _dist().all_gather_into_tensor(self.pack_full, self.pack_local, group=self.group)
                self.collectives += 1
            check(L.nb_stretch_update_packed(ctypes.byref(self._stretch(split)),
                                             eng.ptr(self.pack_full), self.ld, eng.stream()),
                  "nb_stretch_update_packed")
