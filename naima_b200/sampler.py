# -*- coding: utf-8 -*-
"""Affine-invariant ensemble sampler with the slice of the ``emcee`` (>=3.0) API
that naima drives (core.py:127-160, 446-491, 523-536).

emcee is a third-party dependency of the reference and is not installed in this
image; this is a restatement of its published algorithm (Goodman & Weare 2010
stretch move with the red-blue split of Foreman-Mackey et al. 2013) written from
the call sites in naima and memory of emcee 3.1's behaviour -- parity with emcee
itself is UNPINNED (the reference's tests hold no numeric fixture for it).  The
draw order (shuffle of the split, then per half: zz, partner index, accept
uniforms) follows emcee's so that a seeded ``numpy.random.RandomState`` produces
the same kind of stream.

Two execution modes:
  * host-driven (this class): proposals/accepts in NumPy on the host, the
    log-probability of each half-ensemble evaluated in ONE batched device call
    (``vectorize=True`` semantics; the reference maps one lnprob per walker over
    a multiprocessing.Pool);
  * device-resident (:class:`DeviceEnsemble`): coordinates, log-probabilities and
    blobs stay in HBM, pre-drawn random numbers are resident, and a whole step
    (propose -> likelihood plan -> accept, twice) is one CUDA-graph replay.
"""
import numpy as np

__all__ = ["EnsembleSampler", "State", "DeviceEnsemble", "PlanSampler"]


class State:
    """emcee.State: coords [W,P], log_prob [W], blobs, random_state."""

    # PlanSampler extras: the walkers' device blob records and a token saying that the
    # device-resident ensemble still holds exactly this state (continuation without
    # re-evaluating or re-uploading it)
    _rows = None
    _token = None

    def __init__(self, coords, log_prob=None, blobs=None, random_state=None, copy=False):
        if isinstance(coords, State):
            log_prob, blobs, random_state = coords.log_prob, coords.blobs, coords.random_state
            self._rows, self._token = coords._rows, coords._token
            coords = coords.coords
        dc = (lambda x: np.array(x, copy=True)) if copy else (lambda x: x)
        self.coords = dc(np.atleast_2d(np.asarray(coords, dtype=float)))
        self.log_prob = None if log_prob is None else dc(np.asarray(log_prob, dtype=float))
        self.blobs = blobs
        self.random_state = random_state

    @classmethod
    def _make(cls, coords, log_prob, blobs, random_state, rows=None, token=None):
        """Constructor without the input normalisation (the sampling loop's own arrays)."""
        st = object.__new__(cls)
        st.coords, st.log_prob, st.blobs, st.random_state = coords, log_prob, blobs, random_state
        st._rows, st._token = rows, token
        return st

    def __iter__(self):
        # emcee allows `pos, lnp, rstate[, blobs] = state`
        if self.blobs is None:
            return iter((self.coords, self.log_prob, self.random_state))
        return iter((self.coords, self.log_prob, self.random_state, self.blobs))

    def __len__(self):
        return 3 if self.blobs is None else 4

    def __getitem__(self, i):
        return list(iter(self))[i]


def walkers_independent(coords):
    """emcee.ensemble.walkers_independent: condition-number test of the ensemble."""
    if not np.all(np.isfinite(coords)):
        return False
    C = coords - np.mean(coords, axis=0)[None, :]
    C_colmax = np.amax(np.abs(C), axis=0)
    if np.any(C_colmax == 0):
        return False
    C = C / C_colmax
    C_colsum = np.sqrt(np.sum(C**2, axis=0))
    C = C / C_colsum
    return np.linalg.cond(C.astype(float)) <= 1e8


class EnsembleSampler:
    """Stretch-move ensemble sampler.

    log_prob_fn(p, *args, **kwargs) is called per walker (``vectorize=False``,
    returns ``lnp`` or ``(lnp, *blobs)``) or once per half-ensemble with
    ``p[Ns, P]`` (``vectorize=True``, returns ``lnp[Ns]`` or ``(lnp[Ns], blobs)``
    where blobs is a sequence of Ns per-walker blob tuples or a BlobBatch)."""

    def __init__(self, nwalkers, ndim, log_prob_fn, args=None, kwargs=None, pool=None, a=2.0,
                 vectorize=False, blobs_dtype=None, seed=None, live_dangerously=False,
                 moves=None, backend=None):
        self.nwalkers, self.ndim = int(nwalkers), int(ndim)
        self.log_prob_fn = log_prob_fn
        self.args = [] if args is None else list(args)
        self.kwargs = {} if kwargs is None else dict(kwargs)
        self.pool = pool
        self.a = float(a)
        self.vectorize = vectorize
        self.blobs_dtype = blobs_dtype
        self.live_dangerously = live_dangerously
        self._random = np.random.mtrand.RandomState(seed)
        self._previous_state = None
        self.reset()

    # -- storage ---------------------------------------------------------------------
    def reset(self):
        self.iteration = 0
        self._accepted = np.zeros(self.nwalkers)
        self._chain = np.empty((0, self.nwalkers, self.ndim))
        self._log_prob = np.empty((0, self.nwalkers))
        self._blobs = []  # one entry per stored step: list of per-walker blobs | BlobBatch

    @property
    def random_state(self):
        return self._random.get_state()

    @random_state.setter
    def random_state(self, state):
        try:
            self._random.set_state(state)
        except Exception:
            pass

    def _grow(self, n):
        self._chain = np.concatenate([self._chain, np.empty((n, self.nwalkers, self.ndim))])
        self._log_prob = np.concatenate([self._log_prob, np.empty((n, self.nwalkers))])

    def get_chain(self, flat=False, thin=1, discard=0):
        v = self._chain[discard + thin - 1:self.iteration:thin]
        return v.reshape((-1, self.ndim)) if flat else v

    def get_log_prob(self, flat=False, thin=1, discard=0):
        v = self._log_prob[discard + thin - 1:self.iteration:thin]
        return v.reshape(-1) if flat else v

    def get_blobs(self, flat=False, thin=1, discard=0):
        if not self._blobs:
            return None
        steps = self._blobs[discard + thin - 1:self.iteration:thin]
        out = np.empty((len(steps), self.nwalkers), dtype=object)
        for i, st in enumerate(steps):
            for w in range(self.nwalkers):
                out[i, w] = st[w]
        return out.reshape(-1) if flat else out

    def get_last_sample(self):
        return self._previous_state

    @property
    def acceptance_fraction(self):
        return self._accepted / float(max(self.iteration, 1))

    # legacy emcee-2 style accessors used around naima
    @property
    def chain(self):
        return np.swapaxes(self.get_chain(), 0, 1)

    @property
    def flatchain(self):
        return self.get_chain(flat=True)

    @property
    def lnprobability(self):
        return self.get_log_prob().T

    @property
    def blobs(self):
        return self.get_blobs()

    # -- log-probability -------------------------------------------------------------
    def compute_log_prob(self, coords):
        p = np.asarray(coords, dtype=float)
        if np.any(np.isinf(p)):
            raise ValueError("At least one parameter value was infinite")
        if np.any(np.isnan(p)):
            raise ValueError("At least one parameter value was NaN")
        if self.vectorize:
            res = self.log_prob_fn(p, *self.args, **self.kwargs)
            if isinstance(res, tuple):
                log_prob, blobs = np.asarray(res[0], dtype=float), res[1]
            else:
                log_prob, blobs = np.asarray(res, dtype=float), None
        else:
            map_fn = self.pool.map if self.pool is not None else map
            results = list(map_fn(_Wrapper(self.log_prob_fn, self.args, self.kwargs), list(p)))
            try:
                log_prob = np.array([float(r[0]) for r in results])
                blobs = [tuple(r[1:]) for r in results]
            except (IndexError, TypeError):
                log_prob = np.array([float(r) for r in results])
                blobs = None
        if np.any(np.isnan(log_prob)):
            raise ValueError("Probability function returned NaN")
        return log_prob, blobs

    # -- sampling --------------------------------------------------------------------
    def _propose_half(self, state, inds, split):
        """One half of the red-blue stretch move; returns the updated state and the
        acceptance mask of the active walkers."""
        S1 = inds == split
        s = state.coords[S1]
        c = state.coords[~S1]
        Ns, Nc = len(s), len(c)
        zz = ((self.a - 1.0) * self._random.rand(Ns) + 1) ** 2.0 / self.a
        factors = (self.ndim - 1.0) * np.log(zz)
        rint = self._random.randint(Nc, size=(Ns,))
        q = c[rint] - (c[rint] - s) * zz[:, None]
        new_log_prob, new_blobs = self.compute_log_prob(q)
        lnpdiff = factors + new_log_prob - state.log_prob[S1]
        accepted = lnpdiff > np.log(self._random.rand(Ns))
        idx = np.flatnonzero(S1)[accepted]
        state.coords[idx] = q[accepted]
        state.log_prob[idx] = new_log_prob[accepted]
        if new_blobs is not None:
            if state.blobs is None:
                raise ValueError("log_prob_fn returned blobs only for some calls")
            acc_i = np.flatnonzero(accepted)
            for j, i in zip(idx, acc_i):
                state.blobs[j] = new_blobs[i]
        return idx

    def sample(self, initial_state, log_prob0=None, rstate0=None, blobs0=None, iterations=1,
               tune=False, skip_initial_state_check=False, thin_by=1, thin=None, store=True,
               progress=False):
        state = State(initial_state, copy=True)
        if np.shape(state.coords) != (self.nwalkers, self.ndim):
            raise ValueError("incompatible input dimensions {0}".format(np.shape(state.coords)))
        if not skip_initial_state_check and not walkers_independent(state.coords):
            raise ValueError("Initial state has a large condition number. Make sure that your "
                             "walkers are linearly independent for the best performance")
        if self.nwalkers < 2 * self.ndim and not self.live_dangerously:
            raise ValueError("It is unadvisable to use a red-blue move with fewer walkers than "
                             "twice the number of dimensions.")
        if rstate0 is not None:
            self.random_state = rstate0
        if log_prob0 is not None:
            state.log_prob = np.asarray(log_prob0, dtype=float)
        if blobs0 is not None:
            state.blobs = blobs0
        if state.log_prob is None:
            state.log_prob, blobs = self.compute_log_prob(state.coords)
            state.blobs = None if blobs is None else [blobs[w] for w in range(self.nwalkers)]
        if np.shape(state.log_prob) != (self.nwalkers,):
            raise ValueError("incompatible input dimensions")
        if np.any(np.isnan(state.log_prob)):
            raise ValueError("The initial log_prob was NaN")
        if store:
            self._grow(int(iterations))
        all_inds = np.arange(self.nwalkers)
        for _ in range(int(iterations)):
            # emcee draws the move with random.choice(moves, p=weights) every iteration,
            # which consumes exactly one uniform of the stream
            self._random.random_sample()
            inds = all_inds % 2
            self._random.shuffle(inds)
            for split in range(2):
                idx = self._propose_half(state, inds, split)
                self._accepted[idx] += 1
            state.random_state = self.random_state
            if store:
                self._chain[self.iteration] = state.coords
                self._log_prob[self.iteration] = state.log_prob
                if state.blobs is not None:
                    self._blobs.append(list(state.blobs))
            self.iteration += 1
            self._previous_state = state
            yield State(state, copy=True)

    def run_mcmc(self, initial_state, nsteps, **kwargs):
        if initial_state is None:
            if self._previous_state is None:
                raise ValueError("Cannot have `initial_state=None` if run_mcmc has never been "
                                 "called.")
            initial_state = self._previous_state
        results = None
        for results in self.sample(initial_state, iterations=nsteps, **kwargs):
            pass
        self._previous_state = results
        return results


class _Wrapper:
    def __init__(self, f, args, kwargs):
        self.f, self.args, self.kwargs = f, args, kwargs

    def __call__(self, x):
        return self.f(x, *self.args, **self.kwargs)


class BlobBatch:
    """Per-walker blobs of a batched evaluation, materialised lazily: indexing
    with a walker index yields the reference's blob tuple for that walker."""

    def __init__(self, plan, flux=None, blob_arrays=None, rows=None):
        """Either (flux, blob_arrays) or `rows`: per-walker records [W][plan.row_width]
        that are split on first access."""
        self.plan, self._flux, self._arrays, self._rows = plan, flux, blob_arrays, rows

    def _split(self):
        if self._flux is None:
            if isinstance(self._rows, _DeviceRows):
                self._rows = self._rows.fetch()
            self._flux, self._arrays = self.plan.split_rows(self._rows)

    @property
    def flux(self):
        self._split()
        return self._flux

    @property
    def blob_arrays(self):
        self._split()
        return self._arrays

    def __len__(self):
        return (self._rows if self._flux is None else self._flux).shape[0]

    def __getitem__(self, w):
        return _LazyBlob(self, w)


class _DeviceRows:
    """Blob records [W][row_width] of one stored step, still on the device."""

    def __init__(self, tensor):
        self.tensor = tensor
        self.shape = tuple(tensor.shape)

    def fetch(self):
        return self.tensor.cpu().numpy()


class _LazyBlob:
    """Blob tuple of one walker, built on first access (keeps the sampling loop
    free of per-walker Python object construction)."""

    __slots__ = ("batch", "w", "_t")

    def __init__(self, batch, w):
        self.batch, self.w, self._t = batch, w, None

    def _get(self):
        if self._t is None:
            b = self.batch
            self._t = b.plan.blobs_for(b.flux, b.blob_arrays, self.w)
        return self._t

    def __getitem__(self, i):
        return self._get()[i]

    def __len__(self):
        return len(self._get())

    def __iter__(self):
        return iter(self._get())


class DeviceEnsemble:
    """Device-resident stretch-move loop over a LikelihoodPlan.

    Positions, log-probabilities and model-flux blobs live in HBM; the random
    draws of a block of steps are generated with the same NumPy stream as the
    host sampler and uploaded once per block; each ensemble step is one CUDA
    graph replay (2 x [propose, plan, accept] + store).  The chain is read back
    once at the end of the block.
    """

    use_host_lib = True  # C helper for the random draws (NumPy calls when it is absent)

    def __init__(self, plan, nwalkers, a=2.0, seed=None, store_blobs=True, use_graph=True,
                 timeline=False):
        import torch

        from . import engine as eng

        if nwalkers % 2:
            raise ValueError("DeviceEnsemble needs an even number of walkers")
        if nwalkers < 2 * plan.P:
            raise ValueError("It is unadvisable to use a red-blue move with fewer walkers than "
                             "twice the number of dimensions.")
        self.plan, self.W, self.P, self.a = plan, int(nwalkers), plan.P, float(a)
        self.Ns = self.W // 2
        self._all_inds = np.arange(self.W)
        self.nb = plan.row_width if store_blobs else 0  # [model flux | further blobs]
        self._random = np.random.mtrand.RandomState(seed)
        self.ex = plan.executable(self.Ns)
        self.coords = eng.zeros(self.W, self.P)
        self.lp = eng.zeros(self.W)
        self.blobs = eng.zeros(self.W, max(self.nb, 1))
        self.n_acc = eng.zeros(self.W, dtype=torch.int32)
        self.step = eng.zeros(1, dtype=torch.int32)
        self.sync = eng.zeros(1, dtype=torch.int32)
        self._timeline = None
        if timeline:  # diagnostic: device time stamps per half-step (nb_stretch.timeline)
            from ._lib import NB_TIMELINE_CAP, NB_TIMELINE_COLS
            self._timeline = eng.zeros(NB_TIMELINE_CAP, NB_TIMELINE_COLS, dtype=torch.int64)
            import weakref

            from ._lib import lib

            def _forget(L=lib()):  # before the buffer goes: no kernel may stamp into it
                try:
                    torch.cuda.synchronize()
                    L.nb_timeline_reset()
                except Exception:  # interpreter / CUDA context shutting down
                    pass

            weakref.finalize(self, _forget)
        self.before_step = None  # host hook before every graph replay (bench.py's L2 flush)
        self.read_rows = True
        self.min_block = 0
        self.use_graph = use_graph
        self.steps_per_graph = 1  # steps captured into one graph (set before the first run)
        self._graph = None
        self._block = 0
        self.kernel_launches_per_step = 2 * plan.launches_per_eval

    def set_state(self, coords, log_prob=None, rows=None):
        """Upload positions; their log-probabilities and blob records are evaluated in
        one batched call unless given."""
        from . import engine as eng

        import torch

        coords = np.ascontiguousarray(coords, dtype=float)
        if log_prob is None or (self.nb and rows is None):
            log_prob, rows = self.plan.eval_rows(coords)
        if np.any(np.isnan(log_prob)):
            raise ValueError("The initial log_prob was NaN")
        self.coords.copy_(eng.to_dev(coords))
        self.lp.copy_(eng.to_dev(np.ascontiguousarray(log_prob, dtype=float)))
        if self.nb:  # records: a host array or a tensor that already lives on the device
            self.blobs.copy_(rows if isinstance(rows, torch.Tensor)
                             else eng.to_dev(np.ascontiguousarray(rows, dtype=float)))
        self.n_acc.zero_()

    def _draw_step(self, s_idx, c_idx, zz, lnu):
        """Draws of one ensemble step into [2][Ns] views, in emcee's order: the move
        choice (one uniform), the shuffle of the red-blue split, then per half zz, the
        partner index and the accept uniform."""
        rs, Ns = self._random, self.Ns
        rs.random_sample()  # random.choice(moves, p=weights)
        inds = self._all_inds % 2
        rs.shuffle(inds)
        half = (np.flatnonzero(inds == 0), np.flatnonzero(inds == 1))
        for split in range(2):
            s_idx[split] = half[split]
            t = (self.a - 1.0) * rs.rand(Ns) + 1
            zz[split] = t * t / self.a  # == t ** 2.0 / a (numpy squares for exponent 2)
            c_idx[split] = half[1 - split][rs.randint(Ns, size=(Ns,))]
            lnu[split] = np.log(rs.rand(Ns))

    def _draw_into(self, s_idx, c_idx, zz, lnu, state=None):
        """Draws of len(zz) consecutive ensemble steps into [n][2][Ns] arrays (C-contiguous;
        lnu receives log(accept uniform)).  Uses the C helper (nb_host_draw_steps: the same
        MT19937 stream and algorithms as numpy.random.RandomState, bit for bit) when it is
        built, else the NumPy calls of _draw_step.  `state`: the generator's current
        get_state() if the caller already has it (saves one 85 us call)."""
        import ctypes

        from ._lib import host_lib

        n = len(zz)
        H = host_lib() if self.use_host_lib else None
        ok = (H is not None and n > 0 and self.W % 2 == 0
              and all(x.flags.c_contiguous for x in (s_idx, c_idx, zz, lnu))
              and s_idx.dtype == np.int32 and c_idx.dtype == np.int32)
        if ok:
            st = state if state is not None else self._random.get_state()
            if st[0] == "MT19937":
                key = np.array(st[1], dtype=np.uint32)  # private copy, advanced in place
                pos = ctypes.c_int(int(st[2]))
                rc = H.nb_host_draw_steps(key.ctypes.data, ctypes.byref(pos), self.W, n,
                                          float(self.a), s_idx.ctypes.data, c_idx.ctypes.data,
                                          zz.ctypes.data, lnu.ctypes.data)
                if rc == 0:
                    self._random.set_state((st[0], key, pos.value, st[3], st[4]))
                    np.log(lnu, out=lnu)
                    return
        for t in range(n):
            self._draw_step(s_idx[t], c_idx[t], zz[t], lnu[t])

    def _draw_block(self, n):
        Ns = self.Ns
        s_idx = np.empty((n, 2, Ns), dtype=np.int32)
        c_idx = np.empty((n, 2, Ns), dtype=np.int32)
        zz = np.empty((n, 2, Ns))
        lnu = np.empty((n, 2, Ns))
        self._draw_into(s_idx, c_idx, zz, lnu)
        return s_idx, c_idx, zz, lnu

    def _alloc_block(self, n):
        import torch

        from . import engine as eng

        if self._block >= n:
            return
        self._block = n
        W, Ns = self.W, self.Ns
        self.s_idx = eng.zeros(n, 2, Ns, dtype=torch.int32)
        self.c_idx = eng.zeros(n, 2, Ns, dtype=torch.int32)
        self.zz = eng.zeros(n, 2, Ns)
        self.lnu = eng.zeros(n, 2, Ns)
        self.chain, self.chain_lp, self.chain_blobs = self._alloc_chain(n)
        self._graph = None

    def _alloc_chain(self, n):
        """Chain buffers for n steps (overridden where they live in symmetric memory)."""
        from . import engine as eng

        return (eng.zeros(n, self.W, self.P), eng.zeros(n, self.W),
                eng.zeros(n, self.W, self.nb) if self.nb else None)

    def _sync_ranks(self):
        """Rendezvous of all ranks that share this ensemble's state (single GPU: no-op)."""

    def _wait_pushes(self):
        """Enqueue a wait for the peers' writes into this rank's copy of the state and the
        chain (replicated-state sharding only; single GPU: no-op)."""

    def _stretch(self, split):
        """nb_stretch descriptor of the active half `split` over the current buffers."""
        from ._lib import nb_stretch

        mv = nb_stretch()
        for name, t in (("coords", self.coords), ("lp", self.lp),
                        ("blobs", self.blobs if self.nb else None), ("step", self.step),
                        ("sync", self.sync), ("s_idx", self.s_idx), ("c_idx", self.c_idx),
                        ("zz", self.zz), ("lnu", self.lnu), ("n_accepted", self.n_acc),
                        ("chain", self.chain), ("chain_lp", self.chain_lp),
                        ("chain_blobs", self.chain_blobs if self.nb else None)):
            setattr(mv, name, t.data_ptr() if t is not None else None)
        mv.nb, mv.W, mv.P, mv.Ns, mv.split = self.nb, self.W, self.P, self.Ns, split
        if self._timeline is not None:
            mv.timeline = self._timeline.data_ptr()
        return mv

    def timeline(self):
        """Diagnostic (timeline=True): %globaltimer stamps of this rank, ns, one row per
        half-step (row (2 t + split) mod NB_TIMELINE_CAP): first kernel entered, its wait for
        the peers over, accept kernel entered, last CTA before its release (sharded runs),
        accept kernel done."""
        if self._timeline is None:
            raise RuntimeError("construct the ensemble with timeline=True")
        return self._timeline.cpu().numpy()

    def _enqueue_step(self):
        """One ensemble step: per half, set-up (+ proposal) -> components -> combine
        (+ accept + chain append)."""
        for split in range(2):
            self.plan._enqueue(self.ex, mv=self._stretch(split))

    def load_draws(self, nsteps):
        """Draw and upload the random numbers of the next `nsteps` steps."""
        import torch

        from . import engine as eng

        self._alloc_block(nsteps)
        s_idx, c_idx, zz, lnu = self._draw_block(nsteps)
        self.s_idx[:nsteps].copy_(eng.to_dev(s_idx, dtype=torch.int32))
        self.c_idx[:nsteps].copy_(eng.to_dev(c_idx, dtype=torch.int32))
        self.zz[:nsteps].copy_(eng.to_dev(zz))
        self.lnu[:nsteps].copy_(eng.to_dev(lnu))
        self._sync_ranks()  # no peer is still writing rows of the previous block
        self.step.zero_()

    def run_loaded(self, nsteps):
        """Run `nsteps` steps on the loaded draws (asynchronous; no host sync)."""
        import torch

        if self.use_graph and self._graph is None:
            # warm-up outside capture, then rewind the step counter and state
            c0, l0, b0, a0, s0 = (self.coords.clone(), self.lp.clone(), self.blobs.clone(),
                                  self.n_acc.clone(), self.step.clone())
            self._enqueue_step()
            torch.cuda.synchronize()
            self._sync_ranks()  # nobody is still pushing into the state restored below
            from . import engine as eng

            with eng.capture_graph() as g:
                for _ in range(self.steps_per_graph):
                    self._enqueue_step()
            self._graph = g
            self.coords.copy_(c0)
            self.lp.copy_(l0)
            self.blobs.copy_(b0)
            self.n_acc.copy_(a0)
            self.step.copy_(s0)
            torch.cuda.synchronize()
            self._sync_ranks()  # every copy is restored before anybody steps again
        k = 0
        while k < nsteps:
            if self.before_step is not None:
                self.before_step()  # measurement hook (bench.py flushes L2 here)
            if self._graph is not None and nsteps - k >= self.steps_per_graph:
                self._graph.replay()
                k += self.steps_per_graph
            else:
                self._enqueue_step()
                k += 1

    # -- pipelined execution for the host-facing sampler ------------------------------
    def begin_chunk(self, n):
        """Prepare buffers (device + pinned host mirrors) for a chunk of up to n steps and
        rewind the step counter.  The caller must have consumed the previous chunk."""
        import torch

        n = max(n, self.min_block)  # full-size once: no reallocation / recapture later
        first = self._block < n or not hasattr(self, "_pin")
        self._alloc_block(n)
        if first:
            nn, W, Ns = self._block, self.W, self.Ns
            pin = lambda *shape, dtype=torch.float64: torch.empty(*shape, dtype=dtype).pin_memory()
            self._pin = dict(
                s_idx=pin(nn, 2, Ns, dtype=torch.int32), c_idx=pin(nn, 2, Ns, dtype=torch.int32),
                zz=pin(nn, 2, Ns), lnu=pin(nn, 2, Ns), chain=pin(nn, W, self.P),
                lp=pin(nn, W), rows=pin(nn, W, max(self.nb, 1)))
            self._pin_np = {k: v.numpy() for k, v in self._pin.items()}
        self._sync_ranks()  # every rank has read the previous chunk's rows back
        self.step.zero_()

    def enqueue_steps(self, t0, t1, dst=None, rows_dev=None):
        """Steps [t0, t1) of the current chunk: draw on the host, upload from pinned
        memory, replay, read the chain rows back into pinned memory -- all asynchronous.
        dst: optional (chain, lp, rows|None) pinned host tensors of t1 - t0 steps that
        receive the rows instead of this ensemble's staging buffers (the public sampler
        passes slices of its own storage: no host copy afterwards).  rows_dev: optional
        DEVICE tensor of t1 - t0 steps that receives the blob records instead of the host
        (they then stay in HBM until somebody asks for them).
        Returns (event, rng0) with the generator state before the block."""
        import torch

        n = t1 - t0
        pn, hp = self._pin, self._pin_np
        rng0 = self._random.get_state()  # generator state before the block
        self._draw_into(hp["s_idx"][t0:t1], hp["c_idx"][t0:t1], hp["zz"][t0:t1],
                        hp["lnu"][t0:t1], state=rng0)
        for name, dev in (("s_idx", self.s_idx), ("c_idx", self.c_idx), ("zz", self.zz),
                          ("lnu", self.lnu)):
            dev[t0:t1].copy_(pn[name][t0:t1], non_blocking=True)
        self.run_loaded(n)
        self._wait_pushes()  # replicated-state sharding: the peers' rows of these steps
        d_chain, d_lp, d_rows = dst if dst is not None else (
            pn["chain"][t0:t1], pn["lp"][t0:t1], pn["rows"][t0:t1])
        d_chain.copy_(self.chain[t0:t1], non_blocking=True)
        d_lp.copy_(self.chain_lp[t0:t1], non_blocking=True)
        if self.nb and rows_dev is not None:
            rows_dev.copy_(self.chain_blobs[t0:t1], non_blocking=True)
        elif self.nb and self.read_rows and d_rows is not None:
            d_rows.copy_(self.chain_blobs[t0:t1], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return ev, rng0

    def rng_state_after(self, rng0, k):
        """Generator state after k steps drawn from state rng0 (replays the draws)."""
        keep = self._random.get_state()
        self._random.set_state(rng0)
        self._draw_block(k)
        out = self._random.get_state()
        self._random.set_state(keep)
        return out

    def io_bytes_per_step(self, rows_to_host=None):
        """(host->device, device->host) bytes moved per ensemble step by enqueue_steps."""
        rows_to_host = self.read_rows if rows_to_host is None else rows_to_host
        h2d = 2 * self.Ns * (4 + 4 + 8 + 8)
        d2h = 8 * self.W * (self.P + 1 + (self.nb if rows_to_host else 0))
        return h2d, d2h

    def run(self, nsteps):
        """Draw, upload, run and read back `nsteps` steps: returns (chain, log_prob,
        records|None) as host arrays [nsteps, W, ...]; plan.split_rows(records) gives
        the model flux and the further blobs."""
        import torch

        self.load_draws(nsteps)
        self.run_loaded(nsteps)
        self._wait_pushes()
        torch.cuda.synchronize()
        chain = self.chain[:nsteps].cpu().numpy()
        lp = self.chain_lp[:nsteps].cpu().numpy()
        blobs = self.chain_blobs[:nsteps].cpu().numpy() if self.nb else None
        if np.any(np.isnan(lp)):
            raise ValueError("Probability function returned NaN")
        return chain, lp, blobs

    @property
    def acceptance_counts(self):
        return self.n_acc.cpu().numpy()


class PlanSampler(EnsembleSampler):
    """The EnsembleSampler API (what naima's run_sampler drives, core.py:127-160) over the
    device-resident stretch-move loop of a LikelihoodPlan.

    Per ensemble step the host draws the random numbers with the same NumPy stream and
    draw order as the host-driven sampler, uploads them from pinned memory, replays one
    CUDA graph (2 x [set-up + proposal -> components -> combine + accept + chain append])
    and reads the step's chain row, log-probabilities and blob records back.  Steps are
    enqueued `block` at a time so that the host's drawing overlaps the device's stepping;
    the chain is identical to the host-driven sampler's for the same seed."""

    def __init__(self, nwalkers, ndim, plan, a=2.0, seed=None, block=32, chunk=256,
                 blobs_dtype=None, group=None, sharded=None, transport="auto",
                 steps_per_graph=1, **kwargs):
        self.plan = plan
        self.block, self.chunk = int(block), int(chunk)
        # ensemble steps captured into one CUDA graph: a graph launch costs ~15 us of device
        # idle time on B200 whatever it contains; 4 steps per graph
        # measured 7 % faster end to end (0.118 vs 0.127 ms/step, warm L2); steps that do not
        # fill a graph are launched kernel by kernel
        self.steps_per_graph = max(1, int(steps_per_graph))
        self._de = None
        self.group = group
        self.transport = transport
        if sharded is None:
            import torch.distributed as dist

            sharded = dist.is_available() and dist.is_initialized() and \
                dist.get_world_size(group) > 1
        if sharded and seed is None:
            raise ValueError("a sharded PlanSampler needs an explicit seed shared by all ranks")
        self.sharded = bool(sharded)
        # sharded: only rank 0 reads the blob records back (every rank still keeps the
        # chain and the log-probabilities); get_blobs() is None on the other ranks
        self.read_rows = True
        if self.sharded:
            import torch.distributed as dist

            self.read_rows = dist.get_rank(group) == 0
        super().__init__(nwalkers, ndim, self._log_prob, a=a, vectorize=True,
                         blobs_dtype=blobs_dtype, seed=seed, **kwargs)

    def reset(self):
        super().reset()
        self._store_t = None   # pinned torch tensors behind _chain / _log_prob
        self._rows_dev = None  # [capacity][W][nb] blob records of the stored steps, in HBM
        self._rows_host = None  # (n, host copy of the first n steps' records), on demand

    def _grow(self, n):
        """Storage for n more steps.  Chain and log-probabilities: PINNED host memory that the
        device copies each block's rows straight into (capacity doubles, so that continuing a
        run rarely re-pins).  Blob records: a device buffer -- the per-walker model fluxes and
        blobs of every step stay in HBM and come to the host when get_blobs() / a State's
        blobs are looked at."""
        import torch

        if getattr(self, "_host_loop", False):
            return super()._grow(n)
        from . import engine as eng

        de = self._device()
        tot = self.iteration + n
        cap = 0 if self._store_t is None else self._store_t[0].shape[0]
        if tot > cap:
            # at least one chunk (256 steps) from the start: pinning host memory and a fresh
            # device allocation cost milliseconds -- on a bad day tens of them -- and a
            # burn-in followed by the run proper should not pay them a second time
            cap = max(tot, 2 * cap, self.chunk)
            new = (torch.empty(cap, self.nwalkers, self.ndim, dtype=torch.float64,
                               pin_memory=True),
                   torch.empty(cap, self.nwalkers, dtype=torch.float64, pin_memory=True))
            views = [t.numpy() for t in new]
            it = self.iteration
            if it:
                views[0][:it] = self._chain[:it]
                views[1][:it] = self._log_prob[:it]
            self._store_t = new
            self._chain_full, self._log_prob_full = views
            if de.nb and self.read_rows:
                rows = eng.empty(cap, self.nwalkers, de.nb)
                if it and self._rows_dev is not None:
                    rows[:it].copy_(self._rows_dev[:it])
                self._rows_dev = rows
        # the base class slices [discard:iteration] of these
        self._chain, self._log_prob = self._chain_full, self._log_prob_full

    def _rows_upto(self, n):
        """Host copy of the blob records of the first n stored steps (one D2H, cached)."""
        if self._rows_dev is None:
            return None
        if self._rows_host is None or self._rows_host[0] < n:
            import torch

            torch.cuda.current_stream().synchronize()
            self._rows_host = (n, self._rows_dev[:n].cpu().numpy())
        return self._rows_host[1]

    def get_blobs(self, flat=False, thin=1, discard=0):
        if getattr(self, "_host_loop", False):
            return super().get_blobs(flat=flat, thin=thin, discard=discard)
        rows = self._rows_upto(self.iteration)
        if rows is None:
            return None
        steps = range(discard + thin - 1, self.iteration, thin)
        out = np.empty((len(steps), self.nwalkers), dtype=object)
        for i, t in enumerate(steps):
            batch = BlobBatch(self.plan, rows=rows[t])
            for w in range(self.nwalkers):
                out[i, w] = batch[w]
        return out.reshape(-1) if flat else out

    def _log_prob(self, p):
        lnp, rows = self.plan.eval_rows(p)
        flux, arrays = self.plan.split_rows(rows)
        return lnp, BlobBatch(self.plan, flux, arrays)

    def _device(self):
        if self._de is None:
            if self.sharded:  # proposals of each half-ensemble sharded over the ranks
                from .parallel import ShardedDeviceEnsemble

                self._de = ShardedDeviceEnsemble(self.plan, self.nwalkers, a=self.a, seed=0,
                                                 group=self.group, transport=self.transport)
            else:
                self._de = DeviceEnsemble(self.plan, self.nwalkers, a=self.a, seed=0)
            self._de._random = self._random  # one stream, shared with the host-side API
            self._de.min_block = self.chunk
            self._de.read_rows = self.read_rows
            self._de.steps_per_graph = self.steps_per_graph
        return self._de

    def sample(self, initial_state, log_prob0=None, rstate0=None, blobs0=None, iterations=1,
               tune=False, skip_initial_state_check=False, thin_by=1, thin=None, store=True,
               progress=False):
        if self.nwalkers % 2 or self.nwalkers < 2 * self.ndim or self.ndim > 32:
            # odd ensembles / live_dangerously / more parameters than the fused proposal
            # kernel maps (NB_MAX_MOVE_PAR): the host-driven loop handles them
            self._host_loop = True
            yield from super().sample(initial_state, log_prob0=log_prob0, rstate0=rstate0,
                                      blobs0=blobs0, iterations=iterations,
                                      skip_initial_state_check=skip_initial_state_check,
                                      store=store)
            return
        state = State(initial_state, copy=True)
        if np.shape(state.coords) != (self.nwalkers, self.ndim):
            raise ValueError("incompatible input dimensions {0}".format(np.shape(state.coords)))
        if not skip_initial_state_check and not walkers_independent(state.coords):
            raise ValueError("Initial state has a large condition number. Make sure that your "
                             "walkers are linearly independent for the best performance")
        if np.any(~np.isfinite(state.coords)):
            raise ValueError("At least one parameter value was infinite or NaN")
        if rstate0 is not None:
            self.random_state = rstate0
        de = self._device()
        # continuation: the device-resident ensemble still holds exactly this state (the
        # last State of a completed sample() call on this sampler) -- nothing to evaluate
        # or upload
        token, self._dev_token = getattr(self, "_dev_token", None), None
        resident = (token is not None and state._token is token and log_prob0 is None
                    and blobs0 is None)
        if not resident:
            if log_prob0 is not None:
                state.log_prob = np.asarray(log_prob0, dtype=float)
            rows = state._rows
            if rows is None or state.log_prob is None or log_prob0 is not None:
                # the device loop needs the walkers' blob records: evaluate the ensemble once
                lnp, rows = self.plan.eval_rows(state.coords)
                if state.log_prob is None:
                    state.log_prob = lnp
            if np.shape(state.log_prob) != (self.nwalkers,):
                raise ValueError("incompatible input dimensions")
            if np.any(np.isnan(state.log_prob)):
                raise ValueError("The initial log_prob was NaN")
            de.set_state(state.coords, state.log_prob, rows)
        iterations = int(iterations)
        if store:
            self._grow(iterations)
        prev = state.coords
        done = 0
        make = State._make
        while done < iterations:
            nchunk = min(self.chunk, iterations - done)
            de.begin_chunk(nchunk)
            pending = []  # (t0, t1, event, rng state before the block), in flight
            t_enq = 0
            t_out = 0
            while t_out < nchunk:
                while t_enq < nchunk and len(pending) < 2:
                    t1 = min(t_enq + self.block, nchunk)
                    dst = rows_dev = None
                    if store:  # the device writes into the sampler's (pinned) storage
                        i0 = self.iteration + (t_enq - t_out)
                        st = self._store_t
                        dst = (st[0][i0:i0 + t1 - t_enq], st[1][i0:i0 + t1 - t_enq], None)
                        if self._rows_dev is not None:
                            rows_dev = self._rows_dev[i0:i0 + t1 - t_enq]
                    ev, rng0 = de.enqueue_steps(t_enq, t1, dst=dst, rows_dev=rows_dev)
                    pending.append((t_enq, t1, ev, rng0))
                    t_enq = t1
                t0, t1, ev, rng0 = pending.pop(0)
                ev.synchronize()
                nblk = t1 - t0
                if store:
                    it0 = self.iteration
                    chain, lps = self._chain[it0:it0 + nblk], self._log_prob[it0:it0 + nblk]
                    recs = None
                else:
                    hp = de._pin_np
                    chain, lps = hp["chain"][t0:t1].copy(), hp["lp"][t0:t1].copy()
                    recs = hp["rows"][t0:t1].copy() if (de.nb and self.read_rows) else None
                if np.isnan(lps).any():
                    raise ValueError("Probability function returned NaN")
                moved = np.empty((nblk, self.nwalkers), dtype=bool)
                np.any(prev[None] != chain[:1], axis=2, out=moved[:1])
                if nblk > 1:
                    np.any(chain[:-1] != chain[1:], axis=2, out=moved[1:])
                prev = chain[-1]
                for k in range(nblk):
                    self._accepted += moved[k]
                    blobs = rows_k = None
                    if recs is not None:
                        blobs, rows_k = BlobBatch(self.plan, rows=recs[k]), recs[k]
                    elif store and self._rows_dev is not None:
                        rows_k = self._rows_dev[self.iteration]  # device view, fetched lazily
                        blobs = BlobBatch(self.plan, rows=_DeviceRows(rows_k))
                    self.iteration += 1
                    last = done + t0 + k + 1 == iterations
                    if last:
                        self._dev_token = token = object()
                    out = make(chain[k].copy(), lps[k].copy(), blobs,
                               self._random.get_state() if last else None,
                               rows_k, token if last else None)
                    self._previous_state = out
                    try:
                        yield out
                    except GeneratorExit:
                        # the consumer stopped here: rewind the stream to this step (the
                        # device ensemble is ahead: it no longer holds this state)
                        self._dev_token = None
                        self._random.set_state(de.rng_state_after(rng0, k + 1))
                        raise
                t_out = t1
            done += nchunk
