# -*- coding: utf-8 -*-
"""Likelihood and sampler driver with the reference's API surface
(src/naima/core.py:34-538): priors, ``lnprobmodel``, ``lnprob``,
``get_sampler``, ``run_sampler``.

What changes underneath: the reference hands ``lnprob`` to emcee, which maps it
over a multiprocessing.Pool one walker at a time (core.py:446-457).  Here a whole
half-ensemble is one batched device evaluation:

  1. if the user's ``model``/``prior`` callbacks can be traced
     (naima_b200.fused), the likelihood is a CUDA-graph LikelihoodPlan;
  2. otherwise the callbacks are called once with ``pars`` of shape ``[P, W]``
     (every radiative class is batch-aware) and the Gaussian likelihood of the
     returned ``[W, N_E]`` model runs in the lnprob kernel;
  3. ``vectorize=False`` restores the reference's one-call-per-walker behaviour
     (each call still evaluates on the device).
"""
import logging
import os
import warnings
from collections.abc import Iterable

import numpy as np

from . import engine as eng
from . import units as u
from .fused import (PRIOR_LOGUNIFORM, PRIOR_NORMAL, PRIOR_UNIFORM, LikelihoodPlan, SymPrior,
                    TraceError, is_sym)
from .sampler import BlobBatch, EnsembleSampler, PlanSampler, State
from .units import Quantity
from .utils import sed_conversion, validate_data_table

__all__ = ["normal_prior", "uniform_prior", "log_uniform_prior", "get_sampler", "run_sampler",
           "lnprob", "lnprobmodel"]

log = logging.getLogger("naima_b200.core")


# ------------------------------------------------------------------------------
# priors (core.py:34-58); array-valued `value` = one entry per walker
# ------------------------------------------------------------------------------
def uniform_prior(value, umin, umax):
    """Uniform prior distribution."""
    if is_sym(value):
        return SymPrior.term(value, PRIOR_UNIFORM, umin, umax)
    if np.ndim(value) == 0:
        if umin <= value <= umax:
            return 0.0
        return -np.inf
    value = np.asarray(value, dtype=float)
    return np.where((umin <= value) & (value <= umax), 0.0, -np.inf)


def normal_prior(value, mean, sigma):
    """Normal prior distribution (the reference's literal formula)."""
    if is_sym(value):
        return SymPrior.term(value, PRIOR_NORMAL, mean, sigma)
    return -0.5 * (2 * np.pi * sigma) - (value - mean) ** 2 / (2.0 * sigma)


def log_uniform_prior(value, umin=0, umax=None):
    """Log-uniform prior distribution (returns 1/value inside the support, as the
    reference does)."""
    if is_sym(value):
        return SymPrior.term(value, PRIOR_LOGUNIFORM, umin, np.inf if umax is None else umax)
    if np.ndim(value) == 0:
        if value > 0 and value >= umin:
            if umax is not None:
                if value <= umax:
                    return 1 / value
                return -np.inf
            return 1 / value
        return -np.inf
    value = np.asarray(value, dtype=float)
    ok = (value > 0) & (value >= umin)
    if umax is not None:
        ok &= value <= umax
    with np.errstate(divide="ignore"):
        return np.where(ok, 1 / value, -np.inf)


# ------------------------------------------------------------------------------
# likelihood (core.py:64-121)
# ------------------------------------------------------------------------------
_DATA_CACHE = {}


def _device_data(data):
    key = id(data)
    ent = _DATA_CACHE.get(key)
    if ent is None or ent[0] is not data:
        fl = Quantity(data["flux"])
        dd = eng.DeviceData(fl.value, Quantity(data["flux_error_lo"]).to(fl.unit).value,
                            Quantity(data["flux_error_hi"]).to(fl.unit).value,
                            np.asarray(data["ul"], dtype=bool), np.asarray(data["cl"], dtype=float))
        if len(_DATA_CACHE) > 8:
            _DATA_CACHE.clear()
        _DATA_CACHE[key] = ent = (data, dd)
    return ent[1]


def lnprobmodel(model, data, prior=None):
    """Gaussian (asymmetric errors) + upper-limit log-likelihood of ``model``
    (Quantity ``[N_E]`` or ``[W, N_E]``) given ``data``; evaluated by the lnprob
    kernel.  Returns a float or an array of W values."""
    model = Quantity(model)
    d_unit = Quantity(data["flux"]).unit
    model_is_sed = model.unit.physical_type in ["power", "flux"]
    data_is_sed = d_unit.physical_type in ["power", "flux"]
    if model_is_sed != data_is_sed:
        unit, sed_factor = sed_conversion(data["energy"], model.unit, data_is_sed)
        model = (model * sed_factor).to(d_unit)
    else:
        model = model.to(d_unit)
    mv = np.atleast_2d(np.asarray(model.value, dtype=float))
    W, N_E = mv.shape
    dd = _device_data(data)
    if N_E != dd.N_E:
        raise ValueError("model and data have different numbers of points")
    src = eng.to_dev(mv)
    lnp = eng.empty(W)
    pr = None if prior is None else eng.to_dev(np.broadcast_to(np.asarray(prior, float), (W,)))
    eng.combine([(src, 0, True, 1.0, None)], W, N_E, eng.to_dev(np.ones(N_E)), data=dd,
                prior_d=pr, lnp_out=lnp)
    out = lnp.cpu().numpy()
    return out if np.ndim(model.value) == 2 else float(out[0])


def _split_modelout(modelout):
    if isinstance(modelout, Iterable) and not isinstance(modelout, (np.ndarray, Quantity)):
        return modelout[0], tuple(modelout)
    return modelout, (modelout, np.nan)


def lnprob(pars, data, modelfunc, priorfunc):
    """core.py:97-121.  ``pars`` is ``[P]`` (returns ``(lnp, *blobs)`` like the
    reference) or ``[W, P]`` (returns ``(lnp[W], blobs)`` with blobs a list of W
    per-walker tuples): the callbacks then see ``pars.T`` so that ``pars[k]`` is
    the vector of the k-th parameter over walkers."""
    pars = np.asarray(pars, dtype=float)
    if pars.ndim == 1:
        lnprob_priors = 0.0 if priorfunc is None else priorfunc(pars)
        modelout = modelfunc(pars, data)
        model, blob = _split_modelout(modelout)
        if not np.isinf(lnprob_priors):
            total_lnprob = lnprobmodel(model, data) + lnprob_priors
        else:
            total_lnprob = lnprob_priors
        return (total_lnprob, *blob)
    W = pars.shape[0]
    pt = np.ascontiguousarray(pars.T)
    pri = np.zeros(W) if priorfunc is None else np.broadcast_to(
        np.asarray(priorfunc(pt), dtype=float), (W,))
    modelout = modelfunc(pt, data)
    model, blob = _split_modelout(modelout)
    # the kernel returns the prior itself where it is infinite (core.py:115-119)
    total = lnprobmodel(model, data, prior=pri)
    blobs = [tuple(_take(b, w, W) for b in blob) for w in range(W)]
    return total, blobs


def _take(b, w, W):
    """Walker w's slice of a batched blob (arrays with a leading axis of W)."""
    if isinstance(b, tuple):
        return tuple(_take(x, w, W) for x in b)
    if isinstance(b, Quantity):
        return b[w] if b.ndim >= 1 and b.shape[0] == W else b
    if isinstance(b, np.ndarray) and b.ndim >= 1 and b.shape[0] == W:
        return b[w]
    return b


class PlanLogProb:
    """Vectorised log-probability backed by a LikelihoodPlan (picklable-free; the
    sampler calls it with ``q[Ns, P]``)."""

    def __init__(self, plan):
        self.plan = plan
        self.calls = 0

    def __call__(self, p):
        self.calls += 1
        lnp, flux, blob_arrays = self.plan(p)
        return lnp, BlobBatch(self.plan, flux, blob_arrays)


class BatchedLogProb:
    """Vectorised log-probability calling the user's callbacks with batched pars."""

    def __init__(self, data, model, prior):
        self.data, self.model, self.prior = data, model, prior

    def __call__(self, p):
        return lnprob(p, self.data, self.model, self.prior)


# ------------------------------------------------------------------------------
# sampler driver (core.py:127-538)
# ------------------------------------------------------------------------------
def _run_mcmc(sampler, pos, nrun):
    """core.py:127-160 incl. the 5 % progress report."""
    state = None
    for i, state in enumerate(sampler.sample(pos, iterations=nrun, store=True)):
        progress = 100.0 * float(i) / float(nrun)
        if progress % 5 < (5.0 / float(nrun)):
            print("\nProgress of the run: {0:.0f} percent ({1} of {2} steps)".format(
                int(progress), i, nrun))
            npars = sampler.get_chain().shape[-1]
            paravg = [np.median(state.coords[:, k]) for k in range(npars)]
            parstd = [np.std(state.coords[:, k]) for k in range(npars)]
            print("                           "
                  + (" ".join(["{%i:-^15}" % k for k in range(npars)])).format(*sampler.labels))
            print("  Last ensemble median : "
                  + (" ".join(["{%i:^15.3g}" % k for k in range(npars)])).format(*paravg))
            print("  Last ensemble std    : "
                  + (" ".join(["{%i:^15.3g}" % k for k in range(npars)])).format(*parstd))
            print("  Last ensemble lnprob :  avg: {0:.3f}, max: {1:.3f}".format(
                np.average(state.log_prob), np.max(state.log_prob)))
    return sampler, state


def _prefit(p0, data, model, prior):
    """core.py:163-217: Nelder-Mead maximum-likelihood prefit (flat prior)."""
    from .minimize import minimize

    P0_IS_ML = False

    def flat_prior(*args):
        return 0.0

    if prior is None:
        prior = flat_prior

    def nll(*args):
        return -lnprob(*args)[0]

    # the simplex's candidate points of an iteration go through the likelihood kernels in
    # one batch: the traced plan when the callbacks can be traced, the batched callbacks when
    # they are batch-aware, else one call per point as in the reference
    nll_batch = None
    try:
        plan = LikelihoodPlan(model, None, data, len(p0))
        nll_batch = lambda X: -plan(X, want_blobs=False)[0]  # noqa: E731
    except TraceError:
        if _batch_safe(p0, data, model, flat_prior):
            nll_batch = lambda X: -np.asarray(  # noqa: E731
                lnprob(X, data, model, flat_prior)[0], dtype=float)

    log.info("Finding Maximum Likelihood parameters through Nelder-Mead fitting...")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        result = minimize(nll, p0, args=(data, model, flat_prior), method="Nelder-Mead",
                          options={"maxfev": 500, "xtol": 1e-1, "ftol": 1e-3},
                          batch_func=nll_batch)
        ll_prior = lnprob(result["x"], data, model, prior)[0]
    if (result["success"] or result["status"] == 1) and not np.isinf(ll_prior):
        if result["status"] != 1:
            P0_IS_ML = True
        p0 = result["x"]
    elif np.isinf(ll_prior):
        log.warning("Maximum Likelihood procedure converged on a parameter vector forbidden "
                    "by prior, using original parameters for MCMC")
    else:
        log.warning("Maximum Likelihood procedure failed to converge, using original "
                    "parameters for MCMC")
    return p0, P0_IS_ML


def _batch_safe(p0, data, model, prior, n=3):
    """True if the callbacks give, for a [n, P] batch of parameter vectors (seen by them as
    ``pars[P, n]``), the same log-probabilities as n separate reference-style calls.  Valid
    reference callbacks that branch on ``pars``, interpolate with scalar parameters or
    broadcast ``pars[k]`` against ``data['energy']`` fail or mis-broadcast when batched; they
    keep the reference's one-call-per-walker semantics instead."""
    p0 = np.asarray(p0, dtype=float)
    P = p0 * (1.0 + 1e-3 * np.arange(1, n + 1)[:, None])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            single = np.array([float(lnprob(p, data, model, prior)[0]) for p in P])
        except Exception:
            return True  # broken either way: let the sampler surface the error as before
        try:
            batched = np.asarray(lnprob(P, data, model, prior)[0], dtype=float)
        except Exception:
            return False
    if batched.shape != (n,):
        return False
    return bool(np.allclose(batched, single, rtol=1e-8, atol=0.0, equal_nan=True))


def get_sampler(data_table=None, p0=None, model=None, prior=None, nwalkers=500, nburn=100,
                guess=True, interactive=False, prefit=False, labels=None, threads=None,
                data_sed=None, vectorize=True, fused=True, seed=None):
    """Generate a new MCMC sampler (core.py:220-493).

    Extra keyword arguments over the reference: ``vectorize`` (batched
    half-ensemble evaluation, default on), ``fused`` (trace the callbacks into a
    LikelihoodPlan when possible, default on) and ``seed``.  ``threads`` is
    accepted for compatibility and ignored (there is no process pool)."""
    if data_table is None:
        raise TypeError("Data table is missing!")
    data = validate_data_table(data_table, sed=data_sed)
    if model is None:
        raise TypeError("Model function is missing!")
    p0 = np.array(p0, dtype=float)
    if labels is None:
        labels = ["norm"] + ["par{0}".format(i) for i in range(1, len(p0))]
    elif len(labels) < len(p0):
        labels += ["par{0}".format(i) for i in range(len(labels), len(p0))]

    # Check that the model returns fluxes in same physical type as data
    modelout = model(p0, data)
    spec = modelout[0] if type(modelout) in (tuple, list) else modelout
    try:
        sed_conversion(data["energy"], spec.unit, False)
        sed_conversion(data["energy"], data["flux"].unit, False)
    except u.UnitsError:
        raise u.UnitsError(
            "The physical type of the model and data units are not compatible, please modify "
            "your model or data so they match:\n Model units: {0} [{1}]\n Data units: {2} "
            "[{3}]\n".format(spec.unit, spec.unit.physical_type, data["flux"].unit,
                             data["flux"].unit.physical_type))

    if guess:
        normNames = ["norm", "ampl", "we", "wp"]
        normNames += ["log({0}".format(n) for n in normNames[:4]] + \
                     ["log10({0}".format(n) for n in normNames[:4]]
        idxs = []
        for normName in normNames:
            for l2 in labels:
                if l2.lower().startswith(normName):
                    idxs.append(labels.index(l2))
        if len(idxs) == 1:
            E = Quantity(data["energy"])
            nunit, sedf = sed_conversion(E, spec.unit, False)
            currFlux = np.trapezoid((E * (spec * sedf).to(nunit)).value, E.value)
            nunit, sedf = sed_conversion(E, data["flux"].unit, False)
            dataFlux = np.trapezoid((E * (data["flux"] * sedf).to(nunit)).value, E.value)
            ratio = dataFlux / currFlux
            if labels[idxs[0]].startswith("log("):
                p0[idxs[0]] += np.log(ratio)
            elif labels[idxs[0]].startswith("log10("):
                p0[idxs[0]] += np.log10(ratio)
            else:
                p0[idxs[0]] *= ratio
        elif len(idxs) == 0:
            log.warning("No label starting with [{0}] found: not applying normalization guess."
                        .format(",".join(normNames)))
        else:
            log.warning("More than one label starting with [{0}] found: not applying "
                        "normalization guess.".format(",".join(normNames)))

    P0_IS_ML = False
    if interactive:
        log.warning("Interactive fitting is not available in naima_b200")
    if prefit and not P0_IS_ML:
        p0, P0_IS_ML = _prefit(p0, data, model, prior)

    plan = None
    if vectorize and fused:
        try:
            plan = LikelihoodPlan(model, prior, data, len(p0))
        except TraceError as e:
            log.info("model/prior callbacks are not traceable (%s); using batched callbacks", e)
    if plan is None and vectorize and not _batch_safe(p0, data, model, prior):
        log.info("model/prior callbacks are not batch-aware; calling them once per walker")
        vectorize = False
    if plan is not None:
        # device-resident stepping behind the EnsembleSampler API
        sampler = PlanSampler(nwalkers, len(p0), plan, blobs_dtype=np.dtype(object), seed=seed)
    elif vectorize:
        sampler = EnsembleSampler(nwalkers, len(p0), BatchedLogProb(data, model, prior),
                                  vectorize=True, blobs_dtype=np.dtype(object), seed=seed)
    else:
        sampler = EnsembleSampler(nwalkers, len(p0), lnprob, args=[data, model, prior],
                                  blobs_dtype=np.dtype(object), seed=seed)
    sampler._naima_pool = None
    sampler.plan = plan
    sampler.data_table = data_table
    sampler.data = data
    sampler.labels = labels
    sampler.modelfn = model
    sampler.run_info = {"n_walkers": nwalkers, "n_burn": nburn,
                        "p0": [float(p) for p in p0], "guess": guess}

    # ball of relative size 0.5% if the parameters were fit to their ML values, 10% otherwise
    spread = 0.005 if P0_IS_ML else 0.1
    p0var = np.array([spread * pp for pp in p0])
    rng = np.random if seed is None else np.random.RandomState(seed + 1)
    p0 = np.vstack([p0 + p0var * rng.normal(size=len(p0)) for i in range(nwalkers)])

    if nburn > 0:
        print("Burning in the {0} walkers with {1} steps...".format(nwalkers, nburn))
        sampler, state = _run_mcmc(sampler, p0, nburn)
    else:
        state = State(p0)
    sampler.run_info["p0_burn_median"] = [float(p) for p in np.median(state.coords, axis=0)]
    return sampler, state


def run_sampler(nrun=100, sampler=None, pos=None, **kwargs):
    """Run an MCMC sampler (core.py:496-538)."""
    if sampler is None or pos is None:
        sampler, pos = get_sampler(**kwargs)
    sampler.run_info["n_run"] = nrun
    print("\nWalker burn in finished, running {0} steps...".format(nrun))
    sampler.reset()
    sampler, pos = _run_mcmc(sampler, pos, nrun)
    if getattr(sampler, "_naima_pool", None) is not None:
        sampler._naima_pool.close()
        sampler._naima_pool.join()
        sampler._naima_pool = None
    return sampler, pos
