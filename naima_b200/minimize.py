# -*- coding: utf-8 -*-
"""Nelder-Mead simplex minimiser with *relative* x/f tolerances, as used by the
reference's prefit (core.py:181-187 calling extern/minimize.py, a variant of
SciPy's ``_minimize_neldermead``).  Written from the textbook algorithm
(Nelder & Mead 1965; reflection 1, expansion 2, contraction 0.5, shrink 0.5; the
usual 5 % initial simplex) -- host code, a caller of the hot path.
"""
import numpy as np

__all__ = ["minimize"]


def minimize(func, x0, args=(), method="Nelder-Mead", options=None, batch_func=None):
    """batch_func(X[k, n]) -> k values: when given, every simplex iteration evaluates its
    candidate points (reflection, expansion, both contractions) in ONE batched call and the
    initial simplex / a shrink in one call each; the decisions, the trajectory and the
    evaluation count `nfev` (what the serial algorithm would have spent: the reference stops
    at maxfev = 500) are those of the serial algorithm."""
    if method != "Nelder-Mead":
        raise ValueError("only the Nelder-Mead method is available")
    if batch_func is not None:
        return _minimize_batched(batch_func, x0, options)
    opt = {"xtol": 1e-4, "ftol": 1e-4, "maxiter": None, "maxfev": None}
    opt.update(options or {})
    x0 = np.asarray(x0, dtype=float).ravel()
    n = x0.size
    maxiter = opt["maxiter"] if opt["maxiter"] is not None else n * 200
    maxfev = opt["maxfev"] if opt["maxfev"] is not None else n * 200
    ncalls = [0]

    def f(x):
        ncalls[0] += 1
        return float(func(x, *args))

    sim = np.empty((n + 1, n))
    sim[0] = x0
    for k in range(n):
        y = x0.copy()
        y[k] = (1 + 0.05) * y[k] if y[k] != 0 else 0.00025
        sim[k + 1] = y
    fsim = np.array([f(s) for s in sim])
    order = np.argsort(fsim)
    sim, fsim = sim[order], fsim[order]
    it = 1
    while ncalls[0] < maxfev and it < maxiter:
        with np.errstate(all="ignore"):
            if (np.max(np.abs((sim[1:] - sim[0]) / sim[0])) <= opt["xtol"]
                    and np.max(np.abs((fsim[0] - fsim[1:]) / fsim[0])) <= opt["ftol"]):
                break
        xbar = sim[:-1].sum(axis=0) / n
        xr = 2 * xbar - sim[-1]
        fxr = f(xr)
        shrink = False
        if fxr < fsim[0]:
            xe = 3 * xbar - 2 * sim[-1]
            fxe = f(xe)
            if fxe < fxr:
                sim[-1], fsim[-1] = xe, fxe
            else:
                sim[-1], fsim[-1] = xr, fxr
        elif fxr < fsim[-2]:
            sim[-1], fsim[-1] = xr, fxr
        elif fxr < fsim[-1]:
            xc = 1.5 * xbar - 0.5 * sim[-1]
            fxc = f(xc)
            if fxc <= fxr:
                sim[-1], fsim[-1] = xc, fxc
            else:
                shrink = True
        else:
            xcc = 0.5 * xbar + 0.5 * sim[-1]
            fxcc = f(xcc)
            if fxcc < fsim[-1]:
                sim[-1], fsim[-1] = xcc, fxcc
            else:
                shrink = True
        if shrink:
            for j in range(1, n + 1):
                sim[j] = sim[0] + 0.5 * (sim[j] - sim[0])
                fsim[j] = f(sim[j])
        order = np.argsort(fsim)
        sim, fsim = sim[order], fsim[order]
        it += 1
    status = 1 if ncalls[0] >= maxfev else (2 if it >= maxiter else 0)
    return {"x": sim[0], "fun": fsim[0], "nit": it, "nfev": ncalls[0], "status": status,
            "success": status == 0,
            "message": ("Optimization terminated successfully.",
                        "Maximum number of function evaluations has been exceeded.",
                        "Maximum number of iterations has been exceeded.")[status]}


def _minimize_batched(batch_func, x0, options):
    opt = {"xtol": 1e-4, "ftol": 1e-4, "maxiter": None, "maxfev": None}
    opt.update(options or {})
    x0 = np.asarray(x0, dtype=float).ravel()
    n = x0.size
    maxiter = opt["maxiter"] if opt["maxiter"] is not None else n * 200
    maxfev = opt["maxfev"] if opt["maxfev"] is not None else n * 200

    def fb(X):
        return np.asarray(batch_func(np.ascontiguousarray(X, dtype=float)), dtype=float).ravel()

    sim = np.empty((n + 1, n))
    sim[0] = x0
    for k in range(n):
        y = x0.copy()
        y[k] = (1 + 0.05) * y[k] if y[k] != 0 else 0.00025
        sim[k + 1] = y
    fsim = fb(sim)
    ncalls, launches = n + 1, 1
    order = np.argsort(fsim)
    sim, fsim = sim[order], fsim[order]
    it = 1
    while ncalls < maxfev and it < maxiter:
        with np.errstate(all="ignore"):
            if (np.max(np.abs((sim[1:] - sim[0]) / sim[0])) <= opt["xtol"]
                    and np.max(np.abs((fsim[0] - fsim[1:]) / fsim[0])) <= opt["ftol"]):
                break
        xbar = sim[:-1].sum(axis=0) / n
        cand = np.array([2 * xbar - sim[-1],            # reflection
                         3 * xbar - 2 * sim[-1],        # expansion
                         1.5 * xbar - 0.5 * sim[-1],    # outside contraction
                         0.5 * xbar + 0.5 * sim[-1]])   # inside contraction
        fxr, fxe, fxc, fxcc = fb(cand)
        launches += 1
        ncalls += 1  # the reflection is always evaluated
        shrink = False
        if fxr < fsim[0]:
            ncalls += 1
            if fxe < fxr:
                sim[-1], fsim[-1] = cand[1], fxe
            else:
                sim[-1], fsim[-1] = cand[0], fxr
        elif fxr < fsim[-2]:
            sim[-1], fsim[-1] = cand[0], fxr
        elif fxr < fsim[-1]:
            ncalls += 1
            if fxc <= fxr:
                sim[-1], fsim[-1] = cand[2], fxc
            else:
                shrink = True
        else:
            ncalls += 1
            if fxcc < fsim[-1]:
                sim[-1], fsim[-1] = cand[3], fxcc
            else:
                shrink = True
        if shrink:
            sim[1:] = sim[0] + 0.5 * (sim[1:] - sim[0])
            fsim[1:] = fb(sim[1:])
            launches += 1
            ncalls += n
        order = np.argsort(fsim)
        sim, fsim = sim[order], fsim[order]
        it += 1
    status = 1 if ncalls >= maxfev else (2 if it >= maxiter else 0)
    return {"x": sim[0], "fun": fsim[0], "nit": it, "nfev": ncalls, "status": status,
            "success": status == 0, "launches": launches,
            "message": ("Optimization terminated successfully.",
                        "Maximum number of function evaluations has been exceeded.",
                        "Maximum number of iterations has been exceeded.")[status]}
