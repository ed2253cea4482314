# -*- coding: utf-8 -*-
"""Device engine: grids, emissivity tables and kernel launches.

Thin, stateless-looking Python over the C ABI (include/naima_b200.h).  PyTorch
tensors are used only as device-memory containers and for the stream handle.
Everything is float64 in the fixed units of the C ABI (eV, G, K, cm, rad).

Layout in HBM:
  grid        x[pitch], dlx[pitch], invdlx[pitch]           per particle grid
  table       K[R][pitch], lrs[R][pitch], coef[R]           per (process, grid, photon energies)
  operands    xn[W][pitch], ds1[W][pitch] (, nraw[W][pitch]) per launch
  results     spec[W][R]                                      per launch
Row r = component * N_E + photon-energy index.
"""
import ctypes
import os
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from ._lib import PD_KIND, PP_MODEL, NB_PD_MAXPAR, check, lib, nb_term

# CODATA 2018 / astropy>=6.1 values in cgs, composed as the reference composes
# them (radiative.py:36-40)
c_cgs = 29979245800.0
m_e_g = 9.1093837015e-28
sigma_sb_cgs = 5.6703744191844314e-05
eV_erg = 1.602176634e-12
erg_eV = 1e-7 / 1.602176634e-19
mec2_erg = m_e_g * c_cgs**2
mec2_eV = mec2_erg / eV_erg
ar_cgs = 4 * sigma_sb_cgs / c_cgs
mpc2_GeV = 0.9382720881604903
T_TH = 0.27966184
T_CMB = 2.72548

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

# sentinel slope of intervals with a zero end point (NB_BIG_SLOPE in csrc/nb_math.cuh)
BIG_SLOPE = 1.5 * 2.0 ** 1022

# reference operation order (log10/pow per interval) instead of the hoisted
# contraction; settable at run time (tests exercise both)
EXACT = False
# hoisted contraction: lean cell (8 fp64 instructions per interval, irregular slopes detected
# and redone with the careful cell) where the table allows it; False: careful cell everywhere
LEAN = True


def device():
    if not torch.cuda.is_available():
        raise RuntimeError("naima_b200 needs a CUDA device (sm_100a); there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


class capture_graph:
    """``with capture_graph() as g:`` -- capture the launches of the block into a
    torch.cuda.CUDAGraph on a side stream.  The cyclic garbage collector is paused for the
    duration: a collection that happens to run inside a capture can free pinned-memory or
    graph objects of earlier plans, and those runtime calls are not permitted while a
    stream is capturing (the capture dies with cudaErrorStreamCaptureInvalidated)."""

    def __enter__(self):
        import gc

        self._gc = gc.isenabled()
        gc.collect()
        gc.disable()
        self.graph = torch.cuda.CUDAGraph()
        self._s = torch.cuda.Stream()
        self._s.wait_stream(torch.cuda.current_stream())
        self._sc = torch.cuda.stream(self._s)
        self._sc.__enter__()
        self._gc_ctx = torch.cuda.graph(self.graph, stream=self._s,
                                        capture_error_mode="thread_local")
        self._gc_ctx.__enter__()
        return self.graph

    def __exit__(self, *exc):
        import gc

        try:
            self._gc_ctx.__exit__(*exc)
        finally:
            self._sc.__exit__(*exc)
            if self._gc:
                gc.enable()
        torch.cuda.current_stream().wait_stream(self._s)
        return False


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def to_dev(a, dtype=torch.float64):
    a = np.ascontiguousarray(a)
    if not a.flags.writeable:  # torch.from_numpy warns on read-only views (np.broadcast_to ...)
        a = a.copy()
    return torch.from_numpy(a).to(device=device(), dtype=dtype)


def empty(*shape, dtype=torch.float64):
    return torch.empty(*shape, dtype=dtype, device=device())


def zeros(*shape, dtype=torch.float64):
    return torch.zeros(*shape, dtype=dtype, device=device())


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def even(n):
    return (n + 1) & ~1


# ------------------------------------------------------------------------------
# particle grids
# ------------------------------------------------------------------------------
class Grid:
    """Log-spaced particle grid (radiative.py:147-154, 1002-1009) on the device."""

    def __init__(self, x, species):
        self.x = np.asarray(x, dtype=float)
        self.key = (species, self.x.size, hash(self.x.tobytes()))
        self.N = self.x.size
        self.pitch = even(self.N)
        self.species = species
        pad = np.zeros(self.pitch)
        pad[: self.N] = self.x
        dlx = np.zeros(self.pitch)
        dlx[: self.N - 1] = np.log(self.x[1:] / self.x[:-1])
        inv = np.zeros(self.pitch)
        inv[: self.N - 1] = 1.0 / dlx[: self.N - 1]
        lnx = np.zeros(self.pitch)
        lnx[: self.N] = np.log(self.x)
        # synchrotron node tables: 1/Ec = (1/kB) x^-2, cbrt(1/Ec) = cbrt(1/kB) cbrt(x^-2)
        gm2 = np.zeros(self.pitch)
        gm2[: self.N] = 1.0 / (self.x * self.x)
        g23 = np.cbrt(gm2)
        self._host = (pad, dlx, inv, lnx, gm2, g23)
        self._dev = None
        if species == "electron":  # e = (gam * mec2[erg]) * (erg -> eV); n per unit gam
            self.e_mul1, self.e_mul2, self.n_scale = mec2_erg, erg_eV, mec2_eV
            self.x_to_erg = mec2_erg
        else:  # protons: x = Ep in GeV; J per GeV
            self.e_mul1, self.e_mul2, self.n_scale = 1e9, 1.0, 1e9
            self.x_to_erg = 1e9 * eV_erg


    def _device(self):
        if self._dev is None:  # uploaded on first use so that tracing needs no GPU
            self._dev = tuple(to_dev(a) for a in self._host)
        return self._dev

    x_d = property(lambda self: self._device()[0])
    dlx_d = property(lambda self: self._device()[1])
    invdlx_d = property(lambda self: self._device()[2])
    lnx_d = property(lambda self: self._device()[3])
    gm2_d = property(lambda self: self._device()[4])
    g23_d = property(lambda self: self._device()[5])


_GRIDS = OrderedDict()


def _cache_get(cache, key, make, maxsize):
    if key in cache:
        cache.move_to_end(key)
        return cache[key]
    val = make()
    cache[key] = val
    while len(cache) > maxsize:
        cache.popitem(last=False)
    return val


def electron_grid(Eemin_eV, Eemax_eV, nEed):
    """radiative.py:147-154 (identical arithmetic to the oracle's electron_grid)."""
    def make():
        l0 = np.log10(Eemin_eV * eV_erg / mec2_erg)
        l1 = np.log10(Eemax_eV * eV_erg / mec2_erg)
        return Grid(np.logspace(l0, l1, max(10, int(nEed * (l1 - l0)))), "electron")
    return _cache_get(_GRIDS, ("e", float(Eemin_eV), float(Eemax_eV), float(nEed)), make, 64)


def proton_grid(Epmin_GeV, Epmax_GeV, nEpd):
    """radiative.py:1002-1009."""
    def make():
        n = max(10, int(nEpd * (np.log10(Epmax_GeV / Epmin_GeV))))
        return Grid(np.logspace(np.log10(Epmin_GeV), np.log10(Epmax_GeV), n), "proton")
    return _cache_get(_GRIDS, ("p", float(Epmin_GeV), float(Epmax_GeV), float(nEpd)), make, 64)


# ------------------------------------------------------------------------------
# particle distributions
# ------------------------------------------------------------------------------
def pd_params_array(kind, params):
    """Broadcast reference-order eval() parameters to a [W][8] host array."""
    cols = [np.atleast_1d(np.asarray(p, dtype=float)) for p in params]
    W = max(c.size for c in cols)
    out = np.zeros((W, NB_PD_MAXPAR))
    for k, c in enumerate(cols):
        if c.size not in (1, W):
            raise ValueError("particle distribution parameters have inconsistent batch sizes")
        out[:, k] = c
    return out


def pdist_eval(kind, params_d, W, e_eV):
    """out[w][i] = PD.eval(e[i]) on the device (models.py eval staticmethods)."""
    e_d = to_dev(np.atleast_1d(e_eV))
    out = empty(W, e_d.numel())
    check(lib().nb_pdist_eval(PD_KIND[kind], ptr(params_d), W, ptr(e_d), e_d.numel(), ptr(out),
                              stream()), "nb_pdist_eval")
    return out


class Prepared:
    __slots__ = ("xn", "ds1", "nraw", "W", "grid")


def pd_prep(grid, kind, params_d, W, need_raw=None, out=None):
    """Integration operands of W walkers on `grid` (nb_pd_prep)."""
    need_raw = EXACT if need_raw is None else need_raw
    pr = out if out is not None else Prepared()
    if out is None:
        pr.xn, pr.ds1 = empty(W, grid.pitch), empty(W, grid.pitch)
        pr.nraw = empty(W, grid.pitch) if need_raw else None
        pr.W, pr.grid = W, grid
    check(lib().nb_pd_prep_ex(PD_KIND[kind], ptr(params_d), W, ptr(grid.x_d), grid.N,
                              grid.e_mul1, grid.e_mul2, grid.n_scale, ptr(grid.invdlx_d),
                              ptr(pr.xn), ptr(pr.ds1), ptr(pr.nraw), grid.pitch, stream()),
          "nb_pd_prep")
    return pr


def particle_energy(grid, kind, params_d, W, out=None):
    """W[w] = trapz_loglog(x n, x * x_to_erg) in erg (We / Wp)."""
    out = empty(W) if out is None else out
    check(lib().nb_particle_energy(PD_KIND[kind], ptr(params_d), W, ptr(grid.x_d), grid.N,
                                   grid.e_mul1, grid.e_mul2, grid.n_scale, grid.x_to_erg,
                                   ptr(out), stream()), "nb_particle_energy")
    return out


# ------------------------------------------------------------------------------
# emissivity tables
# ------------------------------------------------------------------------------
class Table:
    """Walker-independent emissivity rows K[R][pitch] (+ log-slopes, + coef[R])."""

    def __init__(self, grid, R, n_comp, N_E):
        self.grid, self.R, self.n_comp, self.N_E = grid, R, n_comp, N_E
        self.K = zeros(R, grid.pitch)
        self.lrs = zeros(R, grid.pitch)
        self.coef = None
        self.row_j0 = None  # first non-zero node per row (leading zeros are skipped)
        self.clean = False  # no negative / non-finite entry: the lean cell applies

    def finalize(self, coef):
        g = self.grid
        check(lib().nb_table_finalize(ptr(self.K), self.R, g.N, g.pitch, ptr(g.invdlx_d),
                                      ptr(self.lrs), stream()), "nb_table_finalize")
        self.row_j0 = zeros(self.R, dtype=torch.int32)
        flags = zeros(1, dtype=torch.int32)
        check(lib().nb_table_scan(ptr(self.K), self.R, g.N, g.pitch, ptr(self.row_j0), ptr(flags),
                                  stream()), "nb_table_scan")
        self.clean = int(flags.item()) == 0
        self.coef = to_dev(coef)
        return self


_TABLES = OrderedDict()
_TABLES_MAX = 32


def _ekey(E):
    E = np.ascontiguousarray(E, dtype=float)
    return (E.size, hash(E.tobytes()))


def ic_table(grid, E_eV, seeds):
    """IC rows for thermal (iso / aniso) and monochromatic / tabulated seeds.

    seeds: tuple of ("thermal", T_K, u_erg_cm3[, theta_rad]) | ("mono", E_eV, u_erg_cm3)
           | ("array", E_eV[Ns] tuple, dn/dE[1/(eV cm3)] tuple)
    coef[r] = uf * Eph / E_eV  (radiative.py:684-687)."""
    E_eV = np.ascontiguousarray(E_eV, dtype=float)

    def make():
        N_E, S = E_eV.size, len(seeds)
        Eph = E_eV * eV_erg / mec2_erg
        Eph_d = to_dev(Eph)
        tb = Table(grid, S * N_E, S, N_E)
        coef = np.empty(S * N_E)
        keep = tb.inputs = [Eph_d]  # device inputs stay alive with the table
        for s, seed in enumerate(seeds):
            if seed[0] == "thermal":
                T, uu = seed[1], seed[2]
                if uu == 0:
                    uu = ar_cgs * T**4
                uf = uu / (ar_cgs * T**4)
                th = seed[3] if len(seed) > 3 else np.nan
                # named locals: a temporary tensor would be freed (and its block
                # reused by the next upload) before the kernel reads it
                T_d, th_d = to_dev([T]), to_dev([th])
                check(lib().nb_ic_planck_table(
                    ptr(grid.x_d), grid.N, ptr(Eph_d), N_E, ptr(T_d), ptr(th_d),
                    1, ptr(tb.K), grid.pitch, s * N_E, stream()), "nb_ic_planck_table")
                keep.extend([T_d, th_d])
            else:
                uf = 1.0
                if seed[0] == "mono":
                    eps0 = np.atleast_1d(seed[1]) / mec2_eV
                    phn = np.atleast_1d(seed[2]) / mec2_erg
                else:
                    eps0 = np.asarray(seed[1], dtype=float) / mec2_eV
                    phn = np.asarray(seed[2], dtype=float) * mec2_eV
                eps0_d, phn_d = to_dev(eps0), to_dev(phn)
                check(lib().nb_ic_seed_table(
                    ptr(grid.x_d), grid.N, ptr(Eph_d), N_E, ptr(eps0_d), ptr(phn_d),
                    eps0.size, ptr(tb.K), grid.pitch, s * N_E, stream()), "nb_ic_seed_table")
                keep.extend([eps0_d, phn_d])
            coef[s * N_E:(s + 1) * N_E] = uf * Eph / E_eV
        return tb.finalize(coef)

    return _cache_get(_TABLES, ("ic", grid.key, _ekey(E_eV), seeds), make, _TABLES_MAX)


def brems_table(grid, E_eV):
    """Rows [0,N_E): sigma_ee/mec2_eV, rows [N_E,2N_E): sigma_1; coef folds c and
    the per-eV conversion (radiative.py:940-971)."""
    E_eV = np.ascontiguousarray(E_eV, dtype=float)

    def make():
        N_E = E_eV.size
        eps_d = to_dev(E_eV * eV_erg / mec2_erg)
        tb = Table(grid, 2 * N_E, 2, N_E)
        check(lib().nb_brems_table(ptr(grid.x_d), grid.N, ptr(eps_d), N_E, ptr(tb.K),
                                   grid.pitch, 0, stream()), "nb_brems_table")
        coef = np.concatenate([np.full(N_E, c_cgs), np.full(N_E, c_cgs / mec2_eV)])
        return tb.finalize(coef)

    return _cache_get(_TABLES, ("br", grid.key, _ekey(E_eV)), make, _TABLES_MAX)


_LUT = {}


def _pp_lut():
    if "pythia8" not in _LUT:
        f = np.load(os.path.join(DATA_DIR, "pp_kafexhiu14_pythia8_nucenh_bspline.npz"))
        _LUT["pythia8"] = (to_dev(f["tx"]), to_dev(f["ty"]), to_dev(f["c"]))
    return _LUT["pythia8"]


def pp_table(grid, E_eV, useLUT, hiEmodel, nuclear_enhancement):
    """dsigma/dEgamma rows; coef = c * 1e-9 (1/(s GeV) -> 1/(s eV)), radiative.py:1523-1536."""
    E_eV = np.ascontiguousarray(E_eV, dtype=float)

    def make():
        N_E = E_eV.size
        Eg_d = to_dev(E_eV * 1e-9)
        tb = Table(grid, N_E, 1, N_E)
        if useLUT:
            tx, ty, c = _pp_lut()
            check(lib().nb_pp_lut_table(ptr(tx), tx.numel(), ptr(ty), ty.numel(), ptr(c),
                                        ptr(grid.x_d), grid.N, ptr(Eg_d), N_E, ptr(tb.K),
                                        grid.pitch, 0, stream()), "nb_pp_lut_table")
        else:
            check(lib().nb_pp_analytic_table(PP_MODEL[hiEmodel], int(bool(nuclear_enhancement)),
                                             ptr(grid.x_d), grid.N, ptr(Eg_d), N_E, ptr(tb.K),
                                             grid.pitch, 0, stream()), "nb_pp_analytic_table")
        return tb.finalize(np.full(N_E, c_cgs * 1e-9))

    key = ("pp", grid.key, _ekey(E_eV), bool(useLUT), hiEmodel, bool(nuclear_enhancement))
    return _cache_get(_TABLES, key, make, _TABLES_MAX)


# ------------------------------------------------------------------------------
# hot kernels
# ------------------------------------------------------------------------------
def contract(table, prep, out=None, exact=None, coef=None):
    """out[w][r] = coef[r] * trapz_loglog(n[w,:] K[r,:], x); coef: row coefficients in place
    of the table's own (a plan's copy with per-energy factors folded in)."""
    exact = EXACT if exact is None else exact
    g = table.grid
    out = empty(prep.W, table.R) if out is None else out
    if exact and prep.nraw is None:
        raise ValueError("exact contraction needs pd_prep(need_raw=True)")
    mode = 1 if exact else (2 if (LEAN and table.clean) else 0)
    check(lib().nb_contract_ex(ptr(table.K), ptr(table.lrs), table.R, g.N, g.pitch,
                               ptr(table.row_j0), ptr(prep.nraw if exact else prep.xn),
                               ptr(prep.ds1), g.pitch, prep.W, ptr(g.dlx_d), ptr(g.x_d),
                               ptr(table.coef if coef is None else coef), ptr(out), mode,
                               stream()), "nb_contract_ex")
    return out


def photon_energies(E_eV):
    """Device pair (E [erg], cbrt(E)) of photon energies for the synchrotron kernels."""
    E_erg = np.ascontiguousarray(E_eV, dtype=float) * eV_erg
    return to_dev(np.stack([E_erg, np.cbrt(E_erg)]))


def synchrotron(grid, prep, B_d, E2_d, out=None):
    """Synchrotron._spectrum (radiative.py:282-342): out[w][e] in 1/(s eV); E2_d from
    photon_energies()."""
    N_E = E2_d.shape[1]
    out = empty(prep.W, N_E) if out is None else out
    check(lib().nb_synchrotron(ptr(grid.x_d), grid.N, ptr(grid.gm2_d), ptr(grid.g23_d),
                               ptr(prep.xn), ptr(prep.ds1), grid.pitch, ptr(grid.invdlx_d),
                               ptr(grid.dlx_d), ptr(B_d), prep.W, ptr(E2_d[0]), ptr(E2_d[1]),
                               N_E, ptr(out), out.stride(0), stream()), "nb_synchrotron")
    return out


# ------------------------------------------------------------------------------
# synchrotron self-Compton: IC on a per-walker tabulated seed, hoisted
# ------------------------------------------------------------------------------
class SscTable:
    """f_AA81(gam_g, eps0_s, Eph_e) (radiative.py:621-636), s-major, with its log-slopes
    along the seed-energy axis: walker independent, built once per (grid, photon energies,
    seed energies).  Row r = e * N + g, row pitch Rp (multiple of 128)."""

    def __init__(self, grid, E_eV, seed_E_eV):
        E_eV = np.ascontiguousarray(E_eV, dtype=float)
        eps0 = np.ascontiguousarray(seed_E_eV, dtype=float) / mec2_eV
        self.grid, self.N_E, self.Ns = grid, E_eV.size, eps0.size
        if self.Ns < 2:
            raise ValueError("a tabulated seed needs at least two energies")
        self.R = self.N_E * grid.N
        self.Rp = (self.R + 127) & ~127
        self.spitch = even(self.Ns)
        Eph = E_eV * eV_erg / mec2_erg
        dl = np.zeros(self.spitch)
        dl[: self.Ns - 1] = np.log(eps0[1:] / eps0[:-1])
        inv = np.zeros(self.spitch)
        inv[: self.Ns - 1] = 1.0 / dl[: self.Ns - 1]
        self.eps0_d, self.Eph_d = to_dev(eps0), to_dev(Eph)
        self.dlx_s_d, self.invdlx_s_d = to_dev(dl), to_dev(inv)
        self.coef_e_d = to_dev(Eph / E_eV)  # lum = Eph * integral; spec = lum / E (:684-687)
        self.KL = empty(self.Ns - 1, self.Rp, 2)  # (F[s+1][r], slope[s][r]) pairs
        self.F0 = empty(self.Rp)
        self.coef = empty(self.Rp)
        check(lib().nb_ssc_table(ptr(grid.x_d), grid.N, ptr(self.Eph_d), self.N_E,
                                 ptr(self.eps0_d), ptr(self.invdlx_s_d), self.Ns, ptr(self.KL),
                                 ptr(self.F0), ptr(self.coef), self.Rp, stream()), "nb_ssc_table")


def ssc_table(grid, E_eV, seed_E_eV):
    key = ("ssc", grid.key, _ekey(E_eV), _ekey(seed_E_eV))
    return _cache_get(_TABLES, key, lambda: SscTable(grid, E_eV, seed_E_eV), _TABLES_MAX)


def make_ssc_sources(sources):
    """sources: list of (src tensor [W][ld], off, fac)."""
    from ._lib import nb_ssc_src

    arr = (nb_ssc_src * len(sources))()
    for k, (src, off, fac) in enumerate(sources):
        arr[k].src, arr[k].ld, arr[k].off, arr[k].fac = src.data_ptr(), src.stride(0), off, fac
    return arr


def ssc_seed(tb, sources, W, sxn, sds):
    """Seed density operands of W walkers from luminosities: nb_ssc_seed."""
    arr = sources if isinstance(sources, ctypes.Array) else make_ssc_sources(sources)
    check(lib().nb_ssc_seed(arr, len(arr), W, tb.Ns, ptr(tb.invdlx_s_d), ptr(sxn), ptr(sds),
                            tb.spitch, stream()), "nb_ssc_seed")


def ssc_inner(tb, sxn, sds, W, inner):
    check(lib().nb_ssc_inner(ptr(tb.KL), ptr(tb.F0), ptr(tb.coef), tb.Rp, tb.Ns, ptr(sxn),
                             ptr(sds), tb.spitch, W, ptr(tb.dlx_s_d), ptr(inner), stream()),
          "nb_ssc_inner")


def ssc_outer(tb, inner, prep, out, out_off=0):
    g = tb.grid
    check(lib().nb_ssc_outer(ptr(inner), tb.Rp, g.N, tb.N_E, prep.W, ptr(prep.xn), ptr(prep.ds1),
                             g.pitch, ptr(g.dlx_d), ptr(g.invdlx_d), ptr(tb.coef_e_d), ptr(out),
                             out.stride(0), out_off, stream()), "nb_ssc_outer")


def ic_seed_spectrum(grid, prep, E_eV, seed_E_eV, phn_d, per_walker, out, out_off):
    """Fused IC on a tabulated seed whose density may differ per walker (SSC)."""
    Eph = np.ascontiguousarray(E_eV, dtype=float) * eV_erg / mec2_erg
    eps0 = np.ascontiguousarray(seed_E_eV, dtype=float) / mec2_eV
    Eph_d, eps0_d = to_dev(Eph), to_dev(eps0)
    check(lib().nb_ic_seed_spectrum(ptr(grid.x_d), grid.N, ptr(prep.nraw), grid.pitch,
                                    ptr(Eph_d), Eph.size, ptr(eps0_d), ptr(phn_d),
                                    eps0.size, eps0.size if per_walker else 0, prep.W, ptr(out),
                                    out.shape[1], out_off, stream()), "nb_ic_seed_spectrum")
    return out


# ------------------------------------------------------------------------------
# pion decay, Kelner+06
# ------------------------------------------------------------------------------
KELNER_NODES_PER_DECADE = 100
KELNER_DECADES = 7.0


def kelner_table(Eg_TeV, hi):
    """Per-row proton grids and integrand kernels (nb_kelner_table): (Ep[R][N], Kk[R][N])."""
    def make():
        R = Eg_TeV.size
        N = int(KELNER_NODES_PER_DECADE * KELNER_DECADES) + 1
        Eg_d = to_dev(Eg_TeV)
        hi_d = to_dev(hi.astype(np.int32), dtype=torch.int32)
        Ep, Kk = empty(R, N), empty(R, N)
        check(lib().nb_kelner_table(ptr(Eg_d), ptr(hi_d), R, N, KELNER_DECADES, ptr(Ep), ptr(Kk),
                                    stream()), "nb_kelner_table")
        return Ep, Kk, (Eg_d, hi_d)

    key = ("kelner", _ekey(Eg_TeV), hash(hi.tobytes()))
    return _cache_get(_TABLES, key, make, _TABLES_MAX)


def kelner_rows(kind, params_d, W, Ep, Kk):
    """out[w][r] = trapz_loglog(J_w(Ep[r]) Kk[r], Ep[r]) in 1/(s TeV) (nb_kelner_rows)."""
    R, N = Ep.shape
    out = empty(W, R)
    check(lib().nb_kelner_rows(PD_KIND[kind], ptr(params_d), W, ptr(Ep), ptr(Kk), R, N, ptr(out),
                               stream()), "nb_kelner_rows")
    return out


def make_terms(terms):
    """terms: list of (src tensor [W][ld], off, group_end, div, wscale tensor|None)."""
    arr = (nb_term * len(terms))()
    for k, (src, off, ge, div, ws) in enumerate(terms):
        arr[k].src = src.data_ptr()
        arr[k].wscale = ws.data_ptr() if ws is not None else None
        arr[k].ld = src.shape[1]
        arr[k].off = off
        arr[k].group_end = int(ge)
        arr[k].div = float(div)
    return arr


def combine(terms, W, N_E, unit_fac_d, flux_out=None, data=None, prior_d=None, lnp_out=None,
            mv=None, pars_d=None, flux_ld=0, lnp_ld=1, peers=None, nb=0):
    """flux model (radiative.py:102-111) and, with `data`, lnprob (core.py:64-121); with
    `mv` (an nb_stretch) also the accept step and chain append of the half-step."""
    arr = make_terms(terms) if not isinstance(terms, ctypes.Array) else terms
    d = data
    if peers is not None and mv is not None:  # replicated state: accept step writes every copy
        check(lib().nb_combine_lnprob_update_push(
            ctypes.byref(mv), ctypes.byref(peers), ptr(pars_d), arr, len(arr), W, N_E,
            ptr(unit_fac_d), ptr(d.flux), ptr(d.err_lo), ptr(d.err_hi), ptr(d.ul), ptr(d.cl),
            ptr(prior_d), ptr(flux_out), flux_ld, ptr(lnp_out), stream()),
            "nb_combine_lnprob_update_push")
        return
    if mv is not None:
        check(lib().nb_combine_lnprob_update(
            ctypes.byref(mv), ptr(pars_d), arr, len(arr), W, N_E, ptr(unit_fac_d),
            ptr(d.flux), ptr(d.err_lo), ptr(d.err_hi), ptr(d.ul), ptr(d.cl), ptr(prior_d),
            ptr(flux_out), flux_ld, ptr(lnp_out), stream()), "nb_combine_lnprob_update")
        return
    check(lib().nb_combine_lnprob_ld(
        arr, len(arr), W, N_E, ptr(unit_fac_d),
        ptr(d.flux) if d else None, ptr(d.err_lo) if d else None, ptr(d.err_hi) if d else None,
        ptr(d.ul) if d else None, ptr(d.cl) if d else None, ptr(prior_d), ptr(flux_out),
        flux_ld, ptr(lnp_out), max(int(lnp_ld), 1), stream()), "nb_combine_lnprob")


def trapz_loglog(y, x, intervals=False):
    """utils.py:285-355 on the device: y [R][N] host array, x [N] or [R][N]."""
    y = np.ascontiguousarray(y, dtype=float)
    x = np.ascontiguousarray(x, dtype=float)
    R, N = y.shape
    y_d, x_d = to_dev(y), to_dev(x)
    out = empty(R)
    iv = empty(R, max(N - 1, 1)) if intervals else None
    check(lib().nb_trapz_loglog(ptr(y_d), R, N, N, ptr(x_d), N if x.ndim == 2 else 0, ptr(out),
                                ptr(iv), stream()), "nb_trapz_loglog")
    return (iv[:, : N - 1] if intervals else out).cpu().numpy()


class DeviceData:
    """Data columns of core.py:64-94 on the device, in the data table's flux unit."""

    def __init__(self, flux, err_lo, err_hi, ul, cl):
        self.N_E = int(np.size(flux))
        self.flux, self.err_lo, self.err_hi = to_dev(flux), to_dev(err_lo), to_dev(err_hi)
        self.ul = to_dev(np.asarray(ul).astype(np.int32), dtype=torch.int32)
        self.cl = to_dev(np.broadcast_to(np.asarray(cl, dtype=float), (self.N_E,)))


def fallback_counts(reset=False):
    """(contraction (walker, row tile) pairs, self-Compton rows) that the lean cell handed to
    the careful cell since the last reset."""
    torch.cuda.synchronize()
    out = (ctypes.c_ulonglong * 2)()
    check(lib().nb_fallback_counts(out, int(bool(reset))), "nb_fallback_counts")
    return int(out[0]), int(out[1])


def fp64_peak_tflops(iters=4096, reps=5):
    """Measured fp64 FMA throughput of this GPU (2 flops per DFMA), TFLOP/s."""
    sm = torch.cuda.get_device_properties(device()).multi_processor_count
    blocks, threads = sm * 8, 256
    out = empty(blocks * threads)
    fn = lib().nb_fp64_peak_probe
    check(fn(ptr(out), blocks, threads, iters, stream()))
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(fn(ptr(out), blocks, threads, iters, stream()))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = max(best, blocks * threads * iters * 16 * 2 / (ms * 1e-3) / 1e12)
    return best
