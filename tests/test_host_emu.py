"""CPU check of the per-cell device formulas (naima_b200/csrc/nb_math.cuh).

nb_math.cuh is `__host__ __device__` scalar code; tests/host_emu/emu.cpp compiles
it with g++ and replays the kernels' lane decomposition in plain loops.  This
lets the formulas, the hoisted trapezoid algebra and the reduction order be
checked against the oracle without a GPU.  It is a TEST build: the product never
loads it (naima_b200 has no CPU path).
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest
from numpy.testing import assert_allclose

import oracle.naima_oracle as o

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emu", "emu.cpp")
OUT = os.path.join(HERE, "host_emu", "_build", "libemu.so")
HDR = os.path.join(HERE, "..", "naima_b200", "csrc", "nb_math.cuh")

dp = ctypes.POINTER(ctypes.c_double)


def P(a):
    return a.ctypes.data_as(dp)


class Term(ctypes.Structure):
    _fields_ = [("src", ctypes.c_void_p), ("wscale", ctypes.c_void_p), ("ld", ctypes.c_int),
                ("off", ctypes.c_int), ("group_end", ctypes.c_int), ("div", ctypes.c_double)]


@pytest.fixture(scope="module")
def emu():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if (not os.path.exists(OUT)
            or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(HDR))):
        # -ffp-contract=off: no FMA contraction, as numpy; the device build
        # contracts, which is covered by the GPU parity tests' tolerances
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-x", "c++", "-std=c++17",
                               "-shared", "-fPIC", "-o", OUT, SRC])
    return ctypes.CDLL(OUT)


KINDS = {"PowerLaw": 0, "ExponentialCutoffPowerLaw": 1, "BrokenPowerLaw": 2,
         "ExponentialCutoffBrokenPowerLaw": 3, "LogParabola": 4}


def pdpar(pd):
    p = np.zeros(8)
    p[: len(pd.params)] = pd.params
    return p


PDS = [
    o.PDist("PowerLaw", 1.3e33, 1e13, 2.41),
    o.PDist("ExponentialCutoffPowerLaw", 1.3e33, 1e13, 2.41, 4.8e13, 1.0),
    o.PDist("ExponentialCutoffPowerLaw", 1.3e33, 1e13, 1.7, 2e12, 2.0),
    o.PDist("BrokenPowerLaw", 2e30, 2e13, 1e12, 1.5, 2.5),
    o.PDist("ExponentialCutoffBrokenPowerLaw", 3.7e36, 1e12, 2.65e11, 1.5, 3.233, 1.863e15, 2.0),
    o.PDist("LogParabola", 1e30, 2e13, 1.7, 0.2),
]


@pytest.mark.parametrize("pd", PDS, ids=lambda p: p.kind)
def test_pd_eval(emu, pd):
    e = np.logspace(8, 15.5, 200)
    out = np.empty_like(e)
    emu.emu_pd_eval(KINDS[pd.kind], P(pdpar(pd)), P(e), e.size, P(out))
    assert_allclose(out, pd(e), rtol=4e-15, atol=0)


def test_interval_exact(emu, ref_exec):
    x, y = ref_exec["tl_x"], ref_exec["tl_y"]
    for k in range(y.shape[0]):
        out = np.empty(x.size - 1)
        yk = np.ascontiguousarray(y[k])
        emu.emu_interval_exact(P(x), P(yk), x.size, P(out))
        assert_allclose(out, ref_exec["tl_int"][k], rtol=2e-14, atol=0)


def test_gtilde(emu):
    x = np.logspace(-12, 2.8, 300)
    out = np.empty_like(x)
    emu.emu_gtilde(P(x), x.size, P(out))
    assert_allclose(out, o.gtilde(x), rtol=1e-15, atol=0)


def test_ic_planck_tables(emu, ref_exec):
    gam, Eph = ref_exec["ic_gam"], ref_exec["ic_Eph"]
    out = np.empty((Eph.size, gam.size))
    for key, T, th in [("ic_iso_cmb", 2.72548, np.nan), ("ic_iso_nir", 3000.0, np.nan),
                       ("ic_ani_60", 20000.0, np.deg2rad(60.0)),
                       ("ic_ani_135", 30.0, np.deg2rad(135.0))]:
        emu.emu_ic_planck(P(gam), gam.size, P(Eph), Eph.size, ctypes.c_double(T),
                          ctypes.c_double(th), P(out))
        assert_allclose(out, ref_exec[key], rtol=1e-13, atol=0)


def test_ic_seed_tables(emu):
    gam = o.electron_grid(1e9, 1e15, 20)
    Eph = np.logspace(6, 14, 17) / o.mec2_eV
    # monochromatic
    eps0 = np.array([6.3e-3 / o.mec2_eV])
    phn = np.array([1.0 * o.eV_erg / o.mec2_erg])
    out = np.empty((Eph.size, gam.size))
    emu.emu_ic_seed(P(gam), gam.size, P(Eph), Eph.size, P(eps0), P(phn), 1, P(out))
    with np.errstate(all="ignore"):
        ref = o.iso_ic_on_monochromatic(gam, eps0, phn, Eph)
    assert_allclose(out, ref, rtol=1e-13, atol=0)
    # tabulated
    eps0 = np.logspace(-3.5, -1.5, 40) / o.mec2_eV
    phn = 1e3 * (eps0 * o.mec2_eV) ** 2 / np.expm1(eps0 * o.mec2_eV / 2.6e-3) * o.mec2_eV
    emu.emu_ic_seed(P(gam), gam.size, P(Eph), Eph.size, P(eps0), P(phn), eps0.size, P(out))
    with np.errstate(all="ignore"):
        ref = o.iso_ic_on_monochromatic(gam, eps0, phn, Eph)
    assert_allclose(out, ref, rtol=2e-12, atol=0)


def test_brems_tables(emu):
    gam = o.electron_grid(o.mec2_eV, 1e15, 30)
    eps = np.logspace(3, 14, 23) / o.mec2_eV
    see = np.empty((eps.size, gam.size))
    s1 = np.empty_like(see)
    emu.emu_brems(P(gam), gam.size, P(eps), eps.size, P(see), P(s1))
    with np.errstate(all="ignore"):
        ree = (o._sigma_ee(np.vstack(gam), eps) / o.mec2_eV).T
        r1 = o._sigma_1(np.vstack(gam), eps).T
    ok = np.isfinite(ree)
    assert_allclose(see[ok], ree[ok], rtol=1e-11, atol=0)
    assert np.array_equal(np.isnan(see), np.isnan(ree))
    ok = np.isfinite(r1)
    assert_allclose(s1[ok], r1[ok], rtol=1e-12, atol=0)


@pytest.mark.parametrize("model", ["Geant4", "Pythia8", "SIBYLL", "QGSJET"])
@pytest.mark.parametrize("nuc", [1, 0])
def test_pp_diffsigma(emu, ref_exec, model, nuc):
    Ep, Eg = ref_exec["pp_Ep"], ref_exec["pp_Eg"]
    mid = {"Geant4": 0, "Pythia8": 1, "SIBYLL": 2, "QGSJET": 3}[model]
    ref = ref_exec["pp_ds_%s_%d" % (model, nuc)]
    out = np.empty(Ep.size)
    for k, eg in enumerate(Eg):
        emu.emu_pp_diffsigma(P(Ep), Ep.size, ctypes.c_double(eg), mid, nuc, P(out))
        assert_allclose(out, ref[k], rtol=5e-13, atol=0)


def test_bspline(emu, lut_probe):
    f = np.load(os.path.join(HERE, "..", "naima_b200", "data",
                             "pp_kafexhiu14_pythia8_nucenh_bspline.npz"))
    tx, ty, c = f["tx"], f["ty"], f["c"]
    x = np.log10(lut_probe["Ep"])
    out = np.empty(x.size)
    for k, eg in enumerate(lut_probe["Eg"]):
        emu.emu_bspl(P(tx), tx.size, P(ty), ty.size, P(c), P(x), x.size,
                     ctypes.c_double(np.log10(eg)), P(out))
        assert_allclose(out, lut_probe["ds"][k], rtol=1e-12, atol=1e-45)


def test_exp_neg(emu):
    """The device exp(-x) (range reduction + degree-12 polynomial + two-step scaling):
    <= 2 ulp over the whole range incl. the subnormal tail, exact zero beyond, NaN kept."""
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(0, 760, 200000), 10 ** rng.uniform(-300, 3, 50000),
                        [0.0, 1e-320, 708.39, 744.4, 745.2, 746.0, 1e4, 1e300, np.inf]])
    out = np.empty(x.size)
    emu.emu_exp_neg(P(x), x.size, P(out))
    ref = np.exp(-x)
    normal = ref > 2.3e-308
    assert_allclose(out[normal], ref[normal], rtol=4.5e-16)
    assert_allclose(out[~normal], ref[~normal], rtol=0, atol=2e-323)
    assert np.all(out[x > 746] == 0)
    y = np.array([np.nan, -1.0, -800.0])
    out = np.empty(3)
    with np.errstate(all="ignore"):
        emu.emu_exp_neg(P(y), 3, P(out))
        assert np.isnan(out[0]) and out[1] == np.exp(1.0) and out[2] == np.inf


def _prep(emu, pd, x, m1, m2, ns, exact=True):
    """exact: reference-order operands (pow, log of the ratio); else the log-space form
    the hoisted kernels consume."""
    N = x.size
    invdlx = np.zeros(N)
    invdlx[:-1] = 1.0 / np.log(x[1:] / x[:-1])
    xn, ds1, nraw = np.empty(N), np.empty(N), np.empty(N)
    fn = emu.emu_pd_prep if exact else emu.emu_pd_prep_log
    with np.errstate(all="ignore"):
        fn(KINDS[pd.kind], P(pdpar(pd)), P(x), N, ctypes.c_double(m1), ctypes.c_double(m2),
           ctypes.c_double(ns), P(invdlx), P(xn), P(ds1), P(nraw))
    return invdlx, xn, ds1, nraw


@pytest.mark.parametrize("pd", PDS, ids=lambda p: p.kind)
def test_pd_prep_log_vs_reference_order(emu, pd):
    """The log-space operands agree with the reference-order ones: n to a few 1e-14
    (|alpha ln(e/e0)| ulps), the slope term to 1e-11 absolute."""
    gam = o.electron_grid(1e9, 510.9989e12, 100)
    _, xn, ds1, nraw = _prep(emu, pd, gam, o.mec2_erg, o.erg_eV, o.mec2_eV, exact=True)
    _, xn2, ds2, n2 = _prep(emu, pd, gam, o.mec2_erg, o.erg_eV, o.mec2_eV, exact=False)
    ok = nraw > 1e-280
    assert ok.sum() > 100
    # the error grows with the cutoff argument c = (e/ec)**beta (exp(-c) turns the
    # rounding of c into c ulps): 5e-14 over the 40 decades that matter, 1e-12 in the tail
    main = nraw > 1e-40 * nraw.max()
    assert_allclose(n2[main], nraw[main], rtol=5e-14)
    assert_allclose(n2[ok], nraw[ok], rtol=1e-12)
    # where exp(-c) alone underflows the reference returns 0; ln-space keeps A x^-alpha e^-c
    assert np.all(n2[nraw == 0] <= 1e-250 * nraw.max())
    okd = ok[:-1] & ok[1:]
    assert_allclose(ds2[:-1][okd], ds1[:-1][okd], rtol=0, atol=1e-11 * np.max(np.abs(ds1[:-1][okd])))


@pytest.mark.parametrize("pd", PDS[:5], ids=lambda p: p.kind)
@pytest.mark.parametrize("mode", [0, 1, 2], ids=["careful", "exact", "lean"])
def test_contract_ic(emu, pd, mode):
    """Hoisted (careful / lean cell, leading zeros skipped) and reference-order (exact)
    contraction vs the oracle's trapz_loglog(n_e * K, gam) for IC on three grey bodies."""
    exact = int(mode == 1)
    gam = o.electron_grid(100e9, 1e15, 100)
    N = gam.size
    pitch = (N + 1) & ~1
    Eph = np.logspace(8, 14.5, 21) / o.mec2_eV
    rows = []
    with np.errstate(all="ignore"):
        for T in (2.72548, 30.0, 3000.0):
            rows.append(o.iso_ic_on_planck(gam, T, Eph))
        rows.append(o.ani_ic_on_planck(gam, 20000.0, Eph, 2.0))
    Kref = np.concatenate(rows, axis=0)
    R = Kref.shape[0]
    K = np.zeros((R, pitch))
    K[:, :N] = Kref
    invdlx, xn, ds1, nraw = _prep(emu, pd, gam, o.mec2_erg, o.erg_eV, o.mec2_eV, exact=bool(exact))
    nref = o.nelec(pd, gam)
    main = nref > (0 if exact else 1e-40 * nref.max())
    assert_allclose(nraw[main], nref[main], rtol=4e-15 if exact else 5e-14)
    lrs = np.zeros((R, pitch))
    with np.errstate(all="ignore"):
        emu.emu_finalize(P(K), R, N, pitch, P(invdlx), P(lrs))
    dlx = np.zeros(N)
    dlx[:-1] = np.log(gam[1:] / gam[:-1])
    out = np.empty(R)
    nfb = ctypes.c_int(0)
    with np.errstate(all="ignore"):
        emu.emu_contract(P(K), P(lrs), R, N, pitch, P(nraw if exact else xn), P(ds1), P(dlx),
                         P(gam), mode, int(mode != 1), P(out), ctypes.byref(nfb))
        ref = o.trapz_loglog(o.nelec(pd, gam) * Kref, gam)
    assert np.all(np.isfinite(ref))
    nz = ref != 0
    assert nz.sum() > R // 2
    assert_allclose(out[nz], ref[nz], rtol=1e-12 if exact else 2e-10, atol=0)
    assert np.all(out[~nz] == 0)
    if mode == 2:  # clean table, regular slopes: the lean cell alone, and it agrees with
        assert nfb.value == 0  # the careful cell to rounding
        out0 = np.empty(R)
        with np.errstate(all="ignore"):
            emu.emu_contract(P(K), P(lrs), R, N, pitch, P(xn), P(ds1), P(dlx), P(gam), 0, 0,
                             P(out0), None)
        assert_allclose(out[nz], out0[nz], rtol=1e-13)


def test_contract_negative_table(emu):
    """Sign changes in K (the pion-decay spline undershoots below zero) take
    trapz_loglog's log branch through the NaN slope (utils.py:341-345)."""
    rng = np.random.default_rng(5)
    x = o.proton_grid(1.2179, 1e6, 40)
    N = x.size
    pitch = (N + 1) & ~1
    R = 6
    K = np.zeros((R, pitch))
    K[:, :N] = np.exp(rng.normal(size=(R, N))) * 1e-26
    K[1, 10:14] *= -1e-3
    K[2, :30] = 0.0
    K[3, 5] = -K[3, 5]
    K[4, -7:] = 0.0
    pd = o.PDist("PowerLaw", 1e-12 * 1e36, 30e12, 2.34)
    invdlx, xn_e, ds1_e, nraw = _prep(emu, pd, x, 1e9, 1.0, 1e9, exact=True)
    _, xn, ds1, _ = _prep(emu, pd, x, 1e9, 1.0, 1e9, exact=False)
    lrs = np.zeros((R, pitch))
    with np.errstate(all="ignore"):
        emu.emu_finalize(P(K), R, N, pitch, P(invdlx), P(lrs))
    dlx = np.zeros(N)
    dlx[:-1] = np.log(x[1:] / x[:-1])
    for mode in (0, 1, 2):
        exact = mode == 1
        out = np.empty(R)
        nfb = ctypes.c_int(0)
        with np.errstate(all="ignore"):
            emu.emu_contract(P(K), P(lrs), R, N, pitch, P(nraw if exact else xn),
                             P(ds1_e if exact else ds1), P(dlx), P(x), mode, int(mode != 1),
                             P(out), ctypes.byref(nfb))
            ref = o.trapz_loglog(o.Jprot(pd, x) * K[:, :N], x)
        assert_allclose(out, ref, rtol=1e-12 if exact else 1e-9, atol=0)
        if mode == 2:  # the rows with sign changes were detected and redone carefully
            assert nfb.value == 2


@pytest.mark.parametrize("pd", PDS[:4], ids=lambda p: p.kind)
def test_synchrotron(emu, pd):
    gam = o.electron_grid(1e9, 1e9 * o.mec2_eV, 100)
    N = gam.size
    E_eV = np.logspace(-6, 6.5, 26)
    E_erg = E_eV * o.eV_erg
    invdlx, xn, ds1, _ = _prep(emu, pd, gam, o.mec2_erg, o.erg_eV, o.mec2_eV, exact=False)
    dlx = np.zeros(N)
    dlx[:-1] = np.log(gam[1:] / gam[:-1])
    for B in (3.24e-6, 1e-4, 1.0):
        out = np.empty(E_eV.size)
        with np.errstate(all="ignore"):
            emu.emu_synchrotron(P(gam), N, P(xn), P(ds1), P(invdlx), P(dlx),
                                ctypes.c_double(B), P(E_erg), E_eV.size, P(out))
        ref = o.synchrotron_spectrum(pd, E_eV, B)
        big = ref > ref.max() * 1e-200
        assert_allclose(out[big], ref[big], rtol=5e-10, atol=0)


def test_combine_lnprob(emu, rxj_data):
    rng = np.random.default_rng(11)
    W, N_E = 7, rxj_data["hess_flux"].size
    flux = rxj_data["hess_flux"]
    a = flux * np.exp(0.3 * rng.normal(size=(W, N_E))) * 4 * np.pi * 0.5
    b = flux * np.exp(0.3 * rng.normal(size=(W, N_E))) * 4 * np.pi * 0.5
    src = np.ascontiguousarray(np.concatenate([a, b], axis=1))  # [W][2 N_E]
    div = 4 * np.pi
    terms = (Term * 2)(Term(src.ctypes.data, None, 2 * N_E, 0, 0, 1.0),
                       Term(src.ctypes.data, None, 2 * N_E, N_E, 1, div))
    unit = np.full(N_E, 1.0)
    elo = rxj_data["hess_flux_error"].copy()
    ehi = 1.3 * elo
    ul = rxj_data["hess_ul"].astype(np.int32)
    ul[3] = 1
    cl = np.full(N_E, 0.95)
    prior = np.array([0.0, -np.inf, 1.5, 0, 0, 0, 0])
    fm = np.empty((W, N_E))
    lnp = np.empty(W)
    ip = ctypes.POINTER(ctypes.c_int)
    emu.emu_combine_lnprob(terms, 2, W, N_E, P(unit), P(flux), P(elo), P(ehi),
                           ul.ctypes.data_as(ip), P(cl), P(prior), P(fm), P(lnp))
    data = dict(flux=flux, flux_error_lo=elo, flux_error_hi=ehi, ul=ul.astype(bool), cl=cl)
    for w in range(W):
        model = (a[w] + b[w]) / div
        assert_allclose(fm[w], model, rtol=1e-15)
        ref = o.lnprobmodel(model, data) + prior[w] if np.isfinite(prior[w]) else prior[w]
        assert_allclose(lnp[w], ref, rtol=1e-14)


@pytest.mark.parametrize("pd", PDS, ids=lambda p: p.kind)
def test_selfprep_kernels(emu, pd):
    """The self-contained synchrotron kernel's math (operands from the grid's ln x table, no
    set-up arrays; per-node 1/Ec and cbrt(1/Ec) from the walker-independent g^-2 tables)
    against the oracle's synchrotron spectrum, for every particle distribution."""
    gam = o.electron_grid(1e9, 1e9 * o.mec2_eV, 100)
    N = gam.size
    lnx = np.log(gam)
    dlx = np.zeros(N)
    dlx[:-1] = np.log(gam[1:] / gam[:-1])
    invdlx = np.zeros(N)
    invdlx[:-1] = 1.0 / dlx[:-1]
    E_eV = np.logspace(-5, 6.3, 17)
    E_erg = E_eV * o.eV_erg
    for B in (3.24e-6, 2e-5):
        out = np.empty(E_eV.size)
        with np.errstate(all="ignore"):
            emu.emu_synchrotron_fused(KINDS[pd.kind], P(pdpar(pd)), ctypes.c_double(o.mec2_erg),
                                      ctypes.c_double(o.erg_eV), ctypes.c_double(o.mec2_eV),
                                      P(gam), P(lnx), N, P(invdlx), P(dlx), ctypes.c_double(B),
                                      P(E_erg), E_eV.size, P(out))
        ref = o.synchrotron_spectrum(pd, E_eV, B)
        big = ref > ref.max() * 1e-200
        assert_allclose(out[big], ref[big], rtol=5e-10)


def test_ssc_hoisted(emu):
    """Self-Compton seed (examples/CrabNebula_SynSSC.py shapes, reduced grid): the hoisted
    two-level integral (f_AA81 table with sentinel slopes, lean inner cell with the careful
    fall-back, outer trapezoid) against the oracle's iso_ic_on_monochromatic path."""
    pd = o.PDist("ExponentialCutoffBrokenPowerLaw", 3.699e36, 1e12, 0.265e12, 1.5, 3.233,
                 1863e12, 2.0)
    gam = o.electron_grid(1e8, 50e15, 12)
    N = gam.size
    Esy = np.logspace(-7, 9, 33)
    lsy = o.synchrotron_spectrum(pd, Esy, 125e-6, Eemin_eV=1e8, Eemax_eV=50e15, nEed=12)
    Rpwn = 2.1 * 3.0856775814913673e18
    phn_eV = lsy / (4 * np.pi * Rpwn**2 * o.c_cgs) * 2.24  # 1/(eV cm3)
    phn_eV[-3:] = 0.0  # a seed field that underflows at its high-energy end: zero nodes
    E_eV = np.logspace(5, 15, 13)
    Eph = E_eV * o.eV_erg / o.mec2_erg
    eps0 = Esy / o.mec2_eV
    phn = phn_eV * o.mec2_eV
    invdlx, xn, ds1, _ = _prep(emu, pd, gam, o.mec2_erg, o.erg_eV, o.mec2_eV, exact=False)
    dlx = np.zeros(N)
    dlx[:-1] = np.log(gam[1:] / gam[:-1])
    out = np.empty(E_eV.size)
    nfb = ctypes.c_int(0)
    with np.errstate(all="ignore"):
        emu.emu_ssc(P(gam), N, P(Eph), P(E_eV), E_eV.size, P(eps0), P(phn), eps0.size, P(xn),
                    P(ds1), P(dlx), P(invdlx), P(out), ctypes.byref(nfb))
        ref = o.ic_seed_spectrum(pd, ("array", Esy, phn_eV), E_eV, gam)
    nz = ref > ref.max() * 1e-200
    assert nz.sum() >= 8
    assert_allclose(out[nz], ref[nz], rtol=1e-9)
    assert np.all(out[ref == 0] == 0)
    assert nfb.value < 0.05 * N * E_eV.size  # the lean cell carries (almost) all rows


def test_kelner06_fixed_grid_vs_adaptive_quadrature(emu):
    """PionDecayKelner06 (radiative.py:1543-1767): the device formulation (log-log trapezoid
    over a per-row proton grid, 100 nodes per decade) against the oracle's restatement with
    the reference's adaptive QUADPACK calls; both sides are only 1e-3 accurate by the
    reference's own epsrel."""
    pd = o.PDist("ExponentialCutoffPowerLaw", 1e-12, 20e12, 2.0, 10e12, 1.0)
    E_eV = np.logspace(9, 13, 20)
    Eg = np.concatenate([E_eV * 1e-12, [0.1, 0.1]])
    hi = np.concatenate([E_eV * 1e-12 >= 0.1, [True, False]]).astype(np.int32)
    R, N = Eg.size, 701
    out = np.empty(R)
    with np.errstate(all="ignore"):
        emu.emu_kelner(KINDS[pd.kind], P(pdpar(pd)), P(Eg), hi.ctypes.data_as(ctypes.c_void_p), R,
                       N, ctypes.c_double(7.0), P(out))
    nhat = out[-2] / out[-1]
    spec = np.where(hi[:-2] == 1, out[:-2], nhat * out[:-2]) * 1e-12
    pk = o.PionDecayKelner06(pd)
    want = pk.spectrum(E_eV)
    assert_allclose(nhat, pk.nhat, rtol=2e-3)
    assert_allclose(spec, want, rtol=3e-3)
    lum = o.trapz_loglog(spec * E_eV, E_eV) * o.eV_erg
    assert_allclose(lum, 5.54580582494601e-13, rtol=2e-3)  # tests/test_models.py:464
