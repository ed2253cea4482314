"""GPU parity tests: the CUDA path (through the C ABI, via naima_b200's host
layer) against the CPU oracle and the reference's golden vectors.

Tolerances (BASELINE.json north_star): flux rtol <= 1e-6, lnprob rtol <= 1e-8.
The reference's own known-answer numbers are asserted at the reference tests'
tolerance (assert_allclose default rtol 1e-7).  Most comparisons against the
oracle are far tighter; the asserted bound is written next to each.
"""
import os

import numpy as np
import pytest
from numpy.testing import assert_allclose

import oracle.naima_oracle as o
from helpers import (TeV, ElectronIC, ElectronSynIC, lnprior_IC, lnprior_SynIC, oracle_data,
                     oracle_IC, oracle_lnprob_batch, oracle_stretch_sampler, oracle_SynIC,
                     rxj_tables)

pytestmark = pytest.mark.gpu

FLUX_RTOL = 1e-6
LNP_RTOL = 1e-8


@pytest.fixture(scope="module")
def nb():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import naima_b200
    from naima_b200 import _lib

    _lib.lib()  # fail loudly if the extension is missing
    return naima_b200


@pytest.fixture(params=["lean", "careful", "exact"])
def mode(request, nb):
    """The contraction's three cells: hoisted with the lean cell (default), hoisted with the
    careful cell everywhere, reference operation order.  Yields True for the exact mode."""
    from naima_b200 import engine

    old = engine.EXACT, engine.LEAN
    engine.EXACT = request.param == "exact"
    engine.LEAN = request.param == "lean"
    yield engine.EXACT
    engine.EXACT, engine.LEAN = old


ENERGY = np.logspace(0, 15, 1000)  # eV


def _dists(nb, amp=None):
    from naima_b200 import units as u
    from naima_b200.models import BrokenPowerLaw, ExponentialCutoffPowerLaw, PowerLaw

    amp = 1 / u.Unit("mec2") if amp is None else amp
    ECPL = ExponentialCutoffPowerLaw(amp, 20 * u.TeV, 2.0, 10 * u.TeV)
    PL = PowerLaw(amp, 20 * u.TeV, 2.0)
    BPL = BrokenPowerLaw(amp, 20 * u.TeV, 1 * u.TeV, 1.5, 2.5)
    return ECPL, PL, BPL


def _lum(nb, flux, energy):
    return nb.trapz_loglog(flux * energy, energy).to("erg/s").value


# ------------------------------------------------------------------------------
# 1. the reference's known-answer tests, replayed through the public API
# ------------------------------------------------------------------------------
def test_ref_synchrotron_lum(nb, goldens, mode):
    """tests/test_models.py:67-103"""
    from naima_b200 import units as u
    from naima_b200.models import Synchrotron

    energy = ENERGY * u.eV
    props = {"Eemin": 100 * u.GeV, "Eemax": 1 * u.PeV}
    lsys, Wes = [], []
    for pd in _dists(nb):
        sy = Synchrotron(pd, **props)
        Wes.append(sy.We.to("erg").value)
        lsys.append(_lum(nb, sy.flux(energy, 0), energy))
    assert_allclose(lsys, goldens["synchrotron_lum"]["value"], rtol=1e-7)
    assert_allclose(Wes, goldens["We"]["value"], rtol=1e-7)
    sy = Synchrotron(_dists(nb)[0], B=1 * u.G, **props)
    sy.flux({"energy": energy})
    assert_allclose(_lum(nb, sy.flux(energy, 0), energy), goldens["synchrotron_lum_B1G"]["value"],
                    rtol=1e-7)


def test_ref_inverse_compton_lum(nb, goldens, mode):
    """tests/test_models.py:197-226"""
    from naima_b200 import units as u
    from naima_b200.models import InverseCompton

    energy = ENERGY * u.eV
    props = {"Eemin": 100 * u.GeV, "Eemax": 1 * u.PeV}
    lums = [_lum(nb, InverseCompton(pd, **props).flux(energy, 0), energy) for pd in _dists(nb)]
    assert_allclose(lums, goldens["ic_lum"]["value"], rtol=1e-7)
    ic = InverseCompton(_dists(nb)[0], seed_photon_fields=["CMB", "FIR", "NIR"])
    assert_allclose(_lum(nb, ic.flux(energy, 0), energy), goldens["ic_lum_3seeds"]["value"],
                    rtol=1e-7)


def test_ref_anisotropic_ic_lum(nb, goldens, mode):
    """tests/test_models.py:229-252"""
    from naima_b200 import units as u
    from naima_b200.models import InverseCompton

    energy = ENERGY * u.eV
    PL = _dists(nb)[1]
    lums = []
    for ang in goldens["ic_aniso_lum"]["angles_deg"]:
        ic = InverseCompton(
            PL, seed_photon_fields=[["Star", 20000 * u.K, 0.1 * u.erg / u.cm**3, ang * u.deg]],
            Eemin=100 * u.GeV, Eemax=1 * u.PeV)
        lums.append(_lum(nb, ic.flux(energy, 0), energy))
    assert_allclose(lums, goldens["ic_aniso_lum"]["value"], rtol=1e-7)


def test_ref_monochromatic_ic(nb):
    """tests/test_models.py:254-287: grey body vs monochromatic vs tabulated seed"""
    from naima_b200 import units as u
    from naima_b200.models import InverseCompton, PowerLaw

    PL = PowerLaw(1 / u.eV, 1 * u.TeV, 3)
    Ephbb = np.logspace(-3.5, -1.5, 100)
    T, w = 30.0, 1.0 * u.eV / u.cm**3
    kT = 8.617333262145179e-05 * T
    hc = 1.2398419843320028e-04
    bbv = 8 * np.pi * Ephbb**2 / hc**3 / np.expm1(Ephbb / kT)
    Ebbmax = Ephbb[np.argmax(Ephbb**2 * bbv)] * u.eV
    bb = u.Quantity(bbv * (1.0 * o.eV_erg / (o.ar_cgs * T**4)), "1/(cm3 eV)")
    eopts = {"Eemax": 10000 * u.GeV, "Eemin": 10 * u.GeV, "nEed": 1000}
    Eb = Ephbb * u.eV
    IC_khang = InverseCompton(PL, seed_photon_fields=[["bb", T * u.K, w]], **eopts)
    IC_mono = InverseCompton(PL, seed_photon_fields=[["mono", Ebbmax, w]], **eopts)
    IC_bb = InverseCompton(PL, seed_photon_fields=[["bb2", Eb, bb]], **eopts)
    IC_bb_ene = InverseCompton(PL, seed_photon_fields=[["bb2", Eb, Eb**2 * bb]], **eopts)
    Eph = np.logspace(-1, 1, 3) * u.GeV
    ref = IC_khang.sed(Eph).value
    assert_allclose(ref, IC_mono.sed(Eph).value, rtol=1e-2)
    assert_allclose(ref, IC_bb.sed(Eph).value, rtol=1e-2)
    assert_allclose(ref, IC_bb_ene.sed(Eph).value, rtol=1e-2)
    # and against the oracle, tightly
    pd = o.PDist("PowerLaw", 1.0, 1 * TeV, 3.0)
    kw = dict(Eemin_eV=10e9, Eemax_eV=10000e9, nEed=1000)
    E = Eph.to("eV").value
    for ic, seed in [(IC_mono, ("mono", Ebbmax.to("eV").value, 1.0 * o.eV_erg)),
                     (IC_bb, ("array", Ephbb, bb.value))]:
        got = ic.flux(Eph, 0).value
        want = o.ic_spectrum(pd, E, [seed], **kw)
        assert_allclose(got, want, rtol=1e-10)


def test_ref_flux_sed(nb):
    """tests/test_models.py:290-324 distance scaling and SED identities"""
    from naima_b200 import units as u
    from naima_b200.models import InverseCompton

    energy = ENERGY[::10] * u.eV
    d1, d2 = 2.5 * u.kpc, 10.0 * u.kpc
    ic = InverseCompton(_dists(nb)[0], seed_photon_fields=["CMB", "FIR", "NIR"],
                        Eemin=100 * u.GeV, Eemax=1 * u.PeV)
    lum = _lum(nb, ic.flux(energy, 0), energy)
    f1 = nb.trapz_loglog(ic.flux(energy, d1) * energy, energy).to("erg/(s cm2)").value
    f2 = nb.trapz_loglog(ic.flux(energy, d2) * energy, energy).to("erg/(s cm2)").value
    assert_allclose(f1 / f2, 16.0)
    assert_allclose(f1, lum / (4 * np.pi * d1.to("cm").value ** 2))
    sed1 = ic.sed(energy, d1).to("erg/(s cm2)").value
    sed0 = (ic.flux(energy, 0) * energy**2).to("erg/s").value
    assert_allclose(sed1, sed0 / (4 * np.pi * d1.to("cm").value ** 2))


def test_ref_ic_seed_input_and_access(nb):
    """tests/test_models.py:327-387 seed parsing, per-seed flux, errors"""
    from naima_b200 import units as u
    from naima_b200.models import InverseCompton

    ECPL = _dists(nb)[0]
    for spf in ("CMB", ["CMB", "FIR", "NIR"], ["CMB", ["test", 5000 * u.K, 0]],
                ["CMB", ["test2", 5000 * u.K, 15 * u.eV / u.cm**3]]):
        ic = InverseCompton(ECPL, seed_photon_fields=spf)
        ic.flux(ENERGY[::50] * u.eV)
    ic = InverseCompton(ECPL, seed_photon_fields="CMB-FIR-NIR")
    assert list(ic.seed_photon_fields) == ["CMB", "FIR", "NIR"]
    ic = InverseCompton(
        ECPL, seed_photon_fields=["CMB", ["test", 5000 * u.K, 0],
                                  ["array", [1, 2] * u.eV, [1, 1] * u.Unit("1/(eV cm3)")]])
    ene = np.logspace(8, 13, 20) * u.eV
    tot = ic.flux(ene).value
    parts = [ic.flux(ene, seed=n).value for n in ("CMB", "test", "array")]
    assert_allclose(np.sum(parts, axis=0), tot, rtol=1e-14)
    for idx, name in enumerate(["CMB", "test", "array"]):
        assert_allclose(ic.flux(ene, seed=idx).value, ic.flux(ene, seed=name).value)
        assert_allclose(ic.sed(ene, seed=idx).value, ic.sed(ene, seed=name).value)
    with pytest.raises(ValueError):
        ic.flux(ene, seed="FIR")
    with pytest.raises(ValueError):
        ic.flux(ene, seed=10)
    with pytest.raises(TypeError):
        InverseCompton(ECPL, seed_photon_fields=["XYZ"])
    with pytest.raises(TypeError):
        InverseCompton(ECPL, seed_photon_fields=[["a", 5000 * u.K]])


def test_ref_bremsstrahlung_lum(nb, goldens, mode):
    """tests/test_models.py:178-194"""
    from naima_b200 import units as u
    from naima_b200.models import Bremsstrahlung, mec2

    energy2 = np.logspace(8, 14, 100) * u.eV
    brems = Bremsstrahlung(_dists(nb)[0], n0=1 * u.cm**-3, Eemin=mec2)
    assert_allclose(_lum(nb, brems.flux(energy2, 0), energy2),
                    goldens["bremsstrahlung_lum"]["value"], rtol=1e-7)


def test_ref_pion_decay(nb, goldens, mode):
    """tests/test_models.py:388-450 (goldens printed to 9 digits there)"""
    from naima_b200 import units as u
    from naima_b200.models import PionDecay

    energy = np.logspace(-3, 3, 60) * u.TeV
    Wps, lut, nolut = [], [], []
    for pd in _dists(nb, amp=1 / u.TeV):
        pp = PionDecay(pd, useLUT=True, Epmax=1 * u.PeV)
        Wps.append(pp.Wp.to("erg").value)
        lut.append(_lum(nb, pp.flux(energy, 0), energy))
        pp.useLUT = False
        nolut.append(_lum(nb, pp.flux(energy, 0), energy))
    assert_allclose(lut, goldens["pp_lum_LUT"]["value"], rtol=1e-7)
    assert_allclose(nolut, goldens["pp_lum_noLUT"]["value"], rtol=1e-7)
    assert_allclose(Wps, goldens["Wp"]["value"], rtol=1e-7)
    # LUT not found -> analytic (radiative.py:1484-1493)
    pp = PionDecay(_dists(nb, amp=1 / u.TeV)[1], useLUT=True, hiEmodel="Geant4", Epmax=1 * u.PeV)
    pp.flux(energy, 0)
    assert pp.useLUT is False
    energy = np.logspace(9, 13, 20) * u.eV
    pp = PionDecay(_dists(nb, amp=1 / u.TeV)[0], nuclear_enhancement=False, useLUT=False,
                   Epmax=1 * u.PeV)
    assert_allclose(_lum(nb, pp.flux(energy, 0), energy), goldens["pp_lum_no_nuc"]["value"],
                    rtol=1e-7)


def test_ref_compute_set_We(nb):
    """tests/test_models.py:122-177"""
    from naima_b200 import units as u
    from naima_b200.models import InverseCompton, PionDecay, Synchrotron

    ECPL, PL, BPL = _dists(nb)
    sy = Synchrotron(ECPL, B=1 * u.G, Eemin=100 * u.GeV, Eemax=1 * u.PeV)
    Eemin, Eemax = 10 * u.GeV, 100 * u.TeV
    sy.compute_We()
    sy.compute_We(Eemin=Eemin)
    sy.compute_We(Eemax=Eemax)
    sy.compute_We(Eemin=Eemin, Eemax=Eemax)
    assert sy.We.unit.physical_type == "energy"
    sy.set_We(1e48 * u.erg, Eemin=Eemin, Eemax=Eemax)
    assert_allclose(sy.compute_We(Eemin=Eemin, Eemax=Eemax).value, 1e48, rtol=1e-12)
    ic = InverseCompton(PL, Eemin=100 * u.GeV, Eemax=1 * u.PeV)
    ic.set_We(1e40 * u.erg)
    assert_allclose(ic.We.value, 1e40, rtol=1e-12)
    pp = PionDecay(_dists(nb, amp=1 / u.TeV)[1])
    pp.set_Wp(1e48 * u.erg, Epmin=10 * u.GeV, Epmax=100 * u.TeV)
    assert_allclose(pp.compute_Wp(Epmin=10 * u.GeV, Epmax=100 * u.TeV).value, 1e48, rtol=1e-12)


def test_ref_exec_vectors(nb, ref_exec):
    """Outputs of the reference's own functions (tests/golden/make_golden.py)."""
    from naima_b200.models import (BrokenPowerLaw, ExponentialCutoffBrokenPowerLaw,
                                   ExponentialCutoffPowerLaw, LogParabola, PowerLaw)

    r = ref_exec
    with np.errstate(all="ignore"):
        got = nb.trapz_loglog(r["tl_y"], r["tl_x"])
        assert_allclose(got, r["tl_sum"], rtol=1e-13)
        assert_allclose(nb.trapz_loglog(r["tl_y"], r["tl_x"], intervals=True), r["tl_int"],
                        rtol=2e-13)
        assert_allclose(nb.trapz_loglog(r["tl_y"].T.copy(), r["tl_x"], axis=0), r["tl_axis0"],
                        rtol=1e-13)
    e = r["pd_e"]
    rt = 1e-14
    assert_allclose(PowerLaw.eval(e, 1.3e33, 1e13, 2.41), r["pd_pl"], rtol=rt)
    assert_allclose(ExponentialCutoffPowerLaw.eval(e, 1.3e33, 1e13, 2.41, 4.8e13, 1.0),
                    r["pd_ecpl"], rtol=rt)
    # beta = 2: exp(-(e/ec)**2) turns a 1-ulp difference between the device pow and
    # glibc's into |(e/ec)**2| ulps of the result
    assert_allclose(ExponentialCutoffPowerLaw.eval(e, 1.3e33, 1e13, 1.7, 2e12, 2.0),
                    r["pd_ecpl_b2"], rtol=2e-13, atol=1e-300)
    assert_allclose(BrokenPowerLaw.eval(e, 2e30, 2e13, 1e12, 1.5, 2.5), r["pd_bpl"], rtol=rt)
    assert_allclose(ExponentialCutoffBrokenPowerLaw.eval(e, 3.7e36, 1e12, 2.65e11, 1.5, 3.233,
                                                         1.863e15, 2.0), r["pd_ecbpl"], rtol=rt)
    assert_allclose(LogParabola.eval(e, 1e30, 2e13, 1.7, 0.2), r["pd_lp"], rtol=rt)


# ------------------------------------------------------------------------------
# 2. per-energy flux parity against the oracle on random parameter draws
# ------------------------------------------------------------------------------
def _rand_pars(rng, W):
    return dict(amp=10 ** rng.uniform(30, 36, W), alpha=rng.uniform(1.5, 3.2, W),
                ecut=10 ** rng.uniform(0.3, 2.5, W), beta=rng.uniform(0.5, 2.0, W),
                B=10 ** rng.uniform(-6, -3.5, W))


@pytest.mark.parametrize("W", [1, 7, 64])
def test_flux_parity_synchrotron_ic(nb, mode, W):
    from naima_b200 import units as u
    from naima_b200.models import ExponentialCutoffPowerLaw, InverseCompton, Synchrotron

    rng = np.random.default_rng(100 + W)
    p = _rand_pars(rng, W)
    E = np.concatenate([np.logspace(2.7, 4, 9), np.logspace(11.5, 14.2, 12)])
    ECPL = ExponentialCutoffPowerLaw(p["amp"] / u.eV, 10 * u.TeV, p["alpha"], p["ecut"] * u.TeV,
                                     p["beta"])
    seeds = ["CMB", ["FIR", 26.5 * u.K, 0.415 * u.eV / u.cm**3], "NIR"]
    ic = InverseCompton(ECPL, seed_photon_fields=seeds, Eemin=100 * u.GeV)
    sy = Synchrotron(ECPL, B=p["B"] * u.G)
    fic = np.atleast_2d(ic.flux(E * u.eV, 1 * u.kpc).value)
    fsy = np.atleast_2d(sy.flux(E * u.eV, 1 * u.kpc).value)
    We = np.atleast_1d(ic.compute_We(Eemin=1 * u.TeV).value)
    assert fic.shape == (W, E.size)
    oseeds = ["CMB", ("thermal", 26.5, 0.415 * o.eV_erg), "NIR"]
    for w in range(W):
        pd = o.PDist("ExponentialCutoffPowerLaw", p["amp"][w], 10 * TeV, p["alpha"][w],
                     p["ecut"][w] * TeV, p["beta"][w])
        ric = o.flux_from_spectrum(o.ic_spectrum(pd, E, oseeds, Eemin_eV=100e9), o.kpc_cm)
        rsy = o.flux_from_spectrum(o.synchrotron_spectrum(pd, E, p["B"][w]), o.kpc_cm)
        big = rsy > 1e-250
        assert_allclose(fic[w], ric, rtol=FLUX_RTOL)
        assert_allclose(fsy[w][big], rsy[big], rtol=FLUX_RTOL)
        assert_allclose(We[w], o.compute_We(pd, 1e12, 1e9 * o.mec2_eV, 100), rtol=1e-12)


@pytest.mark.parametrize("kind", ["PowerLaw", "BrokenPowerLaw", "ExponentialCutoffBrokenPowerLaw",
                                  "LogParabola"])
def test_flux_parity_other_distributions(nb, mode, kind):
    from naima_b200 import models as M
    from naima_b200 import units as u

    rng = np.random.default_rng(7)
    W = 5
    amp = 10 ** rng.uniform(30, 36, W)
    a1, a2 = rng.uniform(1.2, 2.2, W), rng.uniform(2.3, 3.5, W)
    eb = 10 ** rng.uniform(-1, 1, W)
    if kind == "PowerLaw":
        pd = M.PowerLaw(amp / u.eV, 1 * u.TeV, a2)
        args = lambda w: (amp[w], 1 * TeV, a2[w])
    elif kind == "BrokenPowerLaw":
        pd = M.BrokenPowerLaw(amp / u.eV, 1 * u.TeV, eb * u.TeV, a1, a2)
        args = lambda w: (amp[w], 1 * TeV, eb[w] * TeV, a1[w], a2[w])
    elif kind == "LogParabola":
        pd = M.LogParabola(amp / u.eV, 1 * u.TeV, a1, 0.1 * a2)
        args = lambda w: (amp[w], 1 * TeV, a1[w], 0.1 * a2[w])
    else:
        pd = M.ExponentialCutoffBrokenPowerLaw(amp / u.eV, 1 * u.TeV, eb * u.TeV, a1, a2,
                                               100 * u.TeV, 2.0)
        args = lambda w: (amp[w], 1 * TeV, eb[w] * TeV, a1[w], a2[w], 100 * TeV, 2.0)
    E = np.logspace(9, 14, 15)
    ic = M.InverseCompton(pd, seed_photon_fields=["CMB"])
    got = ic.flux(E * u.eV, 0).value
    for w in range(W):
        want = o.ic_spectrum(o.PDist(kind, *args(w)), E, ["CMB"])
        assert_allclose(got[w], want, rtol=FLUX_RTOL)


def test_flux_parity_bremsstrahlung_piondecay(nb, mode):
    from naima_b200 import models as M
    from naima_b200 import units as u

    rng = np.random.default_rng(9)
    W = 4
    amp = 10 ** rng.uniform(30, 36, W)
    al = rng.uniform(1.8, 2.8, W)
    n0 = rng.uniform(0.1, 100, W)
    E = np.logspace(8, 13.5, 12)
    pdb = M.ExponentialCutoffPowerLaw(amp / u.eV, 1 * u.TeV, al, 30 * u.TeV)
    br = M.Bremsstrahlung(pdb, n0=n0 / u.cm**3)
    got = br.flux(E * u.eV, 0).value
    pdp = M.PowerLaw(amp / u.eV, 30 * u.TeV, al)
    pp = M.PionDecay(pdp, nh=n0 / u.cm**3)
    gpp = pp.flux(E * u.eV, 0).value
    pp2 = M.PionDecay(pdp, nh=n0 / u.cm**3, useLUT=False, hiEmodel="SIBYLL")
    gpp2 = pp2.flux(E * u.eV, 0).value
    f = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "..",
                                           "naima_b200", "data",
                                           "pp_kafexhiu14_pythia8_nucenh_bspline.npz"))
    from scipy.interpolate import bisplev

    tck = (f["tx"], f["ty"], f["c"], 3, 3)
    lut = lambda Ep, Eg: np.atleast_1d(bisplev(np.log10(Ep), np.log10(Eg), tck)).flatten()
    for w in range(W):
        opd = o.PDist("ExponentialCutoffPowerLaw", amp[w], 1 * TeV, al[w], 30 * TeV, 1.0)
        assert_allclose(got[w], o.bremsstrahlung_spectrum(opd, E, n0=n0[w]), rtol=FLUX_RTOL)
        opp = o.PDist("PowerLaw", amp[w], 30 * TeV, al[w])
        assert_allclose(gpp[w], o.piondecay_spectrum(opp, E, nh=n0[w], lut=lut), rtol=FLUX_RTOL)
        assert_allclose(gpp2[w], o.piondecay_spectrum(opp, E, nh=n0[w], useLUT=False,
                                                      hiEmodel="SIBYLL"), rtol=FLUX_RTOL)


def test_flux_parity_ssc(nb):
    """examples/CrabNebula_SynSSC.py:13-51 shape (per-walker tabulated seed), small grid."""
    from naima_b200 import models as M
    from naima_b200 import units as u

    W = 3
    rng = np.random.default_rng(3)
    amp = 3.699e36 * (1 + 0.1 * rng.normal(size=W))
    B = 125e-6 * (1 + 0.1 * rng.normal(size=W))
    kw = dict(Eemax=50 * u.PeV, Eemin=0.1 * u.GeV, nEed=20)
    ECBPL = M.ExponentialCutoffBrokenPowerLaw(amp / u.eV, 1 * u.TeV, 0.265 * u.TeV, 1.5, 3.233,
                                              1863 * u.TeV, 2.0)
    SYN = M.Synchrotron(ECBPL, B=B * u.G, **kw)
    Rpwn = 2.1 * 3.0856775814913673e18
    Esy = np.logspace(-7, 9, 40)
    Lsy = SYN.flux(Esy * u.eV, distance=0 * u.cm)
    phn_sy = Lsy / (4 * np.pi * (Rpwn * u.cm) ** 2 * (o.c_cgs * u.cm / u.s)) * 2.24
    IC = M.InverseCompton(
        ECBPL, seed_photon_fields=["CMB", ["FIR", 70 * u.K, 0.5 * u.eV / u.cm**3],
                                   ["SSC", Esy * u.eV, phn_sy]], **kw)
    E = np.logspace(7, 14, 11)
    got = IC.flux(E * u.eV, 2 * u.kpc).value
    for w in range(W):
        pd = o.PDist("ExponentialCutoffBrokenPowerLaw", amp[w], 1 * TeV, 0.265 * TeV, 1.5, 3.233,
                     1863 * TeV, 2.0)
        okw = dict(Eemin_eV=1e8, Eemax_eV=50e15, nEed=20)
        lsy = o.synchrotron_spectrum(pd, Esy, B[w], **okw)
        assert_allclose(Lsy.value[w], lsy, rtol=FLUX_RTOL)
        phn = lsy / (4 * np.pi * Rpwn**2 * o.c_cgs) * 2.24
        seeds = ["CMB", ("thermal", 70.0, 0.5 * o.eV_erg), ("array", Esy, phn)]
        want = o.flux_from_spectrum(o.ic_spectrum(pd, E, seeds, **okw), 2 * o.kpc_cm)
        assert_allclose(got[w], want, rtol=FLUX_RTOL)


def test_ref_exec_unit_bound_vectors(nb, ref_units, mode):
    """The CUDA path against per-energy outputs of the reference's own unit-bound functions
    (tests/golden/make_golden_units.py): Synchrotron._spectrum, _calc_specic on
    monochromatic / tabulated / grey-body seeds, Bremsstrahlung._spectrum."""
    from naima_b200 import models as M
    from naima_b200 import units as u

    r = ref_units

    def close(got, want, rtol=FLUX_RTOL):
        live = want > np.max(want) * 1e-250
        assert live.sum() > 0.4 * want.size
        assert_allclose(np.asarray(got)[live], want[live], rtol=rtol)
        assert np.all(np.abs(np.asarray(got)[~live]) <= np.max(want) * 1e-240)

    ecpl = M.ExponentialCutoffPowerLaw(1.3e33 / u.eV, 1e13 * u.eV, 2.41, 4.8e13 * u.eV, 1.0)
    bpl = M.BrokenPowerLaw(2e30 / u.eV, 2e13 * u.eV, 1e12 * u.eV, 1.5, 2.5)
    E = r["syn_E_eV"] * u.eV
    for tag, pd in (("ecpl", ecpl), ("bpl", bpl)):
        for Bn, B in (("3uG", 3.24e-6), ("1mG", 1e-3)):
            got = M.Synchrotron(pd, B=B * u.G).flux(E, distance=0).value
            close(got, r["syn_spec_%s_%s" % (tag, Bn)])
    E = r["ic_E_eV"] * u.eV
    kw = dict(Eemin=1e11 * u.eV, Eemax=1e15 * u.eV, nEed=60)
    seeds = [["mono", 0.00235 * u.eV, 0.261 * u.eV / u.cm**3],
             ["tab", r["icm_seed_E_eV"] * u.eV, u.Quantity(r["icm_seed_n"], "1/(eV cm3)")],
             ["FIR", 26.5 * u.K, 0.415 * u.eV / u.cm**3],
             ["star", 25000 * u.K, 3.0 * u.eV / u.cm**3, 2.1 * u.rad]]
    ic = M.InverseCompton(ecpl, seed_photon_fields=seeds, **kw)
    ic.flux(E, distance=0)
    for k, sd in enumerate(seeds):
        close(ic.specic[k].value, r["ic_specic_" + sd[0]])
    # the tabulated seed with a per-walker density (the hoisted self-Compton kernels)
    if not mode:
        n2 = np.vstack([r["icm_seed_n"], 2.0 * r["icm_seed_n"]])
        ic2 = M.InverseCompton(ecpl, seed_photon_fields=[
            ["tab", r["icm_seed_E_eV"] * u.eV, u.Quantity(n2, "1/(eV cm3)")]], **kw)
        got = ic2.flux(E, distance=0).value
        close(got[0], r["ic_specic_tab"])
        close(got[1], 2.0 * r["ic_specic_tab"])
    br = M.Bremsstrahlung(ecpl, n0=3.0 / u.cm**3, Eemin=1e8 * u.eV, nEed=40)
    close(br.flux(r["br_E_eV"] * u.eV, distance=0).value, r["br_spec"])


# ------------------------------------------------------------------------------
# 3. likelihood
# ------------------------------------------------------------------------------
def test_lnprob_known_answer(nb, goldens, mode):
    """docs/_static/RXJ1713_IC_results.ecsv:10-11 through lnprob() and the plan."""
    from naima_b200 import units as u
    from naima_b200.models import ExponentialCutoffPowerLaw, InverseCompton

    _, hess = rxj_tables()
    data = nb.validate_data_table(hess)

    def model(pars, data):  # docs/_static/RXJ1713_IC.py:16-62 (default Eemin)
        ECPL = ExponentialCutoffPowerLaw(pars[0] / u.eV, 10.0 * u.TeV, pars[1],
                                         (10 ** pars[2]) * u.TeV)
        IC = InverseCompton(
            ECPL, seed_photon_fields=["CMB", ["FIR", 26.5 * u.K, 0.415 * u.eV / u.cm**3]])
        return IC.flux(data, distance=1.0 * u.kpc).to(data["flux"].unit)

    g = goldens["lnprob_RXJ1713_IC"]
    pars = np.array(g["ML_pars"])
    lp = nb.lnprob(pars, data, model, lnprior_IC)[0]
    assert_allclose(lp, g["MaxLogLikelihood"], rtol=LNP_RTOL)
    plan = nb.LikelihoodPlan(model, lnprior_IC, data, 3)
    lnp, flux, _ = plan(pars[None, :])
    assert_allclose(lnp[0], g["MaxLogLikelihood"], rtol=LNP_RTOL)


@pytest.mark.parametrize("case", ["IC", "SynIC"])
def test_lnprob_parity_paths(nb, mode, case):
    """Plan (traced), batched callbacks and per-walker calls vs the oracle's lnprob
    on the shipped example shapes (C2/C3), incl. walkers outside the prior."""
    suz, hess = rxj_tables()
    rng = np.random.default_rng(20261017)
    W = 24
    if case == "IC":
        data = nb.validate_data_table(hess)
        model, prior = ElectronIC, lnprior_IC
        p_true = np.array([1.37e32, 2.58, np.log10(50.2)])
        omodel, oprior = oracle_IC()
    else:
        data = nb.validate_data_table([suz, hess])
        model, prior = ElectronSynIC, lnprior_SynIC
        p_true = np.array([33.0, 2.5, np.log10(48.0), 20.0])
        omodel, oprior = oracle_SynIC()
    P = p_true * (1 + 0.05 * rng.normal(size=(W, p_true.size)))
    P[3, 1] = 7.0   # outside the prior: lnprob = -inf, model still evaluated
    P[5, 0] = -1.0 if case == "IC" else P[5, 0]
    od = oracle_data(data)
    want, wflux = oracle_lnprob_batch(P, od, omodel, oprior)
    assert np.isinf(want[3])
    plan = nb.LikelihoodPlan(model, prior, data, p_true.size)
    lnp, flux, blob_arrays = plan(P)
    fin = np.isfinite(want)
    assert np.array_equal(np.isfinite(lnp), fin)
    assert_allclose(lnp[fin], want[fin], rtol=LNP_RTOL)
    ok = np.isfinite(wflux).all(axis=1)
    assert_allclose(flux[ok], wflux[ok], rtol=FLUX_RTOL)
    # batched callbacks
    lnp2, blobs2 = nb.lnprob(P, data, model, prior)
    assert_allclose(lnp2[fin], want[fin], rtol=LNP_RTOL)
    assert np.all(np.isneginf(lnp2[~fin]))
    # per-walker calls (the reference's calling convention)
    for w in (0, 3, 11):
        r = nb.lnprob(P[w], data, model, prior)
        if fin[w]:
            assert_allclose(r[0], want[w], rtol=LNP_RTOL)
        else:
            assert r[0] == -np.inf
        # blob parity between the three paths: model flux and We.  The plan's
        # self-contained synchrotron kernel derives ln(e/e0) from the grid's ln x table, the
        # class path from log(e/e0): few-ulp operand differences, amplified by 1/|b + 1|
        # where the integrand slope b -> -1 (conditioning note)
        b_plan = plan.blobs_for(flux, blob_arrays, w)
        assert_allclose(b_plan[0].value, r[1].value, rtol=1e-10)
        assert_allclose(blobs2[w][0].value, r[1].value, rtol=1e-12)
        assert_allclose(b_plan[-1].value, r[-1].value, rtol=1e-12)
        if case == "IC":
            assert_allclose(b_plan[1][1].value, r[2][1].value, rtol=1e-12)
            pd = o.PDist("ExponentialCutoffPowerLaw", P[w, 0], 10 * TeV, P[w, 1],
                         10 ** P[w, 2] * TeV, 1.0)
            assert_allclose(r[3].value, o.compute_We(pd, 1e12, 1e9 * o.mec2_eV, 100), rtol=1e-12)


def test_lnprobmodel_edge_cases(nb):
    """Asymmetric errors, several upper limits (violated / not), all-UL, no-UL."""
    from naima_b200 import units as u

    rng = np.random.default_rng(5)
    N = 17
    E = np.logspace(-1, 2, N)
    f = 1e-11 * E**-2.2
    # (all upper limits AND all violated indexes cl[N]: IndexError in the reference,
    # core.py:92 -- the batched device path returns NaN for that walker, checked below)
    for ul_idx, scale in [([], 1.0), ([2, 9], 0.5), ([2, 9], 3.0), (list(range(N)), 1.0)]:
        ul = np.zeros(N, dtype=int)
        ul[ul_idx] = 1
        t = nb.DataTable(meta={"keywords": {"cl": {"value": 0.99}}})
        t["energy"] = E * u.TeV
        t["flux"] = u.Quantity(f, "1/(cm2 s TeV)")
        t["flux_error_lo"] = u.Quantity(0.1 * f, "1/(cm2 s TeV)")
        t["flux_error_hi"] = u.Quantity(0.25 * f, "1/(cm2 s TeV)")
        t["ul"] = ul
        data = nb.validate_data_table(t)
        model = f * scale * np.exp(0.2 * rng.normal(size=(6, N)))
        got = nb.lnprobmodel(u.Quantity(model, "1/(cm2 s TeV)"), data)
        od = dict(flux=f, flux_error_lo=0.1 * f, flux_error_hi=0.25 * f, ul=ul.astype(bool),
                  cl=np.full(N, 0.99))
        want = [o.lnprobmodel(m, od) for m in model]
        assert_allclose(got, want, rtol=1e-13, atol=1e-300)
        # SED-valued model against differential data (core.py:66-71)
        sed = (u.Quantity(model[0], "1/(cm2 s TeV)") * (E * u.TeV) ** 2).to("erg/(cm2 s)")
        assert_allclose(nb.lnprobmodel(sed, data), want[0], rtol=1e-10)
    # last table: every point is an upper limit; a model above all of them
    assert np.isnan(nb.lnprobmodel(u.Quantity(10 * f, "1/(cm2 s TeV)"), data))
    with pytest.raises(IndexError):
        o.lnprobmodel(10 * f, od)


def test_priors(nb):
    assert nb.uniform_prior(1.0, 0, 2) == 0.0 and nb.uniform_prior(3.0, 0, 2) == -np.inf
    assert_allclose(nb.normal_prior(1.2, 1.0, 0.5), o.normal_prior(1.2, 1.0, 0.5))
    assert nb.log_uniform_prior(2.0, 1, 4) == 0.5 and nb.log_uniform_prior(5.0, 1, 4) == -np.inf
    assert nb.log_uniform_prior(-1.0) == -np.inf and nb.log_uniform_prior(4.0) == 0.25
    v = np.array([-1.0, 0.5, 2.0, 5.0])
    assert_allclose(nb.uniform_prior(v, 0, 2), [o.uniform_prior(x, 0, 2) for x in v])
    assert_allclose(nb.log_uniform_prior(v, 1, 4), [o.log_uniform_prior(x, 1, 4) for x in v])
    # the device prior kernel through a traced plan
    suz, hess = rxj_tables()
    data = nb.validate_data_table(hess)

    def prior(pars):
        return (nb.uniform_prior(pars[0], 0, np.inf) + nb.normal_prior(pars[1], 2.5, 0.3)
                + nb.log_uniform_prior(pars[2], 0.5, 3.0))

    def oprior(p):
        return (o.uniform_prior(p[0], 0, np.inf) + o.normal_prior(p[1], 2.5, 0.3)
                + o.log_uniform_prior(p[2], 0.5, 3.0))

    plan = nb.LikelihoodPlan(ElectronIC, prior, data, 3)
    P = np.array([[1.37e32, 2.58, 1.7], [1.2e32, 2.2, 0.4], [-1e32, 2.2, 1.0],
                  [1.5e32, 2.9, 2.9]])
    lnp, _, _ = plan(P)
    omodel, _ = oracle_IC()
    want, _ = oracle_lnprob_batch(P, oracle_data(data), omodel, oprior)
    fin = np.isfinite(want)
    assert np.array_equal(np.isfinite(lnp), fin)
    assert_allclose(lnp[fin], want[fin], rtol=LNP_RTOL)


@pytest.mark.parametrize("kinds", [(), ("syn",)], ids=["operand-arrays", "syn-selfprep"])
def test_selfprep_kernels_vs_oracle(nb, kinds):
    """nb_synchrotron_fused (operands derived inside the kernel) and nb_synchrotron (operand
    arrays of the set-up kernel) against the oracle, through the plan."""
    suz, hess = rxj_tables()
    data = nb.validate_data_table([suz, hess])
    rng = np.random.default_rng(11)
    p_true = np.array([33.0, 2.5, np.log10(48.0), 20.0])
    P = p_true * (1 + 0.1 * rng.normal(size=(24, 4)))
    plan = nb.LikelihoodPlan(ElectronSynIC, lnprior_SynIC, data, 4)
    for c in plan.comps:
        c["selfprep"] = c["kind"] in kinds
    lnp, flux, blobs = plan(P)
    omodel, oprior = oracle_SynIC()
    want, wflux = oracle_lnprob_batch(P, oracle_data(data), omodel, oprior)
    assert_allclose(lnp, want, rtol=LNP_RTOL)
    assert_allclose(flux, wflux, rtol=FLUX_RTOL)
    # broken power law with cutoff + log-parabola through the same kernels
    from naima_b200 import units as u
    from naima_b200.models import (ExponentialCutoffBrokenPowerLaw, InverseCompton, LogParabola,
                                   Synchrotron)

    def model_bpl(pars, data):
        pd = ExponentialCutoffBrokenPowerLaw(10 ** pars[0] / u.eV, 1 * u.TeV, 3 * u.TeV, pars[1],
                                             3.2, (10 ** pars[2]) * u.TeV, 2.0)
        return (InverseCompton(pd, seed_photon_fields=["CMB"], Eemin=100 * u.GeV).flux(data)
                + Synchrotron(pd, B=pars[3] * u.uG).flux(data))

    def model_lp(pars, data):
        pd = LogParabola(10 ** pars[0] / u.eV, 10 * u.TeV, pars[1], 0.15)
        return (InverseCompton(pd, seed_photon_fields=["CMB"], Eemin=100 * u.GeV).flux(data)
                + Synchrotron(pd, B=pars[3] * u.uG).flux(data))

    od = oracle_data(data)
    for model, kind, args in (
            (model_bpl, "ExponentialCutoffBrokenPowerLaw",
             lambda p: (10 ** p[0], 1 * TeV, 3 * TeV, p[1], 3.2, 10 ** p[2] * TeV, 2.0)),
            (model_lp, "LogParabola", lambda p: (10 ** p[0], 10 * TeV, p[1], 0.15))):
        plan2 = nb.LikelihoodPlan(model, None, data, 4)
        for c in plan2.comps:
            c["selfprep"] = c["kind"] in kinds
        _, f2, _ = plan2(P[:6])
        for w in range(6):
            pd = o.PDist(kind, *args(P[w]))
            ref = (o.flux_from_spectrum(o.ic_spectrum(pd, od["E_eV"], ["CMB"], Eemin_eV=100e9),
                                        o.kpc_cm)
                   + o.flux_from_spectrum(o.synchrotron_spectrum(pd, od["E_eV"], P[w, 3] * 1e-6),
                                          o.kpc_cm)) * od["unit_fac"]
            assert_allclose(f2[w], ref, rtol=FLUX_RTOL)


# ------------------------------------------------------------------------------
# 4. samplers
# ------------------------------------------------------------------------------
def test_sampler_chain_parity(nb):
    """Host-driven sampler over the plan, the device-resident ensemble and the
    oracle-driven NumPy stretch move, same seed: identical accept decisions,
    chains equal to rounding."""
    suz, hess = rxj_tables()
    data = nb.validate_data_table([suz, hess])
    W, P, nsteps, seed = 16, 4, 6, 1234
    rng = np.random.default_rng(1)
    p_true = np.array([33.0, 2.5, np.log10(48.0), 20.0])
    p0 = p_true * (1 + 0.02 * rng.normal(size=(W, P)))
    plan = nb.LikelihoodPlan(ElectronSynIC, lnprior_SynIC, data, P)
    from naima_b200.core import PlanLogProb

    s = nb.EnsembleSampler(W, P, PlanLogProb(plan), vectorize=True, seed=seed)
    s.run_mcmc(p0, nsteps)
    omodel, oprior = oracle_SynIC()
    od = oracle_data(data)
    chain, lps = oracle_stretch_sampler(
        lambda q: oracle_lnprob_batch(q, od, omodel, oprior)[0], p0, nsteps, seed)
    assert_allclose(s.get_chain(), chain, rtol=1e-9)
    assert_allclose(s.get_log_prob(), lps, rtol=LNP_RTOL)
    assert s.get_blobs().shape == (nsteps, W)
    assert_allclose(s.get_blobs()[-1, 3][0].value / plan.to_model_unit,
                    oracle_lnprob_batch(chain[-1, 3:4], od, omodel, oprior)[1][0],
                    rtol=FLUX_RTOL)
    de = nb.DeviceEnsemble(plan, W, seed=seed)
    de.set_state(p0)
    dchain, dlp, dblobs = de.run(nsteps)
    assert_allclose(dchain, s.get_chain(), rtol=1e-13)
    assert_allclose(dlp, s.get_log_prob(), rtol=1e-13)
    assert np.array_equal(de.acceptance_counts, (s.acceptance_fraction * nsteps).round())
    # a second block continues the same stream
    s.run_mcmc(None, 3)
    dchain2, _, _ = de.run(3)
    assert_allclose(dchain2, s.get_chain()[nsteps:], rtol=1e-13)
    # the public-API sampler over the device loop (what get_sampler builds): same chain,
    # same blobs, same acceptance counts, for block sizes that do / do not divide nsteps
    # ... and for several steps per CUDA graph (steps that do not fill a graph are launched
    # kernel by kernel)
    for block, spg in ((4, 1), (16, 1), (8, 4), (6, 4)):
        ps = nb.PlanSampler(W, P, plan, seed=seed, block=block, chunk=5, steps_per_graph=spg)
        st = ps.run_mcmc(p0, nsteps)
        assert_allclose(ps.get_chain(), s.get_chain()[:nsteps], rtol=1e-13)
        assert_allclose(ps.get_log_prob(), s.get_log_prob()[:nsteps], rtol=1e-13)
        assert_allclose(st.coords, s.get_chain()[nsteps - 1], rtol=1e-13)
        bp, bs = ps.get_blobs()[-1, 5], s.get_blobs()[nsteps - 1, 5]
        assert len(bp) == len(bs) == 2
        assert_allclose(bp[0].value, bs[0].value, rtol=1e-13)
        assert_allclose(bp[1].value, bs[1].value, rtol=1e-13)
        ps.run_mcmc(None, 3)
        assert_allclose(ps.get_chain()[nsteps:], s.get_chain()[nsteps:], rtol=1e-13)
        assert np.array_equal(ps.acceptance_fraction, s.acceptance_fraction)
    # stopping the generator early rewinds the random stream to the last yielded step
    ps = nb.PlanSampler(W, P, plan, seed=seed, block=4)
    gen = ps.sample(p0, iterations=nsteps)
    for _ in range(2):
        state = next(gen)
    gen.close()
    ps.run_mcmc(state, nsteps - 2)
    assert_allclose(ps.get_chain()[:nsteps], s.get_chain()[:nsteps], rtol=1e-13)


def test_get_sampler_run_sampler(nb):
    """tests/test_functionfit.py shapes: API surface of get_sampler/run_sampler."""
    from naima_b200 import units as u

    _, hess = rxj_tables()
    p0 = np.array((1e30, 3.0, np.log10(30)))
    labels = ["norm", "index", "log10(cutoff)"]
    sampler, pos = nb.run_sampler(data_table=hess, p0=p0, labels=labels, model=ElectronIC,
                                  prior=lnprior_IC, nwalkers=10, nburn=2, nrun=3, threads=1,
                                  seed=3)
    assert sampler.plan is not None
    assert sampler.get_chain().shape == (3, 10, 3)
    assert sampler.get_log_prob().shape == (3, 10)
    blobs = sampler.get_blobs()
    assert blobs.shape == (3, 10)
    b = blobs[-1, 0]
    assert len(b) == 3 and b[0].unit.physical_type == "differential flux"
    assert b[1][0].shape == (100,) and b[1][1].shape == (100,)
    assert b[2].unit.physical_type == "energy"
    for key in ("n_walkers", "n_burn", "p0", "guess", "p0_burn_median", "n_run"):
        assert key in sampler.run_info
    assert sampler.labels == labels and sampler.modelfn is ElectronIC
    assert np.all((sampler.acceptance_fraction >= 0) & (sampler.acceptance_fraction <= 1))
    # output-side callers: run archive round trip, ML point, posterior recompute
    import tempfile

    with tempfile.TemporaryDirectory() as tmp:
        fn = os.path.join(tmp, "run.npz")
        nb.save_run(fn, sampler)
        res = nb.read_run(fn, modelfn=ElectronIC)
    assert np.array_equal(res.get_chain(), sampler.get_chain())
    assert np.array_equal(res.get_log_prob(), sampler.get_log_prob())
    rb = res.get_blobs()[-1, 0]
    assert_allclose(rb[0].value, b[0].value, rtol=0)
    assert rb[0].unit.physical_type == b[0].unit.physical_type
    assert_allclose(rb[1][1].value, b[1][1].value, rtol=0)
    assert res.labels == labels and res.run_info["n_run"] == 3
    ML, MLp, MLerr, (mx, my) = nb.find_ML(sampler, 0)
    assert ML == sampler.get_log_prob().max() and my.shape == (28,)
    E, m = nb.model_samples(res, [0.1, 100] * u.TeV, e_npoints=17, n_samples=12, seed=2)
    assert m.shape == (12, 17) and m.unit.physical_type == "differential flux"
    pars = res.get_chain(flat=True)[np.random.RandomState(2).randint(30, size=12)]
    one = ElectronIC(pars[5], {"energy": E, "flux": hess["flux"][:1]})[0]
    assert_allclose(m.value[5], one.to(m.unit).value, rtol=1e-10)
    # ... and against the oracle: the reference evaluates sampler.modelfn(p, bogus data) per
    # sample on the new grid (plot.py:372-391); the same with the oracle's model function
    omodel, _ = oracle_IC()
    E_eV = E.to("eV").value
    od = {"E_eV": E_eV, "unit_fac": u.Quantity(1.0, "1/(s cm2 eV)").to(m.unit).value}
    want = np.array([omodel(p_, od) for p_ in pars])
    assert_allclose(m.value, want, rtol=1e-6)
    # find_ML's model (plot.py:396-430 / analysis.py find_ML) is the maximum-likelihood
    # walker's blob: the oracle on the data energies
    od2 = oracle_data(nb.validate_data_table(hess))
    assert_allclose(my.to(hess["flux"].unit).value, omodel(MLp, od2), rtol=1e-6)
    # continue the run; non-traceable callbacks; per-walker mode; prefit
    sampler, pos = nb.run_sampler(nrun=2, sampler=sampler, pos=pos)
    assert sampler.get_chain().shape == (2, 10, 3)

    def model_untraceable(pars, data):
        # np.log/np.exp on a free parameter cannot be traced, but works on batches
        return ElectronIC([pars[0], np.log(np.exp(pars[1])), pars[2]], data)[0]

    s2, pos2 = nb.get_sampler(data_table=hess, p0=p0, labels=labels, model=model_untraceable,
                              prior=lnprior_IC, nwalkers=8, nburn=1, seed=3)
    assert s2.plan is None and s2.vectorize
    s3, pos3 = nb.get_sampler(data_table=hess, p0=p0, labels=labels, model=ElectronIC,
                              prior=lnprior_IC, nwalkers=8, nburn=1, seed=3, vectorize=False,
                              prefit=True)
    assert not s3.vectorize and s3.get_chain().shape == (1, 8, 3)
    with pytest.raises(TypeError):
        nb.get_sampler(p0=p0, model=ElectronIC)
    with pytest.raises(TypeError):
        nb.get_sampler(data_table=hess, p0=p0)
    with pytest.raises(ValueError):  # fewer walkers than 2 x ndim
        nb.get_sampler(data_table=hess, p0=p0, labels=labels, model=ElectronIC, nwalkers=4,
                       nburn=1)


# ------------------------------------------------------------------------------
# 5. size-independent properties at the benchmark's full size
# ------------------------------------------------------------------------------
def test_full_size_properties(nb):
    """BASELINE C3 shape (256 walkers, N_E = 64, 3 seeds): linearity in the
    amplitude, additivity over seeds, batch-vs-single invariance, permutation
    invariance, idempotence of graph replay."""
    from naima_b200 import units as u
    from naima_b200.models import ExponentialCutoffPowerLaw, InverseCompton

    suz, hess = rxj_tables()
    data = nb.validate_data_table([suz, hess])
    W, P = 256, 4
    rng = np.random.default_rng(20261017)
    p_true = np.array([33.0, 2.5, np.log10(48.0), 20.0])
    X = p_true * (1 + 0.1 * rng.normal(size=(W, P)))
    X[:, 3] = np.abs(X[:, 3])
    plan = nb.LikelihoodPlan(ElectronSynIC, lnprior_SynIC, data, P)
    lnp, flux, _ = plan(X)
    lnp_b, flux_b, _ = plan(X)
    assert np.array_equal(lnp, lnp_b) and np.array_equal(flux, flux_b)
    perm = rng.permutation(W)
    lnp_p, flux_p, _ = plan(X[perm])
    assert np.array_equal(lnp_p, lnp[perm]) and np.array_equal(flux_p, flux[perm])
    l1, f1, _ = plan(X[:1])
    assert np.array_equal(l1[0], lnp[0]) and np.array_equal(f1[0], flux[0])
    X2 = X.copy()
    X2[:, 0] += np.log10(2.0)
    _, flux2, _ = plan(X2)
    # not 1e-13: where the integrand slope b -> -1 the trapezoid (x2 y2 - x1 y1)/(b + 1)
    # amplifies rounding by 1/|b + 1| (in the reference's formula too, utils.py:341-343)
    assert_allclose(flux2, 2 * flux, rtol=1e-10)
    E = u.Quantity(data["energy"])
    ECPL = ExponentialCutoffPowerLaw(10 ** X[:, 0] / u.eV, 10 * u.TeV, X[:, 1],
                                     10 ** X[:, 2] * u.TeV)
    seeds = ["CMB", "FIR", "NIR"]
    tot = InverseCompton(ECPL, seed_photon_fields=seeds, Eemin=100 * u.GeV).flux(E).value
    parts = sum(InverseCompton(ECPL, seed_photon_fields=[s], Eemin=100 * u.GeV).flux(E).value
                for s in seeds)
    assert_allclose(parts, tot, rtol=1e-13)


# ------------------------------------------------------------------------------
# 6. C ABI edge cases
# ------------------------------------------------------------------------------
def test_cabi_edge_cases(nb):
    import ctypes

    import torch

    from naima_b200 import engine as eng
    from naima_b200._lib import lib

    L = lib()
    assert L.nb_version() >= 100
    g = eng.electron_grid(1e11, 1e15, 100)
    par = eng.to_dev(eng.pd_params_array("PowerLaw", [1e30, 1e13, 2.2]))
    out = eng.empty(1, g.pitch)
    st = eng.stream()
    # W == 0 is a no-op
    assert L.nb_pd_prep(0, eng.ptr(par), 0, eng.ptr(g.x_d), g.N, 1.0, 1.0, 1.0,
                        eng.ptr(g.invdlx_d), eng.ptr(out), eng.ptr(out), g.pitch, st) == 0
    # null pointers / bad kind / bad sizes -> NB_EINVAL, not a crash
    assert L.nb_pd_prep(0, None, 1, eng.ptr(g.x_d), g.N, 1.0, 1.0, 1.0, eng.ptr(g.invdlx_d),
                        eng.ptr(out), eng.ptr(out), g.pitch, st) == -1
    assert L.nb_pd_prep(9, eng.ptr(par), 1, eng.ptr(g.x_d), g.N, 1.0, 1.0, 1.0,
                        eng.ptr(g.invdlx_d), eng.ptr(out), eng.ptr(out), g.pitch, st) == -1
    assert L.nb_pdist_eval(0, eng.ptr(par), 1, eng.ptr(g.x_d), -1, eng.ptr(out), st) == -1
    assert L.nb_contract(None, None, 1, 10, 10, 0, None, None, 10, 1, None, None, None, None, 0,
                         st) == -1
    # odd pitch violates the TMA alignment rule
    K = eng.zeros(4, 11)
    assert L.nb_contract(eng.ptr(K), eng.ptr(K), 4, 11, 11, 0, eng.ptr(K), eng.ptr(K), 11, 1,
                         eng.ptr(K), eng.ptr(K), None, eng.ptr(out), 0, st) == -3
    assert L.nb_strerror(-3).decode().startswith("alignment")
    # a grid too large for the shared-memory tiling is refused
    big = 20000
    Kb = eng.zeros(2, big)
    assert L.nb_contract(eng.ptr(Kb), eng.ptr(Kb), 2, big, big, 0, eng.ptr(Kb), eng.ptr(Kb), big,
                         1, eng.ptr(Kb), eng.ptr(Kb), None, eng.ptr(out), 0, st) == -2
    # minimum sizes: N_E = 1, 10-node grid (the reference's floor, radiative.py:152-154)
    from naima_b200 import units as u
    from naima_b200.models import InverseCompton, PowerLaw

    ic = InverseCompton(PowerLaw(1e30 / u.eV, 1 * u.TeV, 2.1), Eemin=1 * u.TeV, Eemax=1.2 * u.TeV)
    assert ic._gam.size == 10
    got = ic.flux(1 * u.GeV, 0)
    want = o.ic_spectrum(o.PDist("PowerLaw", 1e30, 1e12, 2.1), [1e9], ["CMB"], 1e12, 1.2e12)
    assert np.ndim(got.value) == 0
    assert_allclose(got.value, want[0], rtol=FLUX_RTOL)
    torch.cuda.synchronize()
