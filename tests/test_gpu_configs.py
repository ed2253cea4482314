"""GPU parity tests at the shapes of BASELINE.json's five configurations (C1..C5), through
the device-resident likelihood plan and the public sampler -- the CUDA path against the CPU
oracle (oracle/naima_oracle.py via oracle/bench_models.py) on the same seeded inputs.

Tolerances (north_star): flux rtol 1e-6, lnprob rtol 1e-8.  Where the oracle is too slow for
the full walker count (C4: 1.5 s per walker) the full-size launch is checked against a
small-batch launch of the same rows BITWISE (results do not depend on the batch) and the
small batch against the oracle.
"""
import os
import sys

import numpy as np
import pytest
from numpy.testing import assert_allclose

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench_workloads as wl  # noqa: E402
import oracle.naima_oracle as o  # noqa: E402
from oracle import bench_models as bm  # noqa: E402

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


def _setup(nb, name, **model_kw):
    wk = wl.WORKLOADS[name]
    model = (lambda p, d: wk.model(p, d, **model_kw)) if model_kw else wk.model
    data = nb.validate_data_table(wk.tables(
        (lambda E: wl.Workload.device_flux(_Shim(wk, model), E)) if model_kw else None))
    plan = nb.LikelihoodPlan(model, wk.prior, data, wk.P)
    return wk, data, plan


class _Shim:
    """A workload with a substituted model (reduced grids) for Workload.device_flux."""

    def __init__(self, wk, model):
        self.p_true, self.model = wk.p_true, model


def _oracle_batch(name, data, P, **kw):
    model, prior = bm.MODELS[name](**kw)
    od = bm.oracle_data(data)
    lnp, flux = np.empty(len(P)), []
    for w, p in enumerate(P):
        lp, m = o.lnprob(p, od, model, prior)
        lnp[w] = lp
        flux.append(m)
    return lnp, np.array(flux)


def test_c1_synchrotron_flux_call(nb):
    """C1: one Synchrotron.flux() call on 64 photon energies (tests/test_models.py shapes)."""
    E = wl.c1_energies()
    got = wl.c1_flux(E)
    want = bm.c1_flux(E, wl.C1_PARS)
    assert got.unit.to_string() == nb.units.Unit("1/(s cm2 eV)").to_string()
    live = want > want.max() * 1e-250
    assert live.sum() > 50
    assert_allclose(got.value[live], want[live], rtol=FLUX_RTOL)


def test_c2_ic_cmb_32_node_grid(nb):
    """C2: IC on the CMB, 128 walkers, 32-node particle grid (nEed = 8.7)."""
    wk, data, plan = _setup(nb, "C2")
    assert plan.comps[0]["table"].grid.N == 32
    P = wk.walkers(128)
    lnp, flux, blobs = plan(P)
    want, wflux = _oracle_batch("C2", data, P)
    fin = np.isfinite(want)
    assert fin.sum() > 100
    assert np.array_equal(np.isfinite(lnp), fin)
    assert_allclose(lnp[fin], want[fin], rtol=LNP_RTOL)
    assert_allclose(flux, wflux, rtol=FLUX_RTOL)
    # the public sampler on the same plan: 50 steps as BASELINE names them
    s = nb.PlanSampler(128, wk.P, plan, seed=3)
    st = s.run_mcmc(P, 50)
    assert s.get_chain().shape == (50, 128, 3)
    lp_end, _ = _oracle_batch("C2", data, st.coords[::16])
    assert_allclose(st.log_prob[::16], lp_end, rtol=LNP_RTOL)


def test_c5_piondecay_plan_and_sampler(nb):
    """C5: PionDecay (Kafexhiu+14 LUT) + PowerLaw, 512 walkers, through LikelihoodPlan and
    PlanSampler (radiative.py:1495-1536)."""
    wk, data, plan = _setup(nb, "C5")
    assert plan.comps[0]["table"].grid.N == 691
    P = wk.walkers(512, spread=0.02)
    lnp, flux, blobs = plan(P)
    want, wflux = _oracle_batch("C5", data, P)
    assert np.all(np.isfinite(want))
    assert_allclose(lnp, want, rtol=LNP_RTOL)
    assert_allclose(flux, wflux, rtol=FLUX_RTOL)
    # Wp blob (compute_Wp(Epmin = 1 TeV)) against the oracle
    pd = o.PDist("PowerLaw", 10 ** P[7, 0] / 1e12, 30e12, P[7, 1])
    assert_allclose(blobs[0][7], o.compute_Wp(pd, 1e3, 1e7, 100), rtol=1e-10)
    # device-resident sampler == host-driven sampler over the same plan, then the oracle on
    # the final ensemble
    ps = nb.PlanSampler(512, wk.P, plan, seed=9)
    st = ps.run_mcmc(P, 8)
    hs = nb.EnsembleSampler(512, wk.P, lambda q: plan(q, want_blobs=False)[0], vectorize=True,
                            seed=9)
    hs.run_mcmc(P, 8)
    assert_allclose(ps.get_chain(), hs.get_chain(), rtol=1e-12)
    assert_allclose(ps.get_log_prob(), hs.get_log_prob(), rtol=1e-12)
    assert 0.1 < np.mean(ps.acceptance_fraction) < 0.95
    lp_end, _ = _oracle_batch("C5", data, st.coords[::32])
    assert_allclose(st.log_prob[::32], lp_end, rtol=LNP_RTOL)


def test_plan_bremsstrahlung_and_anisotropic_seed(nb):
    """Bremsstrahlung and an anisotropic grey-body seed through LikelihoodPlan + the
    device-resident sampler (radiative.py:576-607, 940-989)."""
    from naima_b200 import units as u
    from naima_b200.models import Bremsstrahlung, ExponentialCutoffPowerLaw, InverseCompton

    theta = 2.1

    def model(pars, data):
        pd = ExponentialCutoffPowerLaw(10 ** pars[0] / u.eV, 10 * u.TeV, pars[1],
                                       (10 ** pars[2]) * u.TeV)
        ic = InverseCompton(pd, seed_photon_fields=[
            "CMB", ["star", 25000 * u.K, 3 * u.eV / u.cm**3, theta * u.rad]], Eemin=10 * u.GeV)
        br = Bremsstrahlung(pd, n0=pars[3] / u.cm**3, nEed=60)
        return ic.flux(data, distance=1.5 * u.kpc) + br.flux(data, distance=1.5 * u.kpc)

    def prior(pars):
        return nb.uniform_prior(pars[1], -1, 5) + nb.uniform_prior(pars[3], 0, np.inf)

    p_true = np.array([33.2, 2.3, 1.4, 5.0])
    E = np.logspace(8.3, 13.7, 24)

    def ofl(p, E):
        pd = o.PDist("ExponentialCutoffPowerLaw", 10 ** p[0], 10e12, p[1], 10 ** p[2] * 1e12, 1.0)
        seeds = ["CMB", ("thermal", 25000.0, 3 * o.eV_erg, theta)]
        ic = o.ic_spectrum(pd, E, seeds, Eemin_eV=10e9)
        br = o.bremsstrahlung_spectrum(pd, E, n0=p[3], nEed=60)
        return o.flux_from_spectrum(ic + br, 1.5 * o.kpc_cm)

    rng = np.random.default_rng(2)
    t = nb.DataTable(meta={"keywords": {"cl": {"value": 0.9}}})
    f = ofl(p_true, E) * (1 + 0.1 * rng.normal(size=E.size))
    t["energy"] = u.Quantity(E, "eV")
    t["flux"] = u.Quantity(f, "1/(s cm2 eV)")
    t["flux_error_lo"] = u.Quantity(0.08 * f, "1/(s cm2 eV)")
    t["flux_error_hi"] = u.Quantity(0.12 * f, "1/(s cm2 eV)")
    ul = np.zeros(E.size, dtype=int)
    ul[-2:] = 1
    t["ul"] = ul
    data = nb.validate_data_table(t)
    plan = nb.LikelihoodPlan(model, prior, data, 4)
    kinds = sorted(type(c["obj"]).__name__ for c in plan.comps)
    assert kinds == ["Bremsstrahlung", "InverseCompton"]
    W = 16
    P = p_true * (1 + 0.03 * rng.normal(size=(W, 4)))
    lnp, flux, _ = plan(P)
    od = bm.oracle_data(data)
    for w in range(W):
        m = ofl(P[w], od["E_eV"]) * od["unit_fac"]
        assert_allclose(flux[w], m, rtol=FLUX_RTOL)
        want = o.lnprobmodel(m, od) + 0.0
        assert_allclose(lnp[w], want, rtol=LNP_RTOL)
    de = nb.DeviceEnsemble(plan, W, seed=4)
    de.set_state(P)
    chain, lp, rows = de.run(5)
    m = ofl(chain[-1, 3], od["E_eV"]) * od["unit_fac"]
    assert_allclose(rows[-1, 3, :E.size], m, rtol=FLUX_RTOL)
    assert_allclose(lp[-1, 3], o.lnprobmodel(m, od), rtol=LNP_RTOL)


def test_c4_ssc_traced_small_grid_vs_oracle(nb):
    """C4 shape on a reduced grid (nEed = 20, 40 seed energies): the traced self-Compton
    plan (auxiliary synchrotron on the seed energies -> seed density -> hoisted two-level
    integral) against the oracle for 12 walkers, and against the untraced class path."""
    kw = dict(nseed=40, nEed=20)
    wk, data, plan = _setup(nb, "C4", **kw)
    assert [c["kind"] for c in plan.comps] == ["table", "ssc", "syn"]
    assert len(plan.aux) == 1
    P = wk.walkers(12, spread=0.03)
    lnp, flux, _ = plan(P)
    want, wflux = _oracle_batch("C4", data, P, **kw)
    assert np.all(np.isfinite(want))
    assert_allclose(flux, wflux, rtol=FLUX_RTOL)
    assert_allclose(lnp, want, rtol=LNP_RTOL)
    # the reference-style call with batched parameters (class path, per-walker seed arrays)
    got = wk.model(np.ascontiguousarray(P.T), data, **kw)
    assert_allclose(got.to(data["flux"].unit).value, wflux, rtol=FLUX_RTOL)


def test_c4_ssc_full_size(nb):
    """C4 at BASELINE's size: 256 walkers, 869-node particle grid, 100 x 100 photon / seed
    energies (8.7 M inner intervals per walker).  Oracle on 6 walkers; the 256-walker launch
    must reproduce the same rows bit for bit (no dependence on the batch), and the
    device-resident sampler must step."""
    wk, data, plan = _setup(nb, "C4")
    ssc = [c for c in plan.comps if c["kind"] == "ssc"][0]
    assert ssc["tb"].grid.N == 869 and ssc["tb"].Ns == 100 and plan.N_E == 100
    P = wk.walkers(256, spread=0.03)
    lnp, flux, _ = plan(P)
    assert np.all(np.isfinite(lnp[np.isfinite(lnp)])) and not np.any(np.isnan(lnp))
    sub = [0, 17, 100, 101, 200, 255]
    lnp6, flux6, _ = plan(P[sub])
    assert np.array_equal(lnp6, lnp[sub]) and np.array_equal(flux6, flux[sub])
    want, wflux = _oracle_batch("C4", data, P[sub])
    assert_allclose(flux6, wflux, rtol=FLUX_RTOL)
    fin = np.isfinite(want)
    assert_allclose(lnp6[fin], want[fin], rtol=LNP_RTOL)
    de = nb.DeviceEnsemble(plan, 256, seed=1)
    de.set_state(P, lnp, None)
    chain, lp, rows = de.run(3)
    assert not np.any(np.isnan(lp))
    lnp_end, _, _ = plan(chain[-1])
    assert_allclose(lp[-1], lnp_end, rtol=1e-13)


def test_nan_lnprob_is_reported_by_the_device_sampler(nb):
    """A NaN log-probability of a proposal must raise emcee's ValueError in the
    device-resident loop as it does in the host-driven one (the accept comparison is False
    for NaN, so it has to be recorded apart).  NaN source: a table of upper limits that are
    ALL violated (core.py:89-92 indexes cl by the violation count)."""
    from naima_b200 import units as u
    from naima_b200.models import InverseCompton, PowerLaw

    def model(pars, data):
        pd = PowerLaw(10 ** pars[0] / u.eV, 10 * u.TeV, pars[1])
        return InverseCompton(pd, seed_photon_fields=["CMB"], Eemin=100 * u.GeV).flux(data)

    E = np.array([1e12, 1e13])
    probe = nb.DataTable(meta={"keywords": {"cl": {"value": 0.9}}})
    probe["energy"] = u.Quantity(E, "eV")
    probe["flux"] = u.Quantity([1.0, 1.0], "1/(s cm2 eV)")
    probe["flux_error"] = u.Quantity([0.1, 0.1], "1/(s cm2 eV)")
    f0 = model(np.array([33.0, 2.5]), nb.validate_data_table(probe)).value
    t = nb.DataTable(meta={"keywords": {"cl": {"value": 0.9}}})
    t["energy"] = u.Quantity(E, "eV")
    # limits: the first is violated by every walker, the second only above 10**33.02
    t["flux"] = u.Quantity([f0[0] * 0.5, f0[1] * 10 ** 0.02], "1/(s cm2 eV)")
    t["flux_error"] = u.Quantity([f0[0] * 0.1, f0[1] * 0.1], "1/(s cm2 eV)")
    t["ul"] = np.array([1, 1])
    data = nb.validate_data_table(t)
    plan = nb.LikelihoodPlan(model, None, data, 2)
    rng = np.random.default_rng(0)
    W = 32
    P = np.column_stack([33.0 + 0.018 * rng.uniform(-1, 1, W), np.full(W, 2.5)
                         + 1e-4 * rng.normal(size=W)])
    lnp0, _, _ = plan(P)
    assert np.all(np.isfinite(lnp0))
    with pytest.raises(ValueError, match="NaN"):
        nb.EnsembleSampler(W, 2, lambda q: plan(q, want_blobs=False)[0], vectorize=True,
                           seed=5).run_mcmc(P, 40)
    de = nb.DeviceEnsemble(plan, W, seed=5)
    de.set_state(P)
    with pytest.raises(ValueError, match="NaN"):
        de.run(40)
    with pytest.raises(ValueError, match="NaN"):
        nb.PlanSampler(W, 2, plan, seed=5).run_mcmc(P, 40)


def test_tablemodel_as_particle_distribution(nb):
    """tests/test_models.py:495-527: a TableModel as the particle distribution of the
    radiative classes; against the analytic distribution it tabulates."""
    from naima_b200 import units as u
    from naima_b200.models import (ExponentialCutoffPowerLaw, InverseCompton, PionDecay,
                                   Synchrotron, TableModel)

    e = np.logspace(-4, 3, 400) * u.TeV  # the table ends where the cut-off has reached 1e-44
    ecpl = ExponentialCutoffPowerLaw(1e36 / u.eV, 1 * u.TeV, 2.0, 10 * u.TeV)
    tm = TableModel(e, ecpl(e), amplitude=1)
    Eph = np.logspace(-3, 1.5, 19) * u.TeV
    for cls, kw in ((InverseCompton, dict(Eemin=1 * u.GeV, Eemax=1 * u.PeV)),
                    (Synchrotron, dict(Eemin=1 * u.GeV, Eemax=1 * u.PeV)),
                    (PionDecay, dict(Epmax=1 * u.PeV))):
        Ep = Eph if cls is not Synchrotron else np.logspace(-2, 4, 19) * u.eV
        want = cls(ecpl, **kw).flux(Ep).value
        got = cls(tm, **kw).flux(Ep).value
        live = want > want.max() * 1e-30
        assert_allclose(got[live], want[live], rtol=2e-3)  # cubic log-log interpolation
    # amplitude batches like a fitted normalisation; We follows
    tmb = TableModel(e, ecpl(e), amplitude=np.array([1.0, 3.0]))
    ic = InverseCompton(tmb, Eemin=1 * u.GeV, Eemax=1 * u.PeV)
    f = ic.flux(Eph).value
    assert f.shape == (2, 19)
    assert_allclose(f[1], 3 * f[0], rtol=1e-12)
    assert_allclose(ic.We.value[1], 3 * ic.We.value[0], rtol=1e-12)
    assert_allclose(ic.We.value[0], InverseCompton(ecpl, Eemin=1 * u.GeV,
                                                   Eemax=1 * u.PeV).We.value, rtol=2e-3)


def test_ebl_absorbed_model_traced(nb):
    """examples/absorbed_SynIC.py with a fixed redshift: the transmission factor folds into
    the IC table's row coefficients of the traced plan; against the class path and the
    oracle times the same transmission."""
    from naima_b200 import units as u
    from naima_b200.models import (BrokenPowerLaw, EblAbsorptionModel, InverseCompton,
                                   Synchrotron)

    def model(pars, data):
        BPL = BrokenPowerLaw(10 ** pars[0] / u.eV, 1.0 * u.TeV, (10 ** pars[1]) * u.TeV, pars[2],
                             pars[3])
        IC = InverseCompton(BPL, seed_photon_fields=["CMB"], Eemin=10 * u.GeV)
        SYN = Synchrotron(BPL, B=pars[4] * u.uG)
        EBL = EblAbsorptionModel(0.06, "Dominguez")
        return (EBL.transmission(data) * IC.flux(data, distance=1.0 * u.kpc)
                + SYN.flux(data, distance=1.0 * u.kpc))

    p0 = np.array((31.0, 1.0, 1.5, 2.3, 0.35))
    E = np.concatenate([np.logspace(2, 4, 8), np.logspace(10, 13.5, 16)])
    trans = EblAbsorptionModel(0.06).transmission(E * u.eV)
    assert trans[0] == 1.0 and trans[-1] < 0.9

    def ofl(p, E):
        pd = o.PDist("BrokenPowerLaw", 10 ** p[0], 1e12, 10 ** p[1] * 1e12, p[2], p[3])
        ic = o.flux_from_spectrum(o.ic_spectrum(pd, E, ["CMB"], Eemin_eV=10e9), o.kpc_cm)
        sy = o.flux_from_spectrum(o.synchrotron_spectrum(pd, E, p[4] * 1e-6), o.kpc_cm)
        return trans * ic + sy

    rng = np.random.default_rng(8)
    t = nb.DataTable()
    f = ofl(p0, E) * (1 + 0.1 * rng.normal(size=E.size))
    t["energy"] = u.Quantity(E, "eV")
    t["flux"] = u.Quantity(f, "1/(s cm2 eV)")
    t["flux_error"] = u.Quantity(0.1 * f, "1/(s cm2 eV)")
    data = nb.validate_data_table(t)
    plan = nb.LikelihoodPlan(model, None, data, 5)
    P = p0 * (1 + 0.01 * rng.normal(size=(9, 5)))
    lnp, flux, _ = plan(P)
    od = bm.oracle_data(data)
    for w in range(9):
        m = ofl(P[w], od["E_eV"]) * od["unit_fac"]
        assert_allclose(flux[w], m, rtol=FLUX_RTOL)
        assert_allclose(lnp[w], o.lnprobmodel(m, od), rtol=LNP_RTOL)
    got = model(np.ascontiguousarray(P.T), data).to(data["flux"].unit).value
    assert_allclose(got, flux, rtol=1e-9)  # set-up kernel vs in-kernel operand evaluation


def test_pion_decay_kelner06(nb, goldens):
    """PionDecayKelner06 (radiative.py:1543-1767) against the reference's golden
    (tests/test_models.py:454-471) and the oracle's QUADPACK restatement, at the accuracy the
    reference's own epsrel = 1e-3 quadrature defines."""
    from naima_b200 import units as u
    from naima_b200.models import ExponentialCutoffPowerLaw, PionDecayKelner06, PowerLaw

    ecpl = ExponentialCutoffPowerLaw(1 / u.TeV, 20 * u.TeV, 2.0, 10 * u.TeV)
    energy = np.logspace(9, 13, 20) * u.eV
    pp = PionDecayKelner06(ecpl)
    flux = pp.flux(energy, 0)
    lum = nb.trapz_loglog(flux * energy, energy).to("erg/s").value
    assert_allclose(lum, 5.54580582494601e-13, rtol=2e-3)
    want = o.PionDecayKelner06(o.PDist("ExponentialCutoffPowerLaw", 1e-12, 20e12, 2.0, 10e12,
                                       1.0)).spectrum(energy.value)
    assert_allclose(flux.value, want, rtol=3e-3)
    # only-high and only-low photon energies (nhat stays 1), a batch of walkers, nh scaling
    hiE = np.logspace(11.2, 13, 5) * u.eV
    assert_allclose(pp.flux(hiE, 0).value, o.PionDecayKelner06(
        o.PDist("ExponentialCutoffPowerLaw", 1e-12, 20e12, 2.0, 10e12, 1.0)).spectrum(hiE.value),
        rtol=3e-3)
    plb = PowerLaw(np.array([1.0, 2.0]) / u.TeV, 20 * u.TeV, np.array([2.2, 2.4]))
    fb = PionDecayKelner06(plb, nh=np.array([1.0, 3.0]) / u.cm**3).flux(energy, 0).value
    assert fb.shape == (2, 20)
    for w, (a, al, nh) in enumerate(((1.0, 2.2, 1.0), (2.0, 2.4, 3.0))):
        ref = o.PionDecayKelner06(o.PDist("PowerLaw", a * 1e-12, 20e12, al), nh=nh)
        assert_allclose(fb[w], ref.spectrum(energy.value), rtol=3e-3)
    assert pp.Wp.unit.to_string() == "erg"
