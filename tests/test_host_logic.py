"""Host-side logic that needs no GPU: units layer, data-table validation, the
tracing of user callbacks into a plan description, the ensemble sampler's
stretch move against the oracle's NumPy restatement, Nelder-Mead."""
import os

import numpy as np
import pytest
from numpy.testing import assert_allclose

import naima_b200 as nb
from helpers import (ElectronIC, ElectronSynIC, lnprior_IC, lnprior_SynIC, oracle_stretch_sampler,
                     rxj_tables)
from naima_b200 import fused
from naima_b200 import units as u


def test_units_conversions():
    assert (100 * u.GeV).to("eV").value == 1e11
    assert (1 * u.PeV).to("eV").value == 1e15
    assert_allclose((1 * u.kpc).to("cm").value, 3.0856775814913673e21, rtol=1e-15)
    assert_allclose((1 * u.Unit("mec2")).to("eV").value, 510998.9499961642, rtol=1e-15)
    q = u.Quantity([1.0, 2.0], "1/(s cm2 eV)")
    assert q.unit.physical_type == "differential flux"
    assert_allclose(q.to("1/(s cm2 TeV)").value, [1e12, 2e12])
    sed = (q * u.Quantity([1.0, 2.0], "TeV") ** 2).to("erg/(cm2 s)")
    assert sed.unit.physical_type == "flux"
    assert_allclose(sed.value, np.array([1.0, 8.0]) * 1e24 * 1.602176634e-12)
    assert (0.415 * u.eV / u.cm**3).unit.physical_type == "pressure"
    assert (1 / (u.eV * u.cm**3)).unit.physical_type == "differential number density"
    assert (10 ** np.float64(3) / u.eV).unit.physical_type == "differential energy"
    assert (3 * u.uG).to("G").value == 3e-6
    assert_allclose((45 * u.deg).to("rad").value, np.pi / 4)
    assert u.Unit("1 / (cm2 s TeV)") == u.Unit("1/(s TeV cm2)")
    with pytest.raises(u.UnitConversionError):
        (1 * u.eV).to("cm")
    with pytest.raises(u.UnitsError):
        (1 * u.eV) + 1.0
    assert ((1 * u.kpc) != 0) and not np.all((0 * u.kpc) != 0)
    assert_allclose(((2 * u.erg) / (4 * u.erg)).decompose().value, 0.5)
    assert_allclose(float(u.Quantity(3.0, "TeV") / u.Quantity(1.5, "GeV")), 2000.0)


def test_read_ipac_and_validate(tmp_path):
    txt = ("\\ comment\n\\cl=0.95\n|energy|   flux|flux_error|  ul|\n|double| double|    double|long|\n"
           "|   TeV|1 / (cm2 s TeV)|1 / (cm2 s TeV)|    |\n"
           " 1.0 1e-11 1e-12 0\n 0.5 4e-11 3e-12 0\n 2.0 2e-12 1e-12 1\n")
    f = tmp_path / "t.dat"
    f.write_text(txt)
    t = nb.read_ipac(str(f))
    assert t.meta["keywords"]["cl"]["value"] == 0.95
    d = nb.validate_data_table(t)
    assert_allclose(d["energy"].value, [0.5, 1.0, 2.0])  # sorted (regression #247)
    assert list(d["ul"]) == [False, False, True] and np.all(d["cl"] == 0.95)
    assert d["flux"].unit == u.Unit("1/(s TeV cm2)")
    assert "energy_error_lo" in d and "flux_error_hi" in d
    with pytest.raises(TypeError):
        nb.validate_data_table({"energy": 1})
    bad = nb.DataTable()
    bad["energy"] = [1, 2] * u.TeV
    bad["flux"] = u.Quantity([1, 2], "1/(cm2 s TeV)")
    with pytest.raises(TypeError):
        nb.validate_data_table(bad)  # no flux_error
    bad["flux_error"] = u.Quantity([1, 2], "TeV")
    with pytest.raises(TypeError):
        nb.validate_data_table(bad)  # wrong physical type


def test_validate_multiple_tables_and_sed_conversion():
    suz, hess = rxj_tables()
    d = nb.validate_data_table([suz, hess])
    assert len(d) == 36 + 28 and d["flux"].unit.physical_type == "flux"
    assert np.all(np.diff(d["energy"].to("eV").value) > 0)
    assert set(np.unique(d["group"])) == {0, 1}
    d2 = nb.validate_data_table([suz, hess], sed=False)
    assert d2["flux"].unit.physical_type == "differential flux"
    E = d["energy"]
    assert_allclose((d2["flux"] * E**2).to("erg/(cm2 s)").value, d["flux"].value, rtol=1e-14)
    funit, sedf = nb.sed_conversion(E, u.Unit("1/(s cm2 eV)"), True)
    assert funit.physical_type == "flux" and sedf.unit.physical_type != "dimensionless"
    with pytest.raises(u.UnitsError):
        nb.sed_conversion(E, u.Unit("cm"), True)
    t = nb.build_data_table([1, 2, 3] * u.TeV, u.Quantity([3, 2, 1], "1/(cm2 s TeV)"),
                            flux_error=u.Quantity([1, 1, 1], "1/(cm2 s TeV)"), ul=[0, 0, 1],
                            cl=0.99)
    assert np.all(nb.validate_data_table(t)["cl"] == 0.99)


def test_trace_synic():
    suz, hess = rxj_tables()
    data = nb.validate_data_table([suz, hess])
    flux, blobs, sp = fused.trace(ElectronSynIC, lnprior_SynIC, data, 4)
    kinds = [type(c).__name__ for c, _, _ in flux.groups]
    assert kinds == ["InverseCompton", "Synchrotron"]
    assert flux.unit == u.Unit("1/(s cm2 eV)") and not flux.sed
    assert_allclose(flux.groups[0][1], 4 * np.pi * 3.0856775814913673e21**2)
    assert blobs[0].kind == "W"
    assert sp.terms == [(0, 0, 0.0, np.inf), (1, 0, -1.0, 5.0), (3, 0, 0.0, np.inf)]
    pd = flux.groups[0][0].particle_distribution
    vals = pd._eval_params(nb.models._val(pd.amplitude, "1/eV"))
    assert (vals[0].src, vals[0].fn, vals[0].scale) == (0, fused.FN_POW10, 1.0)
    assert vals[1] == 1e13 and (vals[2].src, vals[2].fn) == (1, fused.FN_ID)
    assert (vals[3].src, vals[3].fn, vals[3].scale) == (2, fused.FN_POW10, 1e12)
    B = nb.models._val(flux.groups[1][0].B, "G")
    assert (B.src, B.fn) == (3, fused.FN_ID) and B.scale == pytest.approx(1e-6, rel=1e-15)


def test_trace_ic_blobs_and_failures():
    _, hess = rxj_tables()
    data = nb.validate_data_table(hess)
    flux, blobs, sp = fused.trace(ElectronIC, lnprior_IC, data, 3)
    assert flux.unit == data["flux"].unit
    assert isinstance(blobs[0], tuple) and blobs[0][1].kind == "pdist" and blobs[1].kind == "W"

    def bad1(pars, data):
        return ElectronIC([pars[0] + 1.0, pars[1], pars[2]], data)

    def bad2(pars, data):
        if pars[1] > 2:
            return ElectronIC(pars, data)
        return ElectronIC(pars, data)

    def bad3(pars, data):
        return np.ones(len(data["energy"])) * u.Unit("1/(s cm2 TeV)")

    for bad in (bad1, bad2, bad3):
        with pytest.raises(fused.TraceError):
            fused.trace(bad, lnprior_IC, data, 3)
    with pytest.raises(fused.TraceError):
        fused.trace(ElectronIC, lambda p: -0.5 * p[1] ** 2, data, 3)


def _gauss_lnprob(q):
    q = np.atleast_2d(q)
    return -0.5 * np.sum((q - np.array([1.0, -2.0, 0.5])) ** 2 / np.array([1.0, 4.0, 0.25]), axis=1)


def test_ensemble_sampler_matches_oracle_stretch_move():
    rng = np.random.default_rng(2)
    W, P, n = 12, 3, 25
    p0 = rng.normal(size=(W, P))
    s = nb.EnsembleSampler(W, P, _gauss_lnprob, vectorize=True, seed=99)
    st = s.run_mcmc(p0, n)
    chain, lps = oracle_stretch_sampler(_gauss_lnprob, p0, n, 99)
    assert np.array_equal(s.get_chain(), chain) and np.array_equal(s.get_log_prob(), lps)
    assert np.array_equal(st.coords, chain[-1])
    assert s.get_chain(flat=True).shape == (n * W, P)
    assert 0.2 < s.acceptance_fraction.mean() < 0.9
    # per-walker calling convention with blobs
    s2 = nb.EnsembleSampler(W, P, lambda p: (float(_gauss_lnprob(p)[0]), p.sum(), "x"), seed=99)
    s2.run_mcmc(p0, n)
    assert np.array_equal(s2.get_chain(), chain)
    b = s2.get_blobs()
    assert b.shape == (n, W) and b[3, 4][1] == "x"
    assert b[-1, 0][0] == pytest.approx(chain[-1, 0].sum())
    s2.reset()
    assert s2.get_chain().shape == (0, W, P) and s2.iteration == 0


def test_ensemble_sampler_errors():
    p0 = np.random.default_rng(0).normal(size=(8, 3))
    with pytest.raises(ValueError):
        nb.EnsembleSampler(4, 3, _gauss_lnprob, vectorize=True).run_mcmc(p0[:4], 1)
    with pytest.raises(ValueError):
        nb.EnsembleSampler(8, 3, _gauss_lnprob, vectorize=True).run_mcmc(p0[:6], 1)
    with pytest.raises(ValueError):
        nb.EnsembleSampler(8, 3, _gauss_lnprob, vectorize=True).run_mcmc(np.ones((8, 3)), 1)
    with pytest.raises(ValueError):
        nb.EnsembleSampler(8, 3, lambda q: np.full(len(q), np.nan), vectorize=True).run_mcmc(p0, 1)
    bad = p0.copy()
    bad[0, 0] = np.inf
    with pytest.raises(ValueError):
        nb.EnsembleSampler(8, 3, _gauss_lnprob, vectorize=True).run_mcmc(
            bad, 1, skip_initial_state_check=True)


def test_nelder_mead():
    from naima_b200.minimize import minimize

    f = lambda x: (x[0] - 3.0) ** 2 + 10 * (x[1] + 1.0) ** 2 + 2.0
    r = minimize(f, [1.0, 1.0], options={"xtol": 1e-6, "ftol": 1e-9, "maxfev": 2000})
    assert r["success"] and r["status"] == 0
    assert_allclose(r["x"], [3.0, -1.0], rtol=1e-4)
    r = minimize(f, [1.0, 1.0], options={"maxfev": 5})
    assert r["status"] == 1 and not r["success"]


def test_priors_scalar_conventions():
    assert nb.uniform_prior(1, 0, 2) == 0.0 and nb.uniform_prior(-1, 0, 2) == -np.inf
    assert nb.normal_prior(1.0, 1.0, 2.0) == -0.5 * (2 * np.pi * 2.0)
    assert nb.log_uniform_prior(4.0, 1, None) == 0.25
    assert nb.log_uniform_prior(0.5, 1, 3) == -np.inf


def test_device_ensemble_draws_follow_the_host_sampler_stream():
    """DeviceEnsemble pre-draws the random numbers of a block of steps; they must be the
    numbers the host-driven EnsembleSampler (emcee's draw order) consumes, so that both
    produce the same chain from the same seed."""
    from naima_b200.sampler import DeviceEnsemble

    class Draws(DeviceEnsemble):  # no device: only the host-side drawing logic
        def __init__(self, W, seed, a=2.0):
            self.W, self.Ns, self.a = W, W // 2, a
            self._random = np.random.mtrand.RandomState(seed)
            self._all_inds = np.arange(W)

    W, n, a = 14, 5, 2.0
    s_idx, c_idx, zz, lnu = Draws(W, 9)._draw_block(n)
    rs = np.random.mtrand.RandomState(9)
    for t in range(n):
        rs.choice(1, p=[1.0])  # emcee: self._random.choice(self._moves, p=self._weights)
        inds = np.arange(W) % 2
        rs.shuffle(inds)
        for split in range(2):
            S1 = inds == split
            comp = np.flatnonzero(~S1)
            assert np.array_equal(s_idx[t, split], np.flatnonzero(S1))
            assert np.array_equal(zz[t, split], ((a - 1.0) * rs.rand(W // 2) + 1) ** 2.0 / a)
            assert np.array_equal(c_idx[t, split], comp[rs.randint(W // 2, size=(W // 2,))])
            assert np.array_equal(lnu[t, split], np.log(rs.rand(W // 2)))
    # replaying k steps from a saved generator state lands on the same state
    d = Draws(W, 9)
    rng0 = d._random.get_state()
    d._draw_block(3)
    after3 = d._random.get_state()
    d._draw_block(4)
    got = d.rng_state_after(rng0, 3)
    assert got[0] == after3[0] and np.array_equal(got[1], after3[1]) and got[2:] == after3[2:]


@pytest.mark.parametrize("W", [2, 6, 14, 256, 1030])
def test_host_draw_helper_is_bit_identical_to_numpy(W):
    """nb_host_draw_steps (C, MT19937 + numpy's legacy algorithms restated) against the NumPy
    calls it replaces: same numbers, same generator state afterwards, for slices of a larger
    buffer (the sampler writes straight into its pinned staging arrays)."""
    from naima_b200._lib import host_lib
    from naima_b200.sampler import DeviceEnsemble

    if host_lib() is None:
        pytest.skip("host helper library not built")

    class Draws(DeviceEnsemble):
        def __init__(self, W, seed, a, use_c):
            self.W, self.Ns, self.a = W, W // 2, a
            self._random = np.random.mtrand.RandomState(seed)
            self._all_inds = np.arange(W)
            self.use_host_lib = use_c

    for seed, a in ((0, 2.0), (5, 2.0), (11, 1.6)):
        c, py = Draws(W, seed, a, True), Draws(W, seed, a, False)
        c._random.rand(seed + 1), py._random.rand(seed + 1)  # not at a fresh state
        big = [np.zeros((9, 2, W // 2), dtype=np.int32), np.zeros((9, 2, W // 2), dtype=np.int32),
               np.zeros((9, 2, W // 2)), np.zeros((9, 2, W // 2))]
        for t0, t1 in ((0, 1), (1, 4), (4, 9)):
            c._draw_into(*[x[t0:t1] for x in big])
        want = py._draw_block(9)
        for got, ref in zip(big, want):
            assert np.array_equal(got, ref)
        assert c._random.rand() == py._random.rand()
        # the replay used for rewinding after an early generator exit
        rng0 = c._random.get_state()
        c._draw_block(5)
        s3 = py._random.get_state()  # same state as c had at rng0
        py._draw_block(3)
        after3 = py._random.get_state()
        got = c.rng_state_after(rng0, 3)
        assert got[0] == after3[0] and np.array_equal(got[1], after3[1]) and got[2] == after3[2]
        del s3


def test_batched_simplex_is_the_serial_simplex():
    """The batched Nelder-Mead (one launch per iteration: reflection, expansion and both
    contractions evaluated together) makes the serial algorithm's decisions: same trajectory,
    same evaluation count, same stopping status (core.py:181-187 stops at maxfev = 500)."""
    from naima_b200.minimize import minimize

    def f(x):
        return 100 * (x[1] - x[0] ** 2) ** 2 + (1 - x[0]) ** 2 + (x[2] - 3) ** 2 + abs(x[3]) ** 1.5

    def fb(X):
        return np.array([f(x) for x in X])

    for opts in ({"maxfev": 500, "xtol": 1e-1, "ftol": 1e-3},
                 {"maxfev": 500, "xtol": 1e-6, "ftol": 1e-9},
                 {"maxfev": 60, "xtol": 1e-8, "ftol": 1e-12}):
        for x0 in ([1.3, 0.7, 2.0, 0.5], [-1.2, 1.0, 0.0, 4.0]):
            a = minimize(f, x0, options=opts)
            b = minimize(f, x0, options=opts, batch_func=fb)
            assert np.array_equal(a["x"], b["x"]) and a["fun"] == b["fun"]
            assert (a["nfev"], a["nit"], a["status"]) == (b["nfev"], b["nit"], b["status"])
            assert b["launches"] <= b["nit"] + 1 + b["nfev"] // 4


def test_tablemodel_and_ebl_host_side():
    """tests/test_models.py:495-547 (the parts that need no radiative class)."""
    from naima_b200 import units as u
    from naima_b200.models import EblAbsorptionModel, TableModel

    lemin, lemax = -4, 2
    e = np.logspace(lemin, lemax, 50) * u.TeV
    n = (e.value) ** -2 * np.exp(-e.value / 10) / u.eV
    tm = TableModel(e, n, amplitude=1)
    np.testing.assert_allclose(n.to("1/eV").value, tm(e).to("1/eV").value)
    e2 = np.logspace(lemin, lemax, 1000) * u.TeV
    n2 = (e2.value) ** -2 * np.exp(-e2.value / 10) / u.eV
    np.testing.assert_allclose(n2.to("1/eV").value, tm(e2).to("1/eV").value, rtol=1e-1)
    tm2 = TableModel(e, n.value)
    np.testing.assert_allclose(tm2(e2).value, n2.value, rtol=1e-1)
    e3 = np.logspace(lemin - 4, lemin - 2, 100) * u.TeV
    np.testing.assert_allclose(tm(e3).value, 0.0)
    # a batch of amplitudes: one row per walker
    tmb = TableModel(e, n, amplitude=np.array([1.0, 2.0, 0.5]))
    out = tmb(e2).value
    assert out.shape == (3, 1000)
    np.testing.assert_allclose(out[1], 2 * out[0])

    EBL_zero = EblAbsorptionModel(0.0, "Dominguez")
    EBL_moderate = EblAbsorptionModel(0.5, "Dominguez")
    np.testing.assert_allclose(np.ones(50), EBL_zero.transmission(e), rtol=1e-1)
    assert np.all(EBL_zero.transmission(e) - EBL_moderate.transmission(e) > -1e-10)
    t = EBL_moderate.transmission(np.array([0.5, 1e3, 2e5]) * u.GeV)
    assert t[0] == 1.0 and 0 < t[1] < 1 and t[2] == np.exp(-np.log10(6000.0))
    with pytest.raises(ValueError):
        EblAbsorptionModel(0.3, "Franceschini")
    with pytest.raises(ValueError):
        EblAbsorptionModel(-0.1)


def test_transposing_row_sum_tree_has_the_pairing_of_the_plain_tree():
    """nb_kernels.cu: warp_sum_rows<RT> (RT row sums in one shuffle tree whose lane pairs split
    the rows between them) against warp_sum (shfl_down by 16, 8, 4, 2, 1 per row), both
    emulated lane by lane in float64: lane (32 / RT) * r must end up with the bits lane 0 of
    the plain tree has for row r."""
    rng = np.random.default_rng(11)

    def plain(v):  # v[32]: value of each lane; returns lane 0's result
        v = v.copy()
        for o in (16, 8, 4, 2, 1):
            nxt = v.copy()
            for lane in range(32):
                src = lane + o
                nxt[lane] = v[lane] + (v[src] if src < 32 else v[lane])  # out of range: own value
            v = nxt
        return v[0]

    def transposed(acc, RT):  # acc[32][RT]
        acc = acc.copy()
        off, n = 16, RT
        while n > 1:
            new = acc.copy()
            for lane in range(32):
                upper = (lane & off) != 0
                for r in range(n // 2):
                    send_partner = acc[lane ^ off][r] if ((lane ^ off) & off) else acc[lane ^ off][r + n // 2]
                    keep = acc[lane][r + n // 2] if upper else acc[lane][r]
                    new[lane][r] = keep + send_partner
            acc = new
            n //= 2
            off //= 2
        v = acc[:, 0].copy()
        while off > 0:
            v = np.array([v[lane] + v[lane ^ off] for lane in range(32)])
            off //= 2
        return v

    for RT in (8, 4, 2):
        acc = rng.normal(size=(32, RT)) * 10.0 ** rng.integers(-8, 8, size=(32, RT))
        got = transposed(acc, RT)
        for r in range(RT):
            want = plain(acc[:, r])
            assert got[(32 // RT) * r] == want, (RT, r)


def test_ssc_ring_schedule_never_reads_a_stale_or_overwritten_slot():
    """ssc_inner_kernel's cp.async ring (nb_kernels.cu): SSC_STAGES slots per thread, filled
    SSC_STAGES - 1 intervals ahead, one commit group per interval, wait_group<SSC_STAGES - 2>
    before interval s is read.  Emulated here with the group semantics of cp.async: when interval
    s is read, the copy of interval s must have completed and its slot must not have been
    refilled; a refill may only target a slot whose interval has been consumed."""
    STAGES = 8
    for nint in list(range(1, 40)) + [99, 128]:
        slot_holds = [None] * STAGES     # interval whose data the slot holds (once landed)
        groups = []                      # per committed group: list of (slot, interval)
        landed = 0                       # groups [0, landed) have completed

        def commit(copies):
            groups.append(copies)

        def wait_group(n):               # at most n most-recent groups may still be pending
            nonlocal landed
            upto = max(landed, len(groups) - n)
            for g in groups[landed:upto]:
                for slot, itv in g:
                    slot_holds[slot] = itv
            landed = upto

        consumed = -1
        for g in range(STAGES - 1):      # prologue
            commit([(g, g)] if g < nint else [])
        s = 0

        def interval(s, slot):
            nonlocal consumed
            wait_group(STAGES - 2)
            assert slot_holds[slot] == s, (nint, s, slot, slot_holds)
            consumed = s
            sn = s + STAGES - 1
            tgt = (slot + STAGES - 1) % STAGES
            if sn < nint:
                # the slot being refilled held interval s - 1 (or nothing): already consumed
                assert slot_holds[tgt] is None or slot_holds[tgt] <= consumed - 1 or \
                    slot_holds[tgt] == s - 1, (nint, s, tgt, slot_holds)
                commit([(tgt, sn)])
            else:
                commit([])

        while s + STAGES <= nint:        # unrolled rounds: slot numbers are compile-time
            for g in range(STAGES):
                interval(s + g, g)
            s += STAGES
        g = 0
        while s < nint:                  # fewer than STAGES left
            interval(s, g)
            s += 1
            g += 1
        assert consumed == nint - 1
