"""save_run / read_run / find_ML / model_samples on a stand-in sampler (CPU only: these are
host-side callers of the hot path)."""
import numpy as np
import pytest

from naima_b200 import analysis
from naima_b200 import units as u
from naima_b200.utils import DataTable


class FakeSampler:
    def __init__(self, nsteps=5, nwalkers=6, npar=3, seed=0):
        rng = np.random.default_rng(seed)
        self.chain = rng.normal(size=(nsteps, nwalkers, npar))
        self.lp = -rng.random(size=(nsteps, nwalkers))
        self.labels = ["norm", "index", "log10(cutoff)"]
        self.run_info = {"n_walkers": nwalkers, "n_burn": 2, "p0": [1.0, 2.0, 3.0], "guess": True,
                         "n_run": nsteps}
        self.acceptance_fraction = np.full(nwalkers, 0.4)
        d = DataTable()
        d["energy"] = u.Quantity(np.logspace(-1, 1, 7), "TeV")
        d["flux"] = u.Quantity(np.ones(7) * 1e-12, "1/(cm2 s TeV)")
        d["flux_error_lo"] = u.Quantity(np.ones(7) * 1e-13, "1/(cm2 s TeV)")
        d["flux_error_hi"] = u.Quantity(np.ones(7) * 1e-13, "1/(cm2 s TeV)")
        d["ul"] = np.zeros(7, dtype=bool)
        d["cl"] = np.full(7, 0.9)
        self.data = d
        blobs = np.empty((nsteps, nwalkers), dtype=object)
        for s in range(nsteps):
            for w in range(nwalkers):
                flux = u.Quantity(rng.random(7), "1/(cm2 s TeV)")
                pair = (u.Quantity(np.logspace(11, 15, 4), "eV"),
                        u.Quantity(rng.random(4), "1/eV"))
                blobs[s, w] = (flux, pair, u.Quantity(float(rng.random()), "erg"))
        self.blobs = blobs

    def get_chain(self, flat=False):
        return self.chain.reshape(-1, self.chain.shape[-1]) if flat else self.chain

    def get_log_prob(self, flat=False):
        return self.lp.reshape(-1) if flat else self.lp

    def get_blobs(self, flat=False):
        return self.blobs.reshape(-1) if flat else self.blobs


def test_save_read_roundtrip(tmp_path):
    s = FakeSampler()
    fn = tmp_path / "run.npz"
    analysis.save_run(fn, s)
    analysis.save_run(fn, s)  # exists, clobber=False: no error, file untouched
    r = analysis.read_run(fn, modelfn=len)
    assert r.modelfn is len
    assert np.array_equal(r.get_chain(), s.chain) and np.array_equal(r.get_log_prob(), s.lp)
    assert r.get_chain(flat=True).shape == (30, 3)
    assert r.labels == s.labels
    assert r.run_info["n_walkers"] == 6 and r.run_info["p0"] == [1.0, 2.0, 3.0]
    assert abs(r.acceptance_fraction - 0.4) < 1e-15
    b, b0 = r.get_blobs()[3, 2], s.blobs[3, 2]
    assert b[0].unit.to_string() == b0[0].unit.to_string()
    assert np.array_equal(b[0].value, b0[0].value)
    assert np.array_equal(b[1][1].value, b0[1][1].value) and b[1][0].unit.physical_type == "energy"
    assert b[2].value == b0[2].value and b[2].unit.physical_type == "energy"
    assert np.array_equal(u.Quantity(r.data["energy"]).value, s.data["energy"].value)
    assert u.Quantity(r.data["flux"]).unit.to_string() == s.data["flux"].unit.to_string()
    with pytest.raises(ValueError):
        analysis.save_run(tmp_path / "run.txt", s)


def test_find_ML_and_model_samples():
    s = FakeSampler()
    ML, MLp, MLerr, (mx, my) = analysis.find_ML(s, 0)
    idx = np.unravel_index(np.argmax(s.lp), s.lp.shape)
    assert ML == s.lp[idx] and np.array_equal(MLp, s.chain[idx])
    assert np.array_equal(my.value, s.blobs[idx][0].value) and len(MLerr) == 3
    mx2, my2 = analysis.find_ML(s, 1)[3]
    assert mx2.unit.physical_type == "energy" and my2.shape == (4,)

    def modelfn(pars, data):  # batch-aware: pars [P, n]
        E = u.Quantity(data["energy"]).to("TeV").value
        amp = np.atleast_1d(pars[0])[:, None]
        return u.Quantity(amp * E[None, :] ** -2.0, "1/(cm2 s TeV)")

    s.modelfn = modelfn
    E, m = analysis.model_samples(s, u.Quantity([0.1, 100.0], "TeV"), e_npoints=20, n_samples=9,
                                  seed=1)
    assert E.shape == (20,) and m.shape == (9, 20)
    with pytest.raises(TypeError):
        analysis.model_samples(s, u.Quantity([1.0, 2.0], "cm"))


def test_save_run_uses_the_reference_layout_names(tmp_path):
    """Group, dataset and attribute names of save_run against the layout read off the
    reference's save_run source (tests/golden/save_run_layout.json, analysis.py:366-471); the
    data table is the one documented difference."""
    import json
    import os

    with open(os.path.join(os.path.dirname(__file__), "golden", "save_run_layout.json")) as f:
        ref = json.load(f)
    s = FakeSampler()
    fn = tmp_path / "run.npz"
    analysis.save_run(fn, s)
    z = np.load(fn, allow_pickle=False)
    attrs = json.loads(str(z["__attrs__"]))
    (group,) = ref["group"]
    names = [k for k in z.files if k != "__attrs__"]
    assert all(k.startswith(group + "/") for k in names)
    for ds in ref["datasets"]:
        key = "%s/%s" % (group, ds.format(0))
        assert key in names, key
    # blob datasets are flattened over steps x walkers, as the reference's loop builds them
    assert z["mcmc/blob0"].shape == (s.chain.shape[0] * s.chain.shape[1], 7)
    assert z["mcmc/blob1"].shape == (30, 2, 4)
    for a in ref["attrs"]:
        if a == "<run_info key>":
            for k in s.run_info:
                assert "%s@%s" % (group, k) in attrs
        elif a == "unit":
            assert "mcmc/blob0@unit" in attrs and "mcmc/blob2@unit" in attrs
        elif a == "unit{0}":
            assert "mcmc/blob1@unit0" in attrs and "mcmc/blob1@unit1" in attrs
        else:
            assert "%s@%s" % (group, a.format(0)) in attrs, a
    # the data table sits under <group>/<table path>/, one dataset per column
    (tp,) = ref["table_path"]
    assert "%s/%s/energy" % (group, tp) in names
    # a reference-written file (compound mcmc/data) is refused, not misread
    bad = dict(z)
    bad["mcmc/data"] = np.zeros(3)
    np.savez(tmp_path / "foreign.npz", **bad)
    with pytest.raises(ValueError, match="compound dataset"):
        analysis.read_run(tmp_path / "foreign.npz")
