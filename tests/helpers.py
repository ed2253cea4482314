"""Shared builders for the parity tests: the RXJ1713 data tables from the golden
fixture, the shipped example models written against naima_b200 (same source as
the reference's examples/*.py) and their oracle counterparts."""
import os

import numpy as np

import oracle.naima_oracle as o

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TeV = 1e12


def rxj_tables():
    import naima_b200 as naima
    from naima_b200 import units as u

    d = np.load(os.path.join(GOLDEN, "rxj1713_data.npz"))
    hess = naima.DataTable(meta={"keywords": {"cl": {"value": float(d["hess_cl"])}}})
    hess["energy"] = d["hess_energy_TeV"] * u.TeV
    hess["flux"] = u.Quantity(d["hess_flux"], "1/(cm2 s TeV)")
    hess["flux_error"] = u.Quantity(d["hess_flux_error"], "1/(cm2 s TeV)")
    hess["ul"] = d["hess_ul"]
    suz = naima.DataTable()
    suz["energy"] = u.Quantity(d["suzaku_energy_eV"][::5], "eV")
    suz["flux"] = u.Quantity(d["suzaku_flux"][::5], "erg/(cm2 s)")
    suz["flux_error"] = u.Quantity(d["suzaku_flux_error"][::5], "erg/(cm2 s)")
    return suz, hess


# --- examples/RXJ1713_IC.py:16-62 -------------------------------------------------
def ElectronIC(pars, data):
    from naima_b200 import units as u
    from naima_b200.models import ExponentialCutoffPowerLaw, InverseCompton

    amplitude = pars[0] / u.eV
    alpha = pars[1]
    e_cutoff = (10 ** pars[2]) * u.TeV
    ECPL = ExponentialCutoffPowerLaw(amplitude, 10.0 * u.TeV, alpha, e_cutoff)
    IC = InverseCompton(
        ECPL, seed_photon_fields=["CMB", ["FIR", 26.5 * u.K, 0.415 * u.eV / u.cm**3]],
        Eemin=100 * u.GeV)
    model = IC.flux(data, distance=1.0 * u.kpc).to(data["flux"].unit)
    elec_energy = np.logspace(11, 15, 100) * u.eV
    nelec = ECPL(elec_energy)
    We = IC.compute_We(Eemin=1 * u.TeV)
    return model, (elec_energy, nelec), We


def lnprior_IC(pars):
    import naima_b200 as naima

    return naima.uniform_prior(pars[0], 0.0, np.inf) + naima.uniform_prior(pars[1], -1, 5)


def oracle_IC(Eemin_eV=100e9):
    """Oracle model/prior pair for ElectronIC on a validated data dict."""
    seeds = ["CMB", ("thermal", 26.5, 0.415 * o.eV_erg)]

    def model(p, data):
        pd = o.PDist("ExponentialCutoffPowerLaw", p[0], 10 * TeV, p[1], 10 ** p[2] * TeV, 1.0)
        spec = o.ic_spectrum(pd, data["E_eV"], seeds, Eemin_eV=Eemin_eV)
        return o.flux_from_spectrum(spec, 1.0 * o.kpc_cm) * data["unit_fac"]

    def prior(p):
        return o.uniform_prior(p[0], 0.0, np.inf) + o.uniform_prior(p[1], -1, 5)

    return model, prior


# --- examples/RXJ1713_SynIC.py:19-62 ------------------------------------------------
def ElectronSynIC(pars, data):
    from naima_b200 import units as u
    from naima_b200.models import ExponentialCutoffPowerLaw, InverseCompton, Synchrotron

    amplitude = 10 ** pars[0] / u.eV
    alpha = pars[1]
    e_cutoff = (10 ** pars[2]) * u.TeV
    B = pars[3] * u.uG
    ECPL = ExponentialCutoffPowerLaw(amplitude, 10.0 * u.TeV, alpha, e_cutoff)
    IC = InverseCompton(
        ECPL, seed_photon_fields=["CMB", ["FIR", 26.5 * u.K, 0.415 * u.eV / u.cm**3]],
        Eemin=100 * u.GeV)
    SYN = Synchrotron(ECPL, B=B)
    model = IC.flux(data, distance=1.0 * u.kpc) + SYN.flux(data, distance=1.0 * u.kpc)
    return model, IC.compute_We(Eemin=1 * u.TeV)


def lnprior_SynIC(pars):
    import naima_b200 as naima

    return (naima.uniform_prior(pars[0], 0.0, np.inf) + naima.uniform_prior(pars[1], -1, 5)
            + naima.uniform_prior(pars[3], 0, np.inf))


def oracle_SynIC(seeds=None):
    seeds = ["CMB", ("thermal", 26.5, 0.415 * o.eV_erg)] if seeds is None else seeds

    def model(p, data):
        pd = o.PDist("ExponentialCutoffPowerLaw", 10 ** p[0], 10 * TeV, p[1], 10 ** p[2] * TeV,
                     1.0)
        E = data["E_eV"]
        ic = o.flux_from_spectrum(o.ic_spectrum(pd, E, seeds, Eemin_eV=100e9), o.kpc_cm)
        sy = o.flux_from_spectrum(o.synchrotron_spectrum(pd, E, p[3] * 1e-6), o.kpc_cm)
        return (ic + sy) * data["unit_fac"]

    def prior(p):
        return (o.uniform_prior(p[0], 0.0, np.inf) + o.uniform_prior(p[1], -1, 5)
                + o.uniform_prior(p[3], 0, np.inf))

    return model, prior


from oracle.bench_models import oracle_data  # noqa: E402,F401


def oracle_lnprob_batch(P, odata, model, prior):
    out = np.empty(len(P))
    fl = []
    for w, p in enumerate(P):
        lp, m = o.lnprob(p, odata, model, prior)
        out[w] = lp
        fl.append(m)
    return out, np.array(fl)


def oracle_stretch_sampler(lnprob_batch, p0, nsteps, seed, a=2.0):
    """emcee stretch move in plain NumPy with the same draw
    order as naima_b200.sampler -- the oracle for chain parity."""
    rs = np.random.mtrand.RandomState(seed)
    coords = np.array(p0, dtype=float)
    W, ndim = coords.shape
    lp = lnprob_batch(coords)
    chain = np.empty((nsteps, W, ndim))
    lps = np.empty((nsteps, W))
    all_inds = np.arange(W)
    for t in range(nsteps):
        rs.choice(1, p=[1.0])
        inds = all_inds % 2
        rs.shuffle(inds)
        for split in range(2):
            S1 = inds == split
            s, c = coords[S1], coords[~S1]
            Ns = len(s)
            zz = ((a - 1.0) * rs.rand(Ns) + 1) ** 2.0 / a
            factors = (ndim - 1.0) * np.log(zz)
            rint = rs.randint(len(c), size=(Ns,))
            q = c[rint] - (c[rint] - s) * zz[:, None]
            nlp = lnprob_batch(q)
            acc = factors + nlp - lp[S1] > np.log(rs.rand(Ns))
            idx = np.flatnonzero(S1)[acc]
            coords[idx] = q[acc]
            lp[idx] = nlp[acc]
        chain[t], lps[t] = coords, lp
    return chain, lps
