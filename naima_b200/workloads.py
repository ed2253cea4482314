# -*- coding: utf-8 -*-
"""Synthetic workloads of BASELINE.json's configs: seeded
fake spectra of the named shapes, the example model functions written against
this package, and plain-float descriptions of the same models for the CPU
baseline.  No files are read; `data: "synthetic"` in bench.py means this.
"""
import numpy as np

SEED = 20261017
TeV = 1e12


def _u():
    from . import units as u
    return u


# --- C3: RXJ1713_SynIC, joint Synchrotron + IC (CMB + FIR + NIR) ----------------------
C3_PTRUE = np.array([33.0, 2.5, np.log10(48.0), 20.0])  # examples/RXJ1713_SynIC.py:72 (+B)
C3_SEEDS = ("CMB", "FIR", "NIR")


def c3_model(pars, data):
    """examples/RXJ1713_SynIC.py:19-46 with BASELINE's three seed fields."""
    u = _u()
    from .models import ExponentialCutoffPowerLaw, InverseCompton, Synchrotron

    amplitude = 10 ** pars[0] / u.eV
    alpha = pars[1]
    e_cutoff = (10 ** pars[2]) * u.TeV
    B = pars[3] * u.uG
    ECPL = ExponentialCutoffPowerLaw(amplitude, 10.0 * u.TeV, alpha, e_cutoff)
    IC = InverseCompton(ECPL, seed_photon_fields=list(C3_SEEDS), Eemin=100 * u.GeV)
    SYN = Synchrotron(ECPL, B=B)
    model = IC.flux(data, distance=1.0 * u.kpc) + SYN.flux(data, distance=1.0 * u.kpc)
    return model, IC.compute_We(Eemin=1 * u.TeV)


def c3_prior(pars):
    from .core import uniform_prior

    return (uniform_prior(pars[0], 0.0, np.inf) + uniform_prior(pars[1], -1, 5)
            + uniform_prior(pars[3], 0, np.inf))


def c3_energies():
    """36 X-ray + 28 VHE photon energies (N_E = 64), eV."""
    x = np.logspace(np.log10(0.55e3), np.log10(10e3), 36)
    g = np.logspace(np.log10(0.33 * TeV), np.log10(170 * TeV), 28)
    return x, g


def c3_tables(flux_model_fn, seed=SEED):
    """Fake data tables: flux = model(p_true) (1 + 0.1 N(0,1)), sigma = 0.1 flux, last VHE
    point an upper limit at cl = 0.95.  flux_model_fn(E_eV) -> 1/(s cm2 eV)."""
    u = _u()
    from .utils import DataTable

    rng = np.random.default_rng(seed)
    x, g = c3_energies()
    fx = flux_model_fn(x) * (1 + 0.1 * rng.normal(size=x.size))
    fg = flux_model_fn(g) * (1 + 0.1 * rng.normal(size=g.size))
    xt = DataTable()
    sed = fx * x**2 * 1.602176634e-12  # erg/(cm2 s)
    xt["energy"] = u.Quantity(x, "eV")
    xt["flux"] = u.Quantity(sed, "erg/(cm2 s)")
    xt["flux_error"] = u.Quantity(0.1 * np.abs(sed), "erg/(cm2 s)")
    gt = DataTable(meta={"keywords": {"cl": {"value": 0.95}}})
    gt["energy"] = u.Quantity(g / TeV, "TeV")
    gt["flux"] = u.Quantity(fg * TeV, "1/(cm2 s TeV)")
    gt["flux_error"] = u.Quantity(0.1 * np.abs(fg) * TeV, "1/(cm2 s TeV)")
    ul = np.zeros(g.size, dtype=int)
    ul[-1] = 1
    gt["ul"] = ul
    return xt, gt


def c3_device_flux(E_eV, pars=C3_PTRUE):
    """Model flux at p_true from the device path (used to synthesise the data)."""
    u = _u()
    out = c3_model(np.asarray(pars, dtype=float), {"energy": u.Quantity(E_eV, "eV")})
    return out[0].to("1/(s cm2 eV)").value


def walkers(p_true, W, seed=SEED, spread=0.1):
    """The reference's initial ball (core.py:477-481)."""
    rng = np.random.default_rng(seed + 1)
    return p_true * (1 + spread * rng.normal(size=(W, len(p_true))))


# --- oracle-side description of the same model (CPU baseline only) --------------------
def c3_oracle(o):
    """(model, prior) for oracle.naima_oracle.lnprob on a tests/helpers.oracle_data dict."""
    seeds = list(C3_SEEDS)

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


def oracle_data(data):
    """Plain-float view of a validated data table (model values are
    1/(s cm2 eV) * unit_fac -> the table's flux unit)."""
    u = _u()
    E = u.Quantity(data["energy"])
    fl = u.Quantity(data["flux"])
    E_eV = E.to("eV").value
    if fl.unit.physical_type == "flux":
        fac = (u.Quantity(E_eV**2, "eV2") * u.Quantity(1.0, "1/(s cm2 eV)")).to(fl.unit).value
    else:
        fac = u.Quantity(np.ones(E_eV.size), "1/(s cm2 eV)").to(fl.unit).value
    return dict(E_eV=E_eV, unit_fac=fac, flux=fl.value,
                flux_error_lo=u.Quantity(data["flux_error_lo"]).to(fl.unit).value,
                flux_error_hi=u.Quantity(data["flux_error_hi"]).to(fl.unit).value,
                ul=np.asarray(data["ul"], dtype=bool), cl=np.asarray(data["cl"], dtype=float))
