# -*- coding: utf-8 -*-
"""TEST / BENCH INFRASTRUCTURE ONLY (like everything under oracle/): plain-float
descriptions of the BASELINE workloads of bench_workloads.py for the CPU oracle
(oracle/naima_oracle.py), used by the parity tests and by bench.py's CPU legs
(`cpu_baseline`, `--impl reference`).  Each (model, prior) pair follows the reference
example cited in bench_workloads.py; `model(p, data)` returns the model in the data
table's flux unit (`data` is oracle_data(validated table))."""
import numpy as np

from . import naima_oracle as o

TeV = 1e12
PC_CM = 3.0856775814913673e18


def c2(nEed=8.7):
    def model(p, data):
        pd = o.PDist("ExponentialCutoffPowerLaw", p[0], 10 * TeV, p[1], 10 ** p[2] * TeV, 1.0)
        spec = o.ic_spectrum(pd, data["E_eV"], ["CMB"], Eemin_eV=100e9, nEed=nEed)
        return o.flux_from_spectrum(spec, o.kpc_cm) * data["unit_fac"]

    def prior(p):
        return o.uniform_prior(p[0], 0.0, np.inf) + o.uniform_prior(p[1], -1, 5)

    return model, prior


def c3(seeds=("CMB", "FIR", "NIR")):
    seeds = list(seeds)

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


def c4(nseed=100, nEed=100):
    """examples/CrabNebula_SynSSC.py:13-51."""
    okw = dict(Eemin_eV=1e8, Eemax_eV=50e15, nEed=nEed)
    Esy = np.logspace(-7, 9, nseed)
    Rpwn = 2.1 * PC_CM

    def model(p, data):
        pd = o.PDist("ExponentialCutoffBrokenPowerLaw", 10 ** p[0], 1 * TeV, 0.265 * TeV, p[1],
                     p[2], 10 ** p[3] * TeV, 2.0)
        B = p[4] * 1e-6
        E = data["E_eV"]
        lsy = o.synchrotron_spectrum(pd, Esy, B, **okw)
        phn = lsy / (4 * np.pi * Rpwn**2 * o.c_cgs) * 2.24
        seeds = ["CMB", ("thermal", 70.0, 0.5 * o.eV_erg), ("thermal", 5000.0, 1.0 * o.eV_erg),
                 ("array", Esy, phn)]
        d = 2 * o.kpc_cm
        ic = o.flux_from_spectrum(o.ic_spectrum(pd, E, seeds, **okw), d)
        sy = o.flux_from_spectrum(o.synchrotron_spectrum(pd, E, B, **okw), d)
        return (ic + sy) * data["unit_fac"]

    def prior(p):
        return (o.uniform_prior(p[1], -1, 5) + o.uniform_prior(p[2], -1, 8)
                + o.uniform_prior(p[4], 0, np.inf))

    return model, prior


_LUT = {}


def reference_lut():
    """dsigma/dEgamma(Ep, Egamma) [GeV] of the reference's Pythia8 + nuclear-enhancement
    table: scipy FITPACK evaluation (bisplev, as RectBivariateSpline.__call__ does,
    radiative.py:1793-1797) of the spline the reference builds from its .npz table --
    knots and coefficients exported once by tools/make_pp_lut_spline.py into
    naima_b200/data/ (the reference tree itself does not travel to the GPU box)."""
    if "f" not in _LUT:
        import os

        from scipy.interpolate import bisplev

        f = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "naima_b200",
                                 "data", "pp_kafexhiu14_pythia8_nucenh_bspline.npz"))
        tck = (f["tx"], f["ty"], f["c"], 3, 3)
        _LUT["f"] = lambda Ep, Eg: np.atleast_1d(
            bisplev(np.log10(Ep), np.log10(Eg), tck)).flatten()
    return _LUT["f"]


def c5():
    def model(p, data):
        pd = o.PDist("PowerLaw", 10 ** p[0] / TeV, 30 * TeV, p[1])
        spec = o.piondecay_spectrum(pd, data["E_eV"], nh=1.0, lut=reference_lut())
        return o.flux_from_spectrum(spec, o.kpc_cm) * data["unit_fac"]

    def prior(p):
        return o.uniform_prior(p[1], -1, 5)

    return model, prior


MODELS = {"C2": c2, "C3": c3, "C4": c4, "C5": c5}


def c1_flux(E_eV, pars):
    """Synchrotron + ECPL single flux() call (C1): 1/(s cm2 eV) at 1 kpc."""
    amp, e0_TeV, alpha, ec_TeV, B_uG = pars
    pd = o.PDist("ExponentialCutoffPowerLaw", amp, e0_TeV * TeV, alpha, ec_TeV * TeV, 1.0)
    return o.flux_from_spectrum(o.synchrotron_spectrum(pd, E_eV, B_uG * 1e-6), o.kpc_cm)


def oracle_data(data):
    """Plain-float view of a validated naima_b200 data table: model values are
    1/(s cm2 eV) * unit_fac -> the table's flux unit."""
    from naima_b200 import units as u

    E = u.Quantity(data["energy"])
    fl = u.Quantity(data["flux"])
    E_eV = E.to("eV").value
    if fl.unit.physical_type == "flux":  # SED: erg/(cm2 s)
        fac = (u.Quantity(E_eV**2, "eV2") * u.Quantity(1.0, "1/(s cm2 eV)")).to(fl.unit).value
    else:
        fac = u.Quantity(np.ones(E_eV.size), "1/(s cm2 eV)").to(fl.unit).value
    return dict(E_eV=E_eV, unit_fac=fac, flux=fl.value,
                flux_error_lo=u.Quantity(data["flux_error_lo"]).to(fl.unit).value,
                flux_error_hi=u.Quantity(data["flux_error_hi"]).to(fl.unit).value,
                ul=np.asarray(data["ul"], dtype=bool), cl=np.asarray(data["cl"], dtype=float))
