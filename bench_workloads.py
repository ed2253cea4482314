# -*- coding: utf-8 -*-
"""Synthetic workloads of BASELINE.json's five configurations:
seeded fake spectra of the named shapes and the example model functions written against
the naima_b200 API exactly as a user of the reference would write them (same source as
examples/*.py).  Bench / test fixtures, not part of the product package; the plain-float
descriptions of the same models for the CPU arm live in oracle/bench_models.py.
No files are read; `data: "synthetic"` in bench.py means this.
"""
import numpy as np

SEED = 20261017
TeV = 1e12


def _u():
    from naima_b200 import units as u
    return u


# --- C1: Synchrotron + ExponentialCutoffPowerLaw, 64 photon energies, one flux() call ----
C1_PARS = (1e36, 1.0, 2.0, 13.0, 30.0)  # amplitude [1/eV], e_0 [TeV], alpha, e_cutoff [TeV], B [uG]


def c1_energies():
    return np.logspace(0, 6, 64)  # eV: radio .. hard X-rays


def c1_flux(E_eV, B_uG=C1_PARS[4]):
    """One reference-style call: tests/test_models.py:67-103 shapes."""
    u = _u()
    from naima_b200.models import ExponentialCutoffPowerLaw, Synchrotron

    ECPL = ExponentialCutoffPowerLaw(C1_PARS[0] / u.eV, C1_PARS[1] * u.TeV, C1_PARS[2],
                                     C1_PARS[3] * u.TeV)
    return Synchrotron(ECPL, B=B_uG * u.uG).flux(u.Quantity(E_eV, "eV"), distance=1 * u.kpc)


# --- C2: RXJ1713_IC: IC on the CMB, 32-node particle grid ---------------------------------
C2_PTRUE = np.array([1.37e32, 2.58, np.log10(50.2)])  # docs/_static/RXJ1713_IC_results.ecsv ML
C2_NEED = 8.7  # int(8.7 * log10(1e9 mec2 / 100 GeV)) = 32 nodes (BASELINE "32-point grid")


def c2_model(pars, data):
    """examples/RXJ1713_IC.py:16-49 with BASELINE's single CMB seed and 32-node grid."""
    u = _u()
    from naima_b200.models import ExponentialCutoffPowerLaw, InverseCompton

    amplitude = pars[0] / u.eV
    alpha = pars[1]
    e_cutoff = (10 ** pars[2]) * u.TeV
    ECPL = ExponentialCutoffPowerLaw(amplitude, 10.0 * u.TeV, alpha, e_cutoff)
    IC = InverseCompton(ECPL, seed_photon_fields=["CMB"], Eemin=100 * u.GeV, nEed=C2_NEED)
    model = IC.flux(data, distance=1.0 * u.kpc).to(data["flux"].unit)
    return model, IC.compute_We(Eemin=1 * u.TeV)


def c2_prior(pars):
    from naima_b200.core import uniform_prior

    return uniform_prior(pars[0], 0.0, np.inf) + uniform_prior(pars[1], -1, 5)


def c2_energies():
    return (np.logspace(np.log10(0.33 * TeV), np.log10(170 * TeV), 28),)


# --- C3: RXJ1713_SynIC, joint Synchrotron + IC (CMB + FIR + NIR) ----------------------
C3_PTRUE = np.array([33.0, 2.5, np.log10(48.0), 20.0])  # examples/RXJ1713_SynIC.py:72 (+B)
C3_SEEDS = ("CMB", "FIR", "NIR")


def c3_model(pars, data):
    """examples/RXJ1713_SynIC.py:19-46 with BASELINE's three seed fields."""
    u = _u()
    from naima_b200.models import ExponentialCutoffPowerLaw, InverseCompton, Synchrotron

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
    from naima_b200.core import uniform_prior

    return (uniform_prior(pars[0], 0.0, np.inf) + uniform_prior(pars[1], -1, 5)
            + uniform_prior(pars[3], 0, np.inf))


def c3_energies():
    """36 X-ray + 28 VHE photon energies (N_E = 64), eV."""
    x = np.logspace(np.log10(0.55e3), np.log10(10e3), 36)
    g = np.logspace(np.log10(0.33 * TeV), np.log10(170 * TeV), 28)
    return x, g


# --- C4: CrabNebula_SynSSC: Synchrotron + IC(CMB, FIR, NIR, SSC), ECBPL electrons --------
# examples/CrabNebula_SynSSC.py:13-51; free parameters: log10 amplitude, the two indices,
# log10 cutoff and B (the example itself evaluates one SED at fixed values)
C4_PTRUE = np.array([np.log10(3.699e36), 1.5, 3.233, np.log10(1863.0), 125.0])
C4_NSEED = 100
C4_NEED = 100


def c4_model(pars, data, nseed=None, nEed=None):
    u = _u()
    from naima_b200.constants import c
    from naima_b200.models import ExponentialCutoffBrokenPowerLaw, InverseCompton, Synchrotron

    nseed = C4_NSEED if nseed is None else nseed
    nEed = C4_NEED if nEed is None else nEed
    amplitude = 10 ** pars[0] / u.eV
    ECBPL = ExponentialCutoffBrokenPowerLaw(amplitude, 1 * u.TeV, 0.265 * u.TeV, pars[1], pars[2],
                                            (10 ** pars[3]) * u.TeV, 2.0)
    B = pars[4] * u.uG
    eopts = {"Eemax": 50 * u.PeV, "Eemin": 0.1 * u.GeV, "nEed": nEed}
    SYN = Synchrotron(ECBPL, B=B, **eopts)
    # photon density of the synchrotron emission inside a sphere of R = 2.1 pc
    Rpwn = 2.1 * u.pc
    Esy = np.logspace(-7, 9, nseed) * u.eV
    Lsy = SYN.flux(Esy, distance=0 * u.cm)  # distance 0: luminosity
    phn_sy = Lsy / (4 * np.pi * Rpwn**2 * c) * 2.24
    IC = InverseCompton(
        ECBPL,
        seed_photon_fields=["CMB", ["FIR", 70 * u.K, 0.5 * u.eV / u.cm**3],
                            ["NIR", 5000 * u.K, 1 * u.eV / u.cm**3], ["SSC", Esy, phn_sy]],
        **eopts)
    return IC.sed(data, 2 * u.kpc) + SYN.sed(data, 2 * u.kpc)


def c4_prior(pars):
    from naima_b200.core import uniform_prior

    return (uniform_prior(pars[1], -1, 5) + uniform_prior(pars[2], -1, 8)
            + uniform_prior(pars[4], 0, np.inf))


def c4_energies(n=100):
    return (np.logspace(-7, 15, n),)


# --- C5: PionDecay (Kafexhiu+14 LUT) + PowerLaw protons -----------------------------------
C5_PTRUE = np.array([46.0, 2.34])  # examples/model_examples.py:19 without the cutoff


def c5_model(pars, data):
    """examples/model_examples.py:26-45 with a PowerLaw."""
    u = _u()
    from naima_b200.models import PionDecay, PowerLaw

    amplitude = 10 ** pars[0] / u.TeV
    PL = PowerLaw(amplitude, 30 * u.TeV, pars[1])
    PP = PionDecay(PL, nh=1.0 * u.cm**-3)
    model = PP.flux(data, distance=1.0 * u.kpc)
    return model, PP.compute_Wp(Epmin=1 * u.TeV)


def c5_prior(pars):
    from naima_b200.core import uniform_prior

    return uniform_prior(pars[1], -1, 5)


def c5_energies():
    return (np.logspace(8.5, 14, 64),)  # 0.3 GeV .. 100 TeV (Fermi-LAT + IACT range)


# --- registry -------------------------------------------------------------------------------
class Workload:
    """One BASELINE configuration: model/prior callbacks, truth, photon-energy groups."""

    def __init__(self, name, title, model, prior, p_true, energies, walkers_per_gpu, steps,
                 scaling="weak", sed_groups=(), describe="", spread=0.005):
        self.name, self.title, self.spread = name, title, spread
        self.model, self.prior = model, prior
        self.p_true = None if p_true is None else np.asarray(p_true, dtype=float)
        self.energies, self.walkers_per_gpu, self.steps = energies, walkers_per_gpu, steps
        self.scaling, self.sed_groups, self.describe = scaling, tuple(sed_groups), describe

    @property
    def P(self):
        return self.p_true.size

    def total_walkers(self, n_gpus, per_gpu=None):
        """weak: per-GPU count fixed; strong: the configuration's total is fixed."""
        if self.scaling == "strong":
            return self.walkers_per_gpu if per_gpu is None else per_gpu
        return (self.walkers_per_gpu if per_gpu is None else per_gpu) * n_gpus

    def device_flux(self, E_eV, pars=None):
        """Model flux at p_true from the device path (synthesises the data), 1/(s cm2 eV)."""
        u = _u()
        pars = self.p_true if pars is None else np.asarray(pars, dtype=float)
        out = self.model(pars, {"energy": u.Quantity(E_eV, "eV"),
                                "flux": u.Quantity(np.ones(np.size(E_eV)), "1/(s cm2 eV)")})
        out = out[0] if isinstance(out, tuple) else out
        if out.unit.physical_type in ("power", "flux"):  # SED -> differential
            return (out / u.Quantity(E_eV, "eV") ** 2).to("1/(s cm2 eV)").value
        return out.to("1/(s cm2 eV)").value

    def tables(self, flux_model_fn=None, seed=SEED):
        """Fake data tables: flux = model(p_true) (1 + 0.1 N(0,1)), sigma = 0.1 flux, the
        last point of the last group an upper limit at cl = 0.95.  Groups named
        in `sed_groups` are written as SEDs in erg/(cm2 s), the others as 1/(cm2 s TeV).
        flux_model_fn(E_eV) -> 1/(s cm2 eV); default: the device path."""
        u = _u()
        from naima_b200.utils import DataTable

        fn = self.device_flux if flux_model_fn is None else flux_model_fn
        rng = np.random.default_rng(seed)
        groups = self.energies()
        out = []
        for k, E in enumerate(groups):
            f = fn(E) * (1 + 0.1 * rng.normal(size=E.size))
            last = k == len(groups) - 1
            t = DataTable(meta={"keywords": {"cl": {"value": 0.95}}}) if last else DataTable()
            if k in self.sed_groups:
                sed = f * E**2 * 1.602176634e-12  # erg/(cm2 s)
                t["energy"] = u.Quantity(E, "eV")
                t["flux"] = u.Quantity(sed, "erg/(cm2 s)")
                t["flux_error"] = u.Quantity(0.1 * np.abs(sed), "erg/(cm2 s)")
            else:
                t["energy"] = u.Quantity(E / TeV, "TeV")
                t["flux"] = u.Quantity(f * TeV, "1/(cm2 s TeV)")
                t["flux_error"] = u.Quantity(0.1 * np.abs(f) * TeV, "1/(cm2 s TeV)")
            if last:
                ul = np.zeros(E.size, dtype=int)
                ul[-1] = 1
                t["ul"] = ul
            out.append(t)
        return out

    def walkers(self, W, seed=SEED, spread=None):
        return walkers(self.p_true, W, seed=seed, spread=self.spread if spread is None else spread)


def walkers(p_true, W, seed=SEED, spread=0.005):
    """The reference's initial ball (core.py:474-481): p0 + spread * p0 * N(0, 1) with the
    0.5 % spread naima uses when p0 is the maximum-likelihood point (P0_IS_ML) -- the
    synthetic data are generated at p_true, so it is.  The 10 % ball of an unfitted p0 spans
    +-3 decades on a log10(norm) of 33 (+-5 on 46): the ensemble then disperses along the
    flat low-amplitude direction instead of converging, and within a few hundred steps
    proposes amplitudes beyond 1e308 -- inf * 0 in the integrand, a NaN likelihood and
    emcee's "Probability function returned NaN" (observed at 512 walkers; the reference
    would stop the same way)."""
    rng = np.random.default_rng(seed + 1)
    return p_true * (1 + spread * rng.normal(size=(W, len(p_true))))


WORKLOADS = {
    "C2": Workload("C2", "RXJ1713_IC: InverseCompton(CMB) on ExponentialCutoffPowerLaw "
                   "electrons, N_E=28, particle grid 32 nodes, P=3",
                   c2_model, c2_prior, C2_PTRUE, c2_energies, 128, 50),
    "C3": Workload("C3", "RXJ1713_SynIC: Synchrotron + InverseCompton(CMB+FIR+NIR) on "
                   "ExponentialCutoffPowerLaw electrons, N_E=64 (36 X-ray + 28 VHE), "
                   "IC grid 370 nodes, synchrotron grid 570 nodes, P=4",
                   c3_model, c3_prior, C3_PTRUE, c3_energies, 256, 200, sed_groups=(0,)),
    "C4": Workload("C4", "CrabNebula_SynSSC: Synchrotron + InverseCompton(CMB+FIR+NIR+SSC) on "
                   "ExponentialCutoffBrokenPowerLaw electrons, N_E=100, particle grid 869 nodes, "
                   "SSC seed field on 100 synchrotron photon energies per walker, P=5",
                   c4_model, c4_prior, C4_PTRUE, c4_energies, 256, 20, sed_groups=(0,)),
    "C5": Workload("C5", "PionDecay (Kafexhiu+14 Pythia8 LUT, nuclear enhancement) on PowerLaw "
                   "protons, N_E=64, proton grid 691 nodes, P=2; 512 walkers in total",
                   c5_model, c5_prior, C5_PTRUE, c5_energies, 512, 500, scaling="strong"),
}

# C3 helpers under their round-1 names (tools/, tests/multi)
c3_device_flux = WORKLOADS["C3"].device_flux


def c3_tables(flux_model_fn=None, seed=SEED):
    return WORKLOADS["C3"].tables(flux_model_fn, seed)
