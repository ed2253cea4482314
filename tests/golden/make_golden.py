#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the reference itself.

Run ONCE in the build container (``/root/reference`` exists only there):

    python tests/golden/make_golden.py

The reference (zblz/naima) cannot be imported here because astropy/emcee are
absent, but its *unit-free* numeric functions can be executed straight from
its source text: this script parses ``src/naima/{utils,radiative,models}.py``
with ``ast``, compiles ONLY those function/class bodies (no reference source is
copied into this repository) and evaluates them on seeded inputs.  The input
and output vectors are stored as ``ref_exec.npz``; ``tests/test_oracle_golden.py``
replays them against ``oracle/naima_oracle.py``.

Also written:
  * ``reference_goldens.json`` -- the known-answer numbers held by the
    reference's own tests/docs, with file:line provenance;
  * ``rxj1713_data.npz`` -- the RXJ1713 HESS / Suzaku data columns of
    ``examples/*.dat`` (observational data tables, needed for the lnprob pin);
  * ``pp_lut_probe.npz`` -- LookupTable outputs on a probe grid.
"""
import ast
import json
import os
import sys
import warnings

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _module_ast(relpath):
    with open(os.path.join(REF, relpath)) as f:
        return ast.parse(f.read())


def _compile_nodes(nodes, ns):
    mod = ast.Module(body=nodes, type_ignores=[])
    ast.fix_missing_locations(mod)
    exec(compile(mod, "<reference-slice>", "exec"), ns)


def load_functions(relpath, names, ns):
    tree = _module_ast(relpath)
    nodes = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert len(nodes) == len(names), (relpath, names)
    _compile_nodes(nodes, ns)


def load_static_methods(relpath, cls, names, ns):
    """Compile staticmethods of a class as plain functions ``<cls>_<name>``."""
    tree = _module_ast(relpath)
    (cnode,) = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls]
    nodes = []
    for n in cnode.body:
        if isinstance(n, ast.FunctionDef) and n.name in names:
            n.decorator_list = []
            n.name = "%s_%s" % (cls, n.name)
            nodes.append(n)
    assert len(nodes) == len(names), (cls, names)
    _compile_nodes(nodes, ns)


def load_piondecay(ns):
    """PionDecay numerics (radiative.py:1175-1482) as a stand-alone class."""
    tree = _module_ast("src/naima/radiative.py")
    (cnode,) = [
        n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "PionDecay"
    ]
    body = []
    for n in cnode.body:
        if isinstance(n, ast.FunctionDef) and n.name in (
            "__init__",
            "_loadLUT",
            "_spectrum",
        ):
            continue
        if isinstance(n, ast.Expr):  # docstring
            continue
        if (
            isinstance(n, ast.Assign)
            and isinstance(n.targets[0], ast.Name)
            and n.targets[0].id == "_m_p"
        ):
            # (m_p * c**2).to("GeV").value with CODATA 2018
            n = ast.parse("_m_p = 0.9382720881604903").body[0]
        body.append(n)
    cnode.body = body
    cnode.bases = []
    cnode.name = "RefPionDecay"
    _compile_nodes([cnode], ns)


def main():
    ns = {"np": np, "warnings": warnings}
    load_functions("src/naima/utils.py", ["trapz_loglog"], ns)
    load_functions("src/naima/radiative.py", ["G12", "G34", "heaviside"], ns)
    load_static_methods(
        "src/naima/radiative.py",
        "InverseCompton",
        ["_iso_ic_on_planck", "_ani_ic_on_planck"],
        ns,
    )
    for cls in (
        "PowerLaw",
        "ExponentialCutoffPowerLaw",
        "BrokenPowerLaw",
        "ExponentialCutoffBrokenPowerLaw",
        "LogParabola",
    ):
        load_static_methods("src/naima/models.py", cls, ["eval"], ns)
    load_piondecay(ns)

    rng = np.random.default_rng(20261017)
    out = {}

    # ---- trapz_loglog (utils.py:285-355), incl. zero / negative / b=-1 cases
    x = np.logspace(0.3, 6.1, 37)
    y = np.exp(rng.normal(size=(5, 37))) * x**-2.2
    y[1, 10:13] = 0.0  # zero nodes
    y[2, 20] = -y[2, 20]  # sign change -> NaN slope -> log branch
    y[3] = 3.0 / x  # local index exactly -1
    y[4, -5:] = 0.0
    out["tl_x"] = x
    out["tl_y"] = y
    with np.errstate(all="ignore"):
        out["tl_sum"] = ns["trapz_loglog"](y, x)
        out["tl_int"] = ns["trapz_loglog"](y, x, intervals=True)
        out["tl_axis0"] = ns["trapz_loglog"](y.T.copy(), x, axis=0)

    # ---- particle distributions (models.py eval staticmethods)
    e = np.logspace(8, 15.5, 64)
    out["pd_e"] = e
    out["pd_pl"] = ns["PowerLaw_eval"](e, 1.3e33, 1e13, 2.41)
    out["pd_ecpl"] = ns["ExponentialCutoffPowerLaw_eval"](e, 1.3e33, 1e13, 2.41, 4.8e13, 1.0)
    out["pd_ecpl_b2"] = ns["ExponentialCutoffPowerLaw_eval"](e, 1.3e33, 1e13, 1.7, 2e12, 2.0)
    out["pd_bpl"] = ns["BrokenPowerLaw_eval"](e, 2e30, 2e13, 1e12, 1.5, 2.5)
    out["pd_ecbpl"] = ns["ExponentialCutoffBrokenPowerLaw_eval"](
        e, 3.7e36, 1e12, 2.65e11, 1.5, 3.233, 1.863e15, 2.0
    )
    out["pd_lp"] = ns["LogParabola_eval"](e, 1e30, 2e13, 1.7, 0.2)

    # ---- Khangulyan kernels (radiative.py:345-367, 547-607)
    xx = np.logspace(-6, 3, 50)
    out["g_x"] = xx
    out["g12_a1"] = ns["G12"](xx, [0.857, 0.153, 1.840, 0.254])
    out["g34_a3"] = ns["G34"](xx, [0.606, 0.443, 1.481, 0.540, 0.319])
    gam = np.logspace(np.log10(0.5), 9, 120)
    Eph = np.logspace(-8, 8.5, 40)
    out["ic_gam"] = gam
    out["ic_Eph"] = Eph
    with np.errstate(all="ignore"):
        out["ic_iso_cmb"] = ns["InverseCompton__iso_ic_on_planck"](gam, 2.72548, Eph)
        out["ic_iso_nir"] = ns["InverseCompton__iso_ic_on_planck"](gam, 3000.0, Eph)
        out["ic_ani_60"] = ns["InverseCompton__ani_ic_on_planck"](
            gam, 20000.0, Eph, np.deg2rad(60.0)
        )
        out["ic_ani_135"] = ns["InverseCompton__ani_ic_on_planck"](
            gam, 30.0, Eph, np.deg2rad(135.0)
        )
    out["heaviside"] = ns["heaviside"](np.array([-2.0, -0.0, 0.0, 3.0]))

    # ---- PionDecay numerics (radiative.py:1215-1482)
    Ep = np.logspace(np.log10(0.9382720881604903 + 0.27966184 + 1e-4), 7, 150)
    out["pp_Ep"] = Ep
    Egs = np.array([1e-2, 0.1, 0.7, 3.0, 50.0, 1e3, 1e5])
    out["pp_Eg"] = Egs
    for model in ("Pythia8", "Geant4", "SIBYLL", "QGSJET"):
        for nuc in (True, False):
            pp = ns["RefPionDecay"]()
            pp.hiEmodel = model
            pp.nuclear_enhancement = nuc
            with np.errstate(all="ignore"):
                ds = np.array([pp._diffsigma(Ep, eg) for eg in Egs])
            out["pp_ds_%s_%d" % (model, nuc)] = ds
    pp = ns["RefPionDecay"]()
    pp.hiEmodel = "Pythia8"
    Tp = Ep - 0.9382720881604903
    with np.errstate(all="ignore"):
        out["pp_sigma_inel"] = pp._sigma_inel(Tp)
        out["pp_sigma_pi"] = pp._sigma_pi(Tp)
        out["pp_Amax"] = pp._Amax(Tp)
        out["pp_Egmax"] = pp._calc_Egmax(Tp)
        out["pp_nuc"] = pp._nuclear_factor(Tp)

    np.savez_compressed(os.path.join(HERE, "ref_exec.npz"), **out)

    # ---- LookupTable probe (radiative.py:1770-1797) on the packaged LUT
    from scipy.interpolate import RectBivariateSpline

    lutf = np.load(
        os.path.join(REF, "src/naima/data/PionDecayKafexhiu14_LUT_NucEnh_Pythia8.npz")
    )
    spl = RectBivariateSpline(lutf["X"], lutf["Y"], 10 ** lutf["lut"], kx=3, ky=3, s=0)
    Ep_probe = np.logspace(np.log10(1.2179), 7.2, 97)  # incl. clamped region
    Eg_probe = np.logspace(-2.5, 6.3, 41)
    vals = np.array([spl(np.log10(Ep_probe), np.log10(eg)).flatten() for eg in Eg_probe])
    np.savez_compressed(
        os.path.join(HERE, "pp_lut_probe.npz"), Ep=Ep_probe, Eg=Eg_probe, ds=vals
    )

    # ---- data tables of the RXJ1713 examples
    def read_ipac(path):
        rows = [
            line.split()
            for line in open(path)
            if line.strip() and line[0] not in "\\|"
        ]
        return np.array(rows, dtype=float)

    hess = read_ipac(os.path.join(REF, "examples/RXJ1713_HESS_2007.dat"))
    suzaku = read_ipac(os.path.join(REF, "examples/RXJ1713_Suzaku-XIS.dat"))
    np.savez_compressed(
        os.path.join(HERE, "rxj1713_data.npz"),
        hess_energy_TeV=hess[:, 0],
        hess_flux=hess[:, 3],
        hess_flux_error=hess[:, 4],
        hess_ul=hess[:, 5].astype(np.int64),
        hess_cl=np.float64(0.95),
        suzaku_energy_eV=suzaku[:, 0],
        suzaku_flux=suzaku[:, 1],
        suzaku_flux_error=suzaku[:, 2],
    )

    goldens = {
        "_provenance": "numbers held by the reference's own tests/docs "
        "(zblz/naima @ ba20a64); all asserted there with rtol=1e-7",
        "fixture": {
            "cite": "tests/test_models.py:30-40",
            "e_0_TeV": 20,
            "e_cutoff_TeV": 10,
            "alpha": 2.0,
            "e_break_TeV": 1,
            "alpha_1": 1.5,
            "alpha_2": 2.5,
            "Eemin_GeV": 100,
            "Eemax_PeV": 1,
            "Epmax_PeV": 1,
            "energy": "logspace(0,15,1000) eV",
        },
        "synchrotron_lum": {
            "cite": "tests/test_models.py:76-80",
            "value": [0.00025231296225663107, 0.03316715765695228, 0.00044597089198025806],
        },
        "We": {
            "cite": "tests/test_models.py:81",
            "value": [5064124672.902273, 11551172166.866821, 926633861.2898524],
        },
        "synchrotron_lum_B1G": {"cite": "tests/test_models.py:103", "value": 31374131.90312505},
        "bremsstrahlung_lum": {"cite": "tests/test_models.py:194", "value": 2.3064095039069847e-05},
        "ic_lum": {
            "cite": "tests/test_models.py:206-210",
            "value": [0.0002782201669858555, 0.004821189222961136, 0.00012916582897424096],
        },
        "ic_lum_3seeds": {"cite": "tests/test_models.py:226", "value": 0.0005833030059049264},
        "ic_aniso_lum": {
            "cite": "tests/test_models.py:237-239",
            "angles_deg": [45, 90, 135],
            "value": [48901.363932, 111356.423781, 149800.235776],
        },
        "pp_lum_LUT": {
            "cite": "tests/test_models.py:401",
            "value": [9.94070311e-13, 2.30256683e-12, 1.57263936e-13],
        },
        "pp_lum_noLUT": {
            "cite": "tests/test_models.py:403",
            "value": [9.94144387e-13, 2.30264140e-12, 1.57272216e-13],
        },
        "Wp": {
            "cite": "tests/test_models.py:405",
            "value": [5406.36160963, 8727.55086557, 554.13864492],
        },
        "pp_lum_no_nuc": {"cite": "tests/test_models.py:442", "value": 5.693100769654807e-13},
        "lnprob_RXJ1713_IC": {
            "cite": "docs/_static/RXJ1713_IC_results.ecsv:10-11, model docs/_static/RXJ1713_IC.py:16-62",
            "ML_pars": [1.3697204402514948e32, 2.5839150825284958, 1.7002798209378214],
            "MaxLogLikelihood": -17.98655803890747,
        },
    }
    with open(os.path.join(HERE, "reference_goldens.json"), "w") as f:
        json.dump(goldens, f, indent=1)
    print("wrote fixtures to", HERE)


if __name__ == "__main__":
    sys.exit(main())
