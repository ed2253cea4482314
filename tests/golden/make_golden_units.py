#!/usr/bin/env python
"""Golden vectors from the reference's UNIT-BOUND hot-path functions.

    python tests/golden/make_golden_units.py        (build container only)

make_golden.py executes the reference's unit-free numerics straight from its source
text.  The functions here (``Synchrotron._spectrum`` radiative.py:282-342,
``InverseCompton._iso_ic_on_monochromatic`` / ``_calc_specic`` :609-687, all of
``Bremsstrahlung`` :838-989) do their arithmetic on astropy Quantities, and astropy is not
installable here -- so their bodies are compiled from the reference source (``ast``; no
reference source is copied into this repository) and executed with a stand-in for
``astropy.units`` / ``astropy.constants``: naima_b200.units (Quantity algebra: value x unit
bookkeeping only) and the CODATA-2018 numbers astropy >= 6.1 ships.  Every floating point
operation on the VALUES is the reference's own code in the reference's own order; the
stand-in contributes unit-conversion factors (products of the same constants).

Inputs that the reference derives through more unit machinery (the Lorentz-factor grid,
the particle density on it) are fed from the oracle's electron_grid / reference-exec'd
``eval`` -- both already pinned by ref_exec.npz and the reference's IC goldens.

Output: ref_exec_units.npz (per-energy spectra, cross sections and IC kernels);
tests/test_oracle_golden.py replays it against oracle/naima_oracle.py, the GPU tests
against the CUDA path.
"""
import ast
import logging
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from make_golden import _compile_nodes, _module_ast, load_functions, load_static_methods  # noqa: E402

import oracle.naima_oracle as o  # noqa: E402
from naima_b200 import units as u  # noqa: E402


def load_class(relpath, cls, keep, new_name, ns):
    """Compile the methods `keep` of reference class `cls` as a base-less class."""
    tree = _module_ast(relpath)
    (cnode,) = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls]
    body = [n for n in cnode.body if isinstance(n, ast.FunctionDef) and n.name in keep]
    assert len(body) == len(keep), (cls, [n.name for n in body])
    cnode.body, cnode.bases, cnode.name = body, [], new_name
    cnode.decorator_list = []
    _compile_nodes([cnode], ns)


def namespace():
    ns = {"np": np, "warnings": warnings, "u": u, "log": logging.getLogger("ref-exec")}
    load_functions("src/naima/utils.py", ["trapz_loglog"], ns)
    load_functions("src/naima/radiative.py", ["G12", "G34", "heaviside"], ns)
    # astropy.constants (CODATA 2018) as the reference composes them (radiative.py:34-40)
    c = u.Quantity(29979245800.0, "cm/s")
    m_e = u.Quantity(9.1093837015e-28, "g")
    hbar = u.Quantity(1.0545718176461565e-27, "erg s")
    e = types.SimpleNamespace(value=4.803204712570263e-10)  # e.gauss
    sigma_sb = u.Quantity(5.6703744191844314e-05, "erg/(cm2 s K4)")
    mec2 = (m_e * c**2).cgs
    ns.update(c=c, m_e=m_e, hbar=hbar, e=e, alpha=0.0072973525693, mec2=mec2,
              mec2_unit=u.Unit("mec2"), ar=(4 * sigma_sb / c).to("erg/(cm3 K4)"),
              r0=u.Quantity(e.value**2 / mec2.value, "cm"),
              _validate_ene=lambda ene: u.Quantity(ene))
    load_class("src/naima/radiative.py", "Synchrotron", ["_spectrum"], "RefSynchrotron", ns)
    load_class("src/naima/radiative.py", "InverseCompton",
               ["_iso_ic_on_planck", "_ani_ic_on_planck", "_iso_ic_on_monochromatic",
                "_calc_specic"], "RefIC", ns)
    load_class("src/naima/radiative.py", "Bremsstrahlung",
               ["_sigma_1", "_sigma_2", "_sigma_ee_rel", "_F", "_sigma_ee_nonrel", "_sigma_ee",
                "_sigma_ep", "_emiss_ee", "_emiss_ep", "_spectrum"], "RefBrems", ns)
    for cls in ("ExponentialCutoffPowerLaw", "BrokenPowerLaw"):
        load_static_methods("src/naima/models.py", cls, ["eval"], ns)
    return ns


def main():
    ns = namespace()
    out = {}
    mec2_eV = ns["mec2"].to("eV").value
    assert mec2_eV == o.mec2_eV

    def grid_and_density(Eemin_eV, Eemax_eV, nEed, which):
        gam = o.electron_grid(Eemin_eV, Eemax_eV, nEed)
        e_eV = gam * mec2_eV
        if which == "ecpl":
            n = ns["ExponentialCutoffPowerLaw_eval"](e_eV, 1.3e33, 1e13, 2.41, 4.8e13, 1.0)
        else:
            n = ns["BrokenPowerLaw_eval"](e_eV, 2e30, 2e13, 1e12, 1.5, 2.5)
        return gam, n * mec2_eV  # per unit Lorentz factor (radiative.py:156-160)

    # ---- Synchrotron._spectrum: per-energy spectra [1/(s eV)] -----------------------------
    E_syn = np.logspace(-7, 7, 57)
    for tag, which in (("ecpl", "ecpl"), ("bpl", "bpl")):
        gam, nelec = grid_and_density(1e9, 1e9 * mec2_eV, 100, which)
        out["syn_gam_" + tag], out["syn_nelec_" + tag] = gam, nelec
        for Bname, B in (("3uG", 3.24e-6), ("1mG", 1e-3)):
            obj = object.__new__(ns["RefSynchrotron"])
            obj.B, obj._gam, obj._nelec = u.Quantity(B, "G"), gam, nelec
            with np.errstate(all="ignore"):
                spec = obj._spectrum(u.Quantity(E_syn, "eV"))
            assert spec.unit.to_string() == u.Unit("1/(s eV)").to_string()
            out["syn_spec_%s_%s" % (tag, Bname)] = np.asarray(spec.value)
    out["syn_E_eV"] = E_syn

    # ---- InverseCompton on monochromatic / tabulated seeds -----------------------------------
    gam, nelec = grid_and_density(1e11, 1e15, 60, "ecpl")
    E_ic = np.logspace(8, 14.5, 23)
    Eph = (u.Quantity(E_ic, "eV") / ns["mec2"]).decompose().value
    out["ic_gam"], out["ic_nelec"], out["ic_E_eV"], out["icm_Eph"] = gam, nelec, E_ic, Eph
    IC = ns["RefIC"]
    seed_E = np.logspace(-4, 1, 31)  # eV
    seed_n = 3e2 * seed_E ** -1.3 * np.exp(-seed_E / 2.0)  # 1/(eV cm3)
    seed_n[-2:] = 0.0  # a seed field that runs out: zero nodes in the inner integral
    out["icm_seed_E_eV"], out["icm_seed_n"] = seed_E, seed_n
    with np.errstate(all="ignore"):
        out["icm_mono"] = IC._iso_ic_on_monochromatic(
            gam, u.Quantity([0.00235], "eV"), u.Quantity([0.261], "eV/cm3"), Eph)
        out["icm_array"] = IC._iso_ic_on_monochromatic(
            gam, u.Quantity(seed_E, "eV"), u.Quantity(seed_n, "1/(eV cm3)"), Eph)
    obj = object.__new__(IC)
    obj._gam, obj._nelec = gam, nelec
    obj.seed_photon_fields = {
        "mono": {"type": "array", "energy": u.Quantity([0.00235], "eV"),
                 "photon_density": u.Quantity([0.261], "eV/cm3")},
        "tab": {"type": "array", "energy": u.Quantity(seed_E, "eV"),
                "photon_density": u.Quantity(seed_n, "1/(eV cm3)")},
        "FIR": {"type": "thermal", "isotropic": True, "T": u.Quantity(26.5, "K"),
                "u": u.Quantity(0.415, "eV/cm3")},
        "star": {"type": "thermal", "isotropic": False, "T": u.Quantity(25000.0, "K"),
                 "u": u.Quantity(3.0, "eV/cm3"), "theta": u.Quantity(2.1, "rad")},
    }
    for name in obj.seed_photon_fields:
        with np.errstate(all="ignore"):
            spec = obj._calc_specic(name, u.Quantity(E_ic, "eV"))
        out["ic_specic_" + name] = np.asarray(spec.to("1/(s eV)").value)

    # ---- Bremsstrahlung ------------------------------------------------------------------------
    gam, nelec = grid_and_density(1e8, 1e9 * mec2_eV, 40, "ecpl")
    E_br = np.logspace(5.6, 13, 19)  # no node of the electron grid hit exactly
    eps = (u.Quantity(E_br, "eV") / ns["mec2"]).decompose().value
    out["br_gam"], out["br_nelec"], out["br_E_eV"], out["br_eps"] = gam, nelec, E_br, eps
    B = ns["RefBrems"]
    g2 = np.vstack(gam)
    bobj = object.__new__(B)
    bobj._gam, bobj._nelec = gam, nelec
    bobj.n0 = u.Quantity(3.0, "1/cm3")
    Y, Z = np.array([1.0, 9.59e-2]), np.array([1, 2])
    X = Y / np.sum(Y)
    bobj.weight_ee, bobj.weight_ep = np.sum(Z * X), np.sum(Z**2 * X)  # radiative.py:829-834
    with np.errstate(all="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out["br_sigma_1"] = np.asarray(B._sigma_1(g2, eps).to("cm2/mec2").value)
        out["br_sigma_2"] = np.asarray(B._sigma_2(g2, eps).to("cm2/mec2").value)
        see = bobj._sigma_ee(g2, u.Quantity(E_br, "eV"))
        out["br_sigma_ee"] = np.asarray(see.to("cm2/eV").value)
        spec = bobj._spectrum(u.Quantity(E_br, "eV"))
    out["br_spec"] = np.asarray(spec.to("1/(s eV)").value)

    np.savez_compressed(os.path.join(HERE, "ref_exec_units.npz"), **out)
    for k, v in out.items():
        print("%-22s %s" % (k, np.shape(v)))


def save_run_layout():
    """The HDF5 layout the reference's save_run writes (analysis.py:366-471), read off its
    source: group, dataset names and attribute names.  h5py is not installable here, so the
    npz mirror of naima_b200.analysis.save_run is checked against these names."""
    import json

    tree = _module_ast("src/naima/analysis.py")
    (fn,) = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "save_run"]
    groups, datasets, attrs, table_paths = [], [], [], []

    def text(node):
        if isinstance(node, ast.Constant) and isinstance(node.value, str):
            return node.value
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) \
                and node.func.attr == "format" and isinstance(node.func.value, ast.Constant):
            return node.func.value.value  # "blob{0}".format(idx) -> "blob{0}"
        return None

    for node in ast.walk(fn):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute):
            if node.func.attr == "create_group":
                groups.append(text(node.args[0]))
            elif node.func.attr == "create_dataset":
                datasets.append(text(node.args[0]))
        if isinstance(node, ast.Call) and getattr(node.func, "id", "") == "write_table_hdf5":
            for kw in node.keywords:
                if kw.arg == "path":
                    table_paths.append(text(kw.value))
        if isinstance(node, ast.Subscript) and isinstance(node.value, ast.Attribute) \
                and node.value.attr == "attrs" and isinstance(node.ctx, ast.Store):
            t = text(node.slice)
            attrs.append(t if t is not None else "<run_info key>")
    layout = {"group": groups, "datasets": sorted(set(datasets)), "attrs": sorted(set(attrs)),
              "table_path": table_paths,
              "cite": "src/naima/analysis.py:366-471 (zblz/naima @ ba20a64)"}
    with open(os.path.join(HERE, "save_run_layout.json"), "w") as f:
        json.dump(layout, f, indent=1)
    print(layout)


if __name__ == "__main__":
    main()
    save_run_layout()
