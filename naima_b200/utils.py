# -*- coding: utf-8 -*-
"""Data-table glue of the likelihood path (host side; mirrors the subset of
``naima.utils`` that the sampler boundary needs: utils.py:38-355).

astropy tables are not available in this image; a data table is a
:class:`DataTable` (a dict of columns -- Quantity or ndarray -- with ``meta``),
which is also what ``model(pars, data)`` receives.  `read_ipac` reads the IPAC
ASCII files the reference's examples ship.
"""
import ast
import logging
import re

import numpy as np

from . import engine as eng
from . import units as u
from .units import Quantity

__all__ = ["DataTable", "read_ipac", "validate_data_table", "sed_conversion", "trapz_loglog",
           "generate_energy_edges", "build_data_table"]

log = logging.getLogger("naima_b200.utils")


class DataTable(dict):
    """Minimal column table: ``t['flux']`` is a Quantity (or ndarray for flags)."""

    def __init__(self, *a, meta=None, **k):
        super().__init__(*a, **k)
        self.meta = {} if meta is None else meta

    @property
    def colnames(self):
        return list(self.keys())

    def __len__(self):
        for v in self.values():
            return len(v)
        return 0

    def copy(self):
        return DataTable({k: (v.copy() if hasattr(v, "copy") else v) for k, v in self.items()},
                         meta=dict(self.meta))

    def __getitem__(self, key):
        if isinstance(key, str):
            return dict.__getitem__(self, key)
        # row selection (slice / index array / mask)
        return DataTable({k: v[key] for k, v in self.items()}, meta=dict(self.meta))


def _is_table(t):
    return isinstance(t, DataTable) or (hasattr(t, "colnames") and hasattr(t, "meta"))


def read_ipac(path):
    """Read an IPAC ASCII table (header ``|name|``, ``|type|``, ``|unit|`` lines,
    ``\\key=value`` keywords) into a DataTable."""
    meta = {"keywords": {}}
    header, rows = [], []
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if not line.strip():
                continue
            if line.startswith("\\"):
                m = re.match(r"^\\(\w+)\s*=\s*(.+)$", line)
                if m:
                    val = m.group(2).strip()
                    try:
                        val = ast.literal_eval(val)
                    except Exception:
                        pass
                    meta["keywords"][m.group(1)] = {"value": val}
                continue
            if line.startswith("|"):
                header.append([c.strip() for c in line.strip().strip("|").split("|")])
                continue
            rows.append(line.split())
    names = header[0]
    types = header[1] if len(header) > 1 else ["double"] * len(names)
    units = header[2] if len(header) > 2 else [""] * len(names)
    t = DataTable(meta=meta)
    for k, name in enumerate(names):
        col = [r[k] for r in rows]
        if types[k].startswith(("int", "long")):
            t[name] = np.array(col, dtype=int)
        elif types[k].startswith(("double", "float", "real")):
            arr = np.array(col, dtype=float)
            t[name] = Quantity(arr, units[k]) if units[k] else arr
        else:
            t[name] = np.array(col)
    return t


def trapz_loglog(y, x, axis=-1, intervals=False):
    """Composite trapezoid in log-log space (utils.py:285-355), evaluated on the
    device in the reference's operation order."""
    yu = Quantity(y).unit if u._is_quantity(y) else None
    xu = Quantity(x).unit if u._is_quantity(x) else None
    yv = np.asarray(Quantity(y).value if yu is not None else y, dtype=float)
    xv = np.asarray(Quantity(x).value if xu is not None else x, dtype=float)
    if xv.ndim != 1 and xv.shape != yv.shape:
        raise ValueError("x must be 1-d or have the shape of y")
    ym = np.moveaxis(yv, axis, -1)
    lead = ym.shape[:-1]
    N = ym.shape[-1]
    y2 = ym.reshape(-1, N)
    x2 = xv if xv.ndim == 1 else np.moveaxis(xv, axis, -1).reshape(-1, N)
    res = eng.trapz_loglog(y2, x2, intervals=intervals)
    if intervals:
        res = np.moveaxis(res.reshape(lead + (N - 1,)), -1, axis)
    else:
        res = res.reshape(lead) if lead else float(res[0])
    if yu is not None or xu is not None:
        return Quantity(res, (yu or u.Unit()) * (xu or u.Unit()))
    return res


def _validate_column(data_table, key, pt, domain=None):
    try:
        column = data_table[key]
        if not u._is_quantity(column):
            raise TypeError("Column {0} has no unit".format(key))
        column = Quantity(column)
        pts = [pt] if isinstance(pt, str) else pt
        if column.unit.physical_type not in pts:
            raise TypeError("{0} should be given in units of {1}".format(key, ", ".join(pts)))
        if np.ndim(column.value) != 1:
            raise TypeError("{0} should be a 1-d sequence".format(key))
        if domain == "positive" and np.any(column.value < 0):
            raise ValueError("{0} should be positive".format(key))
    except KeyError:
        raise TypeError('Data table does not contain required column "{0}"'.format(key))
    return column


def _generate_energy_edges_single(ene):
    """utils.py:358-366."""
    v = ene.value
    midene = np.sqrt(v[1:] * v[:-1])
    elo, ehi = np.zeros(len(v)), np.zeros(len(v))
    elo[1:] = v[1:] - midene
    ehi[:-1] = midene - v[:-1]
    elo[0] = v[0] * (1 - v[0] / (v[0] + ehi[0]))
    ehi[-1] = elo[-1]
    return Quantity(np.array([elo, ehi]), ene.unit)


def generate_energy_edges(ene, groups=None):
    """utils.py:369-398."""
    ene = Quantity(ene)
    if groups is None or len(ene) != len(groups):
        return _generate_energy_edges_single(ene)
    eloehi = np.zeros((2, len(ene)))
    groups = np.asarray(groups)
    for g in np.unique(groups):
        eloehi[:, groups == g] = _generate_energy_edges_single(ene[groups == g]).value
    return Quantity(eloehi, ene.unit)


def _validate_single_data_table(data_table, group=0):
    """utils.py:108-213."""
    data = DataTable()
    flux_types = ["flux", "differential flux", "power", "differential power"]
    data["energy"] = _validate_column(data_table, "energy", "energy")
    data["flux"] = _validate_column(data_table, "flux", flux_types)
    keys = list(data_table.keys())
    if "flux_error" in keys:
        dflux = _validate_column(data_table, "flux_error", flux_types)
        data["flux_error_lo"] = dflux
        data["flux_error_hi"] = dflux
    elif "flux_error_lo" in keys and "flux_error_hi" in keys:
        data["flux_error_lo"] = _validate_column(data_table, "flux_error_lo", flux_types)
        data["flux_error_hi"] = _validate_column(data_table, "flux_error_hi", flux_types)
    else:
        raise TypeError('Data table does not contain required column "flux_error" or columns '
                        '"flux_error_lo" and "flux_error_hi"')
    n = len(data["energy"])
    data["group"] = np.asarray(data_table["group"]) if "group" in keys else np.full(n, group)
    if "energy_width" in keys:
        w = _validate_column(data_table, "energy_width", "energy")
        data["energy_error_lo"] = w / 2.0
        data["energy_error_hi"] = w / 2.0
    elif "energy_error" in keys:
        w = _validate_column(data_table, "energy_error", "energy")
        data["energy_error_lo"] = w
        data["energy_error_hi"] = w
    elif "energy_error_lo" in keys and "energy_error_hi" in keys:
        data["energy_error_lo"] = _validate_column(data_table, "energy_error_lo", "energy")
        data["energy_error_hi"] = _validate_column(data_table, "energy_error_hi", "energy")
    elif "energy_lo" in keys and "energy_hi" in keys:
        data["energy_error_lo"] = data["energy"] - _validate_column(data_table, "energy_lo", "energy")
        data["energy_error_hi"] = _validate_column(data_table, "energy_hi", "energy") - data["energy"]
    else:
        lo_hi = generate_energy_edges(data["energy"], groups=data["group"])
        data["energy_error_lo"], data["energy_error_hi"] = lo_hi[0], lo_hi[1]

    if "ul" in keys:
        ul_col = np.asarray(data_table["ul"])
        if ul_col.dtype.kind in "ib":
            data["ul"] = np.array(ul_col, dtype=bool)
        elif ul_col.dtype.kind in "US":
            if all(s in ("True", "False") for s in ul_col):
                data["ul"] = np.array([ast.literal_eval(s) for s in ul_col], dtype=bool)
            else:
                raise TypeError("UL column is in wrong format")
        else:
            raise TypeError("UL column is in wrong format")
    else:
        data["ul"] = np.array([False] * n)
    if "flux_ul" in keys:
        fl = data["flux"].copy()
        fl.value[data["ul"]] = Quantity(data_table["flux_ul"]).to(fl.unit).value[data["ul"]]
        data["flux"] = fl

    HAS_CL = False
    meta = getattr(data_table, "meta", {}) or {}
    if "keywords" in meta and "cl" in meta["keywords"]:
        HAS_CL = True
        CL = meta["keywords"]["cl"]["value"]
        if not np.isscalar(CL) or not np.isreal(CL):
            raise TypeError("cl should be a scalar floating point value")
        data["cl"] = np.full(n, float(CL))
    if not HAS_CL:
        data["cl"] = np.full(n, 0.9)
        if np.sum(data["ul"]) > 0:
            log.warning('"cl" keyword not provided in input data table, upper limits will be '
                        "assumed to be at 90% confidence level")
    return data


def validate_data_table(data_table, sed=None):
    """Validate (and concatenate, sort by energy) data tables; utils.py:38-105."""
    if _is_table(data_table):
        data_table = [data_table]
    try:
        for dt in data_table:
            if not _is_table(dt):
                raise TypeError("An object passed as data_table is not a table!")
    except TypeError:
        raise TypeError("Argument passed to validate_data_table is not a table and not a list")

    def dt_sed_conversion(dt, sed):
        f_unit, sedf = sed_conversion(dt["energy"], dt["flux"].unit, sed)
        ndt = dt.copy()
        for col in ["flux", "flux_error_lo", "flux_error_hi"]:
            ndt[col] = (dt[col] * sedf).to(f_unit)
        return ndt

    data_list = [_validate_single_data_table(dt, group=g) for g, dt in enumerate(data_table)]
    data_new = data_list[0].copy()
    f_pt = data_new["flux"].unit.physical_type
    if sed is None:
        sed = f_pt in ["flux", "power"]
    data_new = dt_sed_conversion(data_new, sed)
    for dt in data_list[1:]:
        nf_pt = dt["flux"].unit.physical_type
        if ("flux" in nf_pt and "power" in f_pt) or ("power" in nf_pt and "flux" in f_pt):
            raise TypeError("The physical types of the data tables could not be matched: Some "
                            "are in flux and others in luminosity units")
        dt = dt_sed_conversion(dt, sed)
        for k in list(data_new.keys()):
            a, b = data_new[k], dt[k]
            if u._is_quantity(a):
                data_new[k] = Quantity(np.concatenate([a.value, Quantity(b).to(a.unit).value]),
                                       a.unit)
            else:
                data_new[k] = np.concatenate([np.asarray(a), np.asarray(b)])
    sort_idx = np.argsort(data_new["energy"].to("eV").value, kind="stable")
    return data_new[sort_idx]


_INTEGRAL = ("power", "energy", "flux")
_DIFFERENTIAL = ("differential flux", "differential power", "differential energy",
                 "differential number density")


def sed_conversion(energy, model_unit, sed):
    """Conversion between differential spectrum and SED (utils.py:219-282)."""
    model_unit = u.Unit(model_unit)
    model_pt = model_unit.physical_type
    is_integral = model_pt in _INTEGRAL
    is_differential = model_pt in _DIFFERENTIAL
    energy = Quantity(energy)
    ones = np.ones(energy.shape)
    if (sed and is_integral) or (not sed and is_differential):
        sedf = ones
    elif sed and is_differential:
        sedf = energy**2
    elif not sed and is_integral:
        sedf = 1 / (energy**2)
    else:
        raise u.UnitsError("Model physical type ({0}) is not supported".format(model_pt),
                           "Supported physical types are: power, flux, differential power, "
                           "differential flux")
    is_energy_flux = model_pt in ("energy", "differential energy")
    is_particle_flux = model_pt in ("flux", "differential flux")
    if sed:
        f_unit = u.erg if is_energy_flux else (u.erg / u.s / u.cm**2 if is_particle_flux
                                               else u.erg / u.s)
    else:
        f_unit = u.Unit("1/TeV") if is_energy_flux else (
            u.Unit("1/(s TeV cm2)") if is_particle_flux else u.Unit("1/(s TeV)"))
    return f_unit, sedf


def build_data_table(energy, flux, flux_error=None, flux_error_lo=None, flux_error_hi=None,
                     energy_width=None, energy_lo=None, energy_hi=None, ul=None, cl=None):
    """utils.py:401-487: assemble and validate a data table from arrays."""
    table = DataTable()
    if cl is not None:
        cl = float(cl)
        table.meta["keywords"] = {"cl": {"value": cl}}
    table["energy"] = energy
    if energy_width is not None:
        table["energy_width"] = energy_width
    elif energy_lo is not None and energy_hi is not None:
        table["energy_lo"] = energy_lo
        table["energy_hi"] = energy_hi
    table["flux"] = flux
    if flux_error is not None:
        table["flux_error"] = flux_error
    elif flux_error_lo is not None and flux_error_hi is not None:
        table["flux_error_lo"] = flux_error_lo
        table["flux_error_hi"] = flux_error_hi
    else:
        raise TypeError("Flux error not provided!")
    if ul is not None:
        ul = np.array(ul, dtype=int)
        table["ul"] = ul
    table.meta["comments"] = ["Table generated with naima_b200.build_data_table"]
    validate_data_table(table)
    return table
