# -*- coding: utf-8 -*-
"""Whole-ensemble likelihood on the device.

The reference evaluates ``lnprob(pars, data, model, prior)`` (core.py:97-121) once
per walker on a CPU core.  Here the user's *unchanged* ``model(pars, data)`` and
``prior(pars)`` callbacks are traced ONCE with symbolic parameters
(:class:`SymPar`): the arithmetic they do on ``pars`` (``10 ** pars[0] / u.eV``,
``pars[3] * u.uG`` ...), the radiative classes they instantiate and the
``flux``/``sed``/``compute_We`` calls they make are recorded as a
:class:`LikelihoodPlan`.  A plan is a fixed sequence of kernel launches
(parameter map + priors -> particle-distribution operands -> contraction /
synchrotron kernels -> combine + Gaussian likelihood) over device-resident
tables; it is captured in a CUDA graph per batch size and replayed for every
half-ensemble.  Callbacks that cannot be traced (data-dependent Python control
flow, unsupported arithmetic) raise :class:`TraceError` and the caller falls
back to calling the callback with batched numpy parameters.
"""
import ctypes
import math

import numpy as np
import torch

from . import engine as eng
from . import units as u
from ._lib import (NB_PD_MAXPAR, PD_KIND, check, lib, nb_parmap, nb_pd_desc, nb_prep_job, nb_prior,
                   nb_walker_src)
from .units import Quantity, Unit

FN_ID, FN_POW10, FN_EXP = 0, 1, 2
PRIOR_UNIFORM, PRIOR_NORMAL, PRIOR_LOGUNIFORM = 0, 1, 2


class TraceError(TypeError):
    """The callback does something a LikelihoodPlan cannot express."""


# ------------------------------------------------------------------------------
# symbolic values
# ------------------------------------------------------------------------------
class SymPar:
    """``scale * f(pars[src]) [unit]`` with f in {x, 10**x, e**x}."""

    __array_priority__ = 30000
    __array_ufunc__ = None

    def __init__(self, src, fn=FN_ID, scale=1.0, unit=None):
        self.src, self.fn, self.scale = src, fn, float(scale)
        self.unit = Unit() if unit is None else Unit(unit)

    def _new(self, scale=None, unit=None, fn=None):
        return SymPar(self.src, self.fn if fn is None else fn,
                      self.scale if scale is None else scale, self.unit if unit is None else unit)

    @property
    def value(self):
        return self._new(unit=Unit())

    def to(self, unit):
        return self._new(scale=self.scale * self.unit._factor_to(unit), unit=Unit(unit))

    def plain(self, unit=None):
        """Dimensionless SymPar holding the value in `unit`."""
        if unit is None:
            if any(x != 0 for x in self.unit._dims()):
                raise TraceError("parameter with unit '%s' used where a number is expected"
                                 % self.unit)
            return self._new(scale=self.scale * self.unit._cgs_factor(), unit=Unit())
        return self.to(unit).value

    def __rpow__(self, base):
        if self.fn != FN_ID or self.unit.atoms or self.unit.scale != 1.0 or self.scale != 1.0:
            raise TraceError("only 10**pars[k] / e**pars[k] can be traced")
        if base == 10:
            return self._new(fn=FN_POW10)
        if base == math.e:
            return self._new(fn=FN_EXP)
        raise TraceError("only base 10 and e can be traced")

    def __mul__(self, o):
        if isinstance(o, Unit):
            return self._new(unit=self.unit * o)
        if isinstance(o, Quantity):
            if np.ndim(o.value) != 0:
                raise TraceError("parameter times array")
            return self._new(scale=self.scale * float(o.value), unit=self.unit * o.unit)
        if isinstance(o, SymPar):
            raise TraceError("product of two free parameters")
        if np.ndim(o) != 0:
            raise TraceError("parameter times array")
        return self._new(scale=self.scale * float(o))

    __rmul__ = __mul__

    def __truediv__(self, o):
        if isinstance(o, Unit):
            return self._new(unit=self.unit / o)
        if isinstance(o, Quantity):
            if np.ndim(o.value) != 0:
                raise TraceError("parameter divided by array")
            return self._new(scale=self.scale / float(o.value), unit=self.unit / o.unit)
        if isinstance(o, SymPar) or np.ndim(o) != 0:
            raise TraceError("unsupported division")
        return self._new(scale=self.scale / float(o))

    def _no(self, *a, **k):
        raise TraceError("operation on a free parameter that a LikelihoodPlan cannot express")

    __rtruediv__ = __add__ = __radd__ = __sub__ = __rsub__ = __neg__ = __pow__ = _no
    __lt__ = __le__ = __gt__ = __ge__ = __float__ = __int__ = __bool__ = __abs__ = _no
    __getitem__ = __iter__ = __len__ = _no

    def __repr__(self):
        f = ("%s", "10**%s", "exp(%s)")[self.fn] % ("pars[%d]" % self.src)
        return "<SymPar %g * %s %s>" % (self.scale, f, self.unit)


class SymPars:
    """Stands in for ``pars`` while tracing."""

    def __init__(self, P):
        self._P = P

    def __len__(self):
        return self._P

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [SymPar(i) for i in range(*k.indices(self._P))]
        k = int(k)
        if k < 0:
            k += self._P
        if not 0 <= k < self._P:
            raise IndexError(k)
        return SymPar(k)

    def __iter__(self):
        return iter(SymPar(i) for i in range(self._P))


def is_sym(x):
    return isinstance(x, SymPar)


class SymPrior:
    """Sum of prior terms on raw parameters (core.py:34-58)."""

    def __init__(self, terms=()):
        self.terms = list(terms)

    @staticmethod
    def term(value, kind, a, b):
        if value.fn != FN_ID or value.scale != 1.0 or value.unit.atoms or value.unit.scale != 1.0:
            raise TraceError("priors are traced on raw parameters only")
        return SymPrior([(value.src, kind, float(a), float(b))])

    def __add__(self, o):
        if isinstance(o, SymPrior):
            return SymPrior(self.terms + o.terms)
        if np.ndim(o) == 0 and float(o) == 0.0:
            return self
        raise TraceError("prior plus a non-zero constant")

    __radd__ = __add__


class SymBlob:
    def __init__(self, kind, **kw):
        self.kind = kind
        self.__dict__.update(kw)


class SymFlux:
    """Sum of groups ``(sum of component spectra) / (4 pi d^2)`` in a common unit.  A group
    is ``(component, 4 pi d^2, scale[, efac])``: efac is an optional walker-independent
    factor per photon energy (an absorption ``transmission(E)`` multiplied to the flux)."""

    __array_ufunc__ = None  # ndarray * SymFlux -> SymFlux.__rmul__
    __array_priority__ = 30000

    def __init__(self, groups, E, unit, sed, base_unit):
        self.groups, self.E, self.unit, self.sed = groups, E, Unit(unit), sed
        self.base_unit = Unit(base_unit)  # of spectrum / (4 pi d^2)

    @staticmethod
    def from_component(comp, E, distance, sed):
        if eng_nonzero(distance):
            d = Quantity(distance)
            if d.unit.physical_type != "length":
                raise TypeError("distance should be given in units of length")
            div = 4 * np.pi * d.to("cm").value ** 2
            unit, base = ("erg/(cm2 s)" if sed else "1/(s cm2 eV)"), "1/(s cm2 eV)"
        else:
            div, unit, base = 1.0, ("erg/s" if sed else "1/(s eV)"), "1/(s eV)"
        return SymFlux([(comp, float(div), 1.0)], E, unit, sed, base)

    def to(self, unit):
        self.unit._factor_to(unit)  # raises on incompatible units
        return SymFlux(self.groups, self.E, unit, self.sed, self.base_unit)

    def __add__(self, o):
        if not isinstance(o, SymFlux):
            if np.ndim(o) == 0 and float(o) == 0.0:
                return self
            raise TraceError("flux plus a constant")
        if o.sed != self.sed or not np.array_equal(o.E.to("eV").value, self.E.to("eV").value):
            raise TraceError("cannot add fluxes on different energy grids / representations")
        o.unit._factor_to(self.unit)
        f = o.base_unit._factor_to(self.base_unit)  # groups are held in units of base_unit
        og = o.groups if f == 1.0 else [(g[0], g[1], g[2] * f) + tuple(g[3:]) for g in o.groups]
        return SymFlux(self.groups + og, self.E, self.unit, self.sed, self.base_unit)

    __radd__ = __add__

    def _scaled(self, k, unit_op):
        """Product with a scalar (plain number or scalar Quantity) k; unit_op(unit, k.unit)."""
        if isinstance(k, (SymPar, SymFlux)):
            raise TraceError("product of a flux with a free parameter or another flux")
        if isinstance(k, Unit):
            kv, ku = 1.0, k
        elif isinstance(k, Quantity):
            if np.ndim(k.value) != 0:
                raise TraceError("flux times array")
            kv, ku = float(k.value), k.unit
        elif np.ndim(k) == 0:
            kv, ku = float(k), None
        else:
            raise TraceError("flux times array")
        return kv, ku

    def _times_energy_factor(self, f):
        """Product with a dimensionless array f[N_E] (one factor per photon energy)."""
        f = np.asarray(f.value if isinstance(f, Quantity) else f, dtype=float)
        if isinstance(f, Quantity) and any(x != 0 for x in f.unit._dims()):
            raise TraceError("flux times a dimensional array")
        if f.shape != np.shape(self.E.value):
            raise TraceError("flux times an array that is not one factor per photon energy")
        groups = [(g[0], g[1], g[2], f if len(g) < 4 or g[3] is None else g[3] * f)
                  for g in self.groups]
        return SymFlux(groups, self.E, self.unit, self.sed, self.base_unit)

    def __mul__(self, o):
        if not isinstance(o, (SymPar, SymFlux, Unit)) and np.ndim(
                o.value if isinstance(o, Quantity) else o) == 1:
            return self._times_energy_factor(o)
        kv, ku = self._scaled(o, None)
        unit = self.unit if ku is None else self.unit * ku
        base = self.base_unit if ku is None else self.base_unit * ku
        return SymFlux([(g[0], g[1], g[2] * kv) + tuple(g[3:]) for g in self.groups], self.E,
                       unit, self.sed, base)

    __rmul__ = __mul__

    def __truediv__(self, o):
        kv, ku = self._scaled(o, None)
        unit = self.unit if ku is None else self.unit / ku
        base = self.base_unit if ku is None else self.base_unit / ku
        return SymFlux([(g[0], g[1], g[2] / kv) + tuple(g[3:]) for g in self.groups], self.E,
                       unit, self.sed, base)

    def check_seed_density(self, name, energy):
        """Usable as the photon density of a tabulated IC seed field on `energy`
        (radiative.py:509-519 conventions)?"""
        if self.sed:
            raise TraceError("seed photon field given as an SED of the fit parameters")
        if self.unit.physical_type != "differential number density":
            raise TypeError("{0}-density should be given in units of differential number "
                            "density".format(name))
        if not np.array_equal(np.atleast_1d(self.E.to("eV").value),
                              np.atleast_1d(Quantity(energy).to("eV").value)):
            raise TraceError("seed photon density is not evaluated at the seed energies")

    def seed_sources(self):
        """[(component, factor)]: density [1/(mec2 cm3)] = sum factor * spectrum[1/(s eV)]."""
        f = self.base_unit._factor_to("1/(eV cm3)") * eng.mec2_eV
        if any(len(g) > 3 and g[3] is not None for g in self.groups):
            raise TraceError("seed photon density with a per-energy factor is not traced")
        return [(g[0], g[2] / g[1] * f) for g in self.groups]


def comp_B(comp):
    from . import models as M
    return M._val(comp.B, "G")


def eng_nonzero(distance):
    v = Quantity(distance).value if u._is_quantity(distance) else distance
    return bool(np.all(np.asarray(v) != 0))


# ------------------------------------------------------------------------------
# tracing
# ------------------------------------------------------------------------------
def trace(model, prior, data, P):
    """Run the callbacks on symbolic parameters; returns (SymFlux, blobs, SymPrior|None)."""
    pars = SymPars(P)
    try:
        out = model(pars, data)
    except TraceError:
        raise
    except Exception as e:  # arbitrary failures inside user code on symbolic values
        raise TraceError("model callback could not be traced: %r" % (e,))
    if isinstance(out, (tuple, list)):
        flux, blobs = out[0], list(out[1:])
    else:
        flux, blobs = out, []
    if not isinstance(flux, SymFlux):
        raise TraceError("model callback did not return a flux of the radiative classes")
    sp = None
    if prior is not None:
        try:
            sp = prior(pars)
        except TraceError:
            raise
        except Exception as e:
            raise TraceError("prior callback could not be traced: %r" % (e,))
        if not isinstance(sp, SymPrior):
            if np.ndim(sp) == 0 and float(sp) == 0.0:
                sp = SymPrior()
            else:
                raise TraceError("prior callback returned a constant")
    return flux, blobs, sp


# ------------------------------------------------------------------------------
# the plan
# ------------------------------------------------------------------------------
class _Exec:
    """Buffers + captured graph for one batch size W."""


class LikelihoodPlan:
    def __init__(self, model, prior, data, P, use_graph=True, selfprep=True):
        self.P = int(P)
        self.data = data
        self.use_graph = use_graph
        self.selfprep = selfprep
        flux, blobs, sp = trace(model, prior, data, self.P)
        self.flux, self.blobs, self.prior = flux, blobs, sp
        E = Quantity(data["energy"])
        self.E_eV = np.ascontiguousarray(E.to("eV").value, dtype=float)
        if not np.array_equal(flux.E.to("eV").value, self.E_eV):
            raise TraceError("model flux is not evaluated at data['energy']")
        self.N_E = self.E_eV.size
        self._units()
        self._components()
        self._exec = {}
        self._side = []
        self.launches_per_eval = 0

    # -- unit bookkeeping (core.py:64-71, utils.py:219-282) ------------------------
    def _units(self):
        from .utils import sed_conversion

        d_unit = Quantity(self.data["flux"]).unit
        m_unit = self.flux.unit
        model_is_sed = m_unit.physical_type in ["power", "flux"]
        data_is_sed = d_unit.physical_type in ["power", "flux"]
        E = Quantity(self.data["energy"])
        # spectrum [1/(s eV)] (/(4 pi d^2) cm^-2) -> model unit
        base = Quantity(self.E_eV**2 if self.flux.sed else np.ones(self.N_E),
                        self.flux.base_unit * (u.eV**2 if self.flux.sed else Unit()))
        model_fac = base.to(m_unit).value  # per-energy factor
        if model_is_sed != data_is_sed:
            _, sedf = sed_conversion(E, m_unit, data_is_sed)
            conv = (Quantity(np.ones(self.N_E), m_unit) * sedf).to(d_unit).value
        else:
            conv = Quantity(np.ones(self.N_E), m_unit).to(d_unit).value
        self.unit_fac = model_fac * conv      # device: spectrum -> data unit
        self.model_unit = m_unit
        self.to_model_unit = 1.0 / conv       # host: data unit -> model unit (blob 0)
        ul = np.asarray(self.data["ul"], dtype=bool)
        self.ddata = eng.DeviceData(
            Quantity(self.data["flux"]).value,
            Quantity(self.data["flux_error_lo"]).to(d_unit).value,
            Quantity(self.data["flux_error_hi"]).to(d_unit).value, ul,
            np.asarray(self.data["cl"], dtype=float))
        self.unit_fac_d = eng.to_dev(self.unit_fac)

    # -- components -----------------------------------------------------------------
    def _components(self):
        from . import models as M

        self.pds, self.preps, self.comps = [], [], []
        self.aux = []      # auxiliary synchrotron launches (seed fields of the "ssc" components)
        self.scalars = []  # per-walker scalar columns: SymPar or float

        def pd_index(pd):
            if isinstance(pd, M.TableModel):
                raise TraceError("TableModel particle distributions are not traced")
            for i, (p, _) in enumerate(self.pds):
                if p is pd:
                    return i
            vals = pd._eval_params(M._val(pd.amplitude, "1/eV"))
            self.pds.append((pd, vals))
            return len(self.pds) - 1

        def prep_index(ipd, grid, raw):
            for i, pr in enumerate(self.preps):
                if pr["pd"] == ipd and pr["grid"].key == grid.key:
                    pr["raw"] = pr["raw"] or raw
                    return i
            self.preps.append({"pd": ipd, "grid": grid, "raw": raw})
            return len(self.preps) - 1

        def scalar_index(v):
            self.scalars.append(v)
            return len(self.scalars) - 1

        exact = eng.EXACT
        for gi, grp in enumerate(self.flux.groups):
            comp, div, scale = grp[:3]
            efac = grp[3] if len(grp) > 3 else None
            if efac is not None and not isinstance(comp, (M.InverseCompton, M.Bremsstrahlung,
                                                          M.PionDecay)):
                raise TraceError("a per-energy factor on %s is not traced"
                                 % type(comp).__name__)
            ipd = pd_index(comp.particle_distribution)
            grid = comp._grid()
            c = {"group": gi, "div": div / scale, "obj": comp, "efac": efac}
            if isinstance(comp, M.Synchrotron):
                c["kind"] = "syn"
                c["prep"] = prep_index(ipd, grid, False)
                c["B"] = scalar_index(M._val(comp.B, "G"))
            elif isinstance(comp, M.InverseCompton):
                c["prep"] = prep_index(ipd, grid, exact)
                seeds, sym = [], []
                for name, sd in comp.seed_photon_fields.items():
                    if sd.get("symbolic"):
                        sym.append(sd)
                        continue
                    if sd["type"] == "array" and sd["photon_density"].ndim == 2:
                        raise TraceError("per-walker seed photon fields given as arrays are "
                                         "not traced")
                    seeds.append(comp._seed_tuple(sd))
                if sym and efac is not None:
                    raise TraceError("a per-energy factor on self-Compton emission is not traced")
                if sym and exact:
                    raise TraceError("seed fields computed from the fit parameters are traced "
                                     "in the hoisted mode only")
                # seed fields that depend on the fit parameters (synchrotron self-Compton):
                # one "ssc" component each, fed by auxiliary synchrotron launches on the seed
                # energies; summed after the tabulated seeds of the same InverseCompton
                extra = []
                for sd in sym:
                    seed_E = np.atleast_1d(Quantity(sd["energy"]).to("eV").value).astype(float)
                    srcs = []
                    for scomp, fac in sd["photon_density"].seed_sources():
                        if not isinstance(scomp, M.Synchrotron):
                            raise TraceError("seed photon fields computed from %s are not "
                                             "traced" % type(scomp).__name__)
                        srcs.append((self._aux_index(scomp, seed_E, pd_index, prep_index,
                                                     scalar_index), fac))
                    extra.append({"group": gi, "div": div / scale, "obj": comp, "kind": "ssc",
                                  "prep": c["prep"], "sources": srcs,
                                  "tb": eng.ssc_table(grid, self.E_eV, seed_E)})
                if seeds:
                    c["kind"] = "table"
                    c["table"] = eng.ic_table(grid, self.E_eV, tuple(seeds))
                    c["rows"] = [(s * self.N_E, None) for s in range(len(seeds))]
                    self.comps.append(c)
                self.comps.extend(extra)
                continue
            elif isinstance(comp, M.Bremsstrahlung):
                c["kind"] = "table"
                c["prep"] = prep_index(ipd, grid, exact)
                c["table"] = eng.brems_table(grid, self.E_eV)
                n0 = M._val(comp.n0, "1/cm3")
                c["rows"] = []
                if comp.weight_ee != 0.0:
                    c["rows"].append((0, scalar_index(n0 * comp.weight_ee)))
                if comp.weight_ep != 0.0:
                    c["rows"].append((self.N_E, scalar_index(n0 * comp.weight_ep)))
                if not c["rows"]:
                    raise TraceError("bremsstrahlung with both weights zero")
            elif isinstance(comp, M.PionDecay):
                c["kind"] = "table"
                c["prep"] = prep_index(ipd, grid, exact)
                useLUT = bool(comp.useLUT) and (comp.hiEmodel, bool(comp.nuclear_enhancement)) \
                    in comp._LUT_MODELS
                c["table"] = eng.pp_table(grid, self.E_eV, useLUT, comp.hiEmodel,
                                          comp.nuclear_enhancement)
                c["rows"] = [(0, scalar_index(M._val(comp.nh, "1/cm3")))]
            else:
                raise TraceError("unsupported radiative class %r" % type(comp).__name__)
            self.comps.append(c)
        # per-energy factors (absorption) fold into a private copy of the table's row
        # coefficients: coef_eff[c * N_E + e] = coef * efac[e]
        for c in self.comps:
            if c.get("efac") is not None and c["kind"] == "table":
                tb = c["table"]
                c["coef_eff"] = tb.coef * eng.to_dev(np.tile(c["efac"], tb.n_comp))
        # Synchrotron derives the walker's operands itself (no set-up launch in front of it:
        # its CTAs own one walker, so the operands cost one pass over the nodes); the
        # reference-order (exact) mode keeps the operand arrays
        for c in self.comps:
            c["selfprep"] = (self.selfprep and not exact and self.P <= 32 and c["kind"] == "syn")
        for a in self.aux:
            a["selfprep"] = a["selfprep"] and self.selfprep and not exact
        # blobs
        self.blob_specs = []
        for b in self.blobs:
            self.blob_specs.append(self._blob_spec(b, pd_index))
        # columns of the per-walker record: model flux first, then the device-side blobs
        self.blob_cols = []
        off = self.N_E
        for spec in self._flat_blob_specs():
            width = 1 if spec["kind"] == "W" else int(spec["e_eV"].size)
            self.blob_cols.append((off, width))
            off += width
        self.row_width = off

    def _aux_index(self, comp, E_eV, pd_index, prep_index, scalar_index):
        """Auxiliary synchrotron spectrum of `comp` on photon energies E_eV [1/(s eV)]."""
        for i, a in enumerate(self.aux):
            if a["obj"] is comp and np.array_equal(a["E_eV"], E_eV):
                return i
        ipd = pd_index(comp.particle_distribution)
        self.aux.append({"kind": "syn", "obj": comp, "E_eV": np.ascontiguousarray(E_eV),
                         "prep": prep_index(ipd, comp._grid(), False),
                         "B": scalar_index(comp_B(comp)), "selfprep": self.P <= 32})
        return len(self.aux) - 1

    def _blob_spec(self, b, pd_index):
        if isinstance(b, SymBlob):
            if b.kind == "W":
                return {"kind": "W", "pd": pd_index(b.comp.particle_distribution),
                        "grid": b.grid, "unit": u.erg}
            if b.kind == "pdist":
                return {"kind": "pdist", "pd": pd_index(b.pd), "e_eV": b.e_eV, "unit": b.unit,
                        "shape": b.shape}
            raise TraceError("unknown symbolic blob")
        if isinstance(b, (tuple, list)):
            return {"kind": "tuple", "items": [self._blob_spec(x, pd_index) for x in b]}
        if isinstance(b, (SymPar, SymFlux)):
            raise TraceError("free parameters / extra fluxes as blobs are not traced")
        return {"kind": "const", "value": b}

    # -- per-W executable -------------------------------------------------------------
    def pack_width(self):
        """Columns of a packed per-proposal record [blob record | lnprob | proposal]."""
        return self.row_width + 1 + self.P

    def _build(self, W, pack=None):
        """pack: optional device tensor [W][>= pack_width()] whose rows receive the packed
        per-proposal records (walker sharding: the all-gather buffer's local slice)."""
        ex = _Exec()
        ex.W = W
        P = self.P
        ex.pack = pack
        ex.pars = eng.zeros(W, P) if pack is None else \
            pack[:, self.row_width + 1:self.row_width + 1 + P]
        n_pd, n_sc = len(self.pds), len(self.scalars)
        ex.pm = eng.zeros(n_pd * W * NB_PD_MAXPAR + max(n_sc, 1) * W)
        entries = []
        for i, (pd, vals) in enumerate(self.pds):
            for k, v in enumerate(vals):
                entries.append((v, i * W * NB_PD_MAXPAR + k, NB_PD_MAXPAR))
        sc_base = n_pd * W * NB_PD_MAXPAR
        for j, v in enumerate(self.scalars):
            entries.append((v, sc_base + j * W, 1))
        if len(entries) > 32:
            raise TraceError("too many mapped parameters")
        n_pd_entries = sum(len(vals) for _, vals in self.pds)
        ex.scalar_entry = [n_pd_entries + j for j in range(len(self.scalars))]
        ex.map = (nb_parmap * max(len(entries), 1))()
        for k, (v, off, stride) in enumerate(entries):
            m = ex.map[k]
            if is_sym(v):
                v = v.plain()
                m.src, m.fn, m.scale = v.src, v.fn, v.scale
            else:
                if np.ndim(v) != 0:
                    raise TraceError("array-valued parameter in a traced model")
                m.src, m.fn, m.scale = -1, 0, float(v)
            m.dst_off, m.dst_stride = off, stride
        ex.n_map = len(entries)
        pr = self.prior.terms if self.prior is not None else []
        ex.pri = (nb_prior * max(len(pr), 1))()
        for k, (par, kind, a, b) in enumerate(pr):
            ex.pri[k].par, ex.pri[k].kind, ex.pri[k].a, ex.pri[k].b = par, kind, a, b
        ex.n_pri = len(pr)
        ex.prior = eng.zeros(W)

        def pd_block(i):
            return ex.pm[i * W * NB_PD_MAXPAR:(i + 1) * W * NB_PD_MAXPAR]

        def scalar_col(j):
            return ex.pm[sc_base + j * W: sc_base + (j + 1) * W]

        ex.preps = []
        for prd in self.preps:
            g = prd["grid"]
            p = eng.Prepared()
            p.xn, p.ds1 = eng.zeros(W, g.pitch), eng.zeros(W, g.pitch)
            p.nraw = eng.zeros(W, g.pitch) if prd["raw"] else None
            p.W, p.grid = W, g
            ex.preps.append(p)
        terms = []
        ex.outs = []
        ncomp = len(self.comps)
        for ic, c in enumerate(self.comps):
            # a group (one radiative component's flux()) closes with its last term
            last_of_group = ic == ncomp - 1 or self.comps[ic + 1]["group"] != c["group"]
            if c["kind"] in ("syn", "ssc"):
                out = eng.zeros(W, self.N_E)
                terms.append((out, 0, last_of_group, c["div"], None))
            else:
                out = eng.zeros(W, c["table"].R)
                rows = c["rows"]
                for k, (off, sc) in enumerate(rows):
                    terms.append((out, off, last_of_group and k == len(rows) - 1, c["div"],
                                  scalar_col(sc) if sc is not None else None))
            ex.outs.append(out)
        # auxiliary synchrotron spectra (seed luminosities) and the self-Compton work buffers
        ex.aux_outs = [eng.zeros(W, a["E_eV"].size) for a in self.aux]
        ex.aux_E = [eng.photon_energies(a["E_eV"]) for a in self.aux]
        ex.ssc = {}
        for ic, c in enumerate(self.comps):
            if c["kind"] != "ssc":
                continue
            tb = c["tb"]
            ex.ssc[ic] = dict(
                sxn=eng.zeros(W, tb.spitch), sds=eng.zeros(W, tb.spitch),
                inner=eng.empty(W, tb.Rp),
                sources=eng.make_ssc_sources([(ex.aux_outs[ia], 0, fac)
                                              for ia, fac in c["sources"]]))
        ex.terms = eng.make_terms(terms)
        # one record per walker: [model flux (N_E) | further blobs], so that a sampler
        # moves a walker's blobs with one row copy (row pitch = self.row_width)
        ex.row = eng.zeros(W, self.row_width) if pack is None else pack[:, :self.row_width]
        ex.row_ld = ex.row.stride(0)
        ex.flux = ex.row[:, :self.N_E]
        ex.lnp = eng.zeros(W) if pack is None else pack[:, self.row_width]
        ex.E_erg = eng.photon_energies(self.E_eV)  # [2][N_E]: E in erg, cbrt(E)
        ex.blob_bufs = []
        for spec, (off, width) in zip(self._flat_blob_specs(), self.blob_cols):
            ex.blob_bufs.append(ex.row[:, off] if spec["kind"] == "W"
                                else ex.row[:, off:off + width])
            if spec["kind"] == "pdist":
                spec["e_d"] = eng.to_dev(spec["e_eV"])
        # descriptors for the self-contained component kernels
        ex.src = nb_walker_src()
        ex.src.pars, ex.src.P, ex.src.n_map = ex.pars.data_ptr(), P, ex.n_map
        ex.src.map_host = ctypes.cast(ex.map, ctypes.POINTER(nb_parmap))
        ex.pd_desc = []
        for prd in self.preps:
            g = prd["grid"]
            d = nb_pd_desc()
            d.kind = PD_KIND[self.pds[prd["pd"]][0]._kind]
            d.pd_off = prd["pd"] * W * NB_PD_MAXPAR
            d.e_mul1, d.e_mul2, d.n_scale = g.e_mul1, g.e_mul2, g.n_scale
            d.lnx, d.invdlx = g.lnx_d.data_ptr(), g.invdlx_d.data_ptr()
            ex.pd_desc.append(d)
        # operand arrays are only needed by components that do not prepare their own
        needed = set(c["prep"] for c in self.comps if not c["selfprep"])
        needed |= set(a["prep"] for a in self.aux if not a["selfprep"])
        # jobs of the fused per-walker set-up kernel
        jobs = []
        for ip, (prd, p) in enumerate(zip(self.preps, ex.preps)):
            if ip not in needed:
                continue
            g = prd["grid"]
            jobs.append(dict(kind=PD_KIND[self.pds[prd["pd"]][0]._kind], N=g.N,
                             pd_off=prd["pd"] * W * NB_PD_MAXPAR, x=g.x_d, invdlx=g.invdlx_d,
                             e_mul1=g.e_mul1, e_mul2=g.e_mul2, n_scale=g.n_scale, xn=p.xn,
                             ds1=p.ds1, nraw=p.nraw, wpitch=g.pitch))
        for spec, buf in zip(self._flat_blob_specs(), ex.blob_bufs):
            if spec["kind"] == "W":
                g = spec["grid"]
                jobs.append(dict(kind=PD_KIND[self.pds[spec["pd"]][0]._kind], N=g.N,
                                 pd_off=spec["pd"] * W * NB_PD_MAXPAR, x=g.x_d,
                                 e_mul1=g.e_mul1, e_mul2=g.e_mul2, n_scale=g.n_scale,
                                 x_to_energy=g.x_to_erg, energy_out=buf,
                                 energy_stride=ex.row_ld))
        # two launches of the set-up kernel: the operand jobs (and the published parameter
        # map / priors / proposals) gate the contraction; the total-energy blobs (reference-
        # order log10/pow integration, the slowest CTAs by far) only gate the combine and run
        # on a side branch with scratch outputs for the duplicate parameter map
        def job_array(js):
            if len(js) > 8:
                raise TraceError("too many particle-distribution jobs in one plan")
            arr = (nb_prep_job * max(len(js), 1))()
            for k, jd in enumerate(js):
                for name, v in jd.items():
                    setattr(arr[k], name, v.data_ptr() if hasattr(v, "data_ptr") else v)
            return arr, len(js)
        ex.jobs, ex.n_jobs = job_array([j for j in jobs if "xn" in j])
        ex.blob_jobs, ex.n_blob_jobs = job_array([j for j in jobs if "xn" not in j])
        ex.pm2 = torch.empty_like(ex.pm) if ex.n_blob_jobs else None
        ex.pars2 = eng.zeros(W, P) if ex.n_blob_jobs else None
        # pinned staging for the host-facing call
        ex.pars_pin = torch.empty(W, P, dtype=torch.float64).pin_memory()
        ex.lnp_pin = torch.empty(W, dtype=torch.float64).pin_memory()
        ex.row_pin = torch.empty(W, self.row_width, dtype=torch.float64).pin_memory()
        ex.pd_block, ex.scalar_col = pd_block, scalar_col
        ex.graph = None
        if pack is not None:
            return ex  # driven by a sharded ensemble step (proposals computed on the device)
        # warm-up launch outside capture (function attributes, lazy module load)
        self._enqueue(ex)
        torch.cuda.synchronize()
        if self.use_graph:
            with eng.capture_graph() as g:
                self._enqueue(ex)
            ex.graph = g
        return ex

    def _flat_blob_specs(self):
        out = []

        def walk(s):
            if s["kind"] == "tuple":
                for x in s["items"]:
                    walk(x)
            elif s["kind"] in ("W", "pdist"):
                out.append(s)
        for s in self.blob_specs:
            walk(s)
        return out

    def _src(self, ex, mv):
        """nb_walker_src of ex: dense parameters, or the proposals described by mv."""
        if mv is None:
            return ex.src
        src = nb_walker_src()
        src.pars, src.P, src.n_map, src.map_host = None, ex.src.P, ex.src.n_map, ex.src.map_host
        src.mv_host = ctypes.pointer(mv)
        return src

    def _launch_prep(self, ex, mv):
        """Parameter map + priors (+ proposals) published for the combine kernel, the
        total-energy blobs, and the operand arrays of components that need them."""
        L, W, st = lib(), ex.W, eng.stream()
        if mv is None:
            check(L.nb_walker_prep(eng.ptr(ex.pars), W, self.P, ex.map, ex.n_map,
                                   eng.ptr(ex.pm), ex.pri, ex.n_pri, eng.ptr(ex.prior),
                                   ex.jobs, ex.n_jobs, st), "nb_walker_prep")
        else:
            check(L.nb_walker_prep_move(ctypes.byref(mv), eng.ptr(ex.pars), W, self.P,
                                        ex.map, ex.n_map, eng.ptr(ex.pm), ex.pri, ex.n_pri,
                                        eng.ptr(ex.prior), ex.jobs, ex.n_jobs, st),
                  "nb_walker_prep_move")

    def _launch_blob_prep(self, ex, mv):
        """The total-energy blobs (same kernel, scratch parameter-map outputs)."""
        L, W, st = lib(), ex.W, eng.stream()
        if mv is None:
            check(L.nb_walker_prep(eng.ptr(ex.pars), W, self.P, ex.map, ex.n_map,
                                   eng.ptr(ex.pm2), ex.pri, 0, None, ex.blob_jobs,
                                   ex.n_blob_jobs, st), "nb_walker_prep")
        else:
            mv2 = type(mv).from_buffer_copy(mv)
            mv2.pars_ld = 0
            check(L.nb_walker_prep_move(ctypes.byref(mv2), eng.ptr(ex.pars2), W, self.P,
                                        ex.map, ex.n_map, eng.ptr(ex.pm2), ex.pri, 0, None,
                                        ex.blob_jobs, ex.n_blob_jobs, st), "nb_walker_prep_move")

    def _launch_syn(self, ex, c, E_erg, out, src):
        """Synchrotron of component / auxiliary descriptor c on photon energies E_erg."""
        L, W = lib(), ex.W
        p = ex.preps[c["prep"]]
        g = p.grid
        if c["selfprep"]:
            d = ex.pd_desc[c["prep"]]
            check(L.nb_synchrotron_fused(
                ctypes.byref(src), ctypes.byref(d), ex.scalar_entry[c["B"]], eng.ptr(g.x_d), g.N,
                eng.ptr(g.gm2_d), eng.ptr(g.g23_d), eng.ptr(g.dlx_d), W, eng.ptr(E_erg[0]),
                eng.ptr(E_erg[1]), E_erg.shape[1], eng.ptr(out), out.stride(0), eng.stream()),
                "nb_synchrotron_fused")
        else:
            eng.synchrotron(g, p, ex.scalar_col(c["B"]), E_erg, out=out)

    def _nodes(self, ex, mv):
        """The launches of one likelihood evaluation before the combine kernel, as a DAG:
        [(name, launch function, [names it depends on])], in a valid launch order with the
        longest chain (self-Compton) first."""
        src = self._src(ex, mv)
        nodes = []
        for ia, a in enumerate(self.aux):
            nodes.append(("aux_syn%d" % ia,
                          lambda a=a, ia=ia: self._launch_syn(ex, a, ex.aux_E[ia],
                                                              ex.aux_outs[ia], src),
                          [] if a["selfprep"] else ["walker_prep"]))
        nodes.append(("walker_prep", lambda: self._launch_prep(ex, mv), []))
        if any(n[2] for n in nodes[:-1]):  # an auxiliary launch needs the operand arrays
            nodes.insert(0, nodes.pop())
        if ex.n_blob_jobs:
            nodes.append(("blob_prep", lambda: self._launch_blob_prep(ex, mv), []))
        for ic, (c, out) in enumerate(zip(self.comps, ex.outs)):
            if c["kind"] == "ssc":
                tb, b = c["tb"], ex.ssc[ic]
                nodes.append(("ssc_seed%d" % ic,
                              lambda tb=tb, b=b: eng.ssc_seed(tb, b["sources"], ex.W, b["sxn"],
                                                              b["sds"]),
                              ["aux_syn%d" % ia for ia, _ in c["sources"]]))
                nodes.append(("ssc_inner%d" % ic,
                              lambda tb=tb, b=b: eng.ssc_inner(tb, b["sxn"], b["sds"], ex.W,
                                                               b["inner"]),
                              ["ssc_seed%d" % ic]))
                nodes.append(("ssc_outer%d" % ic,
                              lambda tb=tb, b=b, c=c, out=out: eng.ssc_outer(
                                  tb, b["inner"], ex.preps[c["prep"]], out),
                              ["ssc_inner%d" % ic, "walker_prep"]))
            elif c["kind"] == "syn":
                nodes.append(("syn%d" % ic,
                              lambda c=c, out=out: self._launch_syn(ex, c, ex.E_erg, out, src),
                              [] if c["selfprep"] else ["walker_prep"]))
            else:
                nodes.append(("table%d" % ic,
                              lambda c=c, out=out: eng.contract(c["table"], ex.preps[c["prep"]],
                                                                out=out, coef=c.get("coef_eff")),
                              ["walker_prep"]))
        return nodes

    def _launch_combine(self, ex, mv=None, fuse_update=True, peers=None):
        L, st, W = lib(), eng.stream(), ex.W
        n = 0
        for spec, buf in zip(self._flat_blob_specs(), ex.blob_bufs):
            if spec["kind"] != "pdist":
                continue  # particle energies are jobs of nb_walker_prep
            pdobj = self.pds[spec["pd"]][0]
            check(L.nb_pdist_eval_ld(PD_KIND[pdobj._kind], eng.ptr(ex.pd_block(spec["pd"])), W,
                                     eng.ptr(spec["e_d"]), spec["e_eV"].size, eng.ptr(buf),
                                     ex.row_ld, st), "nb_pdist_eval_ld")
            n += 1
        # last: with `mv` the combine kernel also moves the accepted walkers' records
        eng.combine(ex.terms, W, self.N_E, self.unit_fac_d, flux_out=ex.row, data=self.ddata,
                    prior_d=ex.prior if self.prior is not None else None, lnp_out=ex.lnp,
                    mv=mv if fuse_update else None, pars_d=ex.pars, flux_ld=ex.row_ld,
                    lnp_ld=ex.lnp.stride(0), peers=peers, nb=self.row_width)
        return n + 1

    def _enqueue(self, ex, mv=None, fuse_update=True, peers=None):
        """The launch sequence of one likelihood evaluation of ex.W walkers.  With `mv`
        (an nb_stretch describing a device-resident ensemble) the parameters are the
        stretch-move proposals of the active half, computed by the set-up kernel, and
        (fuse_update) the combine kernel also accepts/rejects them and appends the chain.
        Independent launches go to side streams (parallel branches of the captured graph):
        a node continues the stream of a dependency whose stream is still free, else it
        forks; everything joins before the combine."""
        if mv is None and ex.pack is not None:
            raise ValueError("a packed executable is driven by device-side proposals only")
        nodes = self._nodes(ex, mv)
        # kernels of parallel branches only share an SM when they ask for the same L1 /
        # shared-memory split (nb_launch_carveout): half and half lets contraction CTAs run
        # beside synchrotron CTAs (C3: 0.1075 -> 0.098 ms per step)
        kinds = {c["kind"] for c in self.comps}
        carve = 50 if ("syn" in kinds and "table" in kinds and not self.aux) else -1
        L = lib()
        check(L.nb_launch_carveout(carve), "nb_launch_carveout")
        try:
            return self._enqueue_nodes(ex, nodes, mv, fuse_update, peers)
        finally:
            L.nb_launch_carveout(-1)

    def _enqueue_nodes(self, ex, nodes, mv, fuse_update, peers):
        main = torch.cuda.current_stream()
        streams, tails = [main], [None]  # stream k, name of the last node launched on it
        where, done = {}, {}
        fork = torch.cuda.Event()
        fork.record(main)  # before the first launch: side branches depend on nothing earlier
        for name, fn, deps in nodes:
            k = None
            for d in deps:  # continue a dependency's stream when nothing followed it there
                if tails[where[d]] == d:
                    k = where[d]
                    break
            if k is None:
                free = [i for i, t in enumerate(tails) if t is None]
                if free:
                    k = free[0]
                else:
                    k = len(streams)
                    if k - 1 >= len(self._side):
                        self._side.append(torch.cuda.Stream())
                    streams.append(self._side[k - 1])
                    tails.append(None)
            st = streams[k]
            if k != 0 and tails[k] is None:
                st.wait_event(fork)
            for d in deps:
                if where[d] != k:
                    st.wait_event(done[d])
            with torch.cuda.stream(st):
                fn()
                ev = torch.cuda.Event()
                ev.record(st)
            where[name], done[name], tails[k] = k, ev, name
        for k in range(1, len(streams)):
            if tails[k] is not None:
                main.wait_event(done[tails[k]])
        n = len(nodes) + self._launch_combine(ex, mv, fuse_update, peers)
        self.launches_per_eval = n
        return n

    def stages(self, ex):
        """[(name, launch)] of one evaluation on ex.pars, in launch order on the current
        stream (no forking): measurement aid for bench.py."""
        nodes = self._nodes(ex, None)
        # the DAG order may put a selfprep root first; any topological order is valid here
        out, seen = [], set()
        pending = list(nodes)
        while pending:
            for i, (name, fn, deps) in enumerate(pending):
                if all(d in seen for d in deps):
                    out.append((name, fn))
                    seen.add(name)
                    pending.pop(i)
                    break
            else:
                raise RuntimeError("cyclic launch graph")
        out.append(("combine", lambda: self._launch_combine(ex)))
        return out

    def kernel_figures(self, ex):
        """Per stage of stages(ex): kernel name, cells and ALGORITHMIC bytes per launch
         and the reference-operation-order flops per cell of SURVEY 8d
        -- the inputs of bench.py's roofline objects."""
        Wh, N_E = ex.W, self.N_E
        out = {}
        for i, c in enumerate(self.comps):
            g = ex.preps[c["prep"]].grid
            if c["kind"] == "syn":
                out["syn%d" % i] = {
                    "kernel": "synchrotron_fused_kernel", "cells_per_launch": Wh * N_E * (g.N - 1),
                    "algorithmic_bytes_per_launch": 8 * (5 * g.N + N_E + Wh * (self.P + N_E)),
                    "ref_order_flops_per_cell": 320}
            elif c["kind"] == "ssc":
                tb = c["tb"]
                out["ssc_inner%d" % i] = {
                    "kernel": "ssc_inner_kernel (self-Compton, %d x %d rows, %d seed energies)"
                              % (N_E, g.N, tb.Ns),
                    "cells_per_launch": Wh * tb.R * (tb.Ns - 1),
                    "algorithmic_bytes_per_launch":
                        8 * (2 * tb.Ns * tb.R + Wh * (2 * tb.Ns + tb.R)),
                    "ref_order_flops_per_cell": 295, "is_ic": True}
            else:
                tb = c["table"]
                kind = type(c["obj"]).__name__
                out["table%d" % i] = {
                    "kernel": "contract_kernel (%s, %d x %d rows)" % (kind, tb.n_comp, N_E),
                    "cells_per_launch": Wh * tb.R * (g.N - 1),
                    "algorithmic_bytes_per_launch":
                        8 * (2 * tb.R * g.N + 2 * Wh * g.N + g.N + tb.R + Wh * tb.R),
                    "ref_order_flops_per_cell": 164, "is_ic": kind == "InverseCompton"}
        for ia, a in enumerate(self.aux):
            g = ex.preps[a["prep"]].grid
            ne = a["E_eV"].size
            out["aux_syn%d" % ia] = {
                "kernel": "synchrotron_fused_kernel (seed luminosities)",
                "cells_per_launch": Wh * ne * (g.N - 1),
                "algorithmic_bytes_per_launch": 8 * (5 * g.N + ne + Wh * (self.P + ne)),
                "ref_order_flops_per_cell": 320}
        return out

    def executable(self, W, pack=None):
        key = W if pack is None else (W, pack.data_ptr())
        if key not in self._exec:
            self._exec[key] = self._build(W, pack)
        return self._exec[key]

    # -- device-resident evaluation ---------------------------------------------------
    def run(self, ex):
        """Evaluate on the parameters already in ex.pars (device)."""
        if ex.graph is not None:
            ex.graph.replay()
        else:
            self._enqueue(ex)

    # -- host-facing evaluation (the sampler's vectorised lnprob) -----------------------
    def __call__(self, pars, want_blobs=True):
        """pars: host array [W][P] -> (lnp[W], flux[W][N_E] in data units, blob arrays)."""
        pars = np.ascontiguousarray(pars, dtype=float)
        if pars.ndim == 1:
            pars = pars[None, :]
        W = pars.shape[0]
        ex = self.executable(W)
        ex.pars_pin.numpy()[...] = pars
        ex.pars.copy_(ex.pars_pin, non_blocking=True)
        self.run(ex)
        ex.lnp_pin.copy_(ex.lnp, non_blocking=True)
        if want_blobs:
            ex.row_pin.copy_(ex.row, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        lnp = ex.lnp_pin.numpy().copy()
        if not want_blobs:
            return lnp, None, None
        flux, blob_arrays = self.split_rows(ex.row_pin.numpy().copy())
        return lnp, flux, blob_arrays

    def eval_rows(self, pars):
        """pars: host array [W][P] -> (lnp[W], per-walker records [W][row_width])."""
        pars = np.ascontiguousarray(pars, dtype=float)
        ex = self.executable(pars.shape[0])
        ex.pars_pin.numpy()[...] = pars
        ex.pars.copy_(ex.pars_pin, non_blocking=True)
        self.run(ex)
        ex.lnp_pin.copy_(ex.lnp, non_blocking=True)
        ex.row_pin.copy_(ex.row, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return ex.lnp_pin.numpy().copy(), ex.row_pin.numpy().copy()

    def split_rows(self, rows):
        """Per-walker records [..., row_width] -> (flux [..., N_E], [blob arrays])."""
        out = []
        for spec, (off, width) in zip(self._flat_blob_specs(), self.blob_cols):
            out.append(rows[..., off] if spec["kind"] == "W" else rows[..., off:off + width])
        return rows[..., :self.N_E], out

    def io_bytes(self, W, want_blobs=True):
        """(host->device, device->host) bytes of one host-facing evaluation."""
        d2h = 8 * W
        if want_blobs:
            d2h += 8 * W * self.row_width
        return 8 * W * self.P, d2h

    # -- blob reconstruction -----------------------------------------------------------
    def blobs_for(self, flux, blob_arrays, w):
        """The reference's per-walker blob tuple (core.py:106-121): model flux in the
        model's unit followed by the user's blobs."""
        flat = {id(s): a for s, a in zip(self._flat_blob_specs(), blob_arrays)}

        def build(s):
            if s["kind"] == "tuple":
                return tuple(build(x) for x in s["items"])
            if s["kind"] == "const":
                return s["value"]
            a = flat.get(id(s))
            if a is None:  # blob array not gathered (sharded evaluation keeps the flux only)
                return None
            a = a[w]
            if s["kind"] == "pdist":
                a = a.reshape(s["shape"])
            return Quantity(a if np.ndim(a) else float(a), s["unit"])
        model = Quantity(flux[w] * self.to_model_unit, self.model_unit)
        if not self.blob_specs:
            return (model, np.nan)
        return (model,) + tuple(build(s) for s in self.blob_specs)
