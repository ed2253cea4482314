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
    """Sum of groups ``(sum of component spectra) / (4 pi d^2)`` in a common unit."""

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
        return SymFlux(self.groups + o.groups, self.E, self.unit, self.sed, self.base_unit)

    __radd__ = __add__

    def __mul__(self, o):
        if np.ndim(o) == 0 and not isinstance(o, (Quantity, Unit, SymPar)):
            return SymFlux([(c, d, s * float(o)) for c, d, s in self.groups], self.E, self.unit,
                           self.sed, self.base_unit)
        raise TraceError("unsupported product with a flux")

    __rmul__ = __mul__


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
        self.selfprep_kinds = ("syn",)
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
        self.scalars = []  # per-walker scalar columns: SymPar or float

        def pd_index(pd):
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
        for gi, (comp, div, scale) in enumerate(self.flux.groups):
            ipd = pd_index(comp.particle_distribution)
            grid = comp._grid()
            c = {"group": gi, "div": div / scale, "obj": comp}
            if isinstance(comp, M.Synchrotron):
                c["kind"] = "syn"
                c["prep"] = prep_index(ipd, grid, False)
                c["B"] = scalar_index(M._val(comp.B, "G"))
            elif isinstance(comp, M.InverseCompton):
                c["kind"] = "table"
                c["prep"] = prep_index(ipd, grid, exact)
                seeds = []
                for name, sd in comp.seed_photon_fields.items():
                    if sd["type"] == "array" and sd["photon_density"].ndim == 2:
                        raise TraceError("per-walker seed photon fields are not traced")
                    seeds.append(comp._seed_tuple(sd))
                c["table"] = eng.ic_table(grid, self.E_eV, tuple(seeds))
                c["rows"] = [(s * self.N_E, None) for s in range(len(seeds))]
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
        # components whose kernels derive the walker's operands themselves (no set-up
        # launch in front of them); the reference-order (exact) mode keeps the operand arrays
        # Synchrotron only: its CTAs own one walker, so the operands cost one pass over the
        # nodes; the contraction would repeat them in every row-tile CTA (+30 % work,
        # measured slower than reading the arrays of the set-up kernel)
        for c in self.comps:
            c["selfprep"] = (self.selfprep and not exact and self.P <= 32
                             and c["kind"] in self.selfprep_kinds)
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
        n_groups = len(self.comps)
        for c in self.comps:
            if c["kind"] == "syn":
                out = eng.zeros(W, self.N_E)
                terms.append((out, 0, True, c["div"], None))
            else:
                out = eng.zeros(W, c["table"].R)
                rows = c["rows"]
                for k, (off, sc) in enumerate(rows):
                    terms.append((out, off, k == len(rows) - 1, c["div"],
                                  scalar_col(sc) if sc is not None else None))
            ex.outs.append(out)
        ex.terms = eng.make_terms(terms)
        # one record per walker: [model flux (N_E) | further blobs], so that a sampler
        # moves a walker's blobs with one row copy (row pitch = self.row_width)
        ex.row = eng.zeros(W, self.row_width) if pack is None else pack[:, :self.row_width]
        ex.row_ld = ex.row.stride(0)
        ex.flux = ex.row[:, :self.N_E]
        ex.lnp = eng.zeros(W) if pack is None else pack[:, self.row_width]
        ex.E_erg = eng.to_dev(self.E_eV * eng.eV_erg)
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
            g = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                with torch.cuda.graph(g, stream=s):
                    self._enqueue(ex)
            torch.cuda.current_stream().wait_stream(s)
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

    def _launch_comp(self, ex, c, out, src):
        L, W = lib(), ex.W
        p = ex.preps[c["prep"]]
        g = p.grid
        if c["selfprep"]:
            d = ex.pd_desc[c["prep"]]
            if c["kind"] == "syn":
                check(L.nb_synchrotron_fused(
                    ctypes.byref(src), ctypes.byref(d), ex.scalar_entry[c["B"]],
                    eng.ptr(g.x_d), g.N, eng.ptr(g.dlx_d), W, eng.ptr(ex.E_erg), self.N_E,
                    eng.ptr(out), eng.stream()), "nb_synchrotron_fused")
            else:
                tb = c["table"]
                check(L.nb_contract_fused(
                    ctypes.byref(src), ctypes.byref(d), eng.ptr(tb.K), eng.ptr(tb.lrs), tb.R,
                    g.N, g.pitch, W, eng.ptr(g.dlx_d), eng.ptr(g.x_d), eng.ptr(tb.coef),
                    eng.ptr(out), eng.stream()), "nb_contract_fused")
        elif c["kind"] == "syn":
            eng.synchrotron(g, p, ex.scalar_col(c["B"]), ex.E_erg, out=out)
        else:
            eng.contract(c["table"], p, out=out)

    def _enqueue(self, ex, mv=None, fuse_update=True, peers=None):
        """The launch sequence of one likelihood evaluation of ex.W walkers.  With `mv`
        (an nb_stretch describing a device-resident ensemble) the parameters are the
        stretch-move proposals of the active half, computed by the set-up kernel, and
        (fuse_update) the combine kernel also accepts/rejects them and appends the chain."""
        L, W = lib(), ex.W
        n = 0
        if mv is None and ex.pack is not None:
            raise ValueError("a packed executable is driven by device-side proposals only")

        src = self._src(ex, mv)

        def prep():
            self._launch_prep(ex, mv)

        def launch(c, out):
            self._launch_comp(ex, c, out, src)

        # The set-up kernel and the radiative components fork onto side streams (parallel
        # branches of the captured graph) and join before the combine.  Components that
        # still consume operand arrays must follow the set-up kernel.
        comps = list(zip(self.comps, ex.outs))
        main = torch.cuda.current_stream()
        dependent = [x for x in comps if not x[0]["selfprep"]]
        free = [x for x in comps if x[0]["selfprep"]]
        branches = [[prep] + [lambda c=c, o=o: launch(c, o) for c, o in dependent[:1]]]
        branches += [[lambda c=c, o=o: launch(c, o)] for c, o in free]
        if ex.n_blob_jobs:
            branches.append([lambda: self._launch_blob_prep(ex, mv)])
        extra_dep = dependent[1:]
        joins = []
        if len(branches) > 1 or extra_dep:
            while len(self._side) < len(branches) + len(extra_dep):
                self._side.append(torch.cuda.Stream())
            fork = torch.cuda.Event()
            fork.record(main)
            for br, side in zip(branches[1:], self._side):
                side.wait_event(fork)
                with torch.cuda.stream(side):
                    for f in br:
                        f()
                    done = torch.cuda.Event()
                    done.record(side)
                joins.append(done)
        for f in branches[0]:
            f()
        if extra_dep:  # further operand-consuming components: fork after the set-up
            after = torch.cuda.Event()
            after.record(main)
            for (c, o), side in zip(extra_dep, self._side[len(branches) - 1:]):
                side.wait_event(after)
                with torch.cuda.stream(side):
                    launch(c, o)
                    done = torch.cuda.Event()
                    done.record(side)
                joins.append(done)
        for done in joins:
            main.wait_event(done)
        n += 1 + len(comps) + (1 if ex.n_blob_jobs else 0)
        L, st = lib(), eng.stream()
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
        n += 1
        self.launches_per_eval = n
        return n

    def stages(self, ex):
        """[(name, launch)] of one evaluation on ex.pars, in launch order on the current
        stream (no forking): measurement aid for bench.py / tools/timeline.py."""
        out = [("walker_prep", lambda: self._launch_prep(ex, None))]
        if ex.n_blob_jobs:
            out.append(("blob_prep", lambda: self._launch_blob_prep(ex, None)))
        for i, (c, o) in enumerate(zip(self.comps, ex.outs)):
            name = ("syn%d" if c["kind"] == "syn" else "table%d") % i
            out.append((name, lambda c=c, o=o: self._launch_comp(ex, c, o, ex.src)))
        out.append(("combine", lambda: eng.combine(
            ex.terms, ex.W, self.N_E, self.unit_fac_d, flux_out=ex.row, data=self.ddata,
            prior_d=ex.prior if self.prior is not None else None, lnp_out=ex.lnp,
            flux_ld=ex.row_ld, lnp_ld=ex.lnp.stride(0))))
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
