# -*- coding: utf-8 -*-
"""Host-side mirror of ``naima.models`` / ``naima.radiative`` for the likelihood
hot path: same class names, constructor arguments, attributes, methods and
error behaviour; all arithmetic runs in the sm_100a kernels behind the C ABI.

Extension over the reference: every free parameter of a particle distribution
(and ``B``, ``n0``, ``nh``) may be an array of length W -- a batch of walkers.
``flux``/``sed`` then return ``[W, N_E]`` and ``We``/``Wp`` return ``[W]``.  A
user ``model(pars, data)`` written for the reference therefore also works when
``pars`` is a ``[P, W]`` array (see core.lnprob), which is how a whole
half-ensemble is evaluated in one set of launches.

Reference: src/naima/models.py:49-422, src/naima/radiative.py:43-1536.
"""
import logging
from collections import OrderedDict

import numpy as np

from . import engine as eng
from . import units as u
from .units import Quantity

__all__ = [
    "Synchrotron", "InverseCompton", "PionDecay", "Bremsstrahlung", "BrokenPowerLaw",
    "ExponentialCutoffPowerLaw", "PowerLaw", "LogParabola", "ExponentialCutoffBrokenPowerLaw",
    "TableModel", "EblAbsorptionModel", "PionDecayKelner06",
]

log = logging.getLogger("naima_b200.models")

mec2 = Quantity(eng.mec2_erg, u.erg)
mec2_unit = u.Unit("mec2")


# ------------------------------------------------------------------------------
# validators (extern/validator.py:8-85 error conventions)
# ------------------------------------------------------------------------------
def _is_sym(x):
    from .fused import SymPar
    return isinstance(x, SymPar)


def validate_physical_type(name, value, physical_type):
    if physical_type is not None and _is_sym(value):
        pts = [physical_type] if isinstance(physical_type, str) else physical_type
        if value.unit.physical_type not in pts:
            raise TypeError("{0} should be given in units of {1}".format(name, ", ".join(pts)))
        return
    if physical_type is not None:
        if not u._is_quantity(value):
            raise TypeError("{0} should be given as a Quantity object".format(name))
        pt = Quantity(value).unit.physical_type
        if isinstance(physical_type, str):
            if pt != physical_type:
                raise TypeError("{0} should be given in units of {1}".format(name, physical_type))
        elif pt not in physical_type:
            raise TypeError(
                "{0} should be given in units of {1}".format(name, ", ".join(physical_type)))


def validate_scalar(name, value, domain=None, physical_type=None):
    validate_physical_type(name, value, physical_type)
    if _is_sym(value):
        return value
    if u._is_quantity(value):
        value = Quantity(value)
    if not physical_type:
        if not np.isscalar(value) or not np.isreal(value):
            raise TypeError("{0} should be a scalar floating point value".format(name))
    v = value.value if isinstance(value, Quantity) else value
    if domain == "positive" and np.any(np.asarray(v) < 0.0):
        raise ValueError("{0} should be positive".format(name))
    if domain == "strictly-positive" and np.any(np.asarray(v) <= 0.0):
        raise ValueError("{0} should be strictly positive".format(name))
    return value


def validate_array(name, value, domain=None, ndim=1, physical_type=None):
    validate_physical_type(name, value, physical_type)
    if u._is_quantity(value):
        value = Quantity(value)
    arr = np.asarray(value.value if isinstance(value, Quantity) else value)
    if arr.ndim != ndim:
        raise TypeError("{0} should be a {1:d}-d array".format(name, ndim) if ndim != 1
                        else "{0} should be a 1-d sequence".format(name))
    if domain == "positive" and np.any(arr < 0.0):
        raise ValueError("{0} should be positive".format(name))
    return value


def _validate_ene(ene):
    """radiative.py:43-58 / models.py:33-46: Quantity, dict or table with 'energy'."""
    if isinstance(ene, dict) or hasattr(ene, "colnames"):
        try:
            ene = Quantity(ene["energy"])
        except KeyError:
            raise TypeError("Table or dict does not have 'energy' column")
        validate_physical_type("energy", ene, "energy")
    else:
        if not u._is_quantity(ene):
            ene = Quantity(ene)
        else:
            ene = Quantity(ene)
        validate_physical_type("energy", ene, physical_type="energy")
    return ene


def _is_symflux(x):
    from .fused import SymFlux
    return isinstance(x, SymFlux)


def _val(x, unit=None):
    """Plain float/ndarray of a parameter (Quantity converted to `unit`)."""
    if _is_sym(x):
        return x.plain(unit)
    if u._is_quantity(x):
        x = Quantity(x)
        return x.to(unit).value if unit is not None else x._dimensionless_value()
    return np.asarray(x, dtype=float) if np.ndim(x) else float(x)


# ------------------------------------------------------------------------------
# particle distributions (models.py:49-422)
# ------------------------------------------------------------------------------
class _ParticleDistribution:
    _energy_pars = ()

    def _amplitude_unit(self):
        if _is_sym(self.amplitude):
            return self.amplitude.unit
        return Quantity(self.amplitude).unit if u._is_quantity(self.amplitude) else u.Unit()

    def _symbolic(self):
        return any(_is_sym(getattr(self, n)) for n in self.param_names)

    def _eval_params(self, amplitude_value=None):
        """Reference-order eval() arguments after `e`, energies in eV."""
        out = []
        for name in self.param_names:
            v = getattr(self, name)
            if name == "amplitude":
                if amplitude_value is not None:
                    out.append(amplitude_value)
                elif _is_sym(v):
                    out.append(v.value)
                else:
                    out.append(_val(Quantity(v).value if u._is_quantity(v) else v))
            elif name in self._energy_pars:
                out.append(_val(v, "eV"))
            else:
                out.append(_val(v))
        return out

    def _device_params(self, to_unit="1/eV"):
        """(kind, host [W][8]) with the amplitude converted to `to_unit`."""
        amp = self.amplitude
        if not u._is_quantity(amp):
            raise TypeError("Particle distribution should be given as a Quantity object")
        a = Quantity(amp).to(to_unit).value
        return self._kind, eng.pd_params_array(self._kind, self._eval_params(a))

    @property
    def batch(self):
        return self._device_params()[1].shape[0]

    def _calc(self, e):
        kind = self._kind
        par = eng.pd_params_array(kind, self._eval_params())
        W = par.shape[0]
        e_eV = np.atleast_1d(e.to("eV").value)
        out = eng.pdist_eval(kind, eng.to_dev(par), W, e_eV.ravel()).cpu().numpy()
        out = out.reshape((W,) + e_eV.shape)
        if W == 1:
            out = out[0]
        if np.ndim(e.value) == 0:
            out = out[..., 0]
        return Quantity(out, self._amplitude_unit())

    def __call__(self, e):
        e = _validate_ene(e)
        if self._symbolic():
            from .fused import SymBlob
            return SymBlob("pdist", pd=self, e_eV=np.atleast_1d(e.to("eV").value).ravel(),
                           unit=self._amplitude_unit(), shape=np.shape(e.value))
        return self._calc(e)

    @classmethod
    def eval(cls, e, *params):
        """The model function itself on plain numbers (evaluated on the device)."""
        par = eng.pd_params_array(cls._kind, params)
        e = np.atleast_1d(np.asarray(e, dtype=float))
        out = eng.pdist_eval(cls._kind, eng.to_dev(par), par.shape[0], e.ravel()).cpu().numpy()
        return out[0].reshape(e.shape) if par.shape[0] == 1 else out


class PowerLaw(_ParticleDistribution):
    """f(E) = A (E/E0)^-alpha  (models.py:49-106)."""
    param_names = ["amplitude", "e_0", "alpha"]
    _energy_pars = ("e_0",)
    _kind = "PowerLaw"

    def __init__(self, amplitude, e_0, alpha):
        self.amplitude = amplitude
        self.e_0 = validate_scalar("e_0", e_0, domain="positive", physical_type="energy")
        self.alpha = alpha


class ExponentialCutoffPowerLaw(_ParticleDistribution):
    """f(E) = A (E/E0)^-alpha exp(-(E/Ec)^beta)  (models.py:109-177)."""
    param_names = ["amplitude", "e_0", "alpha", "e_cutoff", "beta"]
    _energy_pars = ("e_0", "e_cutoff")
    _kind = "ExponentialCutoffPowerLaw"

    def __init__(self, amplitude, e_0, alpha, e_cutoff, beta=1.0):
        self.amplitude = amplitude
        self.e_0 = validate_scalar("e_0", e_0, domain="positive", physical_type="energy")
        self.alpha = alpha
        self.e_cutoff = validate_scalar("e_cutoff", e_cutoff, domain="positive",
                                        physical_type="energy")
        self.beta = beta


class BrokenPowerLaw(_ParticleDistribution):
    """models.py:180-254."""
    param_names = ["amplitude", "e_0", "e_break", "alpha_1", "alpha_2"]
    _energy_pars = ("e_0", "e_break")
    _kind = "BrokenPowerLaw"

    def __init__(self, amplitude, e_0, e_break, alpha_1, alpha_2):
        self.amplitude = amplitude
        self.e_0 = validate_scalar("e_0", e_0, domain="positive", physical_type="energy")
        self.e_break = validate_scalar("e_break", e_break, domain="positive",
                                       physical_type="energy")
        self.alpha_1 = alpha_1
        self.alpha_2 = alpha_2


class ExponentialCutoffBrokenPowerLaw(_ParticleDistribution):
    """models.py:257-354."""
    param_names = ["amplitude", "e_0", "e_break", "alpha_1", "alpha_2", "e_cutoff", "beta"]
    _energy_pars = ("e_0", "e_break", "e_cutoff")
    _kind = "ExponentialCutoffBrokenPowerLaw"

    def __init__(self, amplitude, e_0, e_break, alpha_1, alpha_2, e_cutoff, beta=1.0):
        self.amplitude = amplitude
        self.e_0 = validate_scalar("e_0", e_0, domain="positive", physical_type="energy")
        self.e_break = validate_scalar("e_break", e_break, domain="positive",
                                       physical_type="energy")
        self.alpha_1 = alpha_1
        self.alpha_2 = alpha_2
        self.e_cutoff = validate_scalar("e_cutoff", e_cutoff, domain="positive",
                                        physical_type="energy")
        self.beta = beta


class LogParabola(_ParticleDistribution):
    """models.py:357-422."""
    param_names = ["amplitude", "e_0", "alpha", "beta"]
    _energy_pars = ("e_0",)
    _kind = "LogParabola"

    def __init__(self, amplitude, e_0, alpha, beta):
        self.amplitude = amplitude
        self.e_0 = validate_scalar("e_0", e_0, domain="positive", physical_type="energy")
        self.alpha = alpha
        self.beta = beta


class TableModel(_ParticleDistribution):
    """A model generated from a table of energy and value arrays, interpolated in log-log
    space with a cubic spline and 0 outside the table (models.py:425-469).

    As the particle distribution of a radiative class the table is interpolated ONCE per
    particle grid on the host (it does not depend on the walker); the walkers differ by
    ``amplitude`` only, so the integration operands are x * amplitude_w * n0(x) and the
    walker-independent log-slope of n0 -- the kernels downstream are the usual ones."""
    param_names = ["amplitude"]
    _kind = "TableModel"

    def __init__(self, energy, values, amplitude=1):
        from scipy.interpolate import interp1d

        self._energy = validate_array("energy", energy, domain="positive",
                                      physical_type="energy")
        self._values = values
        self.amplitude = amplitude
        loge = np.log10(Quantity(self._energy).to("eV").value)
        if u._is_quantity(values):
            self.unit = Quantity(values).unit
            vals = np.asarray(Quantity(values).value, dtype=float)
        else:
            self.unit = u.Unit()
            vals = np.asarray(values, dtype=float)
        with np.errstate(divide="ignore"):
            logy = np.log10(vals)
        self._interplogy = interp1d(loge, logy, fill_value=-np.inf, bounds_error=False,
                                    kind="cubic")

    def _symbolic(self):
        if _is_sym(self.amplitude):
            from .fused import TraceError
            raise TraceError("TableModel particle distributions are not traced")
        return False

    def _amplitude_unit(self):
        return self.unit

    def _table(self, e_eV):
        with np.errstate(all="ignore"):
            return np.power(10, self._interplogy(np.log10(np.asarray(e_eV, dtype=float))))

    @property
    def batch(self):
        return np.size(self.amplitude)

    def _calc(self, e):
        interpy = self._table(e.to("eV").value)
        amp = np.asarray(self.amplitude, dtype=float)
        if amp.ndim:
            interpy = amp.reshape((-1,) + (1,) * np.ndim(interpy)) * interpy
        else:
            interpy = float(amp) * interpy
        return Quantity(interpy, self.unit)

    def _prepared(self, grid, W, need_raw, to_unit="1/eV"):
        """Integration operands of W walkers on `grid` (what nb_pd_prep produces for the
        parametrised distributions), from the host-side interpolation."""
        import torch

        x = grid.x
        fac = self.unit._factor_to(to_unit) * grid.n_scale
        n0 = self._table((x * grid.e_mul1) * grid.e_mul2) * fac
        amp = np.broadcast_to(np.atleast_1d(np.asarray(self.amplitude, dtype=float)), (W,))
        pitch = grid.pitch
        xn0, ds = np.zeros(pitch), np.zeros(pitch)
        xn0[: grid.N] = x * n0
        live = (n0[:-1] != 0) & (n0[1:] != 0)
        dl = np.log(x[1:] / x[:-1])
        with np.errstate(all="ignore"):
            ds[: grid.N - 1] = np.where(live, np.log(n0[1:] / n0[:-1]) / dl + 1.0,
                                        eng.BIG_SLOPE)
        pr = eng.Prepared()
        amp_d = eng.to_dev(amp)
        pr.xn = (amp_d[:, None] * eng.to_dev(xn0)[None, :]).contiguous()
        pr.ds1 = eng.to_dev(ds)[None, :].expand(W, pitch).contiguous()
        pr.nraw = None
        if need_raw:
            nr = np.zeros(pitch)
            nr[: grid.N] = n0
            pr.nraw = (amp_d[:, None] * eng.to_dev(nr)[None, :]).contiguous()
        pr.W, pr.grid = W, grid
        return pr

    def _energy_on(self, grid, W):
        """Total particle energy [erg] on `grid`: trapz_loglog(x n, x) in the reference's
        operation order (the stand-alone device op)."""
        x = grid.x
        n0 = self._table((x * grid.e_mul1) * grid.e_mul2) * \
            (self.unit._factor_to("1/eV") * grid.n_scale)
        amp = np.broadcast_to(np.atleast_1d(np.asarray(self.amplitude, dtype=float)), (W,))
        y = amp[:, None] * (x * n0)[None, :]
        return eng.trapz_loglog(y, x * grid.x_to_erg)


class EblAbsorptionModel(TableModel):
    """Optical depth of the extragalactic background light (Dominguez et al. 2011) at the
    tabulated redshift closest to ``redshift``; ``transmission(e)`` is the dimensionless
    factor to multiply a model with (models.py:472-552).  Host-side: the factor does not
    depend on the walkers unless the redshift is fitted."""

    def __init__(self, redshift, ebl_absorption_model="Dominguez"):
        if _is_sym(redshift):
            from .fused import TraceError
            raise TraceError("a fitted redshift is not traced")
        if not u._is_quantity(redshift):
            redshift = Quantity(redshift, u.dimensionless_unscaled)
        validate_physical_type("redshift", redshift, "dimensionless")
        z = np.asarray(Quantity(redshift).value, dtype=float)
        if z.ndim != 0:
            raise TypeError("redshift should be a scalar floating point value")
        if z < 0:
            raise ValueError("redshift should be positive")
        self.redshift = Quantity(float(z), u.dimensionless_unscaled)
        self.model = ebl_absorption_model
        if self.model != "Dominguez":
            raise ValueError('Model should be one of: ["Dominguez"]')
        import os

        f = np.load(os.path.join(eng.DATA_DIR, "ebl_dominguez11.npz"))
        energy = Quantity(f["energy_TeV"], "TeV")
        redshift_list = np.arange(0.01, 4, 0.01)
        if z >= 0.01:
            table_values = f["tau"][np.abs(redshift_list - float(z)).argmin()].copy()
            table_values[table_values > 150.0] = 150.0  # models.py:530-532
            taus = 10 ** table_values
        else:
            taus = 10 ** np.zeros(energy.size)
        super().__init__(energy, taus)

    def transmission(self, e):
        e = _validate_ene(e)
        E_eV = np.atleast_1d(e.to("eV").value).astype(float)
        taus = np.zeros(E_eV.size)
        mid = (E_eV >= 1e9) & (E_eV <= 100e12)
        taus[E_eV > 100e12] = np.log10(6000.0)
        if np.any(mid):
            with np.errstate(all="ignore"):
                taus[mid] = np.log10(self._table(E_eV[mid]))
        return np.exp(-taus)


# ------------------------------------------------------------------------------
# radiative models
# ------------------------------------------------------------------------------
class BaseRadiative:
    """flux/sed on top of a per-class device spectrum (radiative.py:61-134)."""

    def __init__(self, particle_distribution):
        self.particle_distribution = particle_distribution
        if not isinstance(particle_distribution, _ParticleDistribution):
            raise TypeError(
                "naima_b200 evaluates particle distributions on the device: use PowerLaw, "
                "ExponentialCutoffPowerLaw, BrokenPowerLaw, ExponentialCutoffBrokenPowerLaw "
                "or LogParabola (arbitrary callables are out of scope of the hot path)")
        if isinstance(particle_distribution, TableModel):
            if particle_distribution.unit.physical_type != "differential energy":
                raise TypeError("Particle distribution should be given in units of "
                                "differential energy")
        else:
            validate_physical_type("Particle distribution", particle_distribution.amplitude,
                                   physical_type="differential energy")

    # -- device plumbing ---------------------------------------------------------
    def _pd_device(self):
        kind, par = self.particle_distribution._device_params("1/eV")
        return kind, eng.to_dev(par), par.shape[0]

    def _prep(self, g, W=None, need_raw=None):
        """Integration operands of this model's particle distribution on grid g for W walkers."""
        pd = self.particle_distribution
        if isinstance(pd, TableModel):
            need_raw = eng.EXACT if need_raw is None else need_raw
            return pd._prepared(g, pd.batch if W is None else W, need_raw)
        kind, par_d, Wp = self._pd_device()
        W = Wp if W is None else W
        if Wp != W:
            par_d = par_d.expand(W, par_d.shape[1]).contiguous()
        return eng.pd_prep(g, kind, par_d, W, need_raw=need_raw)

    def _particle_energy(self, grid):
        """Total particle energy [erg] on `grid`, one value per walker (host array)."""
        pd = self.particle_distribution
        if isinstance(pd, TableModel):
            return pd._energy_on(grid, pd.batch)
        kind, par_d, W = self._pd_device()
        return eng.particle_energy(grid, kind, par_d, W).cpu().numpy()

    def _batch(self):
        """Batch size W of this model (1 for the reference's scalar use)."""
        sizes = [self.particle_distribution.batch]
        sizes += [np.size(_val(getattr(self, n))) if not u._is_quantity(getattr(self, n))
                  else np.size(Quantity(getattr(self, n)).value) for n in self._walker_scalars]
        W = max(sizes)
        if any(s not in (1, W) for s in sizes):
            raise ValueError("inconsistent batch sizes among model parameters")
        return W

    _walker_scalars = ()

    def _symbolic(self):
        return self.particle_distribution._symbolic() or any(
            _is_sym(getattr(self, n)) for n in self._walker_scalars)

    def _is_batched(self):
        pd = self.particle_distribution
        vals = [getattr(pd, n) for n in pd.param_names] + [getattr(self, n)
                                                           for n in self._walker_scalars]
        return any(np.ndim(Quantity(v).value if u._is_quantity(v) else v) > 0 for v in vals)

    def _spectrum_dev(self, E_eV):
        """Device tensor [W][N_E] in 1/(s eV)."""
        terms, W = self._terms(E_eV)
        out = eng.empty(W, E_eV.size)
        eng.combine(terms, W, E_eV.size, eng.to_dev(np.ones(E_eV.size)), flux_out=out)
        return out

    def _spectrum(self, photon_energy):
        E = _validate_ene(photon_energy)
        E_eV = np.atleast_1d(E.to("eV").value).astype(float)
        spec = self._spectrum_dev(E_eV).cpu().numpy()
        return self._shape_out(spec, E), E

    def _shape_out(self, arr, E):
        if not self._is_batched():
            arr = arr[0]
        if np.ndim(E.value) == 0:
            arr = arr[..., 0]
        return arr

    def flux(self, photon_energy, distance=1 * u.kpc):
        """Differential flux at `distance` (0: intrinsic luminosity); radiative.py:88-111."""
        if self._symbolic():
            from .fused import SymFlux
            return SymFlux.from_component(self, _validate_ene(photon_energy), distance, False)
        spec, _ = self._spectrum(photon_energy)
        if _nonzero(distance):
            distance = validate_scalar("distance", distance, physical_type="length")
            spec = spec / (4 * np.pi * distance.to("cm").value ** 2)
            return Quantity(spec, "1/(s cm2 eV)")
        return Quantity(spec, "1/(s eV)")

    def sed(self, photon_energy, distance=1 * u.kpc):
        """Spectral energy distribution (radiative.py:113-134)."""
        out_unit = "erg/(cm2 s)" if _nonzero(distance) else "erg/s"
        photon_energy = _validate_ene(photon_energy)
        if self._symbolic():
            from .fused import SymFlux
            return SymFlux.from_component(self, photon_energy, distance, True)
        return (self.flux(photon_energy, distance) * photon_energy**2.0).to(out_unit)


def _nonzero(distance):
    v = Quantity(distance).value if u._is_quantity(distance) else distance
    return bool(np.all(np.asarray(v) != 0))


class BaseElectron(BaseRadiative):
    """Electron grid, We (radiative.py:137-236)."""

    def __init__(self, particle_distribution):
        super().__init__(particle_distribution)
        self.param_names = ["Eemin", "Eemax", "nEed"]

    def _grid(self, Eemin=None, Eemax=None):
        Eemin = self.Eemin if Eemin is None else Eemin
        Eemax = self.Eemax if Eemax is None else Eemax
        return eng.electron_grid(Quantity(Eemin).to("eV").value, Quantity(Eemax).to("eV").value,
                                 self.nEed)

    @property
    def _gam(self):
        return self._grid().x.copy()

    @property
    def _nelec(self):
        g = self._grid()
        pr = self._prep(g, need_raw=True)
        n = pr.nraw[:, : g.N].cpu().numpy()
        return n if self._is_batched() else n[0]

    def _We(self, grid):
        if self._symbolic():
            from .fused import SymBlob
            return SymBlob("W", comp=self, grid=grid)
        We = self._particle_energy(grid)
        return Quantity(We if self._is_batched() else float(We[0]), u.erg)

    @property
    def We(self):
        """Total energy in electrons used for the radiative calculation."""
        return self._We(self._grid())

    def compute_We(self, Eemin=None, Eemax=None):
        """Total energy in electrons between Eemin and Eemax (radiative.py:168-195)."""
        if Eemin is None and Eemax is None:
            return self.We
        return self._We(self._grid(Eemin, Eemax))

    def set_We(self, We, Eemin=None, Eemax=None, amplitude_name=None):
        """Normalise the particle distribution to a total energy (radiative.py:197-236)."""
        We = validate_scalar("We", We, physical_type="energy")
        oldWe = self.compute_We(Eemin=Eemin, Eemax=Eemax)
        ratio = (We / oldWe).decompose().value
        name = "amplitude" if amplitude_name is None else amplitude_name
        try:
            setattr(self.particle_distribution, name,
                    getattr(self.particle_distribution, name) * ratio)
        except AttributeError:
            log.error("The particle distribution does not have an attribute called %s to "
                      "modify its normalization", name)


class Synchrotron(BaseElectron):
    """Synchrotron emission, random magnetic field (radiative.py:239-342)."""
    _walker_scalars = ("B",)

    def __init__(self, particle_distribution, B=3.24e-6 * u.G, **kwargs):
        super().__init__(particle_distribution)
        self.B = validate_scalar("B", B, physical_type="magnetic flux density")
        self.Eemin = 1 * u.GeV
        self.Eemax = 1e9 * mec2
        self.nEed = 100
        self.param_names += ["B"]
        self.__dict__.update(**kwargs)

    def _terms(self, E_eV):
        W = self._batch()
        B = np.broadcast_to(np.atleast_1d(Quantity(self.B).to("G").value).astype(float), (W,))
        g = self._grid()
        pr = self._prep(g, W, need_raw=False)
        out = eng.synchrotron(g, pr, eng.to_dev(B), eng.photon_energies(E_eV))
        return [(out, 0, True, 1.0, None)], W


class InverseCompton(BaseElectron):
    """IC on grey-body, monochromatic and tabulated seed photon fields
    (radiative.py:370-791)."""

    def __init__(self, particle_distribution, seed_photon_fields=["CMB"], **kwargs):
        super().__init__(particle_distribution)
        self.seed_photon_fields = self._process_input_seed(seed_photon_fields)
        self.Eemin = 1 * u.GeV
        self.Eemax = 1e9 * mec2
        self.nEed = 100
        self.param_names += ["seed_photon_fields"]
        self.__dict__.update(**kwargs)

    @staticmethod
    def _process_input_seed(seed_photon_fields):
        """radiative.py:432-545."""
        ar = Quantity(eng.ar_cgs, "erg/(cm3 K4)")
        Tcmb = eng.T_CMB * u.K
        if type(seed_photon_fields) is not list:
            seed_photon_fields = seed_photon_fields.split("-")
        result = OrderedDict()
        for idx, inseed in enumerate(seed_photon_fields):
            seed = {}
            if isinstance(inseed, str):
                name = inseed
                seed["type"] = "thermal"
                seed["isotropic"] = True
                if inseed == "CMB":
                    seed["T"], seed["u"] = Tcmb, ar * Tcmb**4
                elif inseed == "FIR":
                    seed["T"], seed["u"] = 30 * u.K, 0.5 * u.eV / u.cm**3
                elif inseed == "NIR":
                    seed["T"], seed["u"] = 3000 * u.K, 1.0 * u.eV / u.cm**3
                else:
                    log.warning("Will not use seed {0} because it is not CMB, FIR or NIR"
                                .format(inseed))
                    raise TypeError
            elif type(inseed) is list and (len(inseed) == 3 or len(inseed) == 4):
                isotropic = len(inseed) == 3
                if isotropic:
                    name, T, uu = inseed
                    seed["isotropic"] = True
                else:
                    name, T, uu, theta = inseed
                    seed["isotropic"] = False
                    seed["theta"] = validate_scalar("{0}-theta".format(name), theta,
                                                    physical_type="angle")
                if not u._is_quantity(T):
                    raise TypeError("Unable to process seed photon field: {0}".format(inseed))
                T = Quantity(T)
                if T.unit.physical_type == "temperature":
                    seed["type"] = "thermal"
                    validate_scalar("{0}-T".format(name), T, domain="positive",
                                    physical_type="temperature")
                    seed["T"] = T
                    if not u._is_quantity(uu) and uu == 0:
                        seed["u"] = ar * T**4
                    elif u._is_quantity(uu) and np.all(Quantity(uu).value == 0):
                        seed["u"] = ar * T**4
                    else:
                        validate_scalar("{0}-u".format(name), uu, domain="positive",
                                        physical_type="pressure")
                        seed["u"] = Quantity(uu)
                elif _is_symflux(uu):
                    # seed density computed from the fit parameters (synchrotron
                    # self-Compton, examples/CrabNebula_SynSSC.py:24-36): traced
                    seed["type"] = "array"
                    seed["energy"] = validate_array("{0}-energy".format(name), T,
                                                    domain="positive", physical_type="energy")
                    uu.check_seed_density(name, seed["energy"])
                    seed["photon_density"] = uu
                    seed["symbolic"] = True
                else:
                    seed["type"] = "array"
                    T = Quantity(np.atleast_1d(T.value), T.unit)
                    uu = Quantity(uu)
                    if uu.ndim == 0:
                        uu = Quantity(np.atleast_1d(uu.value), uu.unit)
                    seed["energy"] = validate_array("{0}-energy".format(name), T,
                                                    domain="positive", physical_type="energy")
                    if seed["energy"].size == 1:
                        validate_physical_type("{0}-density".format(name), uu, "pressure")
                        seed["photon_density"] = uu
                    else:
                        if uu.unit.physical_type == "pressure":
                            uu = uu / seed["energy"] ** 2
                        validate_physical_type("{0}-density".format(name), uu,
                                               "differential number density")
                        if uu.ndim not in (1, 2) or uu.shape[-1] != seed["energy"].size:
                            raise TypeError("{0}-density should be a 1-d sequence".format(name))
                        seed["photon_density"] = uu
            else:
                raise TypeError("Unable to process seed photon field: {0}".format(inseed))
            result[name] = seed
        return result

    def _symbolic(self):
        return super()._symbolic() or any(sd.get("symbolic")
                                          for sd in self.seed_photon_fields.values())

    def _is_batched(self):
        # a per-walker seed density [W, N_s] batches the model like array-valued parameters do
        return super()._is_batched() or any(
            sd["type"] == "array" and not sd.get("symbolic") and sd["photon_density"].ndim == 2
            for sd in self.seed_photon_fields.values())

    def _seed_tuple(self, seed):
        if seed["type"] == "thermal":
            t = ("thermal", float(seed["T"].to("K").value), float(seed["u"].to("erg/cm3").value))
            if not seed["isotropic"]:
                t += (float(Quantity(seed["theta"]).to("rad").value),)
            return t
        E = np.atleast_1d(seed["energy"].to("eV").value)
        if E.size == 1:
            return ("mono", float(E[0]),
                    float(np.atleast_1d(seed["photon_density"].to("erg/cm3").value)[0]))
        return ("array", tuple(E.tolist()),
                tuple(seed["photon_density"].to("1/(eV cm3)").value.tolist()))

    def _terms(self, E_eV):
        """One table for all walker-independent seeds + a fused launch for each
        per-walker tabulated seed (SSC); rows are kept in seed order so that
        ``self.specic`` and the seed sum follow radiative.py:706-710."""
        W = self.particle_distribution.batch
        g = self._grid()
        N_E = E_eV.size
        names = list(self.seed_photon_fields.keys())
        shared, batched = [], []
        for k, name in enumerate(names):
            sd = self.seed_photon_fields[name]
            if sd["type"] == "array" and sd["photon_density"].ndim == 2:
                batched.append(k)
            else:
                shared.append(k)
        if batched:
            Wb = self.seed_photon_fields[names[batched[0]]]["photon_density"].shape[0]
            if W == 1 and Wb > 1:
                W = Wb
        pr = self._prep(g, W, need_raw=(bool(batched) and eng.EXACT) or None)
        S = len(names)
        out = eng.empty(W, S * N_E)
        if shared:
            tb = eng.ic_table(g, E_eV, tuple(self._seed_tuple(self.seed_photon_fields[names[k]])
                                             for k in shared))
            if len(shared) == S:
                eng.contract(tb, pr, out=out)
            else:
                tmp = eng.contract(tb, pr)
                for i, k in enumerate(shared):
                    out[:, k * N_E:(k + 1) * N_E] = tmp[:, i * N_E:(i + 1) * N_E]
        for k in batched:
            sd = self.seed_photon_fields[names[k]]
            phn = np.ascontiguousarray(sd["photon_density"].to("1/(eV cm3)").value) * eng.mec2_eV
            seed_E = sd["energy"].to("eV").value
            if eng.EXACT:  # reference operation order, one serial inner trapezoid per node
                eng.ic_seed_spectrum(g, pr, E_eV, seed_E, eng.to_dev(phn), True, out, k * N_E)
                out[:, k * N_E:(k + 1) * N_E] /= eng.to_dev(E_eV)
                continue
            # hoisted: walker-independent f_AA81 table, lean inner contraction, outer trapezoid
            tb = eng.ssc_table(g, E_eV, seed_E)
            sxn, sds = eng.empty(W, tb.spitch), eng.empty(W, tb.spitch)
            eng.ssc_seed(tb, [(eng.to_dev(phn), 0, 1.0)], W, sxn, sds)
            inner = eng.empty(W, tb.Rp)
            eng.ssc_inner(tb, sxn, sds, W, inner)
            eng.ssc_outer(tb, inner, pr, out, k * N_E)
        self._specic_dev = (out, N_E, names)
        terms = [(out, k * N_E, k == S - 1, 1.0, None) for k in range(S)]
        return terms, W

    def _spectrum(self, photon_energy):
        res = super()._spectrum(photon_energy)
        out, N_E, names = self._specic_dev
        E = res[1]
        host = out.cpu().numpy()
        self.specic = [Quantity(self._shape_out(host[:, k * N_E:(k + 1) * N_E], E), "1/(s eV)")
                       for k in range(len(names))]
        return res

    def flux(self, photon_energy, distance=1 * u.kpc, seed=None):
        """radiative.py:712-758 (incl. per-seed access by name or index)."""
        model = super().flux(photon_energy, distance=distance)
        if seed is not None and self._symbolic():
            from .fused import TraceError
            raise TraceError("per-seed IC flux is not traced")
        if seed is not None:
            if not isinstance(seed, int):
                if seed not in self.seed_photon_fields:
                    raise ValueError("Provided seed photon field name is not in the definition "
                                     "of the InverseCompton instance")
                seed = list(self.seed_photon_fields.keys()).index(seed)
            elif seed > len(self.seed_photon_fields):
                raise ValueError("Provided seed photon field number is larger than the number "
                                 "of seed photon fields defined in the InverseCompton instance")
            if _nonzero(distance):
                distance = validate_scalar("distance", distance, physical_type="length")
                dfac = 4 * np.pi * distance.to("cm").value ** 2
                model = Quantity(self.specic[seed].value / dfac, "1/(s cm2 eV)")
            else:
                model = self.specic[seed]
        return model

    def sed(self, photon_energy, distance=1 * u.kpc, seed=None):
        """radiative.py:760-791."""
        sed = super().sed(photon_energy, distance=distance)
        if seed is not None:
            out_unit = "erg/(cm2 s)" if _nonzero(distance) else "erg/s"
            photon_energy = _validate_ene(photon_energy)
            sed = (self.flux(photon_energy, distance=distance, seed=seed)
                   * photon_energy**2.0).to(out_unit)
        return sed


class Bremsstrahlung(BaseElectron):
    """Non-thermal e-e and e-p bremsstrahlung, Baring+99 (radiative.py:794-989)."""
    _walker_scalars = ("n0",)

    def __init__(self, particle_distribution, n0=1 / u.cm**3, **kwargs):
        super().__init__(particle_distribution)
        self.n0 = n0
        self.Eemin = 100 * u.MeV
        self.Eemax = 1e9 * mec2
        self.nEed = 300
        # compute ee and ep weights from H and He abundances in ISM assuming ionized medium
        Y = np.array([1.0, 9.59e-2])
        Z = np.array([1, 2])
        N = np.sum(Y)
        X = Y / N
        self.weight_ee = np.sum(Z * X)
        self.weight_ep = np.sum(Z**2 * X)
        self.param_names += ["n0", "weight_ee", "weight_ep"]
        self.__dict__.update(**kwargs)

    def _terms(self, E_eV):
        W = self._batch()
        g = self._grid()
        N_E = E_eV.size
        pr = self._prep(g, W)
        tb = eng.brems_table(g, E_eV)
        out = eng.contract(tb, pr)
        n0 = np.broadcast_to(np.atleast_1d(Quantity(self.n0).to("1/cm3").value).astype(float),
                             (W,))
        # spec = n0 * (w_ee * emiss_ee + w_ep * emiss_ep); zero weights skip the term
        # (radiative.py:943-971)
        terms = []
        if self.weight_ee != 0.0:
            terms.append((out, 0, False, 1.0, eng.to_dev(n0 * self.weight_ee)))
        if self.weight_ep != 0.0:
            terms.append((out, N_E, False, 1.0, eng.to_dev(n0 * self.weight_ep)))
        if not terms:
            terms.append((eng.zeros(W, N_E), 0, False, 1.0, None))
        terms[-1] = terms[-1][:2] + (True,) + terms[-1][3:]
        return terms, W


class BaseProton(BaseRadiative):
    """Proton grid, Wp (radiative.py:992-1096)."""

    def __init__(self, particle_distribution):
        super().__init__(particle_distribution)
        self.param_names = ["Epmin", "Epmax", "nEpd"]

    def _grid(self, Epmin=None, Epmax=None):
        Epmin = self.Epmin if Epmin is None else Epmin
        Epmax = self.Epmax if Epmax is None else Epmax
        return eng.proton_grid(Quantity(Epmin).to("GeV").value, Quantity(Epmax).to("GeV").value,
                               self.nEpd)

    @property
    def _Ep(self):
        return self._grid().x.copy()

    @property
    def _J(self):
        g = self._grid()
        pr = self._prep(g, need_raw=True)
        n = pr.nraw[:, : g.N].cpu().numpy()
        return n if self._is_batched() else n[0]

    def _Wp(self, grid):
        if self._symbolic():
            from .fused import SymBlob
            return SymBlob("W", comp=self, grid=grid)
        Wp = self._particle_energy(grid)
        return Quantity(Wp if self._is_batched() else float(Wp[0]), u.erg)

    @property
    def Wp(self):
        """Total energy in protons."""
        return self._Wp(self._grid())

    def compute_Wp(self, Epmin=None, Epmax=None):
        """radiative.py:1023-1055."""
        if Epmin is None and Epmax is None:
            return self.Wp
        return self._Wp(self._grid(Epmin, Epmax))

    def set_Wp(self, Wp, Epmin=None, Epmax=None, amplitude_name=None):
        """radiative.py:1057-1096."""
        Wp = validate_scalar("Wp", Wp, physical_type="energy")
        oldWp = self.compute_Wp(Epmin=Epmin, Epmax=Epmax)
        ratio = (Wp / oldWp).decompose().value
        name = "amplitude" if amplitude_name is None else amplitude_name
        try:
            setattr(self.particle_distribution, name,
                    getattr(self.particle_distribution, name) * ratio)
        except AttributeError:
            log.error("The particle distribution does not have an attribute called %s to "
                      "modify its normalization", name)


class PionDecay(BaseProton):
    """Pion-decay gamma rays, Kafexhiu+14 (radiative.py:1099-1536)."""
    _walker_scalars = ("nh",)
    _m_p = eng.mpc2_GeV
    _Tth = eng.T_TH
    _LUT_MODELS = (("Pythia8", True),)  # the only table the reference ships

    def __init__(self, particle_distribution, nh=1.0 / u.cm**3, nuclear_enhancement=True,
                 **kwargs):
        super().__init__(particle_distribution)
        self.nh = validate_scalar("nh", nh, physical_type="number density")
        self.nuclear_enhancement = nuclear_enhancement
        self.useLUT = True
        self.hiEmodel = "Pythia8"
        self.Epmin = (self._m_p + self._Tth + 1e-4) * u.GeV
        self.Epmax = 10 * u.PeV
        self.nEpd = 100
        self.param_names += ["nh", "nuclear_enhancement", "useLUT", "hiEmodel"]
        self.__dict__.update(**kwargs)

    def _terms(self, E_eV):
        W = self._batch()
        useLUT = bool(self.useLUT)
        if useLUT and (self.hiEmodel, bool(self.nuclear_enhancement)) not in self._LUT_MODELS:
            # radiative.py:1484-1493: missing table -> analytic parametrisation
            log.warning("LUT for %s (nuclear_enhancement=%s) not found, reverting to "
                        "useLUT = False", self.hiEmodel, self.nuclear_enhancement)
            useLUT = False
            self.useLUT = False
        g = self._grid()
        pr = self._prep(g, W)
        tb = eng.pp_table(g, E_eV, useLUT, self.hiEmodel, self.nuclear_enhancement)
        out = eng.contract(tb, pr)
        nh = np.broadcast_to(np.atleast_1d(Quantity(self.nh).to("1/cm3").value).astype(float),
                             (W,))
        return [(out, 0, True, 1.0, eng.to_dev(nh))], W


class PionDecayKelner06(BaseRadiative):
    """Pion-decay gamma rays after Kelner, Aharonian & Bugayov 2006 (radiative.py:1543-1767):
    the full calculation (Eq. 71) at photon energies >= ``Etrans``, the delta-functional
    approximation (Eq. 78) below, matched at ``Etrans`` through ``nhat``.

    The reference integrates with adaptive QUADPACK at epsrel = 1e-3; here both integrals are
    log-log trapezoids over per-row proton-energy grids of 100 nodes per decade evaluated by
    nb_kelner_rows -- they agree with the reference to its own quadrature accuracy (a few
    1e-4), which is also all the reference's answers are good for."""
    _walker_scalars = ("nh",)
    param_names = ["nh", "Etrans"]

    def __init__(self, particle_distribution, nh=1.0 / u.cm**3, Etrans=0.1 * u.TeV, **kwargs):
        super().__init__(particle_distribution)
        if isinstance(particle_distribution, TableModel):
            raise TypeError("PionDecayKelner06 needs a parametrised particle distribution")
        self.nh = validate_scalar("nh", nh, physical_type="number density")
        self.Etrans = validate_scalar("Etrans", Etrans, domain="positive",
                                      physical_type="energy")
        self.__dict__.update(**kwargs)

    def _symbolic(self):
        if super()._symbolic():
            from .fused import TraceError
            raise TraceError("PionDecayKelner06 is not traced")
        return False

    def _terms(self, E_eV):
        kind, par_d, Wp = self._pd_device()
        W = self._batch()
        if Wp != W:
            par_d = par_d.expand(W, par_d.shape[1]).contiguous()
        Etr = Quantity(self.Etrans).to("TeV").value
        Eg = E_eV * 1e-12
        hi = Eg >= Etr
        mixed = bool(np.any(hi) and np.any(~hi))
        # two more rows at Etrans: the full and the delta-functional value that fix nhat
        Eg_all = np.concatenate([Eg, [Etr, Etr]])
        hi_all = np.concatenate([hi, [True, False]])
        Ep, Kk, _ = eng.kelner_table(Eg_all, hi_all)
        rows = eng.kelner_rows(kind, par_d, W, Ep, Kk)  # [W][N_E + 2], 1/(s TeV), nhat = 1
        N_E = E_eV.size
        spec = rows[:, :N_E].clone()
        if mixed:  # radiative.py:1748-1753
            nhat = rows[:, N_E] / rows[:, N_E + 1]
            lo = eng.to_dev((~hi).astype(float))
            spec = spec * (1.0 + lo[None, :] * (nhat[:, None] - 1.0))
        spec = spec * 1e-12  # 1/(s TeV) -> 1/(s eV)
        nh = np.broadcast_to(np.atleast_1d(Quantity(self.nh).to("1/cm3").value).astype(float),
                             (W,))
        return [(spec.contiguous(), 0, True, 1.0, eng.to_dev(nh))], W

    @property
    def Wp(self):
        """Total energy in protons above the 1.22 GeV threshold (radiative.py:1719-1729)."""
        kind, par_d, W = self._pd_device()
        x = np.logspace(np.log10(1.22e-3), 7, 1001)  # TeV
        n = eng.pdist_eval(kind, par_d, W, x * 1e12).cpu().numpy() * 1e12  # 1/TeV
        Wp = eng.trapz_loglog(x[None, :] * n, x) * (1e12 * eng.eV_erg)
        return Quantity(Wp if self._is_batched() else float(Wp[0]), u.erg)
