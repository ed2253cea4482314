# -*- coding: utf-8 -*-
"""Minimal physical-units layer with the slice of the ``astropy.units`` API that
the naima call conventions use (``u.eV``, ``10 * u.TeV``, ``q.to("1/(s cm2 eV)")``,
``q.value``, ``q.unit.physical_type`` ...).

The reference strips and re-attaches astropy units around every numeric step
(radiative.py:43-58,102-134; core.py:64-71; utils.py:219-282).  astropy is not
available in this image, so the host layer of naima_b200 carries this small
stand-in; all device work is done on plain float64 in the fixed units of
include/naima_b200.h.  Values may be scalars or numpy arrays (a leading walker
axis is how batched parameters travel through user model functions).

A unit is ``scale * prod(atom ** power)`` over named atoms; conversion between
two units first cancels common atoms so that prefixed units convert by exact
powers of ten (``100 GeV -> 1e11 eV`` exactly), then resolves the remaining
atoms to cgs.
"""
import math
import re
from fractions import Fraction

import numpy as np

__all__ = ["Unit", "Quantity", "UnitsError", "UnitConversionError", "def_physical_type"]


class UnitsError(ValueError):
    pass


class UnitConversionError(UnitsError):
    pass


# atom -> (cgs factor, dims) with dims over (cm, g, s, K, rad, G)
_D = {
    "cm": (1, 0, 0, 0, 0, 0),
    "g": (0, 1, 0, 0, 0, 0),
    "s": (0, 0, 1, 0, 0, 0),
    "K": (0, 0, 0, 1, 0, 0),
    "rad": (0, 0, 0, 0, 1, 0),
    "G": (0, 0, 0, 0, 0, 1),
}
_ERG = (2, 1, -2, 0, 0, 0)
_ATOMS = {
    "cm": (1.0, _D["cm"]),
    "g": (1.0, _D["g"]),
    "s": (1.0, _D["s"]),
    "K": (1.0, _D["K"]),
    "rad": (1.0, _D["rad"]),
    "G": (1.0, _D["G"]),
    "erg": (1.0, _ERG),
    "eV": (1.602176634e-12, _ERG),
    "J": (1e7, _ERG),
    "m": (100.0, _D["cm"]),
    "pc": (3.0856775814913673e18, _D["cm"]),
    "AA": (1e-8, _D["cm"]),
    "deg": (math.pi / 180.0, _D["rad"]),
    "sr": (1.0, (0, 0, 0, 0, 2, 0)),
    "T": (1e4, _D["G"]),
    "yr": (31557600.0, _D["s"]),
    "mec2": (9.1093837015e-28 * 29979245800.0**2, _ERG),
    "Hz": (1.0, (0, 0, -1, 0, 0, 0)),
    "W": (1e7, (2, 1, -3, 0, 0, 0)),
}
_PREFIX = {"f": 1e-15, "p": 1e-12, "n": 1e-9, "u": 1e-6, "m": 1e-3, "c": 1e-2, "k": 1e3,
           "M": 1e6, "G": 1e9, "T": 1e12, "P": 1e15, "E": 1e18}
_PREFIXABLE = ("eV", "G", "pc", "m", "s", "g", "K", "J", "Hz", "yr", "erg", "W")

_PHYS = {}


def def_physical_type(unit, name):
    """Register ``name`` for the dimensions of ``unit`` (core.py:22-28)."""
    _PHYS[Unit(unit)._dims()] = name


class Unit:
    __array_priority__ = 20000
    __array_ufunc__ = None

    def __new__(cls, spec=None, _scale=None, _atoms=None):
        if isinstance(spec, Unit):
            return spec
        self = object.__new__(cls)
        if _atoms is not None:
            self.scale = float(_scale)
            self.atoms = {k: v for k, v in _atoms.items() if v != 0}
            return self
        if isinstance(spec, Quantity):
            un = spec.unit
            self.scale = un.scale * float(spec.value)
            self.atoms = dict(un.atoms)
            return self
        if spec is None or spec == "":
            self.scale, self.atoms = 1.0, {}
            return self
        if isinstance(spec, (int, float)):
            self.scale, self.atoms = float(spec), {}
            return self
        if hasattr(spec, "to_string"):  # foreign (astropy) unit
            spec = spec.to_string()
        parsed = _parse(str(spec))
        self.scale, self.atoms = parsed.scale, parsed.atoms
        return self

    # -- algebra ---------------------------------------------------------------
    def _combine(self, other, sign):
        atoms = dict(self.atoms)
        for k, v in other.atoms.items():
            atoms[k] = atoms.get(k, 0) + sign * v
        scale = self.scale * other.scale if sign > 0 else self.scale / other.scale
        return Unit(_scale=scale, _atoms=atoms)

    def __mul__(self, other):
        if isinstance(other, Unit):
            return self._combine(other, 1)
        if isinstance(other, Quantity):
            return Quantity(other.value, self._combine(other.unit, 1))
        if isinstance(other, str):
            return self._combine(Unit(other), 1)
        return Quantity(other, self)

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return self._combine(other, -1)
        if isinstance(other, Quantity):
            return Quantity(1.0 / other.value, self._combine(other.unit, -1))
        if isinstance(other, str):
            return self._combine(Unit(other), -1)
        return Quantity(1.0 / np.asarray(other, dtype=float), self)

    def __rtruediv__(self, other):
        inv = self**-1
        if isinstance(other, Quantity):
            return Quantity(other.value, other.unit._combine(inv, 1))
        return Quantity(other, inv)

    def __pow__(self, p):
        p = Fraction(p).limit_denominator(12) if not isinstance(p, int) else p
        return Unit(_scale=self.scale ** float(p), _atoms={k: v * p for k, v in self.atoms.items()})

    # -- dimensional analysis --------------------------------------------------
    def _dims(self):
        d = [Fraction(0)] * 6
        for k, v in self.atoms.items():
            ad = _ATOMS[k][1]
            for i in range(6):
                d[i] += ad[i] * v
        return tuple(d)

    def _cgs_factor(self):
        f = self.scale
        for k, v in self.atoms.items():
            a = _ATOMS[k][0]
            if a != 1.0:
                f *= a ** float(v)
        return f

    def _factor_to(self, other):
        """Multiplicative factor converting values in self to values in other."""
        other = Unit(other)
        if self.atoms == other.atoms:
            return self.scale / other.scale
        if self._dims() != other._dims():
            raise UnitConversionError("'%s' and '%s' are not convertible" % (self, other))
        # cancel common atoms, resolve the rest to cgs
        num, den = self.scale, other.scale
        keys = set(self.atoms) | set(other.atoms)
        for k in sorted(keys):
            dv = self.atoms.get(k, 0) - other.atoms.get(k, 0)
            if dv != 0 and _ATOMS[k][0] != 1.0:
                if dv > 0:
                    num *= _ATOMS[k][0] ** float(dv)
                else:
                    den *= _ATOMS[k][0] ** float(-dv)
        return num / den

    def to(self, other, value=1.0):
        return value * self._factor_to(other)

    def is_equivalent(self, other):
        return self._dims() == Unit(other)._dims()

    @property
    def physical_type(self):
        d = self._dims()
        if all(x == 0 for x in d):
            return "dimensionless"
        return _PHYS.get(d, "unknown")

    def decompose(self):
        return Unit(_scale=self._cgs_factor(), _atoms=_dims_to_atoms(self._dims()))

    @property
    def cgs(self):
        return self.decompose()

    def __eq__(self, other):
        try:
            other = Unit(other)
        except Exception:
            return False
        if self._dims() != other._dims():
            return False
        f = self._factor_to(other)
        return abs(f - 1.0) < 1e-14

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash(self._dims())

    def to_string(self):
        num = [(k, v) for k, v in self.atoms.items() if v > 0]
        den = [(k, -v) for k, v in self.atoms.items() if v < 0]

        def fmt(items):
            return " ".join(k if v == 1 else "%s%s" % (k, v) for k, v in items)

        s = fmt(num) if num else ("1" if den else "")
        if den:
            s += " / (%s)" % fmt(den) if len(den) > 1 else " / %s" % fmt(den)
        if self.scale != 1.0:
            s = ("%r %s" % (self.scale, s)).strip()
        return s

    __str__ = to_string

    def __repr__(self):
        return 'Unit("%s")' % self.to_string()


def _dims_to_atoms(d):
    names = ("cm", "g", "s", "K", "rad", "G")
    return {n: v for n, v in zip(names, d) if v != 0}


def _atom(name):
    """Resolve ``name`` (optionally SI-prefixed) to a Unit; trailing digits are an exponent."""
    if name in _ATOMS:
        return Unit(_scale=1.0, _atoms={name: 1})
    m = re.match(r"^([A-Za-z_]+?)(-?\d+)?$", name)
    if not m:
        raise ValueError("cannot parse unit '%s'" % name)
    base, power = m.group(1), int(m.group(2)) if m.group(2) else 1
    if base in _ATOMS:
        un = Unit(_scale=1.0, _atoms={base: 1})
    elif base[0] in _PREFIX and base[1:] in _PREFIXABLE:
        un = Unit(_scale=_PREFIX[base[0]], _atoms={base[1:]: 1})
    elif base == "kpc":
        un = Unit(_scale=1e3, _atoms={"pc": 1})
    else:
        raise ValueError("unknown unit '%s'" % base)
    return un**power if power != 1 else un


def _parse(s):
    s = s.strip()
    toks = re.findall(r"\*\*|\^|[()*/]|[A-Za-z_][A-Za-z_]*-?\d*|\d+(?:\.\d+)?(?:[eE][-+]?\d+)?", s)
    pos = [0]

    def peek():
        return toks[pos[0]] if pos[0] < len(toks) else None

    def take():
        t = toks[pos[0]]
        pos[0] += 1
        return t

    def factor():
        t = take()
        if t == "(":
            un = expr()
            if take() != ")":
                raise ValueError("unbalanced parentheses in unit '%s'" % s)
        elif re.match(r"^\d", t):
            un = Unit(_scale=float(t), _atoms={})
        else:
            un = _atom(t)
        if peek() in ("**", "^"):
            take()
            sign = 1
            p = take()
            if p == "(":
                p = take()
                take()
            un = un ** (sign * (int(p) if re.match(r"^-?\d+$", p) else float(p)))
        return un

    def expr():
        un = factor()
        while peek() is not None and peek() != ")":
            t = peek()
            if t == "/":
                take()
                un = un / factor()
            elif t == "*":
                take()
                un = un * factor()
            else:
                un = un * factor()
        return un

    if not toks:
        return Unit(_scale=1.0, _atoms={})
    out = expr()
    if pos[0] != len(toks):
        raise ValueError("cannot parse unit '%s'" % s)
    return out


def _is_quantity(x):
    return isinstance(x, Quantity) or (hasattr(x, "unit") and hasattr(x, "value")
                                       and not isinstance(x, Unit))


class Quantity:
    """A float64 scalar/array with a Unit.  numpy defers to this class for
    binary operators (``__array_ufunc__ = None``)."""

    __array_priority__ = 10000
    __array_ufunc__ = None

    def __init__(self, value, unit=None, copy=True):
        if isinstance(value, Quantity):
            if unit is None:
                self.value, self.unit = value.value, value.unit
            else:
                self.unit = Unit(unit)
                self.value = value.value * value.unit._factor_to(self.unit)
            return
        if _is_quantity(value):  # foreign quantity (astropy): go through strings
            q = Quantity(np.asarray(value.value, dtype=float), Unit(value.unit))
            self.value, self.unit = (q.to(unit).value, Unit(unit)) if unit is not None else (q.value, q.unit)
            return
        if isinstance(value, (list, tuple)) and len(value) and any(_is_quantity(v) for v in value):
            qs = [Quantity(v) for v in value]
            un = Unit(unit) if unit is not None else qs[0].unit
            self.value = np.array([q.to(un).value for q in qs], dtype=float)
            self.unit = un
            return
        if isinstance(value, str):
            m = re.match(r"^\s*([-+0-9.eE]+)\s*(.*)$", value)
            value, unit = float(m.group(1)), (m.group(2) if unit is None else unit)
        v = np.asarray(value, dtype=float)
        self.value = float(v) if v.ndim == 0 else v
        self.unit = Unit(unit)

    # -- conversion ------------------------------------------------------------
    def to(self, unit, equivalencies=None):
        unit = Unit(unit)
        f = self.unit._factor_to(unit)
        return Quantity(self.value * f if f != 1.0 else self.value, unit)

    def to_value(self, unit):
        return self.to(unit).value

    def decompose(self):
        un = self.unit.decompose()
        if not un.atoms:
            return Quantity(self.value * un.scale, Unit())
        return Quantity(self.value * un.scale, Unit(_scale=1.0, _atoms=un.atoms))

    @property
    def cgs(self):
        return self.decompose()

    @property
    def si(self):
        return self.decompose()

    # -- array protocol ----------------------------------------------------------
    @property
    def shape(self):
        return np.shape(self.value)

    @property
    def size(self):
        return np.size(self.value)

    @property
    def ndim(self):
        return np.ndim(self.value)

    @property
    def isscalar(self):
        return np.ndim(self.value) == 0

    @property
    def dtype(self):
        return np.asarray(self.value).dtype

    def __len__(self):
        return len(self.value)

    def __iter__(self):
        for v in self.value:
            yield Quantity(v, self.unit)

    def __getitem__(self, idx):
        return Quantity(np.asarray(self.value)[idx], self.unit)

    def __setitem__(self, idx, val):
        self.value[idx] = Quantity(val, self.unit).value if _is_quantity(val) else val

    def __float__(self):
        return float(self._dimensionless_value())

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.value, dtype=dtype)

    def _dimensionless_value(self):
        d = self.unit._dims()
        if any(x != 0 for x in d):
            raise TypeError("only dimensionless quantities can be converted to plain numbers")
        return self.value * self.unit._cgs_factor()

    def copy(self):
        return Quantity(np.copy(self.value), self.unit)

    def flatten(self):
        return Quantity(np.asarray(self.value).flatten(), self.unit)

    def reshape(self, *shape):
        return Quantity(np.asarray(self.value).reshape(*shape), self.unit)

    def squeeze(self, *a, **k):
        return Quantity(np.squeeze(self.value, *a, **k), self.unit)

    def sum(self, axis=None):
        return Quantity(np.sum(self.value, axis=axis), self.unit)

    def max(self, axis=None):
        return Quantity(np.max(self.value, axis=axis), self.unit)

    def min(self, axis=None):
        return Quantity(np.min(self.value, axis=axis), self.unit)

    def mean(self, axis=None):
        return Quantity(np.mean(self.value, axis=axis), self.unit)

    def argmax(self, axis=None):
        return np.argmax(self.value, axis=axis)

    def argsort(self, axis=-1):
        return np.argsort(self.value, axis=axis)

    @property
    def T(self):
        return Quantity(np.transpose(self.value), self.unit)

    # -- arithmetic ------------------------------------------------------------
    def _other_value(self, other):
        if _is_quantity(other):
            other = Quantity(other)
            return other.value * other.unit._factor_to(self.unit) if other.unit.atoms != self.unit.atoms or other.unit.scale != self.unit.scale else other.value
        if isinstance(other, Unit):
            return other._factor_to(self.unit)
        if not self.unit.atoms and self.unit.scale == 1.0:
            return other
        if not any(x != 0 for x in self.unit._dims()):
            return np.asarray(other, dtype=float) / self.unit._cgs_factor()
        if np.all(np.asarray(other) == 0):  # 0 is unit-agnostic (flux(..., distance=0))
            return other
        raise UnitConversionError("cannot combine '%s' with a dimensionless number" % self.unit)

    def __add__(self, other):
        return Quantity(self.value + self._other_value(other), self.unit)

    def __radd__(self, other):
        return Quantity(self._other_value(other) + self.value, self.unit)

    def __sub__(self, other):
        return Quantity(self.value - self._other_value(other), self.unit)

    def __rsub__(self, other):
        return Quantity(self._other_value(other) - self.value, self.unit)

    def __neg__(self):
        return Quantity(-self.value, self.unit)

    def __abs__(self):
        return Quantity(np.abs(self.value), self.unit)

    def __mul__(self, other):
        if isinstance(other, Unit):
            return Quantity(self.value, self.unit * other)
        if _is_quantity(other):
            other = Quantity(other)
            return Quantity(self.value * other.value, self.unit * other.unit)
        return Quantity(self.value * other, self.unit)

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return Quantity(self.value, self.unit / other)
        if _is_quantity(other):
            other = Quantity(other)
            return Quantity(self.value / other.value, self.unit / other.unit)
        return Quantity(self.value / other, self.unit)

    def __rtruediv__(self, other):
        return Quantity(other / self.value, self.unit**-1)

    def __itruediv__(self, other):
        r = self.__truediv__(other)
        self.value, self.unit = r.value, r.unit
        return self

    def __imul__(self, other):
        r = self.__mul__(other)
        self.value, self.unit = r.value, r.unit
        return self

    def __pow__(self, p):
        return Quantity(self.value**p, self.unit**p)

    def __rpow__(self, base):
        return base ** self._dimensionless_value()

    def _cmp(self, other, op):
        return op(self.value, self._other_value(other))

    def __lt__(self, o):
        return self._cmp(o, np.less)

    def __le__(self, o):
        return self._cmp(o, np.less_equal)

    def __gt__(self, o):
        return self._cmp(o, np.greater)

    def __ge__(self, o):
        return self._cmp(o, np.greater_equal)

    def __eq__(self, o):
        try:
            return self._cmp(o, np.equal)
        except UnitsError:
            return False

    def __ne__(self, o):
        try:
            return self._cmp(o, np.not_equal)
        except UnitsError:
            return True

    __hash__ = None

    def __bool__(self):
        return bool(np.all(self.value))

    def __repr__(self):
        return "<Quantity %r %s>" % (self.value, self.unit)

    __str__ = __repr__


# module-level unit atoms (astropy.units style)
def _export():
    g = globals()
    for name in _ATOMS:
        g[name] = Unit(_scale=1.0, _atoms={name: 1})
    for p in _PREFIX:
        for base in _PREFIXABLE:
            nm = p + base
            if nm not in g:
                g[nm] = Unit(_scale=_PREFIX[p], _atoms={base: 1})
    g["kpc"] = Unit(_scale=1e3, _atoms={"pc": 1})
    g["Mpc"] = Unit(_scale=1e6, _atoms={"pc": 1})
    g["dimensionless_unscaled"] = Unit()
    g["Gauss"] = g["G"]


_export()

# physical types used by the naima call conventions (core.py:22-28, validators)
for _u, _n in [
    ("erg", "energy"), ("cm", "length"), ("s", "time"), ("K", "temperature"), ("rad", "angle"),
    ("G", "magnetic flux density"), ("g", "mass"),
    ("erg/(cm2 s)", "flux"), ("1/(s cm2 erg)", "differential flux"), ("erg/s", "power"),
    ("1/(s erg)", "differential power"), ("1/erg", "differential energy"),
    ("1/cm3", "number density"), ("1/(erg cm3)", "differential number density"),
    ("erg/cm3", "pressure"), ("1/s", "frequency"), ("cm2", "area"), ("cm3", "volume"),
    ("1/(s cm2)", "particle flux"),
]:
    def_physical_type(_u, _n)
