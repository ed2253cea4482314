# -*- coding: utf-8 -*-
"""The handful of ``astropy.constants`` the reference's examples and radiative.py:36-40 use,
as Quantities of the units shim (CODATA 2018, as astropy >= 6.1 ships them)."""
from . import engine as _eng
from .units import Quantity

c = Quantity(_eng.c_cgs, "cm/s")
m_e = Quantity(_eng.m_e_g, "g")
sigma_sb = Quantity(_eng.sigma_sb_cgs, "erg/(cm2 s K4)")
hbar = Quantity(1.0545718176461565e-27, "erg s")
e_esu = 4.803204712570263e-10  # statcoulomb (no charge dimension in the shim)
alpha = 0.0072973525693
m_p_c2 = Quantity(_eng.mpc2_GeV, "GeV")

__all__ = ["c", "m_e", "sigma_sb", "hbar", "alpha", "m_p_c2"]
