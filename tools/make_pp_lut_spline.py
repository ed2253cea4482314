#!/usr/bin/env python
"""Derive the device-ready pion-decay cross-section table from the reference's
packaged lookup table.

The reference (radiative.py:1770-1797 ``LookupTable``) fits
``RectBivariateSpline(X, Y, 10**lut, kx=3, ky=3, s=0)`` at load time and
evaluates it with FITPACK ``bispev``.  The CUDA path evaluates the same
tensor-product cubic B-spline on the device, so what it needs is the knot
vectors and the coefficient matrix, not the raw table.  This script runs ONCE in
the build container (``/root/reference`` exists only there) and writes

    naima_b200/data/pp_kafexhiu14_pythia8_nucenh_bspline.npz
        tx[804], ty[1028]   knots (log10 Ep[GeV], log10 Egamma[GeV])
        c[800*1024]         B-spline coefficients of dsigma/dEgamma [cm2/GeV]

Only the Pythia8 + nuclear-enhancement table exists in the reference; every
other (hiEmodel, nuclear_enhancement) combination falls back to the analytic
parametrisation there (radiative.py:1484-1493) and here.
"""
import os

import numpy as np
from scipy.interpolate import RectBivariateSpline

REF = "/root/reference/src/naima/data/PionDecayKafexhiu14_LUT_NucEnh_Pythia8.npz"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "naima_b200", "data", "pp_kafexhiu14_pythia8_nucenh_bspline.npz")


def main():
    f = np.load(REF)
    spl = RectBivariateSpline(f["X"], f["Y"], 10 ** f["lut"], kx=3, ky=3, s=0)
    tx, ty = spl.get_knots()
    c = spl.get_coeffs()
    assert c.size == (tx.size - 4) * (ty.size - 4)
    np.savez_compressed(OUT, tx=tx, ty=ty, c=c)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
