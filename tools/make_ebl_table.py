#!/usr/bin/env python
"""Repack the reference's EBL optical-depth table (Dominguez et al. 2011; used by
``EblAbsorptionModel``, src/naima/models.py:472-552) for naima_b200.

The reference ships ``data/tau_dominguez11.npz`` as a structured array (an ``energy``
column in TeV and one column of optical depths per redshift 0.01 ... 3.99).  This script
runs ONCE in the build container (``/root/reference`` exists only there) and writes the
same numbers as two plain arrays:

    naima_b200/data/ebl_dominguez11.npz
        energy_TeV[500], tau[399][500]   (tau[k] belongs to redshift 0.01 * (k + 1))
"""
import os

import numpy as np

REF = "/root/reference/src/naima/data/tau_dominguez11.npz"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "naima_b200", "data", "ebl_dominguez11.npz")


def main():
    t = np.load(REF)["arr_0"]
    cols = [n for n in t.dtype.names if n != "energy"]
    assert cols == ["col%d" % k for k in range(2, 2 + len(cols))]
    tau = np.array([t[c] for c in cols], dtype=float)
    np.savez_compressed(OUT, energy_TeV=np.asarray(t["energy"], dtype=float), tau=tau)
    print("wrote", OUT, tau.shape, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
