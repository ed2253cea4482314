# -*- coding: utf-8 -*-
"""ctypes binding of the C ABI declared in include/naima_b200.h.

There is no CPU fallback: importing the compute layer without the built
library, or calling it without a CUDA device, raises.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libnaima_b200.so")

c_int, c_dbl, c_ll, vp = ctypes.c_int, ctypes.c_double, ctypes.c_longlong, ctypes.c_void_p

NB_PD_MAXPAR = 8
NB_MAX_TERMS = 16
NB_MAX_PAR = 16
PD_KIND = {"PowerLaw": 0, "ExponentialCutoffPowerLaw": 1, "BrokenPowerLaw": 2,
           "ExponentialCutoffBrokenPowerLaw": 3, "LogParabola": 4}
PP_MODEL = {"Geant4": 0, "Pythia8": 1, "SIBYLL": 2, "QGSJET": 3}


class nb_term(ctypes.Structure):
    _fields_ = [("src", vp), ("wscale", vp), ("ld", c_int), ("off", c_int),
                ("group_end", c_int), ("div", c_dbl)]


class nb_parmap(ctypes.Structure):
    # out[w][k] = f(pars[w][src]) * scale (or the constant `scale` when src < 0)
    _fields_ = [("src", c_int), ("fn", c_int), ("scale", c_dbl), ("dst_off", c_ll),
                ("dst_stride", c_int)]


class nb_prior(ctypes.Structure):
    _fields_ = [("par", c_int), ("kind", c_int), ("a", c_dbl), ("b", c_dbl)]


class nb_prep_job(ctypes.Structure):
    _fields_ = [("kind", c_int), ("N", c_int), ("pd_off", c_ll), ("x", vp), ("invdlx", vp),
                ("e_mul1", c_dbl), ("e_mul2", c_dbl), ("n_scale", c_dbl), ("xn", vp),
                ("ds1", vp), ("nraw", vp), ("wpitch", c_int), ("pad_", c_int),
                ("x_to_energy", c_dbl), ("energy_out", vp), ("energy_stride", c_ll)]


class nb_stretch(ctypes.Structure):
    _fields_ = [("coords", vp), ("lp", vp), ("blobs", vp), ("nb", c_int), ("W", c_int),
                ("P", c_int), ("Ns", c_int), ("split", c_int), ("i0", c_int),
                ("pars_ld", c_int), ("pad_", c_int), ("step", vp), ("sync", vp),
                ("s_idx", vp), ("c_idx", vp), ("zz", vp), ("lnu", vp), ("n_accepted", vp),
                ("chain", vp), ("chain_lp", vp), ("chain_blobs", vp), ("wait_flags", vp),
                ("wait_gen", vp), ("wait_world", c_int), ("pad2_", c_int),
                ("timeline", vp)]


NB_MAX_PEERS = 16
NB_TIMELINE_CAP = 8192
NB_TIMELINE_COLS = 16


class nb_peers(ctypes.Structure):
    _fields_ = [("world", c_int), ("rank", c_int), ("i0", c_int), ("ld", c_int),
                ("pack", vp * NB_MAX_PEERS), ("flags", vp * NB_MAX_PEERS), ("gen", vp),
                ("ticket", vp), ("mc_pack", vp), ("arena_local", vp * 2), ("arena_mc", vp * 2),
                ("arena_peer", (vp * NB_MAX_PEERS) * 2), ("arena_bytes", ctypes.c_ulonglong * 2),
                ("mc_flags", vp)]


class nb_walker_src(ctypes.Structure):
    _fields_ = [("pars", vp), ("P", c_int), ("n_map", c_int),
                ("map_host", ctypes.POINTER(nb_parmap)), ("mv_host", ctypes.POINTER(nb_stretch))]


class nb_pd_desc(ctypes.Structure):
    _fields_ = [("kind", c_int), ("pad_", c_int), ("pd_off", c_ll), ("e_mul1", c_dbl),
                ("e_mul2", c_dbl), ("n_scale", c_dbl), ("lnx", vp), ("invdlx", vp)]


class nb_ssc_src(ctypes.Structure):
    _fields_ = [("src", vp), ("ld", c_int), ("off", c_int), ("fac", c_dbl)]


# name -> (argtypes); every function returns int
PROTOTYPES = {
    "nb_trapz_loglog": [vp, c_int, c_int, c_int, vp, c_int, vp, vp, vp],
    "nb_pdist_eval": [c_int, vp, c_int, vp, c_int, vp, vp],
    "nb_pdist_eval_ld": [c_int, vp, c_int, vp, c_int, vp, c_int, vp],
    "nb_pd_prep": [c_int, vp, c_int, vp, c_int, c_dbl, c_dbl, c_dbl, vp, vp, vp, c_int, vp],
    "nb_pd_prep_ex": [c_int, vp, c_int, vp, c_int, c_dbl, c_dbl, c_dbl, vp, vp, vp, vp, c_int, vp],
    "nb_particle_energy": [c_int, vp, c_int, vp, c_int, c_dbl, c_dbl, c_dbl, c_dbl, vp, vp],
    "nb_ic_planck_table": [vp, c_int, vp, c_int, vp, vp, c_int, vp, c_int, c_int, vp],
    "nb_ic_seed_table": [vp, c_int, vp, c_int, vp, vp, c_int, vp, c_int, c_int, vp],
    "nb_ic_seed_spectrum": [vp, c_int, vp, c_int, vp, c_int, vp, vp, c_int, c_int, c_int, vp,
                            c_int, c_int, vp],
    "nb_brems_table": [vp, c_int, vp, c_int, vp, c_int, c_int, vp],
    "nb_pp_analytic_table": [c_int, c_int, vp, c_int, vp, c_int, vp, c_int, c_int, vp],
    "nb_pp_lut_table": [vp, c_int, vp, c_int, vp, vp, c_int, vp, c_int, vp, c_int, c_int, vp],
    "nb_table_finalize": [vp, c_int, c_int, c_int, vp, vp, vp],
    "nb_contract": [vp, vp, c_int, c_int, c_int, c_ll, vp, vp, c_int, c_int, vp, vp, vp, vp,
                    c_int, vp],
    "nb_synchrotron": [vp, c_int, vp, vp, vp, vp, c_int, vp, vp, vp, c_int, vp, vp, c_int, vp,
                       c_int, vp],
    "nb_table_scan": [vp, c_int, c_int, c_int, vp, vp, vp],
    "nb_contract_ex": [vp, vp, c_int, c_int, c_int, vp, vp, vp, c_int, c_int, vp, vp, vp, vp,
                       c_int, vp],
    "nb_ssc_table": [vp, c_int, vp, c_int, vp, vp, c_int, vp, vp, vp, c_ll, vp],
    "nb_ssc_seed": [ctypes.POINTER(nb_ssc_src), c_int, c_int, c_int, vp, vp, vp, c_int, vp],
    "nb_ssc_inner": [vp, vp, vp, c_ll, c_int, vp, vp, c_int, c_int, vp, vp, vp],
    "nb_ssc_outer": [vp, c_ll, c_int, c_int, c_int, vp, vp, c_int, vp, vp, vp, vp, c_int, c_int,
                     vp],
    "nb_combine_lnprob": [ctypes.POINTER(nb_term), c_int, c_int, c_int, vp, vp, vp, vp, vp, vp,
                          vp, vp, c_int, vp, vp],
    "nb_walker_prep": [vp, c_int, c_int, ctypes.POINTER(nb_parmap), c_int, vp,
                       ctypes.POINTER(nb_prior), c_int, vp, ctypes.POINTER(nb_prep_job), c_int, vp],
    "nb_walker_prep_move": [ctypes.POINTER(nb_stretch), vp, c_int, c_int,
                            ctypes.POINTER(nb_parmap), c_int, vp, ctypes.POINTER(nb_prior),
                            c_int, vp, ctypes.POINTER(nb_prep_job), c_int, vp],
    "nb_combine_lnprob_update": [ctypes.POINTER(nb_stretch), vp, ctypes.POINTER(nb_term), c_int,
                                 c_int, c_int, vp, vp, vp, vp, vp, vp, vp, vp, c_int, vp, vp],
    "nb_combine_lnprob_ld": [ctypes.POINTER(nb_term), c_int, c_int, c_int, vp, vp, vp, vp, vp,
                             vp, vp, vp, c_int, vp, c_int, vp],
    "nb_stretch_update_packed": [ctypes.POINTER(nb_stretch), vp, c_int, vp],
    "nb_synchrotron_fused": [ctypes.POINTER(nb_walker_src), ctypes.POINTER(nb_pd_desc), c_int, vp,
                             c_int, vp, vp, vp, c_int, vp, vp, c_int, vp, c_int, vp],
    "nb_combine_lnprob_update_push": [ctypes.POINTER(nb_stretch), ctypes.POINTER(nb_peers), vp,
                                      ctypes.POINTER(nb_term), c_int, c_int, c_int, vp, vp, vp,
                                      vp, vp, vp, vp, vp, c_int, vp, vp],
    "nb_peer_wait": [ctypes.POINTER(nb_stretch), vp],
    "nb_fp64_peak_probe": [vp, c_int, c_int, c_int, vp],
    "nb_fallback_counts": [ctypes.POINTER(ctypes.c_ulonglong), c_int],
    "nb_launch_carveout": [c_int],
    "nb_timeline_reset": [],
    "nb_kelner_table": [vp, vp, c_int, c_int, c_dbl, vp, vp, vp],
    "nb_kelner_rows": [c_int, vp, c_int, vp, vp, c_int, c_int, vp, vp],
}

_lib = None


def lib():
    """The loaded shared library (raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "naima_b200: %s is missing -- build it with `python -m naima_b200.build` "
                "(there is no CPU fallback)" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, args in PROTOTYPES.items():
            fn = getattr(L, name)  # AttributeError if the ABI and the binding disagree
            fn.argtypes = args
            fn.restype = c_int
        L.nb_version.restype = c_int
        L.nb_strerror.restype = ctypes.c_char_p
        L.nb_strerror.argtypes = [c_int]
        _lib = L
    return _lib


HOST_LIB_PATH = os.path.join(HERE, "libnaima_b200_host.so")
_host = False


def host_lib():
    """The host-side helper library (random draws of the device-resident sampler), or None
    when it has not been built -- callers then use the equivalent NumPy code."""
    global _host
    if _host is False:
        _host = None
        if os.path.exists(HOST_LIB_PATH):
            try:
                H = ctypes.CDLL(HOST_LIB_PATH)
                H.nb_host_draw_steps.restype = c_int
                H.nb_host_draw_steps.argtypes = [vp, ctypes.POINTER(c_int), c_int, c_int, c_dbl,
                                                 vp, vp, vp, vp]
                _host = H
            except (OSError, AttributeError):
                _host = None
    return _host


class NaimaB200Error(RuntimeError):
    pass


def check(code, what=""):
    if code != 0:
        msg = lib().nb_strerror(code).decode()
        if code == -1:
            raise ValueError("naima_b200 %s: %s" % (what, msg))
        raise NaimaB200Error("naima_b200 %s: %s (code %d)" % (what, msg, code))
