"""The C-ABI library loads and exports every symbol include/naima_b200.h
declares, and the ctypes binding agrees with the header on every parameter
count (no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "naima_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|const char\*)\s+(nb_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        out[m.group(1)] = n
    return out


@pytest.fixture(scope="module")
def built_lib():
    from naima_b200.build import build_library

    return ctypes.CDLL(build_library())


def test_header_declares_the_path():
    fns = header_functions()
    for name in ("nb_trapz_loglog", "nb_pdist_eval", "nb_pd_prep", "nb_particle_energy",
                 "nb_ic_planck_table", "nb_ic_seed_table", "nb_ic_seed_spectrum",
                 "nb_brems_table", "nb_pp_analytic_table", "nb_pp_lut_table",
                 "nb_table_finalize", "nb_contract", "nb_synchrotron", "nb_combine_lnprob",
                 "nb_walker_prep", "nb_walker_prep_move", "nb_combine_lnprob_update",
                 "nb_synchrotron_fused", "nb_version", "nb_strerror"):
        assert name in fns, name


def test_library_exports_every_declared_symbol(built_lib):
    for name in header_functions():
        assert hasattr(built_lib, name), "missing symbol %s" % name


def test_ctypes_binding_matches_header(built_lib):
    from naima_b200 import _lib

    fns = header_functions()
    for name, argtypes in _lib.PROTOTYPES.items():
        assert name in fns, "binding for undeclared function %s" % name
        assert len(argtypes) == fns[name], (name, len(argtypes), fns[name])
    unbound = set(fns) - set(_lib.PROTOTYPES) - {"nb_version", "nb_strerror",
                                                  "nb_contract_smem_bytes"}
    assert not unbound, unbound
    L = _lib.lib()  # loads without a GPU and sets every prototype
    assert L.nb_version() >= 100
    assert L.nb_strerror(-1).decode() == "invalid argument"


def test_struct_layouts():
    from naima_b200 import _lib

    assert ctypes.sizeof(_lib.nb_term) == 40
    assert ctypes.sizeof(_lib.nb_parmap) == 32
    assert ctypes.sizeof(_lib.nb_prior) == 24


def test_product_has_no_cpu_path():
    """Compute entry points refuse to run without a CUDA device."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import naima_b200 as nb
    from naima_b200 import units as u
    from naima_b200.models import InverseCompton, PowerLaw

    ic = InverseCompton(PowerLaw(1e30 / u.eV, 1 * u.TeV, 2.1))
    with pytest.raises(RuntimeError, match="CUDA"):
        ic.flux([1, 10] * u.TeV)
    with pytest.raises(RuntimeError, match="CUDA"):
        nb.trapz_loglog([1.0, 2.0], [1.0, 2.0])


def test_ctypes_struct_layouts_match_the_header(tmp_path):
    """sizeof / offsetof of every struct of include/naima_b200.h, as gcc lays them out,
    against the ctypes mirrors in naima_b200/_lib.py."""
    import ctypes
    import subprocess

    from naima_b200 import _lib

    structs = {
        "nb_term": ["src", "wscale", "ld", "off", "group_end", "div"],
        "nb_parmap": ["src", "fn", "scale", "dst_off", "dst_stride"],
        "nb_prior": ["par", "kind", "a", "b"],
        "nb_prep_job": ["kind", "N", "pd_off", "x", "invdlx", "e_mul1", "xn", "wpitch",
                        "x_to_energy", "energy_out", "energy_stride"],
        "nb_stretch": ["coords", "nb", "split", "i0", "pars_ld", "step", "sync", "s_idx",
                       "n_accepted", "chain_blobs", "wait_flags", "wait_gen", "wait_world",
                       "timeline"],
        "nb_peers": ["world", "rank", "i0", "ld", "pack", "flags", "gen", "ticket", "mc_pack",
                     "arena_local", "arena_mc", "arena_peer", "arena_bytes", "mc_flags"],
        "nb_walker_src": ["pars", "P", "n_map", "map_host", "mv_host"],
        "nb_pd_desc": ["kind", "pd_off", "e_mul1", "n_scale", "lnx", "invdlx"],
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>',
             '#include "%s"' % os.path.join(ROOT, "include", "naima_b200.h"), "int main(void) {"]
    for name, fields in structs.items():
        lines.append('  printf("%s %%zu\\n", sizeof(%s));' % (name, name))
        for f in fields:
            lines.append('  printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (name, f, name, f))
    lines.append("  return 0;\n}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-o", str(exe), str(src)])
    out = subprocess.check_output([str(exe)], text=True).split()
    got = dict(zip(out[::2], map(int, out[1::2])))
    for name, fields in structs.items():
        cls = getattr(_lib, name)
        assert ctypes.sizeof(cls) == got[name], name
        for f in fields:
            assert getattr(cls, f).offset == got["%s.%s" % (name, f)], (name, f)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under naima_b200/ may import, load or execute it
    (the product has no CPU path)."""
    import re

    pkg = os.path.join(ROOT, "naima_b200")
    bad = []
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if not fn.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                continue
            with open(os.path.join(dirpath, fn)) as f:
                for k, line in enumerate(f, 1):
                    if re.search(r"^\s*(from|import)\s+oracle\b|oracle[/.]naima_oracle|oracle/_ref",
                                 line):
                        bad.append("%s:%d %s" % (fn, k, line.strip()))
    assert not bad, bad
