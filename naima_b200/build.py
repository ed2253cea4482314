# -*- coding: utf-8 -*-
"""Build the sm_100a shared library (``naima_b200/libnaima_b200.so``) in-tree.

    python -m naima_b200.build            # build if sources are newer
    python -m naima_b200.build --force

nvcc cross-compiles for sm_100a without a GPU.  The library is git-ignored but
travels to the GPU box with the repository snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libnaima_b200.so")
SOURCES = ["nb_kernels.cu"]
HEADERS = ["nb_math.cuh", os.path.join("..", "..", "include", "naima_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


HOST_LIB = os.path.join(HERE, "libnaima_b200_host.so")
HOST_SRC = os.path.join(CSRC, "nb_host.c")


def build_host_library(force=False):
    """The host-side helper (random draws of the device-resident sampler), plain C.
    -ffp-contract=off: its arithmetic must round like NumPy's."""
    if (not force and os.path.exists(HOST_LIB)
            and os.path.getmtime(HOST_LIB) >= os.path.getmtime(HOST_SRC)):
        return HOST_LIB
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", HOST_LIB, HOST_SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("gcc failed:\n%s\n%s" % (" ".join(cmd), r.stderr))
    return HOST_LIB


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    try:  # optional: the sampler falls back to NumPy draws without the host helper
        build_host_library(force=force)
    except (OSError, RuntimeError) as e:
        import warnings

        warnings.warn("naima_b200: host helper not built (%s); NumPy draws will be used" % (e,))
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else [])
    cmd += ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), r.stderr))
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
