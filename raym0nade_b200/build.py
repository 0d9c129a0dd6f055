"""Build libraym0nade_b200.so in-tree with nvcc for sm_100a (no GPU needed: cross-compiles).

    python -m raym0nade_b200.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libraym0nade_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--ftz=false", "--prec-div=true", "--prec-sqrt=true",      # IEEE semantics: NaN/denormal behaviour is part of the control flow
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-Wall,-Wno-unused-function",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
    "-shared", "-cudart", "static",
    # keep the statically linked C++ runtime private: the host process (python, numpy, torch) has its own
    "-Xlinker", "--exclude-libs,ALL", "-Xlinker", "-Bsymbolic",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cpp")))


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(ROOT, "include", "*.h")) + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


HOST_DIR = os.path.join(HERE, "host")
HOST_OUT = os.path.join(HERE, "raym0nade")          # the console program (C++ host side above the C ABI)
HOST_FLAGS = ["-O2", "-std=c++17", "-ffp-contract=off", "-Wall", "-Wextra"]


HOST_FXAA_OUT = os.path.join(HERE, "raym0nade_fxaa")  # the stand-alone FXAA program (the reference's fxaa.cpp)
HOST_MAINS = {"main.cpp": HOST_OUT, "fxaa_tool.cpp": HOST_FXAA_OUT}


def host_sources(with_main=True, main="main.cpp"):
    """the host translation units shared by every program, plus the one holding `main` if asked for"""
    src = sorted(glob.glob(os.path.join(HOST_DIR, "*.cpp")))
    return [s for s in src if os.path.basename(s) not in HOST_MAINS or (with_main and os.path.basename(s) == main)]


def host_link_flags():
    # $ORIGIN: the binary finds libraym0nade_b200.so next to itself wherever the tree is copied (the GPU box);
    # zlib (PNG export) is linked statically
    return ["-I", os.path.join(ROOT, "include"), "-I", HOST_DIR, "-L", HERE, "-lraym0nade_b200",
            "-Wl,-rpath,$ORIGIN", "-l:libz.a"]


def build_host(force=False, verbose=False):
    """g++ the host side (raym0nade_b200/host) into the `raym0nade` console and `raym0nade_fxaa` next to the library."""
    deps = glob.glob(os.path.join(HOST_DIR, "*.cpp")) + glob.glob(os.path.join(HOST_DIR, "*.hpp")) + glob.glob(os.path.join(ROOT, "include", "*.h")) + [OUT]
    for main, out in HOST_MAINS.items():
        if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
            continue
        cmd = [os.environ.get("CXX", "g++")] + HOST_FLAGS + host_sources(main=main) + ["-o", out] + host_link_flags()
        if verbose:
            print(" ".join(cmd))
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError("g++ failed on the host side (%d)" % r.returncode)
    return HOST_OUT


def build(force=False, verbose=False, extra=()):
    build_library(force, verbose, extra)
    build_host(force, verbose)
    return OUT


def build_library(force=False, verbose=False, extra=()):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + list(extra) + ["-o", OUT] + sources()
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed (%d)" % r.returncode)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(OUT)
