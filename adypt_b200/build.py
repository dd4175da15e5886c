"""In-tree build of libadypt_b200.so (sm_100a only). nvcc cross-compiles without a GPU.

The flags are part of the numerical contract: -fmad=false (no implicit contraction; the kernels place
their fused multiply-adds explicitly) and the defaults -prec-div=true -prec-sqrt=true -ftz=false.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libadypt_b200.so")
BINDIR = os.path.join(HERE, "bin")
CLI = os.path.join(BINDIR, "adypt_headless")

CU_SOURCES = ["scene.cu", "tracer.cu", "group.cu"]
CPP_SOURCES = ["hostmath.cpp", "exr.cpp", "host/bvh_build.cpp", "host/obj_loader.cpp", "host/host_api.cpp", "host/config.cpp", "host/image_decode.cpp", "host/image_decode_more.cpp", "host/jpeg_decode.cpp"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-DADYPT_NO_FMAD", "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off",
    # host code only: signed overflow wraps. The file decoders follow stb_image's int arithmetic (IDCT, colour conversion),
    # which overflows on corrupt input; with -fwrapv that is defined behaviour and gives what stb_image computes on x86
    "-Xcompiler", "-fwrapv",
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _sources():
    return [os.path.join(CSRC, s) for s in CU_SOURCES + CPP_SOURCES]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(d, f) for d, _, fs in os.walk(CSRC) for f in fs] + [os.path.join(HERE, "..", "include", "adypt_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(HERE, "..", "build", "obj")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in _sources():
        obj = os.path.join(objdir, os.path.relpath(src, CSRC).replace(os.sep, "_") + ".o")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-x", "cu", "-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose:
            print(out)
    link = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-lz", "-ldl"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    # headless CLI (replaces the GLFW/ImGui viewer): plain C++ over the C-ABI, finds the library next to itself
    os.makedirs(BINDIR, exist_ok=True)
    cli = ["g++", "-std=c++11", "-O2", "-Wall", "-I", os.path.join(HERE, "..", "include"), os.path.join(CSRC, "cli", "adypt_headless.cpp"),
           "-o", CLI, "-L", LIBDIR, "-ladypt_b200", "-Wl,-rpath,$ORIGIN/../lib"]
    r = subprocess.run(cli, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"CLI build failed:\n{r.stdout}")
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
