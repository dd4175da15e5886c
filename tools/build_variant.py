#!/usr/bin/env python3
"""Builds an experimental copy of the library with extra -D flags into adypt_b200/lib/variants/NAME/ (travels with gpurun,
never loaded by the product): python tools/build_variant.py NAME -DADYPT_X=1 ...  Use with TUNE_LIB=<path> tools/gpu_tune.py."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adypt_b200 import build as B

def main():
    name, defs = sys.argv[1], sys.argv[2:]
    B.build_library()  # objects of the unchanged sources
    out = os.path.join(B.LIBDIR, "variants", name)  # under adypt_b200/lib: git-ignored (*.so, *.o) but shipped by gpurun
    os.makedirs(out, exist_ok=True)
    objdir = os.path.join(B.HERE, "..", "build", "obj")
    objs, procs = [], []
    for src in B._sources():
        base = os.path.relpath(src, B.CSRC).replace(os.sep, "_") + ".o"
        if src.endswith(".cu"):
            obj = os.path.join(out, base)
            procs.append(subprocess.Popen([B._nvcc()] + B.NVCC_FLAGS + defs + ["-x", "cu", "-c", src, "-o", obj]))
        else:
            obj = os.path.join(objdir, base)
        objs.append(obj)
    for p in procs:
        if p.wait() != 0:
            raise SystemExit("nvcc failed")
    lib = os.path.join(out, "libadypt_b200.so")
    subprocess.check_call([B._nvcc(), "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs + ["-lz", "-ldl"])
    print(os.path.relpath(lib, os.path.join(B.HERE, "..")))

if __name__ == "__main__":
    main()
