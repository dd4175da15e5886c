"""Build-time guard for the occupancy the traversal kernel is designed around: 64 registers and 14 KB of
shared memory per 128-thread CTA = 8 CTAs per SM, and no register spills (the only local memory is the 56-entry stack
overflow). Read from the built library with cuobjdump; needs no GPU."""
import os
import re
import shutil
import subprocess

import pytest

import adypt_b200


def _cuobjdump():
    for c in ("/usr/local/cuda/bin/cuobjdump", shutil.which("cuobjdump")):
        if c and os.path.exists(c):
            return c
    return None


@pytest.mark.skipif(_cuobjdump() is None, reason="cuobjdump not found")
def test_traversal_kernels_fit_eight_ctas_per_sm():
    out = subprocess.run([_cuobjdump(), "-res-usage", adypt_b200.LIB_PATH], capture_output=True, text=True, check=True).stdout
    usage = dict(re.findall(r"Function (\S+):\s*\n\s*(REG:.*)", out))
    # the product kernels: closest-hit and any-hit with the default tuning (CVT planes 3, 8 CTAs/SM, triangle batch 12, staged, packed
    # evaluations on 96-byte nodes)
    names = [n for n in usage if re.match(r"_ZN5adypt12trace_kernelILb[01]ELb0ELi3ELi8ELi12ELb1ELb0ELi2EEE", n)]
    assert len(names) == 2, sorted(usage)
    for n in names:
        f = {k: int(v) for k, v in re.findall(r"([A-Z]+):(\d+)", usage[n])}
        assert f["REG"] <= 64, (n, f)                 # 128 threads x 64 registers x 8 CTAs = the whole register file
        assert f["SHARED"] <= 14336 + 1024, (n, f)    # 4 warps x 3.5 KB (+ 1 KB the system reserves per CTA)
        assert f["STACK"] == 56 * 8, (n, f)           # the local stack overflow and nothing else: no spills
    for n, u in usage.items():
        f = {k: int(v) for k, v in re.findall(r"([A-Z]+):(\d+)", u)}
        if "k_shade_bounce" in n:  # 128 threads x 8 CTAs (or 256 x 4) per SM; a handful of spilled values at most
            assert f["REG"] <= 64 and f["STACK"] <= 64, (n, f)
        if "k_shade_primaryILi4E" in n or "k_shade_primary_sorted" in n:  # bounce 0 (default: class-sorted, 128 threads x 8 CTAs per
            assert f["REG"] <= 64 and f["STACK"] <= 384, (n, f)  # SM): local memory = the parked chunk (8 x 2 float4) + a few spills
            assert f["SHARED"] <= 12288, (n, f)                   # 8 CTAs x 12 KB of sort buffers fit beside the L1


@pytest.mark.skipif(_cuobjdump() is None, reason="cuobjdump not found")
def test_product_traversal_kernel_uses_packed_fp32_and_256_bit_node_loads():
    """The SASS of the product kernels must hold what DESIGN.md 4.1 says they are built from: packed FFMA2 / FADD2 / FMUL2 for the slab and
    Woop evaluations, and exactly three 256-bit loads (the 96-byte node) -- a silent fall-back to the scalar form would still pass every
    parity test."""
    for any_hit in (0, 1):
        fn = f"_ZN5adypt12trace_kernelILb{any_hit}ELb0ELi3ELi8ELi12ELb1ELb0ELi2EEEvNS_11TraceParamsE"
        sass = subprocess.run([_cuobjdump(), "-sass", "-fun", fn, adypt_b200.LIB_PATH], capture_output=True, text=True).stdout
        ops = re.findall(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", sass, flags=re.M)
        assert len(ops) > 500, (fn, len(ops))
        count = lambda prefix: sum(1 for o in ops if o.startswith(prefix))
        assert count("FFMA2") >= 36 and count("FADD2") >= 12 and count("FMUL2") >= 6, (fn, count("FFMA2"), count("FADD2"), count("FMUL2"))
        assert count("LDG.E.ENL2.256") == 3, (fn, count("LDG.E.ENL2.256"))
        assert count("I2F.U8") == 24, (fn, count("I2F.U8"))  # three of the six planes on the conversion pipe
