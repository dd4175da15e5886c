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
