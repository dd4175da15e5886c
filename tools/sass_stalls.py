#!/usr/bin/env python3
"""Static look at ptxas's schedule of the closest-hit kernel: per code region (split at the first / last I2F.U8 = the node
step), the number of instructions and the sum of the stall counts in their control words (issue cycles one warp needs
when nothing else holds it up). usage: tools/sass_stalls.py [libadypt_b200.so]"""
import re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "adypt_b200/lib/libadypt_b200.so"
fn = "_ZN5adypt12trace_kernelILb0ELb0ELi4ELi8ELi12ELb1EEEvNS_11TraceParamsE"
out = subprocess.run(["cuobjdump", "-sass", "-fun", fn, lib], capture_output=True, text=True).stdout.splitlines()
ins = []
for i, l in enumerate(out):
    m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/", l)
    if m and i + 1 < len(out):
        m2 = re.match(r"\s+/\* 0x([0-9a-f]{16}) \*/", out[i + 1])
        hi = int(m2.group(1), 16)
        ins.append((m.group(2).strip(), (hi >> 41) & 0xF, (hi >> 45) & 1, (hi >> 52) & 0x3F))
idx = [k for k, (t, *_r) in enumerate(ins) if "I2F.U8" in t]
a, b = idx[0], idx[-1]
# node step = from the load of the node array's base (the LDC.64 before the first conversion) to the BSYNC that closes it
before = [k for k in range(a) if ins[k][0].startswith("LDC.64")]
lo = before[-1] if before else a
hi_ = next(k for k in range(b, len(ins)) if ins[k][0].startswith("BSYNC"))
def summ(name, r):
    seg = ins[r[0]:r[1]]
    print(f"{name:10s} instr {len(seg):4d}  stall-sum {sum(s for _, s, _, _ in seg):5d}  waits {sum(1 for _, _, _, w in seg if w):4d}")
summ("all", (0, len(ins)))
summ("node step", (lo, hi_))
summ("after", (hi_, len(ins)))
