#!/usr/bin/env python3
"""Experiment check (GPU box): the shared-memory ray-pool kernel (variant 12, csrc/traverse_pool.cuh) against the product
kernel on the committed fixture scenes: ids, t, uv and any-hit bits must be identical."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('ADYPT_EXPERIMENTAL', '1')
import adypt_b200 as A
ok = True
for name in ("tiny_two_triangles", "tiny_shared_edge", "tiny_strip", "tiny_deep", "city12"):
    z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", name + ".npz"))
    sc = A.Scene(z["nodes"], z["tri_indices"], z["woop"])
    rays = z["rays"]
    for reps in (1, 37):
        r = np.tile(rays, (reps, 1))
        sc.configure(0, 0, 0)
        a = sc.trace_closest(r); oa = sc.trace_any(r)
        sc.configure(0, 0, 12)
        b = sc.trace_closest(r); ob = sc.trace_any(r)
        same = all(np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)) for k in ("tri", "t", "uv")) and np.array_equal(oa, ob)
        print(name, r.shape[0], "identical" if same else "DIFFERENT", flush=True)
        ok &= same
sys.exit(0 if ok else 1)
