#!/usr/bin/env python3
"""Does the GPU run one tracer's shading (latency-bound, 40-48 % of the issue slots) under another tracer's traversal (issue-bound)
when both are in flight on their own streams? Two independent tracers on one scene render 32 spp each at the same time, against one
tracer rendering 64 spp; with the traversal kernel limited to fewer CTAs per SM so that the other stream's kernels find room."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adypt_b200 as A
from adypt_b200 import workloads as W, host

mesh = W.city(183, 1, mixed_materials=True)
sc = host.build_scene(mesh).upload(0)
cam = W.city_camera(183)
def tracer():
    t = A.Tracer(sc, A.PTConfig.make(sun=(1.0, 1.0, 1.0)), 1920, 1080, bias_seed=7)
    t.look(cam['position'], cam['yaw'], cam['pitch'], cam['fov'])
    t.sample(16); t.sync()
    return t
t1, t2 = tracer(), tracer()
def one():
    t1.trace(False, 0); t1.sync()
    t0 = time.perf_counter(); t1.sample(64); t1.sync(); return time.perf_counter() - t0
def two(split=16):
    t1.trace(False, 0); t2.trace(False, 0); t1.sync(); t2.sync()
    t0 = time.perf_counter()
    for _ in range(32 // split):
        t1.sample(split); t2.sample(split)
    t1.sync(); t2.sync()
    return time.perf_counter() - t0
for ctas in (0, 7, 6, 5, 4):
    sc.configure(ctas, 0, 0)
    a = min(one() for _ in range(3)); b = min(two() for _ in range(3))
    print(f"trace CTAs/SM {ctas or 8}: one tracer 64 spp {a*1e3:.1f} ms; two tracers 2 x 32 spp concurrently {b*1e3:.1f} ms", flush=True)
