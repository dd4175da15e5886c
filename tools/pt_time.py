#!/usr/bin/env python3
"""Times the C3 render (1920x1080, 64 spp, maxBounce 5, mixed-material city) a few times (GPU box only).
TUNE_LIB=<path> loads an experimental build from tools/build_variant.py."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adypt_b200 as A
if os.environ.get('TUNE_LIB'):
    A.LIB_PATH = os.path.abspath(os.environ['TUNE_LIB'])
from adypt_b200 import workloads as W, host

def main():
    mesh = W.city(183, 1, mixed_materials=True)
    sc = host.build_scene(mesh).upload(0)
    cam = W.city_camera(183)
    t = A.Tracer(sc, A.PTConfig.make(sun=(1.0, 1.0, 1.0)), 1920, 1080, bias_seed=7)
    t.look(cam['position'], cam['yaw'], cam['pitch'], cam['fov'])
    t.sample(16); t.sync()
    ts = []
    for _ in range(int(os.environ.get('REPS', '4'))):
        t.trace(False, 0)
        t0 = time.perf_counter(); t.sample(64); t.sync(); ts.append(time.perf_counter() - t0)
    t.set_profiling(stage_times=True); t.profile()
    t.trace(False, 0); t.sample(64); t.sync()
    pr = t.profile(); t.set_profiling()
    print('   stage ms per 64 spp:', {k: round(v, 3) for k, v in pr['stage_ms'].items() if v > 0})
    img = t.read(3)
    print(f"{os.environ.get('TUNE_LIB', 'in-tree')}: C3 64 spp min {min(ts)*1e3:.1f} ms  med {np.median(ts)*1e3:.1f} ms  {1920*1080*64/min(ts)/1e9:.3f} Gsamples/s  mean {float(img.mean()):.9f}")

if __name__ == '__main__':
    main()
