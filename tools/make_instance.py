#!/usr/bin/env python3
"""Writes a ready-to-run Adypt instance (OBJ + MTL + .config) for one of the benchmark scenes.
   python tools/make_instance.py {c1|c2|c3|c4} <directory> [--width W --height H]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adypt_b200 import host, workloads as W

def main():
    ap = argparse.ArgumentParser(); ap.add_argument("scene"); ap.add_argument("directory")
    ap.add_argument("--width", type=int, default=1920); ap.add_argument("--height", type=int, default=1080)
    a = ap.parse_args()
    if a.scene == "c1": mesh, cam = W.sphere_lattice(5), W.lattice_camera()
    elif a.scene == "c2": mesh, cam = W.city(183, 1), W.city_camera(183)
    elif a.scene == "c3": mesh, cam = W.city(183, 1, mixed_materials=True), W.city_camera(183)
    elif a.scene == "c4": mesh, cam = W.city(577, 1), W.city_camera(577)
    else: raise SystemExit("unknown scene")
    obj = mesh.write_obj(a.directory)
    c = host.InstanceConfig.default()
    c.width, c.height = a.width, a.height
    c.obj_filename = os.path.abspath(obj).encode()
    c.bvh_filename = os.path.abspath(os.path.join(a.directory, mesh.name + ".bvh")).encode()
    c.pt.sun[0] = c.pt.sun[1] = c.pt.sun[2] = 1.0
    c.cam.yaw, c.cam.pitch, c.cam.fov = cam["yaw"], cam["pitch"], cam["fov"]
    for i in range(3): c.cam.position[i] = cam["position"][i]
    path = os.path.join(a.directory, a.scene + ".config"); c.save(path); print(path)

if __name__ == "__main__":
    main()
