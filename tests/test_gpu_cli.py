"""The headless CLI (replacement for the GLFW/ImGui viewer) end to end: .config + OBJ/MTL in, EXR out."""
import json
import os
import subprocess

import numpy as np
import pytest

from adypt_b200 import host, workloads as W
from conftest import ROOT
from exr_reader import read_exr

CLI = os.path.join(ROOT, "adypt_b200", "bin", "adypt_headless")


def make_instance(tmp_path, w=80, h=60):
    mesh = W.city(8, 5, mixed_materials=True, name="clicity")
    obj = mesh.write_obj(str(tmp_path))
    cfg = host.InstanceConfig.default()
    cfg.width, cfg.height = w, h
    cfg.obj_filename = obj.encode()
    cfg.bvh_filename = str(tmp_path / "clicity.bvh").encode()
    cfg.pt.sun[0], cfg.pt.sun[1], cfg.pt.sun[2] = 1.0, 0.9, 0.8
    cam = W.city_camera(8)
    cfg.cam.yaw, cfg.cam.pitch, cfg.cam.fov = cam["yaw"], cam["pitch"], cam["fov"]
    for i in range(3):
        cfg.cam.position[i] = cam["position"][i]
    path = str(tmp_path / "clicity.config")
    cfg.save(path)
    return mesh, cfg, path


def test_cli_exists_and_prints_usage():
    assert os.path.exists(CLI)
    r = subprocess.run([CLI], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr


@pytest.mark.gpu
def test_cli_render_matches_api_and_oracle(A, cpu, tmp_path):
    mesh, cfg, path = make_instance(tmp_path)
    out = str(tmp_path / "r.exr")
    r = subprocess.run([CLI, path, "--spp", "24", "--seed", "7", "--out", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "[INSTANCE]Info: Initialized from" in r.stdout and "[PT]INFO: Saved image to" in r.stdout
    info = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert "[INSTANCE]Info: " + path + " saved" in r.stdout  # the destructor rewrites the .config (Instance.cpp:83-86)
    assert info["spp"] == 24 and info["samples_per_s"] > 0
    assert os.path.exists(str(tmp_path / "clicity.bvh"))  # cache written like Instance.cpp:25
    img = read_exr(out)
    got = np.stack([img["data"]["R"], img["data"]["G"], img["data"]["B"]], axis=2)
    # same render through the Python binding
    hs = host.HostScene.from_obj(cfg.obj_filename.decode()).build_bvh()
    sc = hs.upload(0)
    tr = A.Tracer(sc, cfg.pt, cfg.width, cfg.height, bias_seed=7)
    tr.look(tuple(cfg.cam.position), cfg.cam.yaw, cfg.cam.pitch, cfg.cam.fov)
    tr.sample(24)
    assert np.array_equal(got.view(np.uint32), tr.read(3).view(np.uint32))
    # ... and the oracle
    hs.woop = cpu.build_woop(hs.tris, hs.tri_indices)
    m = cpu.camera_matrices(cfg.cam.fov, cfg.cam.yaw, cfg.cam.pitch, cfg.width, cfg.height)
    ocfg = dict(max_bounce=5, subpixel=8, tmp_lifetime=16, ray_tmin=cfg.pt.ray_tmin, clamp=4.0, sun=(1.0, 0.9, 0.8))
    exp, _, _ = cpu.pt_render(hs, tuple(cfg.cam.position), m["inv_proj"], m["inv_view"], cfg.width, cfg.height, ocfg, tr.get_bias(), 0, 24)
    assert np.array_equal(got.reshape(-1, 3).view(np.uint32), exp[:, :3].copy().view(np.uint32))
    # second run: per-frame dispatch (one Trace(true) per sample, like the viewer loop) from the .bvh cache, fp16 AOV
    out2 = str(tmp_path / "r2.exr")
    r2 = subprocess.run([CLI, path, "--spp", "24", "--seed", "7", "--per-frame", "--out", out2, "--keep-config"], capture_output=True, text=True)
    assert r2.returncode == 0 and "[SBVH]" not in r2.stdout  # loaded from cache, not rebuilt
    img2 = read_exr(out2)
    assert np.array_equal(img2["data"]["R"], img["data"]["R"])
    out3 = str(tmp_path / "n.exr")
    r3 = subprocess.run([CLI, path, "--viewer", "normal", "--fp16", "--out", out3, "--keep-config"], capture_output=True, text=True)
    assert r3.returncode == 0
    n = read_exr(out3)
    assert all(t == 1 for _, t in n["channels"])
    nn = np.stack([n["data"]["R"], n["data"]["G"], n["data"]["B"]], axis=2)
    lens = np.linalg.norm(nn, axis=2)
    assert np.all((np.abs(lens - 1) < 5e-3) | (lens == 0))


@pytest.mark.gpu
def test_cli_reports_bad_config(tmp_path):
    p = tmp_path / "bad.config"
    p.write_text('{"width": 10}')
    r = subprocess.run([CLI, str(p)], capture_output=True, text=True)
    assert r.returncode == 1 and "[PARSER]ERR" in r.stdout and "[INSTANCE]Err: Invalid instance" in r.stdout


@pytest.mark.gpu
def test_render_group_matches_single_gpu(A, tmp_path):
    """adypt_group_*: one process, N devices, blocks round-robin + one NCCL reduce. With one device the image equals
    the sum-mode render exactly; with two (when the box has them) it equals it up to float summation order."""
    mesh, cfg, path = make_instance(tmp_path)
    hs = host.HostScene.from_obj(cfg.obj_filename.decode()).build_bvh()
    tr = A.Tracer(hs.upload(0), cfg.pt, cfg.width, cfg.height, bias_seed=7)
    tr.look(tuple(cfg.cam.position), cfg.cam.yaw, cfg.cam.pitch, cfg.cam.fov)
    tr.sample(40)
    mean = tr.read(4)
    ndev = A.device_count()
    for devices in ([[0], [0, 1]] if ndev >= 2 else [[0]]):
        g = host.RenderGroup(hs, cfg.pt, cfg.width, cfg.height, devices, bias_seed=7)
        g.look(tuple(cfg.cam.position), cfg.cam.yaw, cfg.cam.pitch, cfg.cam.fov)
        g.render(40)
        img = g.read(4)
        assert float(np.sqrt(((img - mean) ** 2).mean())) < 1e-6
        assert np.all(img[..., 3] == 1.0)
        g.close()
    if ndev >= 2:
        out = str(tmp_path / "g.exr")
        r = subprocess.run([CLI, path, "--spp", "40", "--seed", "7", "--gpus", "2", "--out", out, "--keep-config"], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        e = read_exr(out)
        got = np.stack([e["data"]["R"], e["data"]["G"], e["data"]["B"]], axis=2)
        assert float(np.sqrt(((got - mean[..., :3]) ** 2).mean())) < 1e-6
