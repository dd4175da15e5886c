"""GPU parity of the tracer (OglPathTracer stand-in): AOV viewer, wavefront path tracer, sharded accumulate,
EXR export -- against the CPU oracle's restatement of primaryray.glsl / pathtracer.glsl."""
import os

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

W_, H_ = 96, 64
OCFG = dict(max_bounce=5, subpixel=8, tmp_lifetime=16, ray_tmin=1e-4, clamp=4.0, sun=(1.0, 0.9, 0.8))
# Image parity bound: ZERO. Every operation of the shading stage is IEEE-identical on both sides (un-fused fp32,
# IEEE sqrt / divide, deterministic double-precision sin/cos/pow -- DESIGN.md §3), so the GPU image must equal the
# oracle's bit for bit. (With libdevice / glibc sin-cos-pow, round 1 measured RMSE 6.5e-5 with 99.99 % of the
# pixels identical; the deterministic recipes removed the remainder.)
RMSE_BOUND = 0.0


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def make(A, g, cfg=None, w=W_, h=H_, seed=7):
    cfg = cfg or OCFG
    sc = A.Scene(g.nodes, g.tri_indices, g.woop, g.tris, g.mats)
    pc = A.PTConfig.make(max_bounce=cfg["max_bounce"], subpixel=cfg["subpixel"], tmp_lifetime=cfg["tmp_lifetime"],
                         ray_tmin=cfg["ray_tmin"], clamp=cfg["clamp"], sun=cfg["sun"])
    tr = A.Tracer(sc, pc, w, h, bias_seed=seed)
    cam = g.extra["cam"]
    tr.look(cam[:3], float(cam[3]), float(cam[4]), float(cam[5]))
    return sc, tr


def oracle_cam(cpu, g, w=W_, h=H_):
    cam = g.extra["cam"]
    return cam[:3], cpu.camera_matrices(float(cam[5]), float(cam[3]), float(cam[4]), w, h)


def test_primary_rays_and_viewer_modes_bit_exact(A, cpu):
    g = load_golden("city12")
    sc, tr = make(A, g)
    origin, m = oracle_cam(cpu, g)
    assert np.array_equal(bits(tr.primary_rays()), bits(cpu.primary_rays(origin, 1e-4, m["inv_proj"], m["inv_view"], W_, H_)))
    for vtype in (A.VIEW_DIFFUSE, A.VIEW_SPECULAR, A.VIEW_EMISSIVE, A.VIEW_NORMAL, A.VIEW_POSITION):
        tr.trace(False, vtype)
        img = tr.read(4).reshape(-1, 4)
        exp = cpu.primary_view(g, origin, 1e-4, m["inv_proj"], m["inv_view"], W_, H_, vtype)
        assert np.array_equal(bits(img), bits(exp)), vtype
        assert tr.spp == 0
    assert img[:, :3].any()


def test_path_tracer_matches_oracle_within_rmse(A, cpu):
    g = load_golden("city12")
    sc, tr = make(A, g)
    origin, m = oracle_cam(cpu, g)
    tr.sample(40)  # 2.5 tmpLifetime blocks: crosses stratum / primary-cache boundaries
    assert tr.spp == 40
    img = tr.read(4).reshape(-1, 4)
    exp, _, cnt = cpu.pt_render(g, origin, m["inv_proj"], m["inv_view"], W_, H_, OCFG, tr.get_bias(), 0, 40)
    d = img[:, :3] - exp[:, :3]
    rmse = float(np.sqrt((d ** 2).mean()))
    assert rmse <= RMSE_BOUND, rmse
    assert np.array_equal(bits(img), bits(exp))
    assert tr.stats()["segments"] == cnt["segments"]  # same number of traced path segments
    assert img[:, :3].mean() > 0.05


def test_batching_is_invisible(A):
    """sample(1) x N == sample(N) bit for bit: the wavefront batch size, the primary-hit cache and the
    queue order never leak into the image (the reference dispatches one sample per frame)."""
    g = load_golden("city12")
    _, a = make(A, g)
    _, b = make(A, g)
    a.sample(21)
    for _ in range(21):
        b.sample(1)
    assert b.spp == 21
    assert np.array_equal(bits(a.read()), bits(b.read()))
    c_cfg = dict(OCFG, tmp_lifetime=5)  # blocks of 5: batches 5,5,5,5,1
    _, c = make(A, g, c_cfg)
    _, d = make(A, g, c_cfg)
    c.sample(21)
    d.sample(7); d.sample(14)
    assert np.array_equal(bits(c.read()), bits(d.read()))
    # restarting from the viewer resets spp and clears the accumulation (OglPathTracer.cpp:39-46, 52-56)
    a.trace(False, A.VIEW_NORMAL)
    assert a.spp == 0
    a.sample(21)
    assert np.array_equal(bits(a.read()), bits(b.read()))


def test_max_bounce_one_and_pass_through_material(A, cpu):
    g = load_golden("city12")
    cfg = dict(OCFG, max_bounce=1)
    _, tr = make(A, g, cfg)
    origin, m = oracle_cam(cpu, g)
    tr.sample(16)
    exp, _, _ = cpu.pt_render(g, origin, m["inv_proj"], m["inv_view"], W_, H_, cfg, tr.get_bias(), 0, 16)
    assert np.array_equal(bits(tr.read(4).reshape(-1, 4)), bits(exp))  # no sampling involved => bit-exact
    # illum 0 (tinyobj default) passes straight through
    g2 = load_golden("city12")
    mats = g2.mats.copy()
    mats[:, 48:52] = 0
    g2.mats = mats
    _, tr2 = make(A, g2)
    tr2.sample(16)
    exp2, _, _ = cpu.pt_render(g2, origin, m["inv_proj"], m["inv_view"], W_, H_, OCFG, tr2.get_bias(), 0, 16)
    assert np.array_equal(bits(tr2.read(4).reshape(-1, 4)), bits(exp2))


def test_sum_accumulator_equals_running_mean(A, cpu):
    """Sharded mode: blocks added into the sum buffer in any order, then resolved, equal the
    reference's running mean up to float summation order."""
    g = load_golden("city12")
    _, a = make(A, g)
    _, b = make(A, g)
    a.sample(48)
    b.clear_sum()
    for first in (32, 0, 16):
        b.accumulate(first, 16)
    b.resolve_sum()
    x, y = a.read(4), b.read(4)
    assert float(np.sqrt(((x - y) ** 2).mean())) < 1e-6
    with pytest.raises(A.AdyptError):
        b.accumulate(3, 4)  # must start on a tmpLifetime boundary


def test_connect_stage_sun_visibility(A, cpu):
    """The wavefront's connect stage = the any-hit sun test the reference has commented out (pathtracer.glsl:132):
    escaped paths get the sun term only if a shadow ray towards normalize(0.6, 1, 0.2) is unoccluded. Off by
    default (reference behaviour); on, the image must still equal the oracle bit for bit."""
    g = load_golden("city12")
    _, tr = make(A, g)
    origin, m = oracle_cam(cpu, g)
    tr.sample(16)
    plain = tr.read(4).reshape(-1, 4).copy()
    tr.primary(A.VIEW_DIFFUSE)
    tr.set_sun_visibility(True, (0.6, 1.0, 0.2))
    l0 = tr.stats()["launches"]
    tr.sample(16)
    with_connect = tr.stats()["launches"] - l0
    tr.sample(8)
    img = tr.read(4).reshape(-1, 4)
    exp, _, _ = cpu.pt_render(g, origin, m["inv_proj"], m["inv_view"], W_, H_, OCFG, tr.get_bias(), 0, 24, sun_visibility=(0.6, 1.0, 0.2))
    assert np.array_equal(bits(img), bits(exp))
    off, _, _ = cpu.pt_render(g, origin, m["inv_proj"], m["inv_view"], W_, H_, OCFG, tr.get_bias(), 0, 24)
    assert not np.array_equal(exp, off) and exp[:, :3].sum() < off[:, :3].sum()  # some sky light is now shadowed
    tr.primary(A.VIEW_DIFFUSE)
    tr.set_sun_visibility(False)
    l0 = tr.stats()["launches"]
    tr.sample(16)
    assert with_connect == (tr.stats()["launches"] - l0) + 2 * OCFG["max_bounce"]  # one any-hit trace + one apply per bounce
    assert np.array_equal(bits(tr.read(4).reshape(-1, 4)), bits(plain))  # switching it off restores the reference image


def test_profiling_hooks_count_the_oracles_work(A, cpu):
    """adypt_tracer_set_profiling: the instrumented traversal kernel run on the wavefront's own queues counts exactly the nodes
    and triangles the oracle's megakernel restatement touches for the same samples (these counts are the C3 roofline's
    algorithmic bytes in bench.py), the image is unchanged by the hooks, and the stage timer returns a time for every stage."""
    g = load_golden("city12")
    sc, tr = make(A, g)
    origin, m = oracle_cam(cpu, g)
    tr.set_profiling(stage_times=True, trace_counters=True)
    tr.sample(40)
    pr = tr.profile()
    exp, _, cnt = cpu.pt_render(g, origin, m["inv_proj"], m["inv_view"], W_, H_, OCFG, tr.get_bias(), 0, 40)
    assert np.array_equal(bits(tr.read(4).reshape(-1, 4)), bits(exp))
    assert pr["trace"]["nodes"] + pr["primary"]["nodes"] == cnt["nodes"]
    assert pr["trace"]["tris"] + pr["primary"]["tris"] == cnt["tris"]
    assert pr["primary"]["rays"] == 3 * W_ * H_  # one primary traversal per tmpLifetime block (pathtracer.glsl:113-127)
    assert pr["primary"]["rays"] + pr["trace"]["rays"] == cnt["segments"] == tr.stats()["segments"]  # traced segments; cached primary hits are not
    assert max(pr["trace"]["max_stack"], 0) <= cnt["max_stack"]
    for stage in ("generate", "trace_primary", "shade_primary", "trace_bounce", "shade_bounce", "accumulate"):
        assert pr["stage_ms"][stage] > 0.0 and pr["stage_launches"][stage] > 0, stage
    assert pr["stage_launches"]["trace_bounce"] == 3 * (OCFG["max_bounce"] - 1)
    tr.set_profiling()
    assert tr.profile()["stage_ms"]["trace_bounce"] == 0.0  # reset by the previous read
    assert "trace_kernel" in sc.kernel_name(False) and "trace_kernel" in sc.kernel_name(True)


def test_config_limits(A):
    g = load_golden("city12")
    sc = A.Scene(g.nodes, g.tri_indices, g.woop, g.tris, g.mats)
    with pytest.raises(A.AdyptError) as e:
        A.Tracer(sc, A.PTConfig.make(max_bounce=5003), 8, 8)
    assert e.value.code == -6  # ADYPT_ERANGE: 2*maxBounce exceeds the reference's 10 005 Sobol dimensions (Sobol.hpp:9)
    A.Tracer(sc, A.PTConfig.make(max_bounce=5002), 8, 8).close()
    with pytest.raises(A.AdyptError):
        A.Tracer(A.Scene(g.nodes, g.tri_indices, g.woop), A.PTConfig.make(), 8, 8)  # traversal-only scene cannot shade


def test_max_bounce_beyond_64_sobol_dimensions(A, cpu):
    """maxBounce 40 (80 Sobol dimensions; round 1 stopped at 64) and 150 (300 dimensions, queue counters sized from the
    configuration): images equal the oracle bit for bit. The reference takes any maxBounce up to 5002 (OglPathTracer.cpp:139,
    Sobol.hpp:9)."""
    g = load_golden("city12")
    for mb, spp in ((40, 20), (150, 4)):
        cfg = dict(OCFG, max_bounce=mb)
        sc = A.Scene(g.nodes, g.tri_indices, g.woop, g.tris, g.mats)
        tr = A.Tracer(sc, A.PTConfig.make(max_bounce=mb, sun=cfg["sun"]), W_, H_, bias_seed=5)
        cam = g.extra["cam"]
        tr.look(cam[:3], float(cam[3]), float(cam[4]), float(cam[5]))
        tr.sample(spp)
        m = cpu.camera_matrices(float(cam[5]), float(cam[3]), float(cam[4]), W_, H_)
        exp, _, cnt = cpu.pt_render(g, cam[:3], m["inv_proj"], m["inv_view"], W_, H_, cfg, tr.get_bias(), 0, spp)
        assert np.array_equal(bits(tr.read(4).reshape(-1, 4)), bits(exp)), mb
        assert tr.stats()["segments"] == cnt["segments"]
        tr.close()
        sc.close()


def test_save_exr(A, tmp_path):
    from exr_reader import read_exr
    g = load_golden("city12")
    _, tr = make(A, g)
    tr.sample(16)
    img = tr.read(3)
    for fp16 in (False, True):
        p = str(tmp_path / f"o{int(fp16)}.exr")
        tr.save_exr(p, fp16)
        r = read_exr(p)
        got = np.stack([r["data"]["R"], r["data"]["G"], r["data"]["B"]], axis=2)
        if fp16:  # the file is written by the same code as A.write_exr (pinned against tinyexr in test_capi_host.py)
            q = str(tmp_path / "w.exr")
            A.write_exr(q, img, fp16=True)
            w = read_exr(q)
            exp = np.stack([w["data"]["R"], w["data"]["G"], w["data"]["B"]], axis=2)
            assert np.abs(exp - img).max() <= np.abs(img).max() * 2.0 ** -10
        else:
            exp = img
        assert np.array_equal(got, exp)
    with pytest.raises(A.AdyptError):
        tr.save_exr("/nonexistent_dir/x.exr")


def test_c3_style_render_on_small_city(A, cpu, city_small):
    """A mixed-material city through the reference builder: every illum branch, 2 blocks, RMSE vs oracle."""
    from adypt_b200 import workloads as W
    mesh, b = city_small
    sc = A.Scene(b.nodes, b.tri_indices, None, b.tris, b.mats)
    w, h = 160, 90
    tr = A.Tracer(sc, A.PTConfig.make(sun=(1.0, 1.0, 1.0)), w, h, bias_seed=7)
    cam = W.city_camera(24)
    tr.look(cam["position"], cam["yaw"], cam["pitch"], cam["fov"])
    tr.sample(32)
    m = cpu.camera_matrices(cam["fov"], cam["yaw"], cam["pitch"], w, h)
    cfg = dict(OCFG, sun=(1.0, 1.0, 1.0))
    exp, _, _ = cpu.pt_render(b, cam["position"], m["inv_proj"], m["inv_view"], w, h, cfg, tr.get_bias(), 0, 32)
    assert np.array_equal(bits(tr.read(4).reshape(-1, 4)), bits(exp))
