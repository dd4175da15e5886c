#!/usr/bin/env python3
"""Headline benchmark: Mrays/s closest-hit on incoherent rays (BASELINE.json configs[1], "C2").

Workload: ~1.0M-triangle procedural box city, 8M incoherent cosine-weighted bounce rays spawned from the hits
of 1000x1000 primary rays. One "step" = one closest-hit pass over the whole 8M-ray batch.

  python bench.py [--gpus N] [--steps K] [--warmup W]            product arm (CUDA, sm_100a)
  python bench.py --impl reference [...]                          reference arm: the reference's traversal.glsl on the CPU
  torchrun --nproc-per-node N bench.py --gpus N ...               one rank per GPU, weak scaling (each rank
                                                                  traces its own 8M-ray batch, no collective)
Prints ONE JSON line on rank 0. See DESIGN.md §6 for how every field is measured.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from adypt_b200 import workloads as W  # noqa: E402

METRIC = "Mrays/s closest-hit (incoherent)"
UNIT = "Mrays/s"
CELLS, SCENE_SEED, RAY_SEED, PRIMARY = 183, 1, 42, 1000
CACHE = os.path.join(ROOT, ".cache", "scenes")


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


def build_inputs(mixed=False, rank=0, barrier=None):
    """C2 scene arrays from the product's own host stages (adypt_b200.host): Triangle[] assembly and the
    from-scratch SBVH -> CWBVH builder, byte-identical to the reference's src/BVH pipeline (tests/
    test_host_builder.py). The node/index arrays are cached in the reference's own .bvh format
    (WideBVH.cpp:9-66), as Instance::Initialize does (Instance.cpp:19-31). Nothing here touches oracle/."""
    from adypt_b200 import host
    mesh = W.city(CELLS, SCENE_SEED, mixed_materials=mixed)
    hs = host.HostScene.from_triangles(mesh.positions(), mesh.face_mat, host.materials_array(mesh.materials))
    os.makedirs(CACHE, exist_ok=True)
    bvh_path = os.path.join(CACHE, mesh.name + ".bvh")
    if barrier is not None and rank != 0:
        barrier()  # rank 0 builds and caches first; the others then load the .bvh instead of building it N times
    if not hs.load_bvh(bvh_path):
        t0 = time.perf_counter()
        hs.build_bvh()
        log(f"built CWBVH in {time.perf_counter() - t0:.1f} s")
        hs.save_bvh(bvh_path)
    if barrier is not None and rank == 0:
        barrier()
    return mesh, hs


def workload_config(n_rays, extra=None):
    c = {"workload": "C2: 1M-tri procedural box city (cells=183, seed=1), 8M incoherent cosine-diffuse bounce rays, closest-hit",
         "triangles": None, "rays_per_step": int(n_rays), "ray_seed": RAY_SEED,
         "timing": "cuda events per step on the launch stream; 256 MiB L2 flush between steps (outside the events)"}
    if extra:
        c.update(extra)
    return c


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons every 200 ms while the timed region runs."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def cpu_tracer(bvh):
    """The CPU implementation timed as the baseline: the reference's own shaders/traversal.glsl compiled for the CPU
    (oracle/_ref/libadypt_glsl.so, kind "reference") when that build exists, else the oracle's C++ port (kind "port")."""
    from oracle import cpu, glsl_ref
    if glsl_ref.available():
        return (lambda r: glsl_ref.trace_closest(bvh.nodes, bvh.tri_indices, bvh.woop, r)[0]), "reference", \
            "the reference's shaders/traversal.glsl compiled for the CPU from its own text (oracle/glsl_transpile.py)"
    return (lambda r: cpu.trace_closest(bvh.nodes, bvh.tri_indices, bvh.woop, r, want_t=False)["tri"]), "port", \
        "multithreaded C++ port of shaders/traversal.glsl (oracle/oracle.cpp)"


def cpu_leg(bvh, rays, budget_s=10.0):
    """The CPU baseline on all host threads over a bounded sample (+ the oracle's counters on the same sample)."""
    from oracle import cpu
    cores = cpu.hardware_threads()
    probe = min(rays.shape[0], 262144)
    bvh.woop = cpu.build_woop(bvh.tris, bvh.tri_indices)
    trace, kind, what = cpu_tracer(bvh)
    t0 = time.perf_counter()
    trace(rays[:probe])
    rate = probe / (time.perf_counter() - t0)
    n = int(min(rays.shape[0], max(probe, rate * budget_s)))
    t0 = time.perf_counter()
    tri = trace(rays[:n])
    dt = time.perf_counter() - t0
    r = cpu.trace_closest(bvh.nodes, bvh.tri_indices, bvh.woop, rays[:n], want_t=False)
    assert np.array_equal(tri, r["tri"]), "CPU baseline and oracle disagree"
    return dict(value=n / dt / 1e6, unit=UNIT, cores=cores, kind=kind, what=what,
                sample=f"first {n} of the {rays.shape[0]} rays of the same batch, {cores} threads",
                per_core=n / dt / 1e6 / cores), r["counters"], n


def bytes_per_ray(nodes, tris, hits, n):
    """Algorithmic bytes: 80 B per node visited + 48 B per triangle tested + 4 B index remap per
    hit + 32 B ray in + 16 B hit out."""
    return 80.0 * nodes / n + 48.0 * tris / n + 4.0 * hits / n + 32.0 + 16.0


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import cpu
    mesh, bvh = build_inputs()
    bvh.woop = cpu.build_woop(bvh.tris, bvh.tri_indices)
    cam = W.city_camera(CELLS)
    m = cpu.camera_matrices(cam["fov"], cam["yaw"], cam["pitch"], PRIMARY, PRIMARY)
    prim = cpu.primary_rays(cam["position"], 1e-4, m["inv_proj"], m["inv_view"], PRIMARY, PRIMARY)
    ph = cpu.trace_closest(bvh.nodes, bvh.tri_indices, bvh.woop, prim)
    rays = W.bounce_rays(mesh.positions(), prim, ph["tri"], ph["uv"], per_hit=8, seed=RAY_SEED)
    cores = cpu.hardware_threads()
    trace, kind, what = cpu_tracer(bvh)
    # bounded sample per step so the whole run ends within minutes on any host
    t0 = time.perf_counter()
    trace(rays[:262144])
    rate = 262144 / (time.perf_counter() - t0)
    budget = 90.0 / max(1, args.steps + args.warmup)
    n = int(min(rays.shape[0], max(262144, rate * min(budget, 15.0))))
    for _ in range(args.warmup):
        trace(rays[:n])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        trace(rays[:n])
    dt = (time.perf_counter() - t0) / args.steps
    v = n / dt / 1e6
    sample = f"first {n} of the {rays.shape[0]} rays per step, {cores} threads"
    cfg = workload_config(rays.shape[0], {"triangles": int(mesh.n_tris), "timing": "wall clock around the timed steps on the host cores (no GPU work in this arm)"})
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference's hot path is GLSL under OpenGL and cannot run headless; timed here on the host cores: " + what,
    }))


def run_native(args, rank, world, local_rank):
    import torch
    import adypt_b200 as A

    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    A.load_library()

    mesh, bvh = build_inputs(rank=rank, barrier=dist.barrier if dist is not None else None)
    scene = bvh.upload(local_rank)  # OglScene::Initialize: Woop rows are built on the GPU
    tracer = A.Tracer(scene, A.PTConfig.make(), PRIMARY, PRIMARY, bias_seed=7)
    cam = W.city_camera(CELLS)
    tracer.look(cam["position"], cam["yaw"], cam["pitch"], cam["fov"])
    prim = tracer.primary_rays()
    ph = scene.trace_closest(prim)
    # weak scaling: every rank gets its own batch of the same size (different hash seed), no collective
    rays = W.bounce_rays(mesh.positions(), prim, ph["tri"], ph["uv"], per_hit=8, seed=RAY_SEED + rank)
    n = rays.shape[0]
    log(f"rank {rank}: {mesh.n_tris} tris, {bvh.nodes.shape[0]} nodes, {bvh.tri_indices.shape[0]} refs, {n} rays")

    d_rays = torch.from_numpy(rays).to(dev)
    d_tri = torch.empty(n, dtype=torch.int32, device=dev)
    d_t = torch.empty(n, dtype=torch.float32, device=dev)
    d_uv = torch.empty((n, 2), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def step():
        scene.trace_closest(d_rays, d_tri, d_t, d_uv, stream=stream.cuda_stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        flush.fill_(1)
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    l0 = A.launch_count()
    evs = []
    for _ in range(args.steps):
        flush.fill_(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step()
        e1.record(stream)
        evs.append((e0, e1))
    barrier()
    launches = A.launch_count() - l0
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(total_ms.item()) / args.steps
    value = world * n / (ms_per_step * 1e-3) / 1e6

    # ---- end to end: host buffers through the C-ABI, H2D + D2H inside the timed region
    h_rays = torch.from_numpy(rays).pin_memory()
    h_tri = torch.empty(n, dtype=torch.int32).pin_memory()
    h_t = torch.empty(n, dtype=torch.float32).pin_memory()
    h_uv = torch.empty((n, 2), dtype=torch.float32).pin_memory()
    for _ in range(2):
        scene.trace_closest(h_rays, h_tri, h_t, h_uv, stream=stream.cuda_stream)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        scene.trace_closest(h_rays, h_tri, h_t, h_uv, stream=stream.cuda_stream)
    barrier()
    e2e_s = torch.tensor([(time.perf_counter() - t0) / args.steps], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    clocks = sampler.stop()
    assert torch.equal(h_tri.to(dev), d_tri), "host-path results differ from device-path results"
    e2e = {"value": world * n / float(e2e_s.item()) / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(n * 32), "d2h_bytes_per_step": int(n * 16),
           "how": "adypt_trace_closest(ADYPT_MEM_HOST) on pinned host arrays (1M-ray chunks pipelined over 3 streams), wall clock around K blocking calls"}

    # ---- roofline of the dominant (only) kernel in the step
    st = scene.trace_stats(d_rays)
    bpr = bytes_per_ray(st["nodes"], st["tris"], st["hits"], n)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    kernel_ms = float(np.mean(step_ms))
    achieved = bpr * n / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "trace_closest_c2_dram.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": "adypt::trace_kernel<false,false>", "bytes_per_ray": bpr, "nodes_per_ray": st["nodes"] / n,
                "tris_per_ray": st["tris"] / n, "hit_fraction": st["hits"] / n, "kernel_ms": kernel_ms, "peak_source": peak_src,
                "note": "algorithmic bytes (80 B/node + 48 B/Woop + 4 B/hit + 48 B ray io); the 66 MB BVH is L2-resident, so DRAM traffic is far below this"}

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "config": workload_config(n, {"triangles": int(mesh.n_tris), "parallelism": f"{world} independent ray batches (one per GPU), no collective"}),
           "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks,
           "target": {"Mrays/s": 1500.0, "met": value / world >= 1500.0}}
    if not args.no_aux:
        out["aux"] = path_tracer_aux(A, torch, dist, rank, local_rank, world)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb, counters, ns = cpu_leg(bvh, rays)
        out["cpu_baseline"] = cb
        out["roofline"]["oracle_nodes_per_ray"] = counters["nodes"] / ns
        out["roofline"]["oracle_tris_per_ray"] = counters["tris"] / ns
    if rank == 0:
        emit(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def path_tracer_aux(A, torch, dist, rank, local_rank, world):
    """Second half of BASELINE.json's metric: 1080p path samples/s (configs[2], "C3"): the same city with
    glossy / mirror / glass / emissive boxes, 1920x1080, maxBounce 5 (4 bounces), 64 spp, no Russian roulette
    (the reference has none). Each rank renders the full image (weak scaling); wall clock around sample()+sync."""
    mesh, hs = build_inputs(mixed=True, rank=rank, barrier=dist.barrier if dist is not None else None)
    scene = hs.upload(local_rank)
    w, h, spp = 1920, 1080, 64
    tr = A.Tracer(scene, A.PTConfig.make(sun=(1.0, 1.0, 1.0)), w, h, bias_seed=7)
    cam = W.city_camera(CELLS)
    tr.look(cam["position"], cam["yaw"], cam["pitch"], cam["fov"])
    tr.sample(16)
    tr.sync()
    tr.primary(0)
    s0 = tr.stats()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    tr.sample(spp)
    tr.sync()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=torch.device("cuda", local_rank))
    if dist is not None:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dt = float(dt.item())
    s1 = tr.stats()
    seg = s1["segments"] - s0["segments"]
    out = {"workload": "C3: 1920x1080, 64 spp, maxBounce 5, mixed-material 1M-tri city, wavefront path tracer; no Russian roulette (the reference's pathtracer.glsl has none)",
           "path_samples_per_s": world * w * h * spp / dt, "path_segments_per_s": world * seg / dt, "segments_per_sample": seg / (w * h * spp),
           "seconds": dt, "gpu_launches": s1["launches"] - s0["launches"], "scaling": "weak"}
    # BASELINE.json words configs[2] "4 bounces + Russian roulette": the same render with the opt-in roulette from bounce 1
    tr.set_russian_roulette(1)
    tr.primary(0)
    tr.sample(16)
    tr.sync()
    tr.primary(0)
    s0 = tr.stats()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    tr.sample(spp)
    tr.sync()
    dt2 = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=torch.device("cuda", local_rank))
    if dist is not None:
        dist.all_reduce(dt2, op=dist.ReduceOp.MAX)
    dt2 = float(dt2.item())
    seg2 = tr.stats()["segments"] - s0["segments"]
    out["with_russian_roulette"] = {"start_bounce": 1, "path_samples_per_s": world * w * h * spp / dt2, "segments_per_sample": seg2 / (w * h * spp),
                                    "seconds": dt2, "note": "opt-in extension (adypt_tracer_set_russian_roulette); unbiased, not the reference's image"}
    return out


_RESULT_FD = None


def claim_stdout():
    """stdout carries exactly ONE JSON line. Native libraries write there too (under torchrun NCCL prints its version
    banner on fd 1, scene loading prints [SCENE] lines), so fd 1 is pointed at stderr for the whole run and the result
    line is written to the original stdout at the end."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line: str):
    sys.stdout.flush()
    os.write(_RESULT_FD if _RESULT_FD is not None else 1, (line + "\n").encode())


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-aux", action="store_true", help="skip the 1080p path-tracing measurement")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_native(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
