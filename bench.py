#!/usr/bin/env python
"""bench.py -- path samples/s of the wavefront PathTracer on BASELINE.json's config 3 (full Disney BSDF, synthetic
~10M-triangle scene, 1920x1080, depth 16), sample passes sharded over N GPUs.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N ...            # the reference's CPU path (tier B) on the host cores

A "step" is one pass of --spp-per-step samples per pixel over the whole 1080p frame on every GPU (weak scaling: per-GPU
work is fixed, rank r renders its own sample indices).  `value` = paths/s with the scene, BVH and film resident in HBM
(the K steps are only ENQUEUED by the host -- selas_cuda_render_pathtracer_device with stats = NULL does not wait --, CUDA
events on the render stream, max over ranks, one NCCL reduce of the film inside the timed region when N > 1).
`e2e` = the same metric through the public C ABI with the result on the host every step: N = 1, selas_cuda_render_pathtracer
(pinned host film, device->host copy inside the call); N > 1, render_pathtracer_device + selas_cuda_reduce_film + the
device->host copy of the reduced film on rank 0, every step.  `job` = the whole configs[2] render (256 spp split over the N
GPUs, the reduce, FrameBuffer_Scale, one device->host copy of the film): strong scaling.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "path_samples_per_s_1080p"
UNIT = "samples/s"
WIDTH, HEIGHT, MAX_DEPTH, RR_START = 1920, 1080, 16, 8
PARITY_FIRST_SAMPLE = 7777  # sample index of the pass both arms render for the 1080p image comparison


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="selas_b200", choices=["selas_b200", "reference"])
    ap.add_argument("--spp-per-step", type=int, default=4)
    ap.add_argument("--grid", type=int, default=2236, help="heightfield cells per side (2236 -> ~10M triangles)")
    ap.add_argument("--cpu-stride", type=int, default=1, help="pixel stride of the bounded CPU sample (1 = whole frame)")
    ap.add_argument("--cpu-spp", type=int, default=1, help="samples per pixel of one CPU step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-raycast", action="store_true", help="skip the config-2 Mrays/s side measurement")
    ap.add_argument("--no-job", action="store_true", help="skip the whole-job (256 spp) strong-scaling measurement")
    ap.add_argument("--job-spp", type=int, default=256, help="samples per pixel of the whole job (BASELINE configs[2]: 256)")
    return ap.parse_args()


def workload_config(args, n_gpus, tris):
    return {
        "workload": "BASELINE configs[2]: full Disney BSDF (metallic/clearcoat/sheen/specTrans/thin), synthetic heightfield+spheres, "
                    f"{tris} triangles, 1920x1080, depth 16, RR after 8, 4 quad lights, constant background",
        "triangles": int(tris), "width": WIDTH, "height": HEIGHT, "max_depth": MAX_DEPTH,
        "spp_per_step_per_gpu": args.spp_per_step, "paths_per_step": WIDTH * HEIGHT * args.spp_per_step * n_gpus,
        "sharding": f"sample passes over {n_gpus} GPU(s), scene replicated, one ncclReduce of the film (selas_cuda_reduce_film)",
        "mis": "reference (weight 1.0, PathTracer.cpp:215)",
        "l2": "no explicit flush: BVH+triangles (~1 GB) and the wavefront queues (~1 GB) exceed the 126 MB L2",
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
                power.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def captured_traffic_per_ray():
    """DRAM bytes per closest-hit ray from the committed `ncu --set full` capture of EVERY traceKernel<0> launch of one 4-spp step of this
    workload on the final tree (profiles/r2_trace_kernel_metrics.csv, summed in profiles/r2_trace_dram_per_step.json):
    (dram__bytes_read.sum + dram__bytes_write.sum over the 16 closest-hit launches) / (closest-hit rays of that step)."""
    p = ROOT / "profiles" / "r2_trace_dram_per_step.json"
    try:
        d = json.loads(p.read_text())
        return float(d["closest"]["dram"]) / float(d["closest"]["rays"]), str(p.relative_to(ROOT))
    except Exception:
        return None, None


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_cpu_arm(data, args, steps, warmup, which="auto"):
    """The reference's CPU path for this workload.  "reference": the reference's OWN integrator loop -- PathTracer.cpp's EvaluatePath /
    RayPick / OcclusionRay, SurfaceParameters.cpp, Framebuffer.cpp and the Shading library compiled unmodified into oracle/_ref behind
    the Embree-API shim (oracle/ref_pathtracer.cpp); "port": the plain-C++ restatement.  Ray casting by the oracle's scalar SAH BVH in
    both (Embree is absent here and on the GPU box: BASELINE.md tier B, profiles/r2_gpu_box_probe.txt).  The camera is made by the
    reference's InitializeRayCastCamera (ref_camera_init), not by the product.  Bounded sample: cpu-spp samples of every
    cpu-stride-th pixel per step."""
    from selas_b200 import capi, scenes
    from tests import oracle_binding

    lib, kind = None, "port"
    if which in ("auto", "reference"):
        lib = oracle_binding.load_ref()
        kind = "reference"
    if lib is None:
        lib, kind = oracle_binding.load_oracle(), "port"
    threads = cpu_threads()
    c = scenes.HEIGHTFIELD_CAMERA
    camera = lib.camera_init(c["position"], c["look_at"], c["up"], c["fov"], WIDTH, HEIGHT)
    t0 = time.perf_counter()
    h = lib.scene_create(data)
    build_s = time.perf_counter() - t0
    rs = capi.render_settings(WIDTH, HEIGHT, args.cpu_spp, MAX_DEPTH, RR_START, layers=1, seed=1, pixel_stride=args.cpu_stride)
    render = lib.render_reference_loop if kind == "reference" else lib.render
    times, paths, rays = [], 0, 0
    last_film = None
    for i in range(warmup + steps):
        rs.firstSample = PARITY_FIRST_SAMPLE if i == warmup + steps - 1 else i * args.cpu_spp
        t0 = time.perf_counter()
        last_film, st = render(h, camera, rs, nthreads=threads)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
            paths += st.pathSamples
            rays += st.closestRays + st.occlusionRays
    lib.scene_destroy(h)
    total = sum(times)
    return {"value": paths / total, "unit": UNIT, "cores": threads, "kind": kind, "last_film": last_film[0],
            "sample": f"pixel stride {args.cpu_stride} over the 1080p frame, {args.cpu_spp} spp per step, {steps} steps after {warmup} warm-up "
                      f"({paths} paths, {total:.1f} s; CPU BVH build {build_s:.1f} s excluded)",
            "integrator": "PathTracer.cpp EvaluatePath + SurfaceParameters.cpp + Framebuffer.cpp + Shading/*.cpp, unmodified (oracle/_ref)"
                          if kind == "reference" else "oracle restatement (oracle/selas_oracle.cpp)",
            "mrays_per_s": rays / total / 1e6, "ms_per_step": 1e3 * total / max(1, len(times)), "paths_per_step": paths // max(1, len(times)),
            "ray_cast": "oracle scalar SAH-BVH2 behind the Embree-API shim (Embree 3.2.0 absent: headers only in the reference tree, "
                        "no libembree on the GPU box)"}


_REAL_STDOUT = None


def protect_stdout():
    """stdout must carry exactly ONE JSON line: libraries (NCCL's version banner, for one) also write to fd 1, so everything is
    sent to stderr and the JSON line goes to a private duplicate of the original stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    args = parse_args()
    protect_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = args.gpus

    from selas_b200 import capi, scenes

    if args.impl == "reference":
        if rank != 0:
            return 0
        t0 = time.perf_counter()
        data = scenes.disney_heightfield(grid=args.grid)
        log(f"[reference arm] scene generated in {time.perf_counter() - t0:.1f} s")
        steps, warmup = max(1, args.steps), max(0, args.warmup)
        r = run_cpu_arm(data, args, steps, warmup)
        cfg = workload_config(args, n_gpus, data.triangle_count())
        # this arm's own step: the same workload, a bounded sample of it per step, on the host cores (no GPU, no sharding)
        cfg.update({"spp_per_step_per_gpu": args.cpu_spp, "paths_per_step": r["paths_per_step"], "sharding": "none: one process, all host threads",
                    "sample": r["sample"]})
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": n_gpus, "steps": steps, "warmup": warmup,
                "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": cfg,
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "integrator", "ray_cast")},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return 0

    import torch
    import torch.distributed as dist

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    t0 = time.perf_counter()
    data = scenes.disney_heightfield(grid=args.grid)
    camera = scenes.heightfield_camera(WIDTH, HEIGHT)
    tris = data.triangle_count()
    if rank == 0:
        log(f"scene generated: {tris} triangles in {time.perf_counter() - t0:.1f} s")
    ctx = capi.Context(local_rank)  # raises if libselas_cuda.so or the GPU is missing: no fallback
    t0 = time.perf_counter()
    scene = capi.Scene(ctx, data)
    upload_s = time.perf_counter() - t0
    bvh = scene.bvh_stats()
    if rank == 0:
        log(f"scene upload + BVH8 build: {upload_s:.2f} s (device build {bvh.buildMs:.1f} ms, {bvh.nodeCount} nodes, "
            f"{bvh.deviceBytes / 1e9:.2f} GB, {bvh.sahClusters} SAH clusters)")

    S = args.spp_per_step
    pixels = WIDTH * HEIGHT
    film = torch.zeros(pixels * 3, dtype=torch.float32, device=dev)
    stream = torch.cuda.Stream(dev)  # explicit stream: the library launches on it and the CUDA events below are recorded on it
    torch.cuda.set_stream(stream)

    def settings(first_sample, spp=S, flags=0):
        return capi.render_settings(WIDTH, HEIGHT, spp, MAX_DEPTH, RR_START, layers=1, seed=1, first_sample=first_sample, flags=flags)

    def step_settings(step_index, flags=0):
        # rank r renders sample indices [(step*world + r) * S, +S): disjoint streams across ranks and steps
        return settings((step_index * world + rank) * S, flags=flags)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def enqueue(rs):  # selas_cuda_render_pathtracer_device with stats = NULL: returns while the GPU works
        scene.render_device(camera, rs, film.data_ptr(), stream.cuda_stream, want_stats=False)

    reducer = None
    if world > 1:
        from selas_b200 import distributed
        reducer = distributed.FilmReducer(ctx, rank, world)  # raw ncclComm_t for selas_cuda_reduce_film

    # ---- device-resident arm: value ---------------------------------------------------------------------------
    for i in range(args.warmup):
        enqueue(step_settings(i))
    if reducer is not None:  # warm the communicator up (first collective sets up the NVLink channels)
        reducer.reduce(film, dst=0, stream=stream.cuda_stream)
    scene.collect_stats()
    film.zero_()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    host_enqueue_ms = None
    for i in range(args.steps):
        t_host = time.perf_counter()
        enqueue(step_settings(args.warmup + i))
        if i == 0:  # host cost of enqueueing one step into an empty queue (later steps block on the driver's launch-queue depth)
            host_enqueue_ms = (time.perf_counter() - t_host) * 1e3
    if world > 1:  # the path's only collective: one ncclReduce of the film, through the C ABI
        reducer.reduce(film, dst=0, stream=stream.cuda_stream)
    e1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    st = scene.collect_stats()  # per-kernel CUDA-event spans and ray totals of exactly the timed steps
    agg = {"closestMs": st.closestMs, "shadeMs": st.shadeMs, "shadowMs": st.shadowMs, "generateMs": st.generateMs,
           "closestRays": int(st.closestRays), "occlusionRays": int(st.occlusionRays), "launches": int(st.kernelLaunches),
           "closestLaunches": int(st.closestLaunches)}
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    paths_total = pixels * S * args.steps * world
    value = paths_total / (total_ms * 1e-3)
    # FrameBuffer_Scale (Framebuffer.cpp:62-70) on the reduced film, outside the timed region like the reference's save path
    mean_radiance, nonfinite_values = None, None
    if rank == 0:
        ctx.check(ctx.lib.selas_cuda_film_scale_device(ctx.handle, capi.C.c_void_p(film.data_ptr()), capi.C.c_size_t(film.numel()),
                                                       capi.f32(1.0 / (S * args.steps * world)), capi.C.c_void_p(stream.cuda_stream)))
        torch.cuda.synchronize(dev)
        # the reference adds NaN/Inf samples to the film like any other (no checks in release builds, BasicTypes.h:28), and so does
        # the default mode of the CUDA path: a handful of pixels may be non-finite after many samples
        finite = torch.isfinite(film)
        nonfinite_values = int((~finite).sum().item())
        mean_radiance = float(film[finite].mean().item())
        assert mean_radiance > 0.0 and nonfinite_values < film.numel() // 1000, "film is empty or mostly not finite"

    # ---- end-to-end arm: the public C ABI with the step's result on the host, every step ------------------------------------------
    host_film = torch.zeros(pixels * 3, dtype=torch.float32).pin_memory()
    host_np = host_film.numpy()
    ptrs = (capi.C.c_void_p * 1)(host_np.ctypes.data)
    e2e_steps = max(1, min(args.steps, 5))
    stats = capi.RenderStats()

    def e2e_step(i):
        rs = step_settings(1000 + i)
        if world == 1:  # PathTracer::GenerateImage's shape: host film out
            ctx.check(ctx.lib.selas_cuda_render_pathtracer(scene.handle, capi.C.byref(camera), capi.C.byref(rs), ptrs, None))
        else:  # N GPUs: render into the device film, reduce onto rank 0, rank 0 brings the reduced film to the host
            film.zero_()
            enqueue(rs)
            reducer.reduce(film, dst=0, stream=stream.cuda_stream)
            if rank == 0:
                host_film.copy_(film, non_blocking=True)
            stream.synchronize()
        return float(host_np[::4099].sum()) if rank == 0 else 0.0  # the step's result is read on the host

    e2e_step(-1)
    barrier()
    t0 = time.perf_counter()
    e2e_each = []
    for i in range(e2e_steps):
        t1 = time.perf_counter()
        e2e_step(i)
        e2e_each.append((time.perf_counter() - t1) * 1e3)
    barrier()
    if rank == 0:
        log("e2e per-step ms:", [round(x, 2) for x in e2e_each])
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = pixels * S * e2e_steps * world / float(e2e_s.item())
    scene.collect_stats()

    # ---- the whole configs[2] job: 256 spp over the N GPUs, reduce, scale, film to the host (strong scaling) ----------------------------
    job = None
    if not args.no_job:
        from selas_b200 import distributed as D
        JOB_SPP = args.job_spp
        first, count = D.shard_samples(JOB_SPP, rank, world)
        film.zero_()
        barrier()
        t0 = time.perf_counter()
        j0, j1, j2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        j0.record(stream)
        done = 0
        while done < count:  # whole sample passes of S spp, the library cuts them into <= 8 M paths in flight
            k = min(S, count - done)
            enqueue(settings(100000 + first + done, spp=k))
            done += k
        j1.record(stream)
        if world > 1:
            reducer.reduce(film, dst=0, stream=stream.cuda_stream)
        j2.record(stream)
        if rank == 0:
            ctx.check(ctx.lib.selas_cuda_film_scale_device(ctx.handle, capi.C.c_void_p(film.data_ptr()), capi.C.c_size_t(film.numel()),
                                                           capi.f32(1.0 / JOB_SPP), capi.C.c_void_p(stream.cuda_stream)))
            host_film.copy_(film, non_blocking=True)
        stream.synchronize()
        barrier()
        wall = torch.tensor([time.perf_counter() - t0, j0.elapsed_time(j1) * 1e-3, j1.elapsed_time(j2) * 1e-3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(wall, op=dist.ReduceOp.MAX)
        jst = scene.collect_stats()
        if rank == 0:
            fin = np.isfinite(host_np)
            job = {"what": f"BASELINE configs[2] whole job: {JOB_SPP} spp at 1920x1080 split over {world} GPU(s) ({count} spp on rank 0), one "
                           "ncclReduce of the film, FrameBuffer_Scale, one device->host copy of the film on rank 0",
                   "spp": JOB_SPP, "paths": pixels * JOB_SPP, "seconds": float(wall[0].item()), "samples_per_s": pixels * JOB_SPP / float(wall[0].item()),
                   "render_seconds_max_rank": float(wall[1].item()), "reduce_ms": float(wall[2].item()) * 1e3 if world > 1 else 0.0,
                   "scaling": "strong", "mean_radiance": float(host_np[fin].mean()), "nonfinite_film_values": int((~fin).sum()),
                   "rank0_closest_rays": int(jst.closestRays), "rank0_occlusion_rays": int(jst.occlusionRays)}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (closest-hit traversal) -------------------------------------------------------
    cst = scene.render_device(camera, step_settings(2000, flags=capi.RENDER_COUNT_TRAVERSAL), film.data_ptr(), stream.cuda_stream)
    n_node = cst.closestNodes / max(1, cst.closestRays)
    n_tri = cst.closestTris / max(1, cst.closestRays)
    s_node = cst.shadowNodes / max(1, cst.occlusionRays)
    s_tri = cst.shadowTris / max(1, cst.occlusionRays)
    bytes_per_ray = 64.0 + 80.0 * n_node + 48.0 * n_tri        # 32 in + 32 out + 80/node + 48/triangle
    shadow_bytes_per_ray = 36.0 + 80.0 * s_node + 48.0 * s_tri
    peak, peak_src = measured_peaks()
    achieved = bytes_per_ray * agg["closestRays"] / (agg["closestMs"] * 1e-3) / 1e9 if agg["closestMs"] > 0 else 0.0
    kernel_ms = agg["closestMs"] + agg["shadeMs"] + agg["shadowMs"] + agg["generateMs"]
    rays_per_launch = agg["closestRays"] / max(1, agg["closestLaunches"])
    dram_per_ray, traffic_src = captured_traffic_per_ray()
    roofline = {"bound": "hbm", "kernel": "BVH8 closest hit: traceKernel<0> for the camera rays, traceCoopKernel<0> (warp-cooperative triangle stage) from the first bounce on", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "frac_of_nominal_8000": achieved / 8000.0, "peak_source": peak_src,
                # bytes per launch: DRAM bytes per ray of the committed ncu capture (all launches of a step) x this run's average rays per launch
                "traffic": dram_per_ray * rays_per_launch if dram_per_ray else None,
                "traffic_source": f"{traffic_src}: {dram_per_ray:.1f} DRAM B/ray (algorithmic {bytes_per_ray:.0f} B/ray: the tree top is served by L1/L2)"
                if dram_per_ray else None,
                "algorithmic_bytes_per_launch": bytes_per_ray * rays_per_launch,
                "algorithmic_bytes_per_ray": bytes_per_ray, "nodes_per_ray": n_node, "tris_per_ray": n_tri,
                "rays_per_launch": agg["closestRays"] / max(1, agg["closestLaunches"]),
                "avg_launch_ms": agg["closestMs"] / max(1, agg["closestLaunches"]),
                "kernel_share_of_step": {"closest": agg["closestMs"] / kernel_ms, "shade": agg["shadeMs"] / kernel_ms,
                                         "shadow": agg["shadowMs"] / kernel_ms, "generate": agg["generateMs"] / kernel_ms},
                "kernel_ms_per_step": {"closest": agg["closestMs"] / args.steps, "shade_incl_material_sort": agg["shadeMs"] / args.steps,
                                       "shadow": agg["shadowMs"] / args.steps, "generate": agg["generateMs"] / args.steps,
                                       "step": total_ms / args.steps},
                "shadow": {"algorithmic_bytes_per_ray": shadow_bytes_per_ray, "nodes_per_ray": s_node, "tris_per_ray": s_tri,
                           "achieved": shadow_bytes_per_ray * agg["occlusionRays"] / (agg["shadowMs"] * 1e-3) / 1e9 if agg["shadowMs"] > 0 else 0.0,
                           "frac": (shadow_bytes_per_ray * agg["occlusionRays"] / (agg["shadowMs"] * 1e-3) / 1e9 / peak) if agg["shadowMs"] > 0 else 0.0},
                "shade": {"algorithmic_bytes_per_vertex": 388.0,
                          "achieved": 388.0 * agg["closestRays"] / (agg["shadeMs"] * 1e-3) / 1e9 if agg["shadeMs"] > 0 else 0.0,
                          "frac": (388.0 * agg["closestRays"] / (agg["shadeMs"] * 1e-3) / 1e9 / peak) if agg["shadeMs"] > 0 else 0.0}}
    mrays = {"closest": agg["closestRays"] / (agg["closestMs"] * 1e-3) / 1e6 if agg["closestMs"] > 0 else 0.0,
             "occlusion": agg["occlusionRays"] / (agg["shadowMs"] * 1e-3) / 1e6 if agg["shadowMs"] > 0 else 0.0,
             "whole_step": (agg["closestRays"] + agg["occlusionRays"]) * world / (total_ms * 1e-3) / 1e6}

    raycast = None
    if not args.no_raycast:
        try:
            raycast = raycast_microbench(ctx, torch, dev, stream)
        except Exception as ex:  # side measurement only: never hide the headline
            raycast = {"error": str(ex)}

    cpu, parity = None, None
    if world == 1 and not args.no_cpu_baseline:
        # the pass the CPU arm renders last, through the CUDA path (same seed, same per-path streams), for the 1080p image comparison
        prs = capi.render_settings(WIDTH, HEIGHT, args.cpu_spp, MAX_DEPTH, RR_START, layers=1, seed=1, first_sample=PARITY_FIRST_SAMPLE,
                                   pixel_stride=args.cpu_stride)
        gfilm, _ = scene.render(camera, prs)
        scene.close()
        r = run_cpu_arm(data, args, steps=6, warmup=1)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "integrator", "ray_cast", "mrays_per_s")}
        parity = image_parity(gfilm[0], r["last_film"], f"configs[2] at 1920x1080, {args.cpu_spp} spp, depth 16: CUDA wavefront vs the CPU arm "
                              f"({r['kind']}: {r['integrator']}) on the same per-path RNG streams")

    if world == 1:
        e2e_bytes = {"h2d_bytes_per_step": capi.C.sizeof(capi.Camera) + capi.C.sizeof(capi.RenderSettings), "d2h_bytes_per_step": pixels * 3 * 4,
                     "api": "selas_cuda_render_pathtracer (pinned host film, scene resident as in the reference's GenerateImage)"}
    else:
        e2e_bytes = {"h2d_bytes_per_step": (capi.C.sizeof(capi.Camera) + capi.C.sizeof(capi.RenderSettings)) * world, "d2h_bytes_per_step": pixels * 3 * 4,
                     "nccl_reduce_bytes_per_step_per_rank": pixels * 3 * 4,
                     "api": "per rank selas_cuda_render_pathtracer_device + selas_cuda_reduce_film every step; rank 0 copies the reduced film to pinned host memory"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args, world, tris), "clocks": clocks,
            "e2e": dict({"value": e2e_value, "unit": UNIT, "steps": e2e_steps}, **e2e_bytes),
            "gpu_launches": int(agg["launches"]), "host_enqueue_ms_per_step": host_enqueue_ms,
            "roofline": roofline, "cpu_baseline": cpu, "parity_1080p": parity, "job": job, "mrays_per_s": mrays, "raycast_config2": raycast,
            "bvh": {"builder": "binned SAH over Morton clusters" if bvh.sahClusters else "LBVH", "build_ms": bvh.buildMs, "nodes": bvh.nodeCount,
                    "triangles": bvh.triangleCount, "device_gb": bvh.deviceBytes / 1e9, "sah_clusters": bvh.sahClusters, "levels": bvh.maxDepth,
                    "upload_plus_build_s": upload_s},
            "mean_radiance": mean_radiance, "nonfinite_film_values": nonfinite_values}
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def image_parity(g, o, what):
    """Same-stream image comparison (SURVEY 8(d)(v) metric, trimmed: the reference estimator has fireflies of 1e4 and a diverged path on
    one of them would decide a plain mean): fraction of pixels within 1e-3, relMSE over the 99 % best pixels, ratio of the means."""
    g, o = np.asarray(g, np.float64), np.asarray(o, np.float64)
    fin = np.isfinite(g).all(axis=2) & np.isfinite(o).all(axis=2)
    close = np.isclose(g, o, rtol=1e-3, atol=1e-4).all(axis=2)
    with np.errstate(all="ignore"):
        err = np.sort(((g - o) ** 2 / (o ** 2 + 1e-2)).mean(axis=2)[fin])
    return {"what": what, "pixels": int(fin.size), "finite_both": float(fin.mean()),
            "same_nonfinite_pixels": bool(np.array_equal(np.isfinite(g).all(axis=2), np.isfinite(o).all(axis=2))),
            "pixels_within_1e-3": float(close.mean()), "relmse_99pct": float(err[: int(len(err) * 0.99)].mean()), "relmse_all": float(err.mean()),
            "mean_ratio": float(g[fin].mean() / o[fin].mean())}


def raycast_microbench(ctx, torch, dev, stream):
    """BASELINE configs[1]: 1M-triangle displaced sphere, 1920x1080 primary rays + one shadow ray per hit toward (5,5,5)
    + incoherent cosine bounce rays; Mrays/s of the device-resident entry points (CUDA events, median of 10)."""
    from selas_b200 import capi, scenes

    data = scenes.raycast_scene()
    sc = capi.Scene(ctx, data)
    cam = scenes.raycast_camera(WIDTH, HEIGHT)
    rays = scenes.primary_rays(cam, WIDTH, HEIGHT)
    hits, ng = sc.intersect(rays, want_ng=True)
    sh = scenes.shadow_rays_toward(rays, hits, ng, (5.0, 5.0, 5.0))
    inc = scenes.cosine_bounce_rays(rays, hits, ng)
    out = {"triangles": data.triangle_count(), "bvh_build_ms": sc.bvh_stats().buildMs, "primary_hit_fraction": float((hits["geomID"] != capi.SELAS_INVALID_ID).mean())}
    peak, _ = measured_peaks()

    def to_dev(a):
        return torch.from_numpy(a.view(np.uint8).reshape(-1)).to(dev)

    for name, r, anyhit in (("primary_closest", rays, False), ("shadow_occlusion", sh, True), ("incoherent_closest", inc, False)):
        d_r = to_dev(r)
        n = r.size
        d_h = torch.empty(n * 32, dtype=torch.uint8, device=dev)
        d_o = torch.empty(n, dtype=torch.uint8, device=dev)
        ts = []
        for i in range(13):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            if anyhit:
                sc.occluded_device(n, d_r.data_ptr(), d_o.data_ptr(), stream.cuda_stream)
            else:
                sc.intersect_device(n, d_r.data_ptr(), d_h.data_ptr(), 0, stream.cuda_stream)
            b.record(stream)
            torch.cuda.synchronize(dev)
            if i >= 3:
                ts.append(a.elapsed_time(b))
        ms = float(np.median(ts))
        _, counts = sc.intersect_counted(r, any_hit=anyhit)
        nn, nt = float(counts[:, 0].mean()), float(counts[:, 1].mean())
        bpr = (36.0 if anyhit else 64.0) + 80.0 * nn + 48.0 * nt
        out[name] = {"rays": int(n), "ms": ms, "mrays_per_s": n / ms / 1e3, "nodes_per_ray": nn, "tris_per_ray": nt,
                     "algorithmic_bytes_per_ray": bpr, "achieved_gbs": bpr * n / (ms * 1e-3) / 1e9, "frac_of_measured_hbm": bpr * n / (ms * 1e-3) / 1e9 / peak}
    sc.close()
    return out


if __name__ == "__main__":
    sys.exit(main())
