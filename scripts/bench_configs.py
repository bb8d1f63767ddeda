#!/usr/bin/env python
"""BASELINE.json configs other than the one bench.py measures (configs[2]):
  C1  Cornell box (Disney diffuse + copper), 512x512, 64 spp, depth 8: GPU vs the reference CPU path, plus image parity
  C4  Ptex-textured quad grid (default 1581^2 = 2,499,561 quads ~ 5M triangles, 8x8 uint8 faces written with the reference's
      PtexWriter), 1920x1080, depth 8, Ptex lookups inside the shading loop
  C5  VCM (light paths + vertex connection + hash-grid merging) on the caustic scene (floor + glass sphere + box under a sun-and-sky
      IBL), 1920x1080, vmCount = width*height light paths per iteration, path length <= 10: GPU iterations/s and paths/s with the
      per-stage device times, against the CPU restatement of VCM.cpp on a reduced frame (VCM.cpp is dead code in the reference, so
      there is no reference binary to time; kind = "port")
One JSON document on stdout.  (C2 and the C5 hash-grid microbench are in scripts/bench_rows.py.)
"""
import argparse
import json
import os
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from selas_b200 import capi, scenes  # noqa: E402
from tests import oracle_binding  # noqa: E402


def relmse(a, b):
    ok = np.isfinite(a).all(axis=2) & np.isfinite(b).all(axis=2)  # the reference adds NaN/Inf samples to the film like any other
    return float(np.mean((a[ok] - b[ok]) ** 2 / (b[ok] ** 2 + 1e-2)))


def config1(ctx):
    data = scenes.cornell_box(copper_roughness=0.01)  # Copper.json as shipped
    cam = scenes.cornell_camera(512, 512)
    sc = capi.Scene(ctx, data)
    rs = capi.render_settings(512, 512, 64, max_path_length=8, rr_start=8, layers=1, seed=1)
    sc.render(cam, rs)
    t0 = time.perf_counter()
    film, st = sc.render(cam, rs)
    e2e_s = time.perf_counter() - t0
    g = film[0] / 64.0
    ref = oracle_binding.load_ref()
    lib, kind = (ref, "reference") if ref is not None else (oracle_binding.load_oracle(), "port")
    threads = len(os.sched_getaffinity(0))
    hs = lib.scene_create(data)
    t0 = time.perf_counter()
    cfilm, cst = lib.render(hs, cam, rs, nthreads=threads)
    cpu_s = time.perf_counter() - t0
    c = cfilm[0] / 64.0
    hi = capi.render_settings(512, 512, 1024, max_path_length=8, rr_start=8, layers=1, seed=99)
    ref_hi = sc.render(cam, hi)[0][0] / 1024.0
    lib.scene_destroy(hs)
    out = {"paths": 512 * 512 * 64, "gpu_device_ms": st.renderMs, "gpu_samples_per_s": st.pathSamples / (st.renderMs * 1e-3),
           "gpu_e2e_samples_per_s": st.pathSamples / e2e_s, "gpu_rays": int(st.closestRays + st.occlusionRays),
           "gpu_kernel_launches": int(st.kernelLaunches),
           "cpu": {"kind": kind, "cores": threads, "seconds": cpu_s, "samples_per_s": cst.pathSamples / cpu_s,
                   "ray_cast": "oracle scalar SAH-BVH2 (Embree absent)"},
           "image_parity": {"relMSE_gpu_vs_cpu_same_seed": relmse(g, c), "mean_ratio_gpu_over_cpu": float(g.mean() / c.mean()),
                            "relMSE_gpu_vs_1024spp": relmse(g, ref_hi), "relMSE_cpu_vs_1024spp": relmse(c, ref_hi),
                            "nonfinite_pixels_gpu_cpu_1024spp": [int((~np.isfinite(x).all(axis=2)).sum()) for x in (g, c, ref_hi)],
                            "fraction_pixels_within_1e-3": float(np.isclose(g, c, rtol=1e-3, atol=1e-4).all(axis=2).mean())}}
    sc.close()
    return out


def config4(ctx, grid, spp):
    ptexref = oracle_binding.load_ptexref()
    if ptexref is None:
        return {"skipped": "oracle/_ref/libptex_ref.so not present"}
    tmp = Path(tempfile.mkdtemp()) / "c4.ptx"
    t0 = time.perf_counter()
    ptexref.write_grid(tmp, grid, grid, 3, 3, 0, 0, 3, 4, True, 0)
    write_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    tex = capi.Ptex(ctx, tmp)
    load_s = time.perf_counter() - t0
    data = scenes.ptex_quad_grid(grid, grid, ptex=tex)
    sc = capi.Scene(ctx, data)
    cam = scenes.ptex_grid_camera(1920, 1080)
    rs = capi.render_settings(1920, 1080, 4, max_path_length=8, rr_start=8, layers=1, seed=1)
    sc.render(cam, rs)
    tot_ms, paths, shade_ms, rays = 0.0, 0, 0.0, 0
    passes = max(1, spp // 4)
    t0 = time.perf_counter()
    for p in range(passes):
        rs.firstSample = 4 * p
        film, st = sc.render(cam, rs)
        tot_ms += st.renderMs
        shade_ms += st.shadeMs
        paths += st.pathSamples
        rays += st.closestRays + st.occlusionRays
    e2e_s = time.perf_counter() - t0
    # CPU leg: the restated integrator with its Ptex hook bound to the reference's own PtexFilter (oracle/_ref/libptex_ref.so), one
    # 4 spp pass of the same frame on every host thread; the same pass on the GPU gives the image parity figure
    orc = oracle_binding.load_oracle()
    threads = len(os.sched_getaffinity(0))
    rs.firstSample = 0
    gfilm, _ = sc.render(cam, rs)
    t0 = time.perf_counter()
    hs = orc.scene_create(data)
    th = ptexref.open(tmp)
    orc.scene_set_ptex(hs, 0, ptexref, th)
    cpu_build_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    cfilm, cst = orc.render(hs, cam, rs, nthreads=threads)
    cpu_s = time.perf_counter() - t0
    orc.scene_destroy(hs)
    ptexref.close(th)
    cpu = {"kind": "port (integrator) + reference Ptex library (lookups)", "cores": threads, "sample": "one 4 spp pass of the 1920x1080 frame",
           "scene_build_s": cpu_build_s, "seconds": cpu_s, "samples_per_s": cst.pathSamples / cpu_s,
           "relMSE_gpu_vs_cpu_same_seed": relmse(gfilm[0] / 4.0, cfilm[0] / 4.0),
           "mean_ratio_gpu_over_cpu": float(np.nanmean(gfilm[0][np.isfinite(gfilm[0]).all(axis=2) & np.isfinite(cfilm[0]).all(axis=2)])
                                            / np.nanmean(cfilm[0][np.isfinite(gfilm[0]).all(axis=2) & np.isfinite(cfilm[0]).all(axis=2)]))}
    # the same scene without the Ptex lookup (constant base colour) isolates the cost of the lookup inside shading
    plain = scenes.ptex_quad_grid(grid, grid, ptex=None)
    sc2 = capi.Scene(ctx, plain)
    sc2.render(cam, rs)
    _, st2 = sc2.render(cam, rs)
    out = {"quads": grid * grid, "triangles": data.triangle_count(), "ptex_file_mb": tmp.stat().st_size / 1e6, "ptex_write_s": write_s,
           "ptex_load_s": load_s, "spp": passes * 4, "paths": paths, "gpu_device_ms": tot_ms, "gpu_samples_per_s": paths / (tot_ms * 1e-3),
           "gpu_e2e_samples_per_s": paths / e2e_s, "gpu_rays": int(rays), "shade_ms_per_4spp_pass": shade_ms / passes,
           "mean_radiance_finite_pixels": float(film[0][np.isfinite(film[0]).all(axis=2)].mean() / 4.0),
           "nonfinite_pixels_last_pass": int((~np.isfinite(film[0]).all(axis=2)).sum()),
           "without_ptex": {"gpu_device_ms_per_4spp_pass": st2.renderMs, "shade_ms_per_4spp_pass": st2.shadeMs,
                            "gpu_samples_per_s": st2.pathSamples / (st2.renderMs * 1e-3)},
           "cpu": cpu}
    sc.close()
    sc2.close()
    tex.close()
    tmp.unlink()
    return out


def config5(ctx, iterations, cpu_width):
    data = scenes.caustic_scene()
    sphere = scenes.bounding_sphere(data)
    w, h = 1920, 1080
    cam = scenes.caustic_camera(w, h)
    sc = capi.Scene(ctx, data)
    sc.render_vcm(cam, capi.vcm_settings(w, h, 1, sphere, seed=1))  # warm-up: allocations, first-launch costs
    t0 = time.perf_counter()
    film, st = sc.render_vcm(cam, capi.vcm_settings(w, h, iterations, sphere, seed=1, first_iteration=1))
    e2e_s = time.perf_counter() - t0
    paths = int(st.lightPaths + st.cameraPaths)
    out = {"triangles": data.triangle_count(), "iterations": iterations, "vmCount": w * h, "max_path_length": 10,
           "gpu_device_ms": st.renderMs, "gpu_ms_per_iteration": st.renderMs / iterations,
           "gpu_paths_per_s": paths / (st.renderMs * 1e-3), "gpu_e2e_paths_per_s": paths / e2e_s,
           "stage_ms_per_iteration": {"light_pass": st.lightPassMs / iterations, "grid_build": st.gridBuildMs / iterations,
                                      "camera_pass_total": st.cameraPassMs / iterations, "of_which_merge": st.mergeMs / iterations,
                                      "of_which_connect": st.connectMs / iterations},
           "per_iteration": {"light_vertices": st.lightVertices / iterations, "closest_rays": st.closestRays / iterations,
                             "occlusion_rays": st.occlusionRays / iterations, "merge_candidates": st.mergeCandidates / iterations,
                             "merges_accepted": st.mergesAccepted / iterations, "connections": st.connections / iterations},
           "merge_candidates_per_s": st.mergeCandidates / max(st.mergeMs * 1e-3, 1e-9),
           "connections_per_s": st.connections / max(st.connectMs * 1e-3, 1e-9),
           "kernel_launches": int(st.kernelLaunches), "last_radius": st.lastRadius,
           "mean_radiance": float(film[np.isfinite(film).all(axis=2)].mean() / iterations),
           "nonfinite_pixels": int((~np.isfinite(film).all(axis=2)).sum())}
    # CPU restatement on a reduced frame (same scene, same settings otherwise): paths/s scale with vmCount
    cw, ch = cpu_width, cpu_width * 9 // 16
    orc = oracle_binding.load_oracle()
    threads = len(os.sched_getaffinity(0))
    hs = orc.scene_create(data)
    ccam = scenes.caustic_camera(cw, ch)
    t0 = time.perf_counter()
    cfilm, cst = orc.render_vcm(hs, ccam, capi.vcm_settings(cw, ch, 1, sphere, seed=1, first_iteration=1), nthreads=threads)
    cpu_s = time.perf_counter() - t0
    orc.scene_destroy(hs)
    gfilm, _ = sc.render_vcm(ccam, capi.vcm_settings(cw, ch, 1, sphere, seed=1, first_iteration=1))
    ok = np.isfinite(gfilm).all(axis=2) & np.isfinite(cfilm).all(axis=2)
    out["cpu"] = {"kind": "port", "cores": threads, "frame": [cw, ch], "seconds": cpu_s,
                  "paths_per_s": float(cst.lightPaths + cst.cameraPaths) / cpu_s,
                  "note": "oracle/orc_vcm.h over the oracle's scalar SAH-BVH2; VCM.cpp is commented out in the reference"}
    out["image_parity_reduced_frame"] = {"mean_ratio_gpu_over_cpu": float(gfilm[ok].mean() / cfilm[ok].mean()),
                                         "relMSE_2x2_binned": relmse(gfilm.reshape(ch // 2, 2, cw // 2, 2, 3).mean(axis=(1, 3)),
                                                                     cfilm.reshape(ch // 2, 2, cw // 2, 2, 3).mean(axis=(1, 3)))}
    sc.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,4,5")
    ap.add_argument("--c5-iterations", type=int, default=8)
    ap.add_argument("--c5-cpu-width", type=int, default=1920)
    ap.add_argument("--c4-grid", type=int, default=1581)
    ap.add_argument("--c4-spp", type=int, default=128)
    args = ap.parse_args()
    ctx = capi.Context(0)
    doc = {}
    if "1" in args.configs.split(","):
        doc["config1_cornell_512x512_64spp_depth8"] = config1(ctx)
    if "4" in args.configs.split(","):
        doc["config4_ptex_1080p_depth8"] = config4(ctx, args.c4_grid, args.c4_spp)
    if "5" in args.configs.split(","):
        doc["config5_vcm_1080p"] = config5(ctx, args.c5_iterations, args.c5_cpu_width)
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main()
