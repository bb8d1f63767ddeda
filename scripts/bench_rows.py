#!/usr/bin/env python
"""Per-row measurements for SURVEY.md section 8: every hot-path kernel family on a B200 against its algorithmic-bytes roofline,
with the reference's CPU path timed beside it (same inputs, bounded samples).  One JSON document on stdout.

  python scripts/bench_rows.py [--ptex-grid 800] [--grid-points 8000000] > gpurun_out/rows.json

Rows: ray cast (BASELINE configs[1]: 1M-triangle displaced sphere, 1080p primary + shadow + incoherent rays), Ptex filtered
lookups (configs[3] lookup microbench: 64M lookups, half random faces, half coherent), VCM hash grid build + range query
(configs[4] microbench: 8M points on a floor + sphere, 8M queries), Disney BSDF batch operators (through the host C ABI).
GPU times are CUDA events on the launching stream, median of 10 after 3 warm-ups; inputs are larger than the 126 MB L2 or
say otherwise.  CPU numbers: kind "reference" = the reference's own objects (oracle/_ref), "port" = the restatement.
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
from tests import inputs, oracle_binding  # noqa: E402


def peak():
    p = ROOT / "MEASURED_PEAKS.json"
    return float(json.loads(p.read_text())["hbm_gbs"]) if p.exists() else 6650.0


def cpu_threads():
    return len(os.sched_getaffinity(0))


def median_ms(torch, stream, fn, reps=10, warm=3):
    ts = []
    for i in range(warm + reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def to_dev(torch, dev, a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)


def row_raycast(ctx, torch, dev, stream, orc):
    data = scenes.raycast_scene()
    sc = capi.Scene(ctx, data)
    cam = scenes.raycast_camera(1920, 1080)
    rays = scenes.primary_rays(cam, 1920, 1080)
    hits, ng = sc.intersect(rays, want_ng=True)
    sh = scenes.shadow_rays_toward(rays, hits, ng, (5.0, 5.0, 5.0))
    inc = scenes.cosine_bounce_rays(rays, hits, ng)
    t0 = time.perf_counter()
    hs = orc.scene_create(data)
    cpu_build = time.perf_counter() - t0
    out = {"triangles": data.triangle_count(), "bvh_build_ms": sc.bvh_stats().buildMs, "bvh_nodes": sc.bvh_stats().nodeCount,
           "cpu_bvh_build_s": cpu_build, "cpu": {"kind": "port", "cores": cpu_threads(),
                                                "what": "oracle scalar SAH-BVH2, float Moller-Trumbore (Embree 3.2.0 absent)"}}
    for name, r, anyhit in (("primary_closest", rays, False), ("shadow_occlusion", sh, True), ("incoherent_closest", inc, False)):
        n = r.size
        d_r = to_dev(torch, dev, r)
        d_h = torch.empty(n * 32, dtype=torch.uint8, device=dev)
        d_o = torch.empty(n, dtype=torch.uint8, device=dev)
        if anyhit:
            ms = median_ms(torch, stream, lambda: sc.occluded_device(n, d_r.data_ptr(), d_o.data_ptr(), stream.cuda_stream))
        else:
            ms = median_ms(torch, stream, lambda: sc.intersect_device(n, d_r.data_ptr(), d_h.data_ptr(), 0, stream.cuda_stream))
        _, counts = sc.intersect_counted(r, any_hit=anyhit)
        nn, nt = float(counts[:, 0].mean()), float(counts[:, 1].mean())
        bpr = (36.0 if anyhit else 64.0) + 80.0 * nn + 48.0 * nt
        # e2e through the host-buffer entry point
        t0 = time.perf_counter()
        for _ in range(3):
            sc.occluded(r) if anyhit else sc.intersect(r)
        e2e = 3 * n / (time.perf_counter() - t0)
        # CPU: same rays, all host threads, float BVH; parity of the results is checked at the same time
        t0 = time.perf_counter()
        if anyhit:
            co = orc.occluded(hs, r, mode=0, nthreads=cpu_threads())
        else:
            co = orc.intersect(hs, r, mode=0, nthreads=cpu_threads())
        cpu_s = time.perf_counter() - t0
        if anyhit:
            agree = float((sc.occluded(r) == co).mean())
        else:
            g = sc.intersect(r)
            agree = float(((g["geomID"] == co["geomID"]) & (g["primID"] == co["primID"])).mean())
        out[name] = {"rays": int(n), "ms": ms, "mrays_per_s": n / ms / 1e3, "nodes_per_ray": nn, "tris_per_ray": nt,
                     "algorithmic_bytes_per_ray": bpr, "achieved_gbs": bpr * n / (ms * 1e-3) / 1e9,
                     "frac_of_measured_hbm": bpr * n / (ms * 1e-3) / 1e9 / peak(), "e2e_host_buffers_mrays_per_s": e2e / 1e6,
                     "cpu_mrays_per_s": n / cpu_s / 1e6, "raw_id_agreement_with_float_cpu_bvh": agree}
    orc.scene_destroy(hs)
    sc.close()
    return out


def row_instancing(ctx, torch, dev, stream, n_inst=20000):
    """Flattened vs two-level BVH (SELAS_SCENE_TWO_LEVEL_BVH) on a forest of instanced spheres: bytes on the device, build time and
    primary closest-hit rate.  The flattened build stores every instance's triangles; the two-level one stores each mesh once."""
    rng = np.random.default_rng(5)
    sph = scenes.displaced_sphere(64, 128, as_quads=True)  # 16,384 triangles
    floor = dict(positions=np.array([[-200, -2, -200], [-200, -2, 200], [200, -2, 200], [200, -2, -200]], np.float32),
                 indices=np.array([[0, 1, 2, 3]], np.uint32), material=0, lightSet=0)
    inst = [(1, np.eye(4, dtype=np.float32))]
    for _ in range(n_inst):
        a, s = rng.uniform(0, 2 * np.pi), rng.uniform(0.4, 1.6)
        c, sn = np.cos(a), np.sin(a)
        inst.append((0, np.array([[s * c, 0, -s * sn, 0], [0, s * rng.uniform(0.7, 1.3), 0, 0], [s * sn, 0, s * c, 0],
                                  [rng.uniform(-180, 180), rng.uniform(-1, 8), rng.uniform(-180, 180), 1]], np.float32)))
    mats = np.array([scenes.material(**scenes.REFERENCE_MATERIALS["RoughDiffuse"])])
    cam = capi.camera_init((0.0, 60.0, -260.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 50.0, 1920, 1080)
    rays = scenes.primary_rays(cam, 1920, 1080)
    n = rays.size
    d_r = to_dev(torch, dev, rays)
    d_h = torch.empty(n * 32, dtype=torch.uint8, device=dev)
    out = {"instances": n_inst, "triangles_per_instance": 64 * 128 * 2, "rays": int(n)}
    hits = {}
    for name, two in (("flattened", False), ("two_level", True)):
        t0 = time.perf_counter()
        sc = capi.Scene(ctx, capi.SceneData([sph, floor], mats, None, None, instances=inst, two_level=two))
        create_s = time.perf_counter() - t0
        st = sc.bvh_stats()
        ms = median_ms(torch, stream, lambda: sc.intersect_device(n, d_r.data_ptr(), d_h.data_ptr(), 0, stream.cuda_stream))
        hits[name] = sc.intersect(rays)
        out[name] = {"scene_create_s": create_s, "bvh_build_ms": st.buildMs, "stored_triangles": int(st.triangleCount), "bvh_nodes": int(st.nodeCount),
                     "device_mb": st.deviceBytes / 1e6, "primary_closest_ms": ms, "mrays_per_s": n / ms / 1e3}
        sc.close()
    a, b = hits["flattened"], hits["two_level"]
    same = (a["geomID"] == b["geomID"]) & (a["primID"] == b["primID"]) & (a["instID0"] == b["instID0"])
    out["id_agreement_two_level_vs_flattened"] = float(same.mean())
    out["hit_fraction"] = float((a["geomID"] != capi.SELAS_INVALID_ID).mean())
    return out


def row_ptex(ctx, torch, dev, stream, grid, lookups):
    ptexref = oracle_binding.load_ptexref()
    if ptexref is None:
        return {"skipped": "oracle/_ref/libptex_ref.so not present (needed to author the .ptx with the reference's PtexWriter)"}
    tmp = Path(tempfile.mkdtemp()) / "c4.ptx"
    t0 = time.perf_counter()
    ptexref.write_grid(tmp, grid, grid, 3, 3, 0, 0, 3, 4, True, 0)
    write_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    tex = capi.Ptex(ctx, tmp)
    load_s = time.perf_counter() - t0
    nfaces = grid * grid
    rng = np.random.default_rng(4)
    half = lookups // 2
    face = np.empty(lookups, np.int32)
    face[:half] = rng.integers(0, nfaces, size=half)
    face[half:] = (np.arange(half, dtype=np.int64) * nfaces // half).astype(np.int32)  # coherent: ~consecutive faces
    u = rng.uniform(0, 1, size=lookups).astype(np.float32)
    v = rng.uniform(0, 1, size=lookups).astype(np.float32)
    d_f, d_u, d_v = to_dev(torch, dev, face), to_dev(torch, dev, u), to_dev(torch, dev, v)
    d_o = torch.empty(lookups * 3, dtype=torch.float32, device=dev)
    out = {"faces": nfaces, "face_res": "8x8 uint8 RGB", "file_mb": tmp.stat().st_size / 1e6, "write_s_reference_PtexWriter": write_s,
           "load_s": load_s, "lookups": lookups}
    for name, lo, hi in (("random_faces", 0, half), ("coherent_faces", half, lookups), ("all", 0, lookups)):
        n = hi - lo
        ms = median_ms(torch, stream, lambda: tex.eval_device(n, d_f.data_ptr() + 4 * lo, d_u.data_ptr() + 4 * lo, d_v.data_ptr() + 4 * lo,
                                                              d_o.data_ptr() + 12 * lo, stream.cuda_stream))
        bytes_per = 32.0 + 48.0 + 12.0 + 12.0  # face record + 16 texels x 3 B + query (faceid,u,v) + rgb out
        out[name] = {"n": n, "ms": ms, "mlookups_per_s": n / ms / 1e3, "algorithmic_bytes_per_lookup": bytes_per,
                     "achieved_gbs": bytes_per * n / (ms * 1e-3) / 1e9, "frac_of_measured_hbm": bytes_per * n / (ms * 1e-3) / 1e9 / peak()}
    # CPU: the reference's own PtexFilter(f_bspline)::eval, one thread (PtexFilter objects are per-thread), 1M-lookup sample
    k = 1_000_000
    sel = np.concatenate([np.arange(k // 2), half + np.arange(k // 2)])
    t0 = time.perf_counter()
    want = ptexref.eval(tmp, face[sel], u[sel], v[sel])
    cpu_s = time.perf_counter() - t0
    got = tex.eval(face[sel], u[sel], v[sel])
    out["cpu"] = {"kind": "reference", "cores": 1, "mlookups_per_s": k / cpu_s / 1e6, "sample": f"{k} of the same lookups (includes cache warm-up)",
                  "max_abs_diff_vs_gpu": float(np.abs(got - want).max())}
    tex.close()
    tmp.unlink()
    return out


def row_hashgrid(ctx, torch, dev, stream, n, m):
    rng = np.random.default_rng(5)
    # points on a floor quad and a sphere (SURVEY.md 8(d) C5 microbench), scene radius ~ 10
    nf = n // 2
    floor = np.stack([rng.uniform(-10, 10, nf), np.zeros(nf), rng.uniform(-10, 10, nf)], 1)
    d = rng.normal(size=(n - nf, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    sphere = d * 2.0 + np.array([[0.0, 2.0, 0.0]])
    pos = np.concatenate([floor, sphere]).astype(np.float32)
    pos = pos[rng.permutation(n)]
    radius = np.float32(0.0025 * 14.0)  # r0 = 0.0025 * sceneRadius (VCM.cpp:34-35)
    q = (pos[rng.integers(0, n, size=m)] + rng.normal(scale=float(radius) * 0.5, size=(m, 3))).astype(np.float32)
    cells = n  # cellCount = vmCount (VCM.cpp:689)
    lib = ctx.lib
    d_p, d_q = to_dev(torch, dev, pos), to_dev(torch, dev, q)
    handle = capi.C.c_void_p()

    def build():
        ctx.check(lib.selas_cuda_hashgrid_build_device(ctx.handle, capi.C.c_size_t(n), capi.C.c_void_p(d_p.data_ptr()), capi.f32(radius),
                                                       capi.u32(cells), capi.C.byref(handle), capi.C.c_void_p(stream.cuda_stream)))
    build_ms = median_ms(torch, stream, build)
    d_c = torch.empty(m, dtype=torch.int32, device=dev)
    kmax = 16
    d_i = torch.empty(m * kmax, dtype=torch.int32, device=dev)

    def query():
        ctx.check(lib.selas_cuda_hashgrid_query_device(handle, capi.C.c_size_t(m), capi.C.c_void_p(d_q.data_ptr()), capi.C.c_void_p(d_c.data_ptr()),
                                                       capi.C.c_void_p(d_i.data_ptr()), capi.u32(kmax), capi.C.c_void_p(stream.cuda_stream)))
    query_ms = median_ms(torch, stream, query)
    counts = d_c.cpu().numpy().astype(np.int64)
    ends = np.zeros(cells, np.uint32)
    idx = np.zeros(n, np.uint32)
    ctx.check(lib.selas_cuda_hashgrid_read(handle, ends.ctypes.data_as(capi.C.c_void_p), idx.ctypes.data_as(capi.C.c_void_p), None))
    occ = np.diff(np.concatenate([[0], ends.astype(np.int64)]))
    aabb = np.zeros(6, np.float32)
    ctx.check(lib.selas_cuda_hashgrid_read(handle, None, None, aabb.ctypes.data_as(capi.C.c_void_p)))
    # candidates actually visited = sum of the eight cell-range lengths per query (host restatement of the addressing,
    # used only to account the bytes: 64-bit hash as VCMHashGrid.cpp:15-22)
    inv = np.float32(1.0) / (np.float32(2.0) * radius)
    cp = (inv * (q - aabb[None, :3])).astype(np.float32)
    fl = np.floor(cp)
    basec = fl.astype(np.int64) - ((cp - fl) < np.float32(0.5))
    inside = np.all(q >= aabb[None, :3], axis=1) & np.all(q <= aabb[None, 3:], axis=1)
    cand = np.zeros(m, np.int64)
    for c8 in range(8):
        xyz = basec + np.array([c8 & 1, (c8 >> 1) & 1, c8 >> 2], np.int64)[None, :]
        h = (xyz[:, 0].astype(np.uint64) * np.uint64(73856093)) ^ (xyz[:, 1].astype(np.uint64) * np.uint64(19349663)) ^ \
            (xyz[:, 2].astype(np.uint64) * np.uint64(83492791))
        cell = (h % np.uint64(cells)).astype(np.int64)
        cand += occ[cell]
    cand[~inside] = 0
    accepted = float(counts.mean())
    build_bytes = 36.0 * n + 8.0 * cells
    query_bytes = (64.0 + 12.0 + 4.0) * m + 16.0 * float(cand.sum()) + 4.0 * np.minimum(counts, kmax).sum()
    out = {"points": n, "queries": m, "cells": cells, "radius": float(radius), "max_cell_occupancy": int(occ.max()), "mean_accepted_per_query": accepted,
           "mean_candidates_per_query": float(cand.mean()),
           "build": {"ms": build_ms, "mpoints_per_s": n / build_ms / 1e3, "algorithmic_bytes": build_bytes,
                     "achieved_gbs": build_bytes / (build_ms * 1e-3) / 1e9, "frac_of_measured_hbm": build_bytes / (build_ms * 1e-3) / 1e9 / peak()},
           "query": {"ms": query_ms, "mqueries_per_s": m / query_ms / 1e3, "algorithmic_bytes": query_bytes,
                     "achieved_gbs": query_bytes / (query_ms * 1e-3) / 1e9, "frac_of_measured_hbm": query_bytes / (query_ms * 1e-3) / 1e9 / peak()}}
    # CPU: the reference's own BuildHashGrid / SearchHashGrid (serial, as in the reference: one grid per thread-iteration)
    ref = oracle_binding.load_ref()
    lib_cpu, kind = (ref, "reference") if ref is not None else (oracle_binding.load_oracle(), "port")
    t0 = time.perf_counter()
    grid_cpu = lib_cpu.hashgrid_build(pos, float(radius), cells)
    b_s = time.perf_counter() - t0
    mq = min(m, 1_000_000)
    t0 = time.perf_counter()
    if kind == "reference":
        cc, ci = lib_cpu.hashgrid_query(pos, float(radius), cells, q[:mq], kmax)
        q_s = time.perf_counter() - t0 - b_s  # the harness rebuilds the grid inside the call
    else:
        cc, ci = lib_cpu.hashgrid_query(pos, float(radius), cells, grid_cpu, q[:mq], kmax)
        q_s = time.perf_counter() - t0
    gi = d_i.cpu().numpy().view(np.uint32).reshape(m, kmax)[:mq]
    same_counts = bool(np.array_equal(cc, counts[:mq].astype(np.uint32)))
    valid = np.arange(kmax)[None, :] < np.minimum(cc, kmax)[:, None]
    out["cpu"] = {"kind": kind, "cores": 1, "build_mpoints_per_s": n / b_s / 1e6, "query_mqueries_per_s": mq / max(q_s, 1e-9) / 1e6,
                  "sample": f"build: all {n} points; query: first {mq} queries", "arrays_bit_identical": bool(np.array_equal(grid_cpu[0], ends) and np.array_equal(grid_cpu[1], idx)),
                  "query_results_identical": bool(same_counts and np.array_equal(gi[valid], ci[valid]))}
    lib.selas_cuda_hashgrid_destroy(handle)
    return out


def row_bsdf(ctx, n=2_000_000):
    s, v, nrm = inputs.random_surfaces(7, n)
    l = inputs.light_directions(7, n, nrm, v)
    u = inputs.uniforms(7, n, capi.SELAS_BSDF_RANDOMS)
    ref = oracle_binding.load_ref()
    lib_cpu, kind = (ref, "reference") if ref is not None else (oracle_binding.load_oracle(), "port")
    out = {"n": n, "note": "operator-level parity entry points through the host-buffer C ABI (H2D/D2H included); the fused path is shadeKernel"}
    for name, g_fn, c_fn in (("eval", lambda: ctx.bsdf_eval(s, v, l), lambda: lib_cpu.bsdf_eval(s[:200000], v[:200000], l[:200000])),
                             ("sample", lambda: ctx.bsdf_sample(s, v, u), lambda: lib_cpu.bsdf_sample(s[:200000], v[:200000], u[:200000]))):
        g_fn()
        t0 = time.perf_counter()
        g_fn()
        g_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        c_fn()
        c_s = time.perf_counter() - t0
        out[name] = {"gpu_e2e_m_per_s": n / g_s / 1e6, "cpu_m_per_s": 200000 / c_s / 1e6, "cpu_kind": kind, "cpu_cores": 1}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ptex-grid", type=int, default=800)
    ap.add_argument("--ptex-lookups", type=int, default=64_000_000)
    ap.add_argument("--grid-points", type=int, default=8_000_000)
    ap.add_argument("--rows", default="raycast,ptex,hashgrid,bsdf,instancing")
    args = ap.parse_args()
    import torch

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    ctx = capi.Context(0)
    orc = oracle_binding.load_oracle()
    doc = {"peak_hbm_gbs": peak(), "gpu": torch.cuda.get_device_name(0), "host_threads": cpu_threads()}
    rows = args.rows.split(",")
    if "raycast" in rows:
        doc["raycast_config2"] = row_raycast(ctx, torch, dev, stream, orc)
    if "ptex" in rows:
        doc["ptex_lookup"] = row_ptex(ctx, torch, dev, stream, args.ptex_grid, args.ptex_lookups)
    if "hashgrid" in rows:
        doc["vcm_hashgrid"] = row_hashgrid(ctx, torch, dev, stream, args.grid_points, args.grid_points)
    if "bsdf" in rows:
        doc["disney_bsdf_operators"] = row_bsdf(ctx)
    if "instancing" in rows:
        doc["instancing_forest"] = row_instancing(ctx, torch, dev, stream)
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main()
