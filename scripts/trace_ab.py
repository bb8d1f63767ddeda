#!/usr/bin/env python
"""A/B harness for build / traversal variants on the bench workload (BASELINE configs[2]) and on configs[1] (ray-cast microbench).

Every variant runs in its own process (the library reads its tuning knobs from the environment once):
    python scripts/trace_ab.py --out gpurun_out/r2_ab.jsonl  "name:ENV=V,ENV2=V2"  "other:ENV=V" ...
    python scripts/trace_ab.py --child          (internal)
Per variant one JSON line: BVH build ms / nodes / levels, ms per 4-spp step, per-kernel ms, nodes and triangles per closest / shadow ray,
the image checksum (mean radiance -- variants that only reorganise the tree must agree), and configs[1] Mrays/s.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def child(args):
    import numpy as np
    import torch

    from selas_b200 import capi, scenes

    W, H = 1920, 1080
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    out = {}
    ctx = capi.Context(0)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    if not args.no_c3:
        data = scenes.disney_heightfield(grid=args.grid)
        cam = scenes.heightfield_camera(W, H)
        t0 = time.perf_counter()
        scene = capi.Scene(ctx, data)
        bvh = scene.bvh_stats()
        out["c3_bvh"] = {"build_ms": bvh.buildMs, "nodes": bvh.nodeCount, "levels": bvh.maxDepth, "clusters": bvh.sahClusters,
                         "generations": bvh.sahGenerations, "upload_plus_build_s": time.perf_counter() - t0}
        film = torch.zeros(W * H * 3, dtype=torch.float32, device=dev)

        def settings(i, flags=0):
            return capi.render_settings(W, H, 4, 16, 8, layers=1, seed=1, first_sample=i * 4, flags=flags)

        for i in range(3):
            scene.render_device(cam, settings(i), film.data_ptr(), stream.cuda_stream, want_stats=False)
        scene.collect_stats()
        film.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_host = time.perf_counter()
        e0.record(stream)
        for i in range(args.steps):
            scene.render_device(cam, settings(10 + i), film.data_ptr(), stream.cuda_stream, want_stats=False)  # enqueue only
        e1.record(stream)
        enqueue_ms = (time.perf_counter() - t_host) * 1e3 / args.steps
        torch.cuda.synchronize()
        st = scene.collect_stats()
        agg = {k: getattr(st, k) for k in ("closestMs", "shadeMs", "shadowMs", "generateMs", "closestRays", "occlusionRays")}
        ms = e0.elapsed_time(e1) / args.steps
        fin = torch.isfinite(film)
        out["c3"] = {"ms_per_step": ms, "msamples_per_s": W * H * 4 / ms / 1e3, "closest_ms": agg["closestMs"] / args.steps,
                     "shadow_ms": agg["shadowMs"] / args.steps, "shade_ms": agg["shadeMs"] / args.steps,
                     "closest_mrays": agg["closestRays"] / max(agg["closestMs"], 1e-9) / 1e3,
                     "shadow_mrays": agg["occlusionRays"] / max(agg["shadowMs"], 1e-9) / 1e3,
                     "mean_radiance": float(film[fin].mean().item()) / (4 * args.steps), "nonfinite": int((~fin).sum().item()),
                     "host_enqueue_ms_per_step": enqueue_ms, "launches_per_step": st.kernelLaunches / args.steps}
        cst = scene.render_device(cam, settings(2000, flags=capi.RENDER_COUNT_TRAVERSAL), film.data_ptr(), stream.cuda_stream)
        out["c3"].update({"nodes_per_closest": cst.closestNodes / max(1, cst.closestRays), "tris_per_closest": cst.closestTris / max(1, cst.closestRays),
                          "nodes_per_shadow": cst.shadowNodes / max(1, cst.occlusionRays), "tris_per_shadow": cst.shadowTris / max(1, cst.occlusionRays)})
        scene.close()
    if not args.no_c2:
        data = scenes.raycast_scene()
        sc = capi.Scene(ctx, data)
        cam = scenes.raycast_camera(W, H)
        rays = scenes.primary_rays(cam, W, H)
        hits, ng = sc.intersect(rays, want_ng=True)
        sh = scenes.shadow_rays_toward(rays, hits, ng, (5.0, 5.0, 5.0))
        inc = scenes.cosine_bounce_rays(rays, hits, ng)
        c2 = {"bvh_build_ms": sc.bvh_stats().buildMs, "nodes": sc.bvh_stats().nodeCount,
              "hit_checksum": int(hits["primID"].astype(np.uint64).sum() % (1 << 61))}
        for name, r, anyhit in (("primary", rays, False), ("shadow", sh, True), ("incoherent", inc, False)):
            d_r = torch.from_numpy(r.view(np.uint8).reshape(-1)).to(dev)
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
            _, counts = sc.intersect_counted(r, any_hit=anyhit)
            c2[name] = {"mrays": n / float(np.median(ts)) / 1e3, "nodes": float(counts[:, 0].mean()), "tris": float(counts[:, 1].mean())}
        out["c2"] = c2
        sc.close()
    ctx.close()
    print("AB_RESULT " + json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--out", default="gpurun_out/trace_ab.jsonl")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--grid", type=int, default=2236)
    ap.add_argument("--no-c2", action="store_true")
    ap.add_argument("--no-c3", action="store_true")
    ap.add_argument("variants", nargs="*")
    args = ap.parse_args()
    if args.child:
        return child(args)
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    for v in args.variants or ["default:"]:
        name, _, envs = v.partition(":")
        env = dict(os.environ)
        for kv in filter(None, envs.split(",")):
            k, _, val = kv.partition("=")
            env[k] = val
        cmd = [sys.executable, __file__, "--child", "--steps", str(args.steps), "--grid", str(args.grid)]
        if args.no_c2:
            cmd.append("--no-c2")
        if args.no_c3:
            cmd.append("--no-c3")
        t0 = time.time()
        p = subprocess.run(cmd, env=env, capture_output=True, text=True)
        line = {"variant": name, "env": envs, "rc": p.returncode, "wall_s": time.time() - t0}
        for ln in p.stdout.splitlines():
            if ln.startswith("AB_RESULT "):
                line.update(json.loads(ln[len("AB_RESULT "):]))
        if p.returncode != 0:
            line["stderr_tail"] = p.stderr[-1500:]
        with open(args.out, "a") as f:
            f.write(json.dumps(line) + "\n")
        c3 = line.get("c3", {})
        print(name, "rc", p.returncode, "ms/step", round(c3.get("ms_per_step", 0), 3), "nodes/ray", round(c3.get("nodes_per_closest", 0), 2),
              round(c3.get("nodes_per_shadow", 0), 2), "build ms", round(line.get("c3_bvh", {}).get("build_ms", 0), 1), flush=True)


if __name__ == "__main__":
    main()
