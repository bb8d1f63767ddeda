#!/usr/bin/env python
"""Round B-spline curve leaves in the trace kernels: a lawn of tapered blades over a floor quad at 1080p -- primary
closest-hit rays, NEE-style shadow rays from the visible points to the quad light, and a short render.  Prints one JSON object.
  python scripts/curves_bench.py [--blades 200000]"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from selas_b200 import capi, scenes  # noqa: E402


def timed(fn, stream, reps=6):
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms[1:]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blades", type=int, default=200_000)
    args = ap.parse_args()
    W, H = 1920, 1080
    t0 = time.perf_counter()
    data = scenes.grass_patch(blades=args.blades, extent=6.0, height=(0.25, 0.7), radius=(0.012, 0.002))
    gen_s = time.perf_counter() - t0
    ctx = capi.Context(0)
    sc = capi.Scene(ctx, data)
    bs = sc.bvh_stats()
    cam = scenes.grass_camera(W, H)
    rays = scenes.primary_rays(cam, W, H)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(dev)
    n = len(rays)
    d_rays = torch.from_numpy(rays.view(np.uint8)).to(dev)
    d_hits = torch.empty(n * 32, dtype=torch.uint8, device=dev)
    out = {"blades": args.blades, "segments": args.blades * 3, "scene_generation_s": gen_s, "bvh_build_ms": float(bs.buildMs),
           "bvh_nodes": int(bs.nodeCount)}
    with torch.cuda.stream(stream):
        ms = timed(lambda: sc.intersect_device(n, d_rays.data_ptr(), d_hits.data_ptr(), 0, stream.cuda_stream), stream)
    hits = d_hits.cpu().numpy().view(capi.HIT_DTYPE)
    curve_geom = len(data.meshes)
    on_curves = hits["geomID"] == curve_geom
    out["primary_closest"] = {"rays": n, "ms": ms, "mrays_per_s": n / ms / 1e3, "fraction_on_curves": float(on_curves.mean()),
                              "fraction_hit": float((hits["geomID"] != 0xFFFFFFFF).mean())}
    # shadow rays from the visible points towards the light centre
    hit = hits["geomID"] != 0xFFFFFFFF
    p = np.stack([rays[k][hit] + hits["t"][hit] * rays["d" + k[1]][hit] for k in ("ox", "oy", "oz")], axis=1)
    to = np.array([0.0, 6.0, 0.0], np.float32) - p
    dist = np.linalg.norm(to, axis=1).astype(np.float32)
    sh = np.zeros(int(hit.sum()), dtype=capi.RAY_DTYPE)
    sh["ox"], sh["oy"], sh["oz"] = p[:, 0], p[:, 1], p[:, 2]
    sh["dx"], sh["dy"], sh["dz"] = (to / dist[:, None]).T
    sh["tnear"], sh["tfar"] = 1e-3, dist
    d_sh = torch.from_numpy(sh.view(np.uint8)).to(dev)
    d_occ = torch.empty(len(sh), dtype=torch.uint8, device=dev)
    with torch.cuda.stream(stream):
        ms = timed(lambda: sc.occluded_device(len(sh), d_sh.data_ptr(), d_occ.data_ptr(), stream.cuda_stream), stream)
    out["shadow_occlusion"] = {"rays": len(sh), "ms": ms, "mrays_per_s": len(sh) / ms / 1e3, "fraction_occluded": float(d_occ.float().mean().item())}
    # a short render (4 spp, depth 6): path samples per second with curve leaves in every bounce
    rs = capi.render_settings(W, H, 4, max_path_length=6, rr_start=8, layers=1, seed=3)
    film = torch.zeros(W * H * 3, dtype=torch.float32, device=dev)
    sc.render_device(cam, rs, film.data_ptr(), 0)
    st = sc.render_device(cam, rs, film.data_ptr(), 0)
    out["render_4spp_depth6"] = {"ms": float(st.renderMs), "msamples_per_s": W * H * 4 / st.renderMs / 1e3, "closest_ms": float(st.closestMs),
                                 "shade_ms": float(st.shadeMs), "shadow_ms": float(st.shadowMs), "closest_rays": int(st.closestRays),
                                 "occlusion_rays": int(st.occlusionRays)}
    print(json.dumps(out, indent=1))
    sc.close()


if __name__ == "__main__":
    main()
