#!/usr/bin/env python
"""PathTracer over a two-level (SELAS_SCENE_TWO_LEVEL_BVH) scene: 4000 instanced displaced spheres (16 384 triangles each) on a floor
under a sky and one quad light, 1080p, 4 spp, depth 8 -- the form every baked Selas scene is loaded in (LoadBakedScene).  Prints one JSON
object with the per-stage times; run with SELAS_COOP_CLOSEST_FROM=99 SELAS_COOP_SHADOW_FROM=99 for the traceKernel-only baseline.
  python scripts/two_level_render.py [--instances 4000]"""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from selas_b200 import capi, scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--instances", type=int, default=4000)
    args = ap.parse_args()
    rng = np.random.default_rng(5)
    sph = scenes.displaced_sphere(64, 128, as_quads=True)
    floor = dict(positions=np.array([[-200, -2, -200], [-200, -2, 200], [200, -2, 200], [200, -2, -200]], np.float32),
                 indices=np.array([[0, 1, 2, 3]], np.uint32), material=0, lightSet=0)
    inst = [(1, np.eye(4, dtype=np.float32))]
    for _ in range(args.instances):
        a, s = rng.uniform(0, 2 * np.pi), rng.uniform(1.0, 3.0)
        c, sn = np.cos(a), np.sin(a)
        inst.append((0, np.array([[s * c, 0, -s * sn, 0], [0, s * rng.uniform(0.7, 1.3), 0, 0], [s * sn, 0, s * c, 0],
                                  [rng.uniform(-150, 150), rng.uniform(-1, 6), rng.uniform(-150, 150), 1]], np.float32)))
    mats = np.array([scenes.material(**scenes.REFERENCE_MATERIALS["RoughDiffuse"])])
    lights = np.array([scenes.quad_light((0.0, 80.0, 0.0), (40.0, 0.0, 0.0), (0.0, 0.0, 40.0), (0.0, -1.0, 0.0), (60.0, 58.0, 52.0))])
    W, H = 1920, 1080
    cam = capi.camera_init((0.0, 60.0, -260.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 50.0, W, H)
    ctx = capi.Context(0)
    out = {"instances": args.instances, "env": {k: v for k, v in os.environ.items() if k.startswith("SELAS_")}}
    for name, two in (("two_level", True), ("flattened", False)):
        sc = capi.Scene(ctx, capi.SceneData([sph, floor], mats, lights, [(0, 1)], background=(0.4, 0.5, 0.7, 0.0), instances=inst, two_level=two))
        film = torch.zeros(W * H * 3, dtype=torch.float32, device="cuda")
        rs = capi.render_settings(W, H, 4, max_path_length=8, rr_start=8, layers=1, seed=3)
        for i in range(3):
            st = sc.render_device(cam, rs, film.data_ptr(), 0)
        out[name] = {"ms": float(st.renderMs), "msamples_per_s": W * H * 4 / st.renderMs / 1e3, "closest_ms": float(st.closestMs), "shadow_ms": float(st.shadowMs),
                     "shade_ms": float(st.shadeMs), "closest_rays": int(st.closestRays), "occlusion_rays": int(st.occlusionRays),
                     "mean_radiance": float(film.mean().item() / 12.0)}
        sc.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
