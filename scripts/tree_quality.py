#!/usr/bin/env python
"""How good is the tree?  The configs[2] triangle soup under three CPU builders (test tooling in oracle/cpu_bvh.h): the oracle's binned
SAH, an LBVH with per-axis Morton normalisation, and an LBVH on cubic cells (what the device build does) -- the SAH cost of each tree and
the box / triangle tests per ray for camera rays and for incoherent rays.  CPU only; prints one JSON object.
  python scripts/tree_quality.py [--grid 2236] [--width 480] [--height 270]"""
import argparse
import ctypes as C
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from selas_b200 import capi, scenes  # noqa: E402
from tests import oracle_binding  # noqa: E402


def soup(data):
    out = []
    for m in data.meshes:
        p, ix = np.asarray(m["positions"], np.float32), np.asarray(m["indices"])
        if ix.shape[1] == 3:
            out.append(p[ix].reshape(-1, 9))
        else:
            out.append(p[ix[:, [0, 1, 3]]].reshape(-1, 9))
            out.append(p[ix[:, [2, 3, 1]]].reshape(-1, 9))
    return np.ascontiguousarray(np.concatenate(out), np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="configs2", choices=["configs2", "configs1", "configs3"],
                    help="configs2 = heightfield + spheres (bench.py), configs1 = 1 M-triangle displaced sphere, configs3 = Ptex sheet")
    ap.add_argument("--grid", type=int, default=2236)
    ap.add_argument("--width", type=int, default=480)
    ap.add_argument("--height", type=int, default=270)
    args = ap.parse_args()
    orc = oracle_binding.load_oracle()
    if args.scene == "configs2":
        data, cam = scenes.disney_heightfield(grid=args.grid), scenes.heightfield_camera(args.width, args.height)
    elif args.scene == "configs1":
        data, cam = scenes.raycast_scene(), scenes.raycast_camera(args.width, args.height)
    else:
        data, cam = scenes.ptex_quad_grid(), scenes.ptex_grid_camera(args.width, args.height)
    tri = soup(data)
    primary = scenes.primary_rays(cam, args.width, args.height)
    rng = np.random.default_rng(12)
    n = primary.size
    inco = np.zeros(n, dtype=capi.RAY_DTYPE)  # bounce-like rays: from points just above the terrain into the upper hemisphere and sideways
    lo, hi = tri.reshape(-1, 3).min(axis=0), tri.reshape(-1, 3).max(axis=0)
    d = rng.normal(size=(n, 3))
    if args.scene == "configs2":
        inco["ox"], inco["oz"] = rng.uniform(-9, 9, n), rng.uniform(-9, 9, n)
        inco["oy"] = rng.uniform(-1.0, 1.5, n)
        d[:, 1] = np.abs(d[:, 1]) * 0.3 - 0.2
    else:  # random chords through the scene's box
        a, b = rng.uniform(lo, hi, size=(n, 3)), rng.uniform(lo, hi, size=(n, 3))
        inco["ox"], inco["oy"], inco["oz"] = a[:, 0], a[:, 1], a[:, 2]
        d = b - a
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    inco["dx"], inco["dy"], inco["dz"] = d[:, 0], d[:, 1], d[:, 2]
    inco["tnear"], inco["tfar"] = 0.0, 3.0e38
    threads = len(os.sched_getaffinity(0))
    f = orc.fn("tree_quality")
    doc = {"scene": args.scene, "triangles": int(tri.shape[0]), "rays_each": int(n), "builders": {}}
    for name, b in (("binned_sah", 0), ("lbvh_per_axis", 1), ("lbvh_cubic", 2), ("morton_order_sah_sweep", 3), ("sah_top_lbvh_below_65536", 65536),
                    ("sah_top_lbvh_below_4096", 4096), ("sah_top_lbvh_below_256", 256), ("ploc_radius_8", -8), ("ploc_radius_32", -32),
                    ("lbvh_top_sah_inside_256", (1 << 24) + 256), ("lbvh_top_sah_inside_4096", (1 << 24) + 4096)):
        row = {}
        for rname, rays in (("camera", primary), ("incoherent", inco)):
            out = (C.c_double * 4)()
            t0 = time.perf_counter()
            rc = f(C.c_size_t(tri.shape[0]), tri.ctypes.data_as(C.c_void_p), C.c_int(b), C.c_size_t(rays.size), rays.ctypes.data_as(C.c_void_p),
                   C.c_int(threads), out)
            assert rc == 0
            row["sah_inner"], row["sah_leaf"] = out[0], out[1]
            row[rname] = {"box_tests_per_ray": out[2], "tri_tests_per_ray": out[3]}
            row["build_plus_trace_s"] = time.perf_counter() - t0
        doc["builders"][name] = row
    base = doc["builders"]["lbvh_cubic"]
    for name, row in doc["builders"].items():
        row["relative_to_lbvh_cubic"] = {"sah_inner": row["sah_inner"] / base["sah_inner"],
                                         "camera_box_tests": row["camera"]["box_tests_per_ray"] / base["camera"]["box_tests_per_ray"],
                                         "incoherent_box_tests": row["incoherent"]["box_tests_per_ray"] / base["incoherent"]["box_tests_per_ray"]}
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main()
