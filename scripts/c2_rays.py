#!/usr/bin/env python
"""configs[1] ray batches through the device entry points, for ncu captures: python scripts/c2_rays.py [primary|shadow|incoherent] [launches]."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from selas_b200 import capi, scenes  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "incoherent"
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 3
W, H = 1920, 1080
ctx = capi.Context(0)
sc = capi.Scene(ctx, scenes.raycast_scene())
cam = scenes.raycast_camera(W, H)
rays = scenes.primary_rays(cam, W, H)
hits, ng = sc.intersect(rays, want_ng=True)
if which == "shadow":
    rays = scenes.shadow_rays_toward(rays, hits, ng, (5.0, 5.0, 5.0))
elif which == "incoherent":
    rays = scenes.cosine_bounce_rays(rays, hits, ng)
dev = torch.device("cuda", 0)
d_r = torch.from_numpy(rays.view(np.uint8).reshape(-1)).to(dev)
d_h = torch.empty(rays.size * 32, dtype=torch.uint8, device=dev)
d_o = torch.empty(rays.size, dtype=torch.uint8, device=dev)
for _ in range(launches):
    if which == "shadow":
        sc.occluded_device(rays.size, d_r.data_ptr(), d_o.data_ptr(), 0)
    else:
        sc.intersect_device(rays.size, d_r.data_ptr(), d_h.data_ptr(), 0, 0)
ctx.synchronize() if hasattr(ctx, "synchronize") else torch.cuda.synchronize()
torch.cuda.synchronize()
print("done", which, rays.size)
