#!/usr/bin/env python
"""One warm-up step + N steps of the bench workload (BASELINE configs[2], 4 spp, 1080p, depth 16), nothing else: the command the ncu
captures of the wavefront kernels wrap.  python scripts/profile_step.py [steps] [grid]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from selas_b200 import capi, scenes  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
grid = int(sys.argv[2]) if len(sys.argv) > 2 else 2236
W, H = 1920, 1080
ctx = capi.Context(0)
scene = capi.Scene(ctx, scenes.disney_heightfield(grid=grid))
cam = scenes.heightfield_camera(W, H)
film = torch.zeros(W * H * 3, dtype=torch.float32, device="cuda")
for i in range(1 + steps):
    st = scene.render_device(cam, capi.render_settings(W, H, 4, 16, 8, layers=1, seed=1, first_sample=4 * i), film.data_ptr(), 0)
print(f"step: {st.renderMs:.2f} ms, {st.kernelLaunches} launches, closest {st.closestMs:.2f} shade {st.shadeMs:.2f} shadow {st.shadowMs:.2f}; "
      f"{st.closestRays} closest + {st.occlusionRays} shadow rays")
