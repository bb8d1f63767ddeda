#!/usr/bin/env python
"""One warm-up + N VCM iterations at 1080p on the configs[4] caustic scene; the command the ncu captures under profiles/ wrap."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from selas_b200 import capi, scenes  # noqa: E402

iterations = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ctx = capi.Context(0)
data = scenes.caustic_scene()
sc = capi.Scene(ctx, data)
w, h = 1920, 1080
cam = scenes.caustic_camera(w, h)
film, st = sc.render_vcm(cam, capi.vcm_settings(w, h, iterations, scenes.bounding_sphere(data), seed=1, first_iteration=1))
print(f"{iterations} iteration(s): {st.renderMs:.2f} ms, light {st.lightPassMs:.2f} grid {st.gridBuildMs:.2f} camera {st.cameraPassMs:.2f} "
      f"(merge {st.mergeMs:.2f}, connect {st.connectMs:.2f}); {st.lightVertices} light vertices, {st.kernelLaunches} launches")
