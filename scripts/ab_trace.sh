set -e
cd /root/repo
python - <<'PY'
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from selas_b200 import capi, scenes
ctx = capi.Context(0)
data = scenes.disney_heightfield()
sc = capi.Scene(ctx, data)
cam = scenes.heightfield_camera(1920, 1080)
import torch
film = torch.zeros(1920*1080*3, dtype=torch.float32, device='cuda')
st = torch.cuda.current_stream().cuda_stream
for i in range(3):
    sc.render_device(cam, capi.render_settings(1920,1080,4,16,8,seed=1,first_sample=4*i), film.data_ptr(), st)
tot=0; c=0; s=0; sh=0; n=10
for i in range(n):
    r = sc.render_device(cam, capi.render_settings(1920,1080,4,16,8,seed=1,first_sample=4*(i+3)), film.data_ptr(), st)
    tot+=r.renderMs; c+=r.closestMs; s+=r.shadowMs; sh+=r.shadeMs
print(f"ms/step {tot/n:.3f} closest {c/n:.3f} shadow {s/n:.3f} shade {sh/n:.3f}  samples/s {1920*1080*4/(tot/n*1e-3)/1e6:.1f} M")
r = sc.render_device(cam, capi.render_settings(1920,1080,4,16,8,seed=1,first_sample=0,flags=1), film.data_ptr(), st)
print("nodes/ray", r.closestNodes/r.closestRays, "tris/ray", r.closestTris/r.closestRays, "shadow nodes", r.shadowNodes/max(1,r.occlusionRays), r.shadowTris/max(1,r.occlusionRays))
PY
