import sys; sys.path.insert(0,'.')
import numpy as np
from selas_b200 import capi
from tests import inputs, oracle_binding
ctx=capi.Context(0); orc=oracle_binding.load_oracle()
data=inputs.small_scene(instances=True)
sc=capi.Scene(ctx,data); hs=orc.scene_create(data)
rays=inputs.scene_rays(31,20000)
g=sc.intersect(rays); o=orc.intersect(hs,rays,mode=2)
same=(g["geomID"]==o["geomID"])&(g["primID"]==o["primID"])&(g["instID0"]==o["instID0"])
print("agree",same.mean(), "bvh", sc.bvh_stats().triangleCount, sc.bvh_stats().nodeCount)
m=~same
import collections
print("oracle (inst,geom) of mismatches", collections.Counter(zip(o["instID0"][m].tolist(), o["geomID"][m].tolist())).most_common(10))
print("gpu (inst,geom) of mismatches", collections.Counter(zip(g["instID0"][m].tolist(), g["geomID"][m].tolist())).most_common(10))
print("all oracle", collections.Counter(zip(o["instID0"].tolist(), o["geomID"].tolist())).most_common(10))
print("all gpu", collections.Counter(zip(g["instID0"].tolist(), g["geomID"].tolist())).most_common(10))
idx=np.nonzero(m)[0][:5]
for i in idx: print(rays[i], g[i], o[i])
