import numpy as np, sys
sys.path.insert(0,'/root/repo')
from selas_b200 import capi
from tests import inputs, oracle_binding
orc = oracle_binding.load_oracle(); ctx = capi.Context(0)
N=200000
s, v, _ = inputs.random_surfaces(22, N); u = inputs.uniforms(22, N, capi.SELAS_BSDF_RANDOMS)
for i in (175978, 10070):
    K=11
    ss=np.repeat(s[i:i+1],K); vv=np.repeat(v[i:i+1],K,axis=0); uu=np.repeat(u[i:i+1],K,axis=0)
    uu[:,1]=(u[i,1]*(1+np.arange(-5,6)*1e-2)).astype(np.float32)
    g=ctx.bsdf_sample(ss,vv,uu); o=orc.bsdf_sample(ss,vv,uu)
    print("sample",i,"sweep u1")
    for k in range(K): print("  u1 %.6e gpu wi %s orc wi %s diff %s pdf %.6f %.6f" % (uu[k,1], g["wi"][k], o["wi"][k], g["wi"][k]-o["wi"][k], g["forwardPdfW"][k], o["forwardPdfW"][k]))
    uu=np.repeat(u[i:i+1],K,axis=0)
    uu[:,1]=np.float32([1e-6,1e-5,1e-4,1e-3,1e-2,0.05,0.1,0.3,0.5,0.7,0.9])
    g=ctx.bsdf_sample(ss,vv,uu); o=orc.bsdf_sample(ss,vv,uu)
    print("sample",i,"u1 decades")
    for k in range(K): print("  u1 %.6e diff wi %s relpdf %.3e" % (uu[k,1], g["wi"][k]-o["wi"][k], g["forwardPdfW"][k]/o["forwardPdfW"][k]-1))
