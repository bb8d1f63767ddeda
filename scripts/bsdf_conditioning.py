import numpy as np, sys
sys.path.insert(0,'/root/repo')
from selas_b200 import capi
from tests import inputs, oracle_binding
orc = oracle_binding.load_oracle()
ctx = capi.Context(0)
N=200000
s, v, _ = inputs.random_surfaces(22, N)
u = inputs.uniforms(22, N, capi.SELAS_BSDF_RANDOMS)
g = ctx.bsdf_sample(s, v, u); o = orc.bsdf_sample(s, v, u)
same = (g["valid"] == o["valid"]) & (g["flags"] == o["flags"]) & (g["randomsUsed"] == o["randomsUsed"]) & (g["mediumPhase"] == o["mediumPhase"])
ok = same & (o["valid"] == 1)
print("discrete mismatches", (~same).sum())
# sensitivity of the oracle to 1-ulp input perturbations (u and v)
def perturbed(k, sign):
    u2 = u.copy(); u2[:, k] = np.nextafter(u[:, k], np.float32(2.0 if sign > 0 else -1.0)).astype(np.float32)
    return orc.bsdf_sample(s, v, u2)
def perturbed_v(k, sign):
    v2 = v.copy(); v2[:, k] = np.nextafter(v[:, k], np.float32(2.0 if sign > 0 else -2.0)).astype(np.float32)
    return orc.bsdf_sample(s, v2, u)
variants = [perturbed(k, sg) for k in range(u.shape[1]) for sg in (1, -1)] + [perturbed_v(k, sg) for k in range(3) for sg in (1, -1)]
for f in ("reflectance", "wi", "forwardPdfW", "reversePdfW"):
    a = np.asarray(g[f], np.float64).reshape(N, -1); b = np.asarray(o[f], np.float64).reshape(N, -1)
    scale = np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-5)
    err = (np.abs(a - b) / scale).max(axis=1)
    sens = np.zeros(N)
    for w in variants:
        wv = (w["valid"] == o["valid"]) & (w["flags"] == o["flags"])
        c = np.asarray(w[f], np.float64).reshape(N, -1)
        d = (np.abs(c - b) / scale).max(axis=1)
        d[~wv] = np.inf  # a 1-ulp nudge flips a branch: infinitely ill-conditioned
        sens = np.maximum(sens, d)
    out = ok & (err > 1e-5)
    ratio = err[out] / np.maximum(sens[out], 1e-30)
    print(f, "beyond 1e-5:", out.sum(), "of", ok.sum(), "max err", err[ok].max(), "| ratio err/sens: median", np.median(ratio) if out.sum() else None,
          "max", ratio.max() if out.sum() else None, "| unexplained (ratio>4):", (ratio > 4).sum() if out.sum() else 0, " >16:", (ratio > 16).sum() if out.sum() else 0)
    if out.sum():
        idx = np.nonzero(out)[0][np.argsort(-ratio)[:5]]
        for i in idx:
            print("   i", i, "err", err[i], "sens", sens[i], "shader", s["shader"][i], "rough", s["roughness"][i], "flags", o["flags"][i], "used", o["randomsUsed"][i])
for i in (175978, 10070, 96086, 9374):
    print("sample", i, "u", u[i], "v", v[i], "rough", s["roughness"][i], "aniso", s["anisotropic"][i], "metal", s["metallic"][i], "spectrans", s["specTrans"][i], "cc", s["clearcoat"][i])
    w2t = s["worldToTangent"][i].reshape(3, 3)
    print("   wo(tangent)", w2t @ v[i] if False else (v[i] @ w2t))
    for name, r in (("gpu", g[i]), ("orc", o[i])):
        print("  ", name, "wi", r["wi"], "refl", r["reflectance"], "fpdf", r["forwardPdfW"], "rpdf", r["reversePdfW"], "flags", r["flags"])
