"""Seeded input generators shared by the CPU (oracle-vs-reference) and GPU (CUDA-vs-oracle) parity tests."""
from __future__ import annotations

import numpy as np

from selas_b200 import capi, scenes


def unit_vectors(rng, n):
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return v.astype(np.float32)


def random_frames(rng, n):
    """Orthonormal (t, n, b) per sample -> worldToTangent rows (t.x,n.x,b.x),(t.y,n.y,b.y),(t.z,n.z,b.z)."""
    nrm = unit_vectors(rng, n).astype(np.float64)
    a = unit_vectors(rng, n).astype(np.float64)
    t = np.cross(nrm, a)
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    b = np.cross(nrm, t)
    w2t = np.stack([t[:, 0], nrm[:, 0], b[:, 0], t[:, 1], nrm[:, 1], b[:, 1], t[:, 2], nrm[:, 2], b[:, 2]], axis=1)
    return w2t.astype(np.float32), nrm.astype(np.float32)


MATERIAL_FAMILIES = ["rough_diffuse", "metal_aniso", "clearcoat", "sheen", "glass_solid", "glass_thin", "difftrans_thin",
                     "difftrans_solid", "mixed", "copper", "default", "dirac"]


def random_surfaces(seed, n, families=None):
    """n SurfaceParameters covering every Disney lobe (metallic/clearcoat/sheen/specTrans/diffTrans, thin and solid)."""
    rng = np.random.default_rng(seed)
    fam = families or MATERIAL_FAMILIES
    s = np.zeros(n, dtype=capi.SURFACE_DTYPE)
    w2t, nrm = random_frames(rng, n)
    s["worldToTangent"] = w2t
    s["position"] = rng.uniform(-100, 100, size=(n, 3)).astype(np.float32)
    s["error"] = (32.0 * 1.19209e-07 * np.max(np.abs(s["position"]), axis=1)).astype(np.float32)
    U = lambda lo=0.0, hi=1.0: rng.uniform(lo, hi, size=n).astype(np.float32)
    which = rng.integers(0, len(fam), size=n)
    base = rng.uniform(0.02, 1.0, size=(n, 3)).astype(np.float32)
    s["baseColor"] = base
    s["transmittanceColor"] = rng.uniform(0.1, 1.0, size=(n, 3)).astype(np.float32)
    s["ior"] = U(1.1, 2.2)
    s["roughness"] = U(0.02, 1.0)
    s["scatterDistance"] = U(0.5, 20.0)
    s["shader"] = capi.SHADER_DISNEY_SOLID
    for k, name in enumerate(fam):
        m = which == k
        c = int(m.sum())
        if c == 0:
            continue
        u = lambda lo=0.0, hi=1.0: rng.uniform(lo, hi, size=c).astype(np.float32)
        if name == "rough_diffuse":
            s["roughness"][m] = u(0.3, 1.0)
        elif name == "metal_aniso":
            s["metallic"][m] = 1.0
            s["anisotropic"][m] = u(0.0, 1.0)
            s["specularTint"][m] = u()
            s["roughness"][m] = u(0.05, 0.8)
        elif name == "clearcoat":
            s["clearcoat"][m] = u(0.1, 1.0)
            s["clearcoatGloss"][m] = u()
            s["metallic"][m] = u(0.0, 0.5)
        elif name == "sheen":
            s["sheen"][m] = u(0.1, 1.0)
            s["sheenTint"][m] = u()
        elif name == "glass_solid":
            s["specTrans"][m] = u(0.5, 1.0)
            s["materialFlags"][m] = capi.MAT_TRANSPARENT
            s["roughness"][m] = u(0.005, 0.5)
        elif name == "glass_thin":
            s["shader"][m] = capi.SHADER_DISNEY_THIN
            s["specTrans"][m] = u(0.5, 1.0)
            s["roughness"][m] = u(0.005, 0.6)
        elif name == "difftrans_thin":
            s["shader"][m] = capi.SHADER_DISNEY_THIN
            s["diffTrans"][m] = u(0.0, 0.5)
            s["flatness"][m] = u()
        elif name == "difftrans_solid":
            s["diffTrans"][m] = u(0.0, 0.5)
            s["materialFlags"][m] = capi.MAT_TRANSPARENT
        elif name == "mixed":
            for f in ("metallic", "specularTint", "anisotropic", "sheen", "sheenTint", "clearcoat", "clearcoatGloss", "specTrans",
                      "flatness"):
                s[f][m] = u()
            s["diffTrans"][m] = u(0.0, 0.5)
            s["shader"][m] = rng.integers(0, 2, size=c)
            s["materialFlags"][m] = np.where(s["shader"][m] == capi.SHADER_DISNEY_SOLID, capi.MAT_TRANSPARENT, 0)
        elif name == "copper":
            s["baseColor"][m] = [0.955008, 0.637427, 0.538163]
            s["metallic"][m] = 1.0
            s["specularTint"][m] = 0.3
            s["roughness"][m] = 0.01
            s["clearcoatGloss"][m] = 1.0
            s["ior"][m] = 1.5
        elif name == "default":
            s["baseColor"][m] = 0.6 ** 2.2
            s["metallic"][m] = 0.3
            s["roughness"][m] = 0.5
            s["ior"][m] = 1.5
        elif name == "dirac":
            s["shader"][m] = capi.SHADER_DIRAC_TRANSPARENT
    # view directions mostly on the normal's side, some below (inside transparent objects)
    v = unit_vectors(rng, n)
    flip = (np.sum(v * nrm, axis=1) < 0) & (rng.uniform(size=n) < 0.8)
    v[flip] *= -1
    s["view"] = v
    transparent = (s["materialFlags"] & capi.MAT_TRANSPARENT) != 0
    inside = np.sum(v * nrm, axis=1) < 0
    s["relativeIOR"] = np.where(transparent & inside, s["ior"], np.float32(1.0) / s["ior"]).astype(np.float32)
    return s, v, nrm


def light_directions(seed, n, nrm, view):
    """Light directions: mostly upper hemisphere, some transmission-side."""
    rng = np.random.default_rng(seed + 1000)
    l = unit_vectors(rng, n)
    flip = (np.sum(l * nrm, axis=1) < 0) & (rng.uniform(size=n) < 0.7)
    l[flip] *= -1
    return l


def uniforms(seed, n, k):
    rng = np.random.default_rng(seed + 2000)
    return rng.uniform(0.0, 1.0, size=(n, k)).astype(np.float32)


def random_lights(seed, count):
    rng = np.random.default_rng(seed + 3000)
    L = np.zeros(count, dtype=capi.LIGHT_DTYPE)
    for i in range(count):
        c = rng.uniform(-5, 5, size=3)
        c[1] = rng.uniform(4, 9)
        x = np.array([rng.uniform(0.5, 3.0), 0.0, rng.uniform(-0.5, 0.5)])
        z = np.array([rng.uniform(-0.5, 0.5), 0.0, rng.uniform(0.5, 3.0)])
        L[i] = scenes.quad_light(c, x, z, (0.0, -1.0, 0.0), rng.uniform(1, 50, size=3))
    return L


def small_scene(seed=7, with_attrs=True, instances=False):
    """A few hundred triangles/quads with normals+tangents on one mesh, none on the other: exercises every
    CalculateSurfaceParams branch."""
    rng = np.random.default_rng(seed)
    sph = scenes.displaced_sphere(12, 24, as_quads=True, with_normals=with_attrs)
    if with_attrs:
        n = sph["normals"].astype(np.float64)
        a = np.where(np.abs(n[:, 1:2]) > 0.9, np.array([[1.0, 0, 0]]), np.array([[0, 1.0, 0]]))
        t = np.cross(a, n)
        t /= np.linalg.norm(t, axis=1, keepdims=True)
        sph["tangents"] = np.concatenate([t, np.where(rng.uniform(size=(t.shape[0], 1)) < 0.5, -1.0, 1.0)], axis=1).astype(np.float32)
    tri = scenes.displaced_sphere(10, 16, as_quads=False)
    tri["positions"] = (tri["positions"] * np.float32(0.5) + np.array([[2.2, 0.3, 0.1]], np.float32)).astype(np.float32)
    tri["material"] = 1
    floor = dict(positions=np.array([[-6, -1.3, -6], [-6, -1.3, 6], [6, -1.3, 6], [6, -1.3, -6]], np.float32),
                 indices=np.array([[0, 1, 2, 3]], np.uint32), material=2, lightSet=0)
    mats = np.array([scenes.material(**scenes.DISNEY_MATERIAL_SET[1]), scenes.material(**scenes.DISNEY_MATERIAL_SET[3]),
                     scenes.material(**scenes.REFERENCE_MATERIALS["RoughDiffuse"])])
    lights = random_lights(seed, 3)
    inst = None
    if instances:
        def xform(tx, ty, tz, s, ang):
            c, sn = np.cos(ang), np.sin(ang)
            return np.array([[s * c, 0, -s * sn, 0], [0, s, 0, 0], [s * sn, 0, s * c, 0], [tx, ty, tz, 1]], np.float32)
        inst = [(0, np.eye(4, dtype=np.float32)), (1, np.eye(4, dtype=np.float32)), (2, np.eye(4, dtype=np.float32)),
                (0, xform(-2.5, 0.4, 0.5, 0.6, 0.7)), (1, xform(-1.0, 1.5, -1.0, 1.3, -0.4))]
    return capi.SceneData([sph, tri, floor], mats, lights, [(0, 3)], background=(0.2, 0.3, 0.4, 0.0), instances=inst)


def scene_rays(seed, n, radius=4.0):
    """Rays from a sphere of `radius` toward points near the origin, plus some grazing/axis-aligned ones."""
    rng = np.random.default_rng(seed + 4000)
    o = unit_vectors(rng, n) * np.float32(radius)
    target = rng.uniform(-1.5, 1.5, size=(n, 3)).astype(np.float32)
    d = target - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    k = n // 16
    d[:k] = 0
    d[:k, 1] = -1.0  # straight down: exercises zero direction components
    o[:k, 1] = 5.0
    o[:k, 0] = rng.uniform(-2, 2, size=k)
    o[:k, 2] = rng.uniform(-2, 2, size=k)
    rays = np.zeros(n, dtype=capi.RAY_DTYPE)
    rays["ox"], rays["oy"], rays["oz"] = o[:, 0], o[:, 1], o[:, 2]
    rays["dx"], rays["dy"], rays["dz"] = d[:, 0], d[:, 1], d[:, 2]
    rays["tnear"] = 0.0
    rays["tfar"] = np.float32(3.402823466e+38)
    return rays


def rel_err(a, b, floor=1e-6):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)


def mip_chain(size=64, comps=3, seed=3):
    """A square power-of-two texture with its full box-filtered mip chain down to 1 x 1 (what the reference's baker stores,
    TextureResourceData::mipCount levels), finest first."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32) / size
    base = np.stack([0.5 + 0.5 * np.sin(6.0 * xx + c) * np.cos(4.0 * yy - c) for c in range(comps)], axis=2).astype(np.float32)
    base = (base + 0.25 * rng.random(base.shape, dtype=np.float32)).astype(np.float32)
    levels = [base]
    while levels[-1].shape[0] > 1:
        a = levels[-1]
        levels.append((0.25 * (a[0::2, 0::2] + a[1::2, 0::2] + a[0::2, 1::2] + a[1::2, 1::2])).astype(np.float32))
    return levels


def filter_queries(n=4000, seed=4):
    """Lookups for the texture filters: st a little outside [0, 1] too (Repeat wrap, incl. negative coordinates), derivative vectors from
    sub-texel to a quarter of the texture, some degenerate (zero minor axis) and some strongly anisotropic."""
    rng = np.random.default_rng(seed)
    st = rng.uniform(-0.5, 1.5, size=(n, 2)).astype(np.float32)
    mag = np.exp(rng.uniform(np.log(1e-3), np.log(0.25), size=(n, 1))).astype(np.float32)
    ang = rng.uniform(0, 2 * np.pi, size=(n, 1)).astype(np.float32)
    d0 = (mag * np.concatenate([np.cos(ang), np.sin(ang)], axis=1)).astype(np.float32)
    ratio = np.exp(rng.uniform(np.log(1.0 / 64.0), 0.0, size=(n, 1))).astype(np.float32)
    d1 = (mag * ratio * np.concatenate([-np.sin(ang), np.cos(ang)], axis=1) + 0.1 * mag * ratio * rng.normal(size=(n, 2))).astype(np.float32)
    swap = rng.random(n) < 0.3
    d0[swap], d1[swap] = d1[swap].copy(), d0[swap].copy()
    d1[::97] = 0.0
    d0[::193] = 0.0
    return st, d0, d1
