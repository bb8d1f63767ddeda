"""Synthetic inputs for the five BASELINE.json configs plus the material-JSON import.

Everything here is host-side input preparation (numpy): it produces a capi.SceneData that is handed both to
the CUDA scene (capi.Scene) and, in tests/bench, to the CPU oracle.  Nothing in this file computes any
part of the hot path.
"""
from __future__ import annotations

import json
import math

import numpy as np

from . import capi

# ImportMaterial defaults, BuildCommon/ImportMaterial.cpp:36-52
ATTRIBUTE_NAMES = ["metallic", "specularTint", "roughness", "anisotropic", "sheen", "sheenTint", "clearcoat",
                   "clearcoatGloss", "specTrans", "diffTrans", "flatness", "ior", "scatterDistance", "displacement"]
ATTRIBUTE_DEFAULTS = [0.3, 0.0, 0.5, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.5, 0.0, 0.0]
SHADER_NAMES = ["DisneyThin", "DisneySolid", "DeltaTransparent"]  # BuildModel.cpp:24-28


def material(shaderName="DisneySolid", baseColor=(0.6, 0.6, 0.6), transmittanceColor=(0.0, 0.0, 0.0), usesPtex=False,
             alphaTesting=False, invertDisplacement=False, **attrs):
    """ImportMaterial (ImportMaterial.cpp:55-90) + BuildMaterial (BuildModel.cpp:41-66) for an in-memory dict."""
    m = np.zeros((), dtype=capi.MATERIAL_DTYPE)
    shader = capi.SHADER_DISNEY_SOLID
    for i, name in enumerate(SHADER_NAMES):
        if name.lower() == str(shaderName).lower():
            shader = i
    m["shader"] = shader
    if isinstance(baseColor, str):  # texture path: constant falls back to 0.6 (ImportMaterial.cpp:65-67)
        baseColor = (0.6, 0.6, 0.6)
    m["baseColor"] = baseColor
    m["transmittanceColor"] = transmittanceColor
    scal = list(ATTRIBUTE_DEFAULTS)
    for k, v in attrs.items():
        if k not in ATTRIBUTE_NAMES:
            raise KeyError(f"unknown material attribute {k!r}")
        if isinstance(v, str):
            continue  # scalar attribute textures are not used by the path tracer
        scal[ATTRIBUTE_NAMES.index(k)] = float(v)
    m["scalars"] = scal
    flags = 0
    if alphaTesting:
        flags |= capi.MAT_ALPHA_TESTED
    if invertDisplacement:
        flags |= capi.MAT_INVERT_DISPLACEMENT
    if usesPtex:
        flags |= capi.MAT_USES_PTEX
    if shader == capi.SHADER_DISNEY_SOLID and (scal[capi.DIFFUSE_TRANS] > 0.0 or scal[capi.SPEC_TRANS] > 0.0):
        flags |= capi.MAT_TRANSPARENT
    m["flags"] = flags
    return m


def material_from_json(text_or_path):
    """Content/Materials/*.json -> 88-byte material payload."""
    s = str(text_or_path)
    if not s.lstrip().startswith("{"):
        with open(s) as f:
            s = f.read()
    d = json.loads(s) if s.strip() else {}
    return material(**d)


# The six material files the reference ships (Content/Materials/*.json), restated as dicts so tests and the
# benchmark do not need the reference tree at run time.
REFERENCE_MATERIALS = {
    "BSDF": dict(shaderName="DisneySolid", baseColor=[0.9, 1.0, 0.9], transmittanceColor=[0.4519, 1.0, 0.7054], ior=1.7,
                 metallic=0.0, specularTint=0.0, roughness=0.01, anisotropic=0.0, sheen=0.0, sheenTint=0.0, clearcoat=0.0,
                 clearcoatGloss=0.0, specTrans=1.0, diffTrans=0.0, flatness=0.0, scatterDistance=15.0),
    "Copper": dict(shaderName="DisneySolid", baseColor=[0.955008, 0.637427, 0.538163], ior=1.5, metallic=1.0,
                   specularTint=0.3, roughness=0.01, anisotropic=0.0, sheen=0.0, sheenTint=0.0, clearcoat=0.0,
                   clearcoatGloss=1.0, specTrans=0.0, diffTrans=0.0, flatness=0.0),
    "Default": dict(),
    "DiffuseTransmissionThin": dict(shaderName="DisneyThin", baseColor=[0.9, 0.62, 0.9], ior=1.5, metallic=0.0,
                                    specularTint=0.0, roughness=0.4, anisotropic=0.0, sheen=0.0, sheenTint=0.0,
                                    clearcoat=0.0, clearcoatGloss=1.0, specTrans=0.0, diffTrans=1.0, flatness=0.0),
    "Paper": dict(shaderName="DisneySolid", baseColor="Textures~Paper~paper02_col.jpg", metallic=0.0, roughness=0.8),
    "RoughDiffuse": dict(shaderName="DisneySolid", baseColor=[0.82, 0.39, 0.37], ior=1.2, metallic=0.0, specularTint=0.0,
                         roughness=0.8, anisotropic=0.0, sheen=0.0, sheenTint=0.0, clearcoat=0.0, clearcoatGloss=0.0,
                         specTrans=0.0, diffTrans=0.0, flatness=0.5),
}


def quad_light(center, x, z, direction, radiance):
    """SceneLight (SceneResource.h:32-40): position is the quad centre, x/z the full edge vectors."""
    L = np.zeros((), dtype=capi.LIGHT_DTYPE)
    L["type"] = 0
    L["position"], L["x"], L["z"], L["direction"], L["radiance"] = center, x, z, direction, radiance
    return L


def _quad(p0, p1, p2, p3):
    return [p0, p1, p2, p3]


# ---- C1: Cornell box -------------------------------------------------------------------------------------------
def cornell_box(copper_roughness=0.01):
    """Config 1: the classic 555-unit box as quad meshes, one mesh per material.
    Winding is chosen so geometric normals face the inside of the room / outside of the blocks."""
    def box(lo, hi, rot_deg, centre_xz):
        # axis-aligned block rotated about y, returns 5 quads (no bottom)
        sx, sy, sz = hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]
        c, s = math.cos(math.radians(rot_deg)), math.sin(math.radians(rot_deg))
        def P(x, y, z):
            x -= sx / 2
            z -= sz / 2
            return (centre_xz[0] + c * x + s * z, lo[1] + y, centre_xz[1] - s * x + c * z)
        v = [P(0, 0, 0), P(sx, 0, 0), P(sx, 0, sz), P(0, 0, sz), P(0, sy, 0), P(sx, sy, 0), P(sx, sy, sz), P(0, sy, sz)]
        return [
            _quad(v[4], v[7], v[6], v[5]),  # top (+y)
            _quad(v[0], v[4], v[5], v[1]),  # -z side
            _quad(v[1], v[5], v[6], v[2]),  # +x side
            _quad(v[2], v[6], v[7], v[3]),  # +z side
            _quad(v[3], v[7], v[4], v[0]),  # -x side
        ]

    X, Y, Z = 556.0, 548.8, 559.2
    white_quads = [
        _quad((0, 0, 0), (0, 0, Z), (X, 0, Z), (X, 0, 0)),      # floor, normal +y
        _quad((0, Y, 0), (X, Y, 0), (X, Y, Z), (0, Y, Z)),      # ceiling, normal -y
        _quad((0, 0, Z), (0, Y, Z), (X, Y, Z), (X, 0, Z)),      # back wall, normal -z
    ]
    white_quads += box((0, 0, 0), (165, 330, 165), 17.0, (368.0, 351.0))   # tall block (white)
    red_quads = [_quad((X, 0, 0), (X, 0, Z), (X, Y, Z), (X, Y, 0))]         # left wall (x = 556), normal -x
    green_quads = [_quad((0, 0, 0), (0, Y, 0), (0, Y, Z), (0, 0, Z))]       # right wall (x = 0), normal +x
    copper_quads = box((0, 0, 0), (165, 165, 165), -18.0, (185.0, 169.0))  # short block

    def mesh(quads, mat):
        pos = np.array([p for q in quads for p in q], dtype=np.float32)
        idx = np.arange(len(quads) * 4, dtype=np.uint32).reshape(-1, 4)
        return dict(positions=pos, indices=idx, material=mat, lightSet=0)

    rough = dict(REFERENCE_MATERIALS["RoughDiffuse"])
    mats = [
        material(**{**rough, "baseColor": [0.725, 0.71, 0.68]}),
        material(**{**rough, "baseColor": [0.63, 0.065, 0.05]}),
        material(**{**rough, "baseColor": [0.14, 0.45, 0.091]}),
        material(**{**REFERENCE_MATERIALS["Copper"], "roughness": copper_roughness}),
    ]
    lights = [quad_light((278.0, 548.7, 279.5), (130.0, 0.0, 0.0), (0.0, 0.0, 105.0), (0.0, -1.0, 0.0), (17.0, 12.0, 4.0))]
    data = capi.SceneData([mesh(white_quads, 0), mesh(red_quads, 1), mesh(green_quads, 2), mesh(copper_quads, 3)],
                          np.array(mats), np.array(lights), [(0, 1)], background=(0, 0, 0, 0))
    return data


def cornell_camera(width=512, height=512):
    return capi.camera_init((278.0, 273.0, -800.0), (278.0, 273.0, 0.0), (0.0, 1.0, 0.0), 37.0, width, height)


# ---- C2: displaced sphere, 1M triangles --------------------------------------------------------------------------
def _hash01(i):
    i = np.asarray(i, dtype=np.uint32).copy()
    i ^= i >> np.uint32(16)
    i *= np.uint32(0x7feb352d)
    i ^= i >> np.uint32(15)
    i *= np.uint32(0x846ca68b)
    i ^= i >> np.uint32(16)
    return (i >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def displaced_sphere(n_theta=500, n_phi=1000, as_quads=True, with_normals=False, material_index=0):
    """Lat-long sphere of n_theta x n_phi quads (default 500x1000 = 1,000,000 triangles), radius
    1 + 0.05 sin(8 theta) sin(8 phi) + 0.01 hash01(vertexIndex ^ 0x5E1A5)."""
    th = np.linspace(0.0, math.pi, n_theta + 1, dtype=np.float64)
    ph = np.linspace(0.0, 2.0 * math.pi, n_phi + 1, dtype=np.float64)[:-1]
    T, P = np.meshgrid(th, ph, indexing="ij")
    vid = np.arange(T.size, dtype=np.uint32).reshape(T.shape)
    r = 1.0 + 0.05 * np.sin(8 * T) * np.sin(8 * P) + 0.01 * _hash01(vid ^ np.uint32(0x5E1A5)).astype(np.float64)
    # poles: one radius for the whole ring so the mesh closes
    r[0, :] = r[0, 0]
    r[-1, :] = r[-1, 0]
    pos = np.stack([r * np.sin(T) * np.cos(P), r * np.cos(T), r * np.sin(T) * np.sin(P)], axis=-1).astype(np.float32)
    i0 = vid[:-1, :]
    i1 = vid[1:, :]
    i0n = np.roll(i0, -1, axis=1)
    i1n = np.roll(i1, -1, axis=1)
    # outward-facing (CCW seen from outside): (theta,phi) -> (theta, phi+1) -> (theta+1, phi+1) -> (theta+1, phi)
    quads = np.stack([i0, i0n, i1n, i1], axis=-1).reshape(-1, 4).astype(np.uint32)
    mesh = dict(positions=pos.reshape(-1, 3), material=material_index, lightSet=0)
    if as_quads:
        mesh["indices"] = quads
    else:
        mesh["indices"] = np.concatenate([quads[:, [0, 1, 3]], quads[:, [2, 3, 1]]], axis=0)
    if with_normals:
        n = pos.reshape(-1, 3).astype(np.float64)
        n /= np.linalg.norm(n, axis=1, keepdims=True)
        mesh["normals"] = n.astype(np.float32)
    return mesh


def raycast_scene(n_theta=500, n_phi=1000):
    mats = np.array([material(**REFERENCE_MATERIALS["RoughDiffuse"])])
    lights = np.array([quad_light((5.0, 5.0, 5.0), (1.0, 0.0, -1.0), (-1.0, 2.0, -1.0), (-0.57735, -0.57735, -0.57735),
                                  (40.0, 40.0, 40.0))])
    return capi.SceneData([displaced_sphere(n_theta, n_phi)], mats, lights, [(0, 1)], background=(0.5, 0.7, 1.0, 0.0))


def raycast_camera(width=1920, height=1080):
    return capi.camera_init((0.0, 0.0, 3.2), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 40.0, width, height)


def primary_rays(camera, width, height):
    """Pixel-centre rays computed on the host in fp32 exactly like ImageToWorldDirection (Camera.cpp:56-60).
    Used only to *author inputs* for the ray-cast microbench; parity of the device camera kernel is tested apart."""
    f = np.float32
    ys, xs = np.meshgrid(np.arange(height, dtype=np.float32), np.arange(width, dtype=np.float32), indexing="ij")
    x = (xs + f(0.5)).ravel()
    y = (ys + f(0.5)).ravel()
    cx = (x / f(width)) * f(2.0) - f(1.0)
    cy = f(1.0) - f(2.0) * y / f(height)
    X, Y, Z = (np.array(camera.cameraX[:], np.float32), np.array(camera.cameraY[:], np.float32),
               np.array(camera.cameraZ[:], np.float32))
    d = Z[None, :] - cx[:, None] * X[None, :] + cy[:, None] * Y[None, :]
    inv = f(1.0) / np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2], dtype=np.float32)
    d = d * inv[:, None]
    rays = np.zeros(x.size, dtype=capi.RAY_DTYPE)
    rays["ox"], rays["oy"], rays["oz"] = camera.position[0], camera.position[1], camera.position[2]
    rays["tnear"] = 0.0
    rays["dx"], rays["dy"], rays["dz"] = d[:, 0], d[:, 1], d[:, 2]
    rays["tfar"] = np.float32(3.402823466e+38)
    return rays


def shadow_rays_toward(rays, hits, ng, target):
    """One occlusion ray per hit toward `target` with the reference's offset rule (PathTracer.cpp:71-92,
    SurfaceParameters.cpp:403-408): origin = p +- 0.1*error*N, tnear = error, tfar = distance."""
    f = np.float32
    hit = hits["geomID"] != capi.SELAS_INVALID_ID
    t = hits["t"][hit]
    o = np.stack([rays["ox"][hit], rays["oy"][hit], rays["oz"][hit]], 1)
    d = np.stack([rays["dx"][hit], rays["dy"][hit], rays["dz"][hit]], 1)
    p = (o + t[:, None] * d).astype(np.float32)
    n = ng[hit, :3].astype(np.float32)
    n = n / np.linalg.norm(n, axis=1, keepdims=True).astype(np.float32)
    err = f(32.0 * 1.19209e-07) * np.maximum(np.max(np.abs(p), axis=1), t)
    tv = (np.asarray(target, np.float32)[None, :] - p).astype(np.float32)
    dist = np.linalg.norm(tv, axis=1).astype(np.float32)
    l = (tv / dist[:, None]).astype(np.float32)
    sign = np.where(np.sum(l * n, axis=1) < 0, f(-1.0), f(1.0))
    org = p + (sign * err * f(0.1))[:, None] * n
    out = np.zeros(p.shape[0], dtype=capi.RAY_DTYPE)
    out["ox"], out["oy"], out["oz"] = org[:, 0], org[:, 1], org[:, 2]
    out["tnear"] = err
    out["dx"], out["dy"], out["dz"] = l[:, 0], l[:, 1], l[:, 2]
    out["tfar"] = dist
    return out


def cosine_bounce_rays(rays, hits, ng, seed=2):
    """Incoherent second-pass rays: cosine-sampled bounce from each hit (counter RNG keyed by seed)."""
    f = np.float32
    hit = hits["geomID"] != capi.SELAS_INVALID_ID
    t = hits["t"][hit]
    o = np.stack([rays["ox"][hit], rays["oy"][hit], rays["oz"][hit]], 1)
    d = np.stack([rays["dx"][hit], rays["dy"][hit], rays["dz"][hit]], 1)
    p = (o + t[:, None] * d).astype(np.float32)
    n = ng[hit, :3].astype(np.float64)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    n = np.where((np.sum(n * d, axis=1) > 0)[:, None], -n, n)
    idx = np.arange(p.shape[0], dtype=np.uint32)
    u1 = _hash01(idx * np.uint32(2) + np.uint32(seed * 7919)).astype(np.float64)
    u2 = _hash01(idx * np.uint32(2) + np.uint32(1 + seed * 7919)).astype(np.float64)
    r, phi = np.sqrt(u1), 2 * math.pi * u2
    a = np.where(np.abs(n[:, 0:1]) > 0.9, np.array([[0.0, 1.0, 0.0]]), np.array([[1.0, 0.0, 0.0]]))
    tx = np.cross(n, a)
    tx /= np.linalg.norm(tx, axis=1, keepdims=True)
    ty = np.cross(n, tx)
    w = (r * np.cos(phi))[:, None] * tx + (r * np.sin(phi))[:, None] * ty + np.sqrt(np.maximum(0, 1 - u1))[:, None] * n
    w /= np.linalg.norm(w, axis=1, keepdims=True)
    err = f(32.0 * 1.19209e-07) * np.maximum(np.max(np.abs(p), axis=1), t)
    org = p + (err[:, None] * n).astype(np.float32)
    out = np.zeros(p.shape[0], dtype=capi.RAY_DTYPE)
    out["ox"], out["oy"], out["oz"] = org[:, 0], org[:, 1], org[:, 2]
    out["tnear"] = 0.0
    out["dx"], out["dy"], out["dz"] = w[:, 0].astype(np.float32), w[:, 1].astype(np.float32), w[:, 2].astype(np.float32)
    out["tfar"] = np.float32(3.402823466e+38)
    return out


# ---- C3: full-Disney heightfield -------------------------------------------------------------------------------
def _fbm(x, y, seed, octaves=5):
    def noise(ix, iy, s):
        h = (ix.astype(np.uint32) * np.uint32(73856093)) ^ (iy.astype(np.uint32) * np.uint32(19349663)) ^ np.uint32(s * 83492791 & 0xFFFFFFFF)
        return _hash01(h).astype(np.float64)
    z = np.zeros_like(x, dtype=np.float64)
    amp, freq = 1.0, 1.0
    for o in range(octaves):
        fx, fy = x * freq, y * freq
        ix, iy = np.floor(fx), np.floor(fy)
        tx, ty = fx - ix, fy - iy
        tx, ty = tx * tx * (3 - 2 * tx), ty * ty * (3 - 2 * ty)
        ix, iy = ix.astype(np.int64), iy.astype(np.int64)
        n00, n10 = noise(ix, iy, seed + o), noise(ix + 1, iy, seed + o)
        n01, n11 = noise(ix, iy + 1, seed + o), noise(ix + 1, iy + 1, seed + o)
        z += amp * ((n00 * (1 - tx) + n10 * tx) * (1 - ty) + (n01 * (1 - tx) + n11 * tx) * ty)
        amp *= 0.5
        freq *= 2.0
    return z


DISNEY_MATERIAL_SET = [
    dict(shaderName="DisneySolid", baseColor=[0.9, 0.85, 0.6], metallic=1.0, roughness=0.2, anisotropic=0.5),
    dict(shaderName="DisneySolid", baseColor=[0.7, 0.1, 0.1], metallic=0.0, roughness=0.5, clearcoat=1.0, clearcoatGloss=0.9),
    dict(shaderName="DisneySolid", baseColor=[0.2, 0.3, 0.8], metallic=0.0, roughness=0.8, sheen=1.0, sheenTint=0.5),
    dict(shaderName="DisneySolid", baseColor=[0.9, 1.0, 0.9], transmittanceColor=[0.45, 1.0, 0.7], metallic=0.0, roughness=0.05,
         specTrans=1.0, ior=1.5, scatterDistance=15.0),
    dict(shaderName="DisneyThin", baseColor=[0.8, 0.9, 1.0], metallic=0.0, roughness=0.1, specTrans=1.0, ior=1.5),
    dict(shaderName="DisneyThin", baseColor=[0.9, 0.62, 0.9], metallic=0.0, roughness=0.4, diffTrans=1.0, flatness=0.5),
    REFERENCE_MATERIALS["RoughDiffuse"],
    REFERENCE_MATERIALS["Copper"],
]


def disney_heightfield(grid=2236, patch=64, seed=3, spheres=9, sphere_res=(50, 100)):
    """Config 3: grid x grid heightfield (2 triangles per cell; 2236 -> 9,990,450 + spheres ~ 10M triangles) split into
    one mesh per material (8 materials cycling per patch x patch block) + `spheres` displaced spheres."""
    n = grid + 1
    xs = np.linspace(-10.0, 10.0, n)
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    Zh = 1.5 * _fbm(X * 0.4 + 17.0, Y * 0.4 + 5.0, seed) - 1.5
    pos = np.stack([X, Zh, Y], axis=-1).astype(np.float32).reshape(-1, 3)
    # smooth normals from central differences
    dzdx = np.gradient(Zh, xs, axis=0)
    dzdy = np.gradient(Zh, xs, axis=1)
    nrm = np.stack([-dzdx, np.ones_like(Zh), -dzdy], axis=-1)
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    nrm = nrm.astype(np.float32).reshape(-1, 3)
    ci, cj = np.meshgrid(np.arange(grid), np.arange(grid), indexing="ij")
    v00 = (ci * n + cj).astype(np.uint32)
    v10 = v00 + np.uint32(n)
    v01 = v00 + np.uint32(1)
    v11 = v10 + np.uint32(1)
    matid = (((ci // patch) + (cj // patch) * 3) % len(DISNEY_MATERIAL_SET)).ravel()
    # CCW seen from +y
    triA = np.stack([v00, v01, v11], axis=-1).reshape(-1, 3)
    triB = np.stack([v00, v11, v10], axis=-1).reshape(-1, 3)
    meshes = []
    for m in range(len(DISNEY_MATERIAL_SET)):
        sel = matid == m
        idx = np.concatenate([triA[sel], triB[sel]], axis=0)
        meshes.append(dict(positions=pos, normals=nrm, indices=idx, material=m, lightSet=0))
    rng = np.random.default_rng(seed)
    for s in range(spheres):
        sm = displaced_sphere(sphere_res[0], sphere_res[1], as_quads=False, with_normals=False, material_index=(s * 3 + 3) % 8)
        c = np.array([rng.uniform(-7, 7), rng.uniform(0.8, 2.0), rng.uniform(-7, 7)], np.float32)
        sm["positions"] = (sm["positions"] * np.float32(0.7) + c[None, :]).astype(np.float32)
        meshes.append(sm)
    mats = np.array([material(**d) for d in DISNEY_MATERIAL_SET])
    lights = np.array([
        quad_light((-6.0, 8.0, -6.0), (3.0, 0.0, 0.0), (0.0, 0.0, 3.0), (0.0, -1.0, 0.0), (60.0, 55.0, 50.0)),
        quad_light((6.0, 8.0, -6.0), (3.0, 0.0, 0.0), (0.0, 0.0, 3.0), (0.0, -1.0, 0.0), (50.0, 55.0, 60.0)),
        quad_light((-6.0, 8.0, 6.0), (3.0, 0.0, 0.0), (0.0, 0.0, 3.0), (0.0, -1.0, 0.0), (55.0, 60.0, 50.0)),
        quad_light((6.0, 8.0, 6.0), (3.0, 0.0, 0.0), (0.0, 0.0, 3.0), (0.0, -1.0, 0.0), (60.0, 50.0, 55.0)),
    ])
    return capi.SceneData(meshes, mats, lights, [(0, 4)], background=(0.5, 0.7, 1.0, 0.0))


HEIGHTFIELD_CAMERA = dict(position=(0.0, 7.0, -16.0), look_at=(0.0, -0.5, 0.0), up=(0.0, 1.0, 0.0), fov=50.0)


def heightfield_camera(width=1920, height=1080):
    c = HEIGHTFIELD_CAMERA
    return capi.camera_init(c["position"], c["look_at"], c["up"], c["fov"], width, height)


# ---- C4: Ptex-textured quad grid ---------------------------------------------------------------------------------
def ptex_quad_grid(nu=1581, nv=1581, ptex=None, seed=4):
    """Config 4: nu x nv quads (1581^2 = 2,499,561 quads ~ 5M triangles) on a gently rolling sheet; face id of quad (i,j)
    is j*nu + i, matching the grid adjacency ptexref_write_grid writes."""
    xs = np.linspace(-10.0, 10.0, nu + 1)
    ys = np.linspace(-10.0, 10.0, nv + 1)
    X, Y = np.meshgrid(xs, ys, indexing="xy")  # [j, i]
    Z = 0.6 * np.sin(X * 0.7) * np.cos(Y * 0.5) + 0.25 * _fbm(X * 0.8 + 3.0, Y * 0.8 + 11.0, seed, 4)
    pos = np.stack([X, Z, Y], axis=-1).astype(np.float32).reshape(-1, 3)
    jj, ii = np.meshgrid(np.arange(nv), np.arange(nu), indexing="ij")
    v0 = (jj * (nu + 1) + ii).astype(np.uint32)
    # Ptex u runs along edge 0 (v0 -> v1) = +x = i, v along +z = j, so (u, v) of a hit and the face id j*nu + i line up with
    # the grid adjacency of the .ptx; with this winding the geometric normal is -y: camera and light sit below the sheet.
    quads = np.stack([v0, v0 + np.uint32(1), v0 + np.uint32(nu + 2), v0 + np.uint32(nu + 1)], axis=-1).reshape(-1, 4)
    mats = np.array([material(shaderName="DisneySolid", baseColor=[0.6, 0.6, 0.6], metallic=0.0, roughness=0.8,
                              usesPtex=ptex is not None)])
    lights = np.array([quad_light((0.0, -9.0, 0.0), (8.0, 0.0, 0.0), (0.0, 0.0, 8.0), (0.0, 1.0, 0.0), (40.0, 40.0, 40.0))])
    mesh = dict(positions=pos, indices=quads, material=0, lightSet=0, ptex=ptex)
    return capi.SceneData([mesh], mats, lights, [(0, 1)], background=(0.3, 0.35, 0.4, 0.0))


def ptex_grid_camera(width=1920, height=1080):
    return capi.camera_init((0.0, -9.0, -14.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 55.0, width, height)


# ---- image based light (config 5 and the background of any scene) ----------------------------------------------------
def build_ibl(light, miss=None, rotation=0.0, exposure=0.0):
    """The bake step's density functions for a lat-long environment image (BuildCommon/BuildImageBasedLight.cpp:24-105):
    intensity = sin(theta) * CIE luma, per-row inclusive CDFs normalised to 1, marginal inclusive CDF over the row sums."""
    light = np.ascontiguousarray(light, dtype=np.float32)
    h, w = light.shape[:2]
    theta = ((np.arange(h, dtype=np.float32) + np.float32(0.5)) * np.float32(math.pi) / np.float32(h)).astype(np.float32)
    luma = light[..., 0] * np.float32(0.299) + light[..., 1] * np.float32(0.587) + light[..., 2] * np.float32(0.114)
    inten = (np.sin(theta).astype(np.float32)[:, None] * luma).astype(np.float32)
    rows = np.cumsum(inten, axis=1, dtype=np.float32)
    row_sum = rows[:, -1].copy()
    cond = np.zeros((h, w), np.float32)
    nz = row_sum > 0
    cond[nz] = np.cumsum((np.float32(1.0) / row_sum[nz])[:, None] * inten[nz], axis=1, dtype=np.float32)
    cond[:, -1] = 1.0
    total = np.cumsum(row_sum, dtype=np.float32)[-1]
    marg = np.cumsum(row_sum * (np.float32(1.0) / total), dtype=np.float32) if total > 0 else row_sum.copy()
    marg[-1] = 1.0
    return dict(light=light, marginal=marg.astype(np.float32), conditional=cond, miss=light if miss is None else np.ascontiguousarray(miss, np.float32),
                rotation=float(rotation), exposure_scale=float(2.0 ** exposure))


def sun_sky_ibl(width=1024, height=512, sun_theta_deg=40.0, sun_nits=5000.0, sky=0.2):
    """sky of constant radiance with a 16x16-texel sun at theta = 40 degrees; a dimmer copy as the miss image."""
    img = np.full((height, width, 3), sky, np.float32)
    y0 = int(sun_theta_deg / 180.0 * height) - 8
    x0 = width // 3
    img[y0:y0 + 16, x0:x0 + 16] = sun_nits
    miss = np.full((height // 4, width // 4, 3), sky * 0.5, np.float32)
    return build_ibl(img, miss=miss, rotation=0.3, exposure=0.0)


# ---- C5: caustic scene for the VCM integrator ----------------------------------------------------------------------
def _uv_sphere(center, radius, n_theta, n_phi, material_index):
    th = np.linspace(0.0, math.pi, n_theta + 1, dtype=np.float64)
    ph = np.linspace(0.0, 2.0 * math.pi, n_phi + 1, dtype=np.float64)[:-1]
    T, P = np.meshgrid(th, ph, indexing="ij")
    vid = np.arange(T.size, dtype=np.uint32).reshape(T.shape)
    n = np.stack([np.sin(T) * np.cos(P), np.cos(T), np.sin(T) * np.sin(P)], axis=-1)
    pos = (np.asarray(center, np.float64)[None, None, :] + radius * n).astype(np.float32)
    i0, i1 = vid[:-1, :], vid[1:, :]
    quads = np.stack([i0, np.roll(i0, -1, axis=1), np.roll(i1, -1, axis=1), i1], axis=-1).reshape(-1, 4).astype(np.uint32)
    return dict(positions=pos.reshape(-1, 3), normals=n.astype(np.float32).reshape(-1, 3), indices=quads, material=material_index, lightSet=0)


def caustic_scene(sphere_res=(96, 192), floor_cells=64, ibl_size=(1024, 512)):
    """Config 5: floor + glass sphere (Content/Materials/BSDF.json: specTrans 1, ior 1.7, roughness 0.01) +
    diffuse box, lit only by a synthetic sun-and-sky IBL (the reference's VCM emits from the IBL alone, VCMCommon.cpp:37-63).
    No quad lights: the light-set range is empty."""
    g = floor_cells
    xs = np.linspace(-6.0, 6.0, g + 1)
    X, Z = np.meshgrid(xs, xs, indexing="ij")
    fpos = np.stack([X, np.zeros_like(X), Z], axis=-1).astype(np.float32).reshape(-1, 3)
    ci, cj = np.meshgrid(np.arange(g), np.arange(g), indexing="ij")
    v00 = (ci * (g + 1) + cj).astype(np.uint32)
    # CCW seen from +y
    fquads = np.stack([v00, v00 + np.uint32(1), v00 + np.uint32(g + 2), v00 + np.uint32(g + 1)], axis=-1).reshape(-1, 4)
    floor = dict(positions=fpos, indices=fquads, material=0, lightSet=0)
    sphere = _uv_sphere((0.0, 1.2, 0.0), 1.2, sphere_res[0], sphere_res[1], 1)

    def box(lo, hi):
        x0, y0, z0 = lo
        x1, y1, z1 = hi
        v = [(x0, y0, z0), (x1, y0, z0), (x1, y0, z1), (x0, y0, z1), (x0, y1, z0), (x1, y1, z0), (x1, y1, z1), (x0, y1, z1)]
        quads = [_quad(v[4], v[7], v[6], v[5]), _quad(v[0], v[4], v[5], v[1]), _quad(v[1], v[5], v[6], v[2]),
                 _quad(v[2], v[6], v[7], v[3]), _quad(v[3], v[7], v[4], v[0])]
        pos = np.array([p for q in quads for p in q], dtype=np.float32)
        return dict(positions=pos, indices=np.arange(20, dtype=np.uint32).reshape(-1, 4), material=2, lightSet=0)

    rough = dict(REFERENCE_MATERIALS["RoughDiffuse"])
    mats = np.array([
        material(**{**rough, "baseColor": [0.8, 0.8, 0.78]}),
        material(**REFERENCE_MATERIALS["BSDF"]),
        material(**rough),
    ])
    ibl = sun_sky_ibl(ibl_size[0], ibl_size[1])
    return capi.SceneData([floor, sphere, box((2.2, 0.0, -1.0), (3.6, 1.6, 0.6))], mats, None, None, background=(0, 0, 0, 0), ibl=ibl)


def caustic_camera(width=1920, height=1080):
    return capi.camera_init((0.0, 4.5, -9.0), (0.4, 0.6, 0.0), (0.0, 1.0, 0.0), 45.0, width, height)


def bounding_sphere(data):
    """SceneResource::boundingSphere as {centre of the scene's AABB, half its diagonal} (the reference declares the field,
    SceneResource.h:78, and reads it in AreaLighting.cpp:274-275 / VCM.cpp:778; the baker's value is not in the tree)."""
    if getattr(data, "instances", None):
        # CalculateSceneBoundingBox (SceneResource.cpp:194-203): each instance's object-space box through its localToWorld (IncludeBox)
        pts = []
        for mi, mat in data.instances:
            p = np.asarray(data.meshes[mi]["positions"], np.float64)
            if p.size == 0:
                continue
            blo, bhi = p.min(axis=0), p.max(axis=0)
            corners = np.array([[bhi[0] if c & 1 else blo[0], bhi[1] if c & 2 else blo[1], bhi[2] if c & 4 else blo[2], 1.0] for c in range(8)])
            pts.append((corners @ np.asarray(mat, np.float64).reshape(4, 4))[:, :3])
        pts = np.concatenate(pts, axis=0)
        lo, hi = pts.min(axis=0), pts.max(axis=0)
    else:
        lo = np.min([m["positions"].min(axis=0) for m in data.meshes], axis=0).astype(np.float64)
        hi = np.max([m["positions"].max(axis=0) for m in data.meshes], axis=0).astype(np.float64)
    c = 0.5 * (lo + hi)
    return (float(c[0]), float(c[1]), float(c[2]), float(0.5 * np.linalg.norm(hi - lo)))


# ---- curves: a patch of grass (RTC_GEOMETRY_TYPE_ROUND_BSPLINE_CURVE, the geometry the reference's island scene grows everywhere) ----
def grass_patch(blades=400, segments=3, seed=9, extent=4.0, height=(0.6, 1.4), radius=(0.03, 0.004), with_floor=True):
    """`blades` tapered B-spline blades of `segments` segments each over a floor quad, one quad light above.  Returns SceneData with
    one curve geometry (geomID = number of meshes).  Control vertices: x, y, z, radius; segment i of a blade starts at its i-th vertex."""
    rng = np.random.default_rng(seed)
    nv = segments + 3
    verts, indices = [], []
    for b in range(blades):
        base = np.array([rng.uniform(-extent, extent), 0.0, rng.uniform(-extent, extent)])
        h = rng.uniform(*height)
        lean = rng.normal(scale=0.35, size=2)
        sway = rng.normal(scale=0.15, size=2)
        first = len(verts)
        for k in range(nv):
            s = (k - 1) / (nv - 3)  # the spline runs from s = 0 (vertex 1) to s = 1 (vertex nv - 2); the outer vertices are phantom ends
            p = base + np.array([lean[0] * s * s + sway[0] * np.sin(3.0 * s), h * s, lean[1] * s * s + sway[1] * np.sin(2.0 * s + 1.0)])
            r = radius[0] + (radius[1] - radius[0]) * min(max(s, 0.0), 1.0)
            verts.append([p[0], p[1], p[2], r])
        indices += [first + k for k in range(segments)]
    meshes = []
    if with_floor:
        e = extent * 1.5
        meshes.append(dict(positions=np.array([[-e, 0, -e], [-e, 0, e], [e, 0, e], [e, 0, -e]], np.float32), indices=np.array([[0, 1, 2, 3]], np.uint32),
                           material=0, lightSet=0))
    mats = np.array([material(**REFERENCE_MATERIALS["RoughDiffuse"]), material(shaderName="DisneySolid", baseColor=[0.25, 0.6, 0.15], roughness=0.6, sheen=0.4)])
    lights = np.array([quad_light((0.0, 6.0, 0.0), (3.0, 0.0, 0.0), (0.0, 0.0, 3.0), (0.0, -1.0, 0.0), (40.0, 38.0, 34.0))])
    curves = [dict(vertices=np.array(verts, np.float32), indices=np.array(indices, np.uint32), material=1, lightSet=0)]
    return capi.SceneData(meshes, mats, lights, [(0, 1)], background=(0.3, 0.45, 0.7, 0.0), curves=curves)


def grass_camera(width=320, height=180):
    return capi.camera_init((0.0, 2.2, -7.5), (0.0, 0.5, 0.0), (0.0, 1.0, 0.0), 45.0, width, height)
