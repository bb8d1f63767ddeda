"""ctypes binding of libselas_cuda.so (include/selas_cuda.h).

This is the Python-side stub a maintainer would write against the C ABI; tests and bench.py call the CUDA
path through it.  It never falls back to anything: if the shared library is missing, or no CUDA device is
usable, the constructors raise.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from pathlib import Path

import numpy as np

LIB_PATH = Path(__file__).resolve().parent / "lib" / "libselas_cuda.so"

SELAS_INVALID_ID = 0xFFFFFFFF
SELAS_BSDF_RANDOMS = 8
SHADER_DISNEY_THIN, SHADER_DISNEY_SOLID, SHADER_DIRAC_TRANSPARENT = 0, 1, 2
MAT_TRANSPARENT, MAT_ALPHA_TESTED, MAT_DISPLACEMENT, MAT_INVERT_DISPLACEMENT, MAT_USES_PTEX = 1, 2, 4, 8, 16
(METALLIC, SPECULAR_TINT, ROUGHNESS, ANISOTROPIC, SHEEN, SHEEN_TINT, CLEARCOAT, CLEARCOAT_GLOSS, SPEC_TRANS,
 DIFFUSE_TRANS, FLATNESS, IOR, SCATTER_DISTANCE, DISPLACEMENT, SCALAR_COUNT) = range(15)
EVENT_SCATTER, EVENT_TRANSMISSION, EVENT_DIRAC = 1, 2, 4

f32, u32, u64 = C.c_float, C.c_uint32, C.c_uint64


class Material(C.Structure):
    _fields_ = [("shader", u32), ("flags", u32), ("baseColor", f32 * 3), ("transmittanceColor", f32 * 3),
                ("scalars", f32 * SCALAR_COUNT)]


class Light(C.Structure):
    _fields_ = [("type", u32), ("position", f32 * 3), ("direction", f32 * 3), ("x", f32 * 3), ("z", f32 * 3),
                ("radiance", f32 * 3)]


class LightSetRange(C.Structure):
    _fields_ = [("start", u32), ("count", u32)]


class Camera(C.Structure):
    _fields_ = [("cameraX", f32 * 3), ("cameraY", f32 * 3), ("cameraZ", f32 * 3), ("position", f32 * 3),
                ("forward", f32 * 3), ("viewportWidth", f32), ("viewportHeight", f32), ("znear", f32), ("zfar", f32),
                ("pad0", u32), ("width", u64), ("height", u64), ("aspect", f32), ("virtualImagePlaneDistance", f32)]


class CameraSettings(C.Structure):
    _fields_ = [("position", f32 * 3), ("fovDegrees", f32), ("lookAt", f32 * 3), ("znear", f32), ("up", f32 * 3),
                ("zfar", f32)]


class TextureDesc(C.Structure):
    _fields_ = [("width", u32), ("height", u32), ("format", u32), ("reserved", u32), ("texels", C.c_void_p)]


class TextureMipsDesc(C.Structure):
    _fields_ = [("mipCount", u32), ("format", u32), ("mipWidths", u32 * 16), ("mipHeights", u32 * 16), ("mipOffsets", C.c_uint64 * 16),
                ("dataSize", C.c_uint64), ("texels", C.c_void_p)]


FILTER_POINT, FILTER_TRIANGLE, FILTER_TRILINEAR, FILTER_EWA, FILTER_EWA_LEVEL = range(5)


def texture_mips_desc(levels):
    """TextureResourceData (TextureLib/TextureResource.h:15-39) from a list of float32 arrays [h, w] or [h, w, comps], finest first; the
    levels are laid out back to back like the baker does.  Returns (desc, keepalive)."""
    levels = [np.ascontiguousarray(lv, np.float32) for lv in levels]
    levels = [lv[:, :, None] if lv.ndim == 2 else lv for lv in levels]
    comps = levels[0].shape[2]
    assert 1 <= comps <= 4 and all(lv.shape[2] == comps for lv in levels) and len(levels) <= 16
    blob = np.concatenate([lv.ravel() for lv in levels])
    d = TextureMipsDesc()
    d.mipCount, d.format = len(levels), comps - 1
    off = 0
    for i, lv in enumerate(levels):
        d.mipWidths[i], d.mipHeights[i], d.mipOffsets[i] = lv.shape[1], lv.shape[0], off
        off += lv.size * 4
    d.dataSize = off
    d.texels = blob.ctypes.data
    return d, blob


class MeshDesc(C.Structure):
    _fields_ = [("positions", C.c_void_p), ("vertexCount", u32), ("indices", C.c_void_p), ("faceCount", u32),
                ("indicesPerFace", u32), ("normals", C.c_void_p), ("tangents", C.c_void_p), ("uvs", C.c_void_p),
                ("materialIndex", u32), ("lightSetIndex", u32), ("ptex", C.c_void_p), ("baseColorTexture", C.POINTER(TextureDesc))]


class InstanceDesc(C.Structure):
    _fields_ = [("meshIndex", u32), ("localToWorld", f32 * 16), ("instID0", u32), ("instID1", u32), ("geomID", u32)]


class IblDesc(C.Structure):
    _fields_ = [("width", u32), ("height", u32), ("marginalDensityFunction", C.c_void_p), ("conditionalDensityFunctions", C.c_void_p),
                ("lightData", C.c_void_p), ("missWidth", u32), ("missHeight", u32), ("missData", C.c_void_p),
                ("rotationRadians", f32), ("exposureScale", f32)]


class CurveDesc(C.Structure):
    _fields_ = [("vertices", C.c_void_p), ("vertexCount", u32), ("indices", C.c_void_p), ("segmentCount", u32), ("materialIndex", u32),
                ("lightSetIndex", u32)]


class SceneDesc(C.Structure):
    _fields_ = [("meshes", C.POINTER(MeshDesc)), ("meshCount", u32), ("instances", C.POINTER(InstanceDesc)),
                ("instanceCount", u32), ("materials", C.POINTER(Material)), ("materialCount", u32),
                ("lights", C.POINTER(Light)), ("lightCount", u32), ("lightSets", C.POINTER(LightSetRange)),
                ("lightSetCount", u32), ("backgroundIntensity", f32 * 4), ("ibl", C.POINTER(IblDesc)), ("flags", u32), ("curveCount", u32),
                ("curves", C.POINTER(CurveDesc))]


class RenderSettings(C.Structure):
    _fields_ = [("width", u32), ("height", u32), ("samplesPerPixel", u32), ("maxPathLength", u32),
                ("rrStartBounce", u32), ("layerCount", u32), ("seed", u32), ("firstSample", u32), ("misMode", u32),
                ("flags", u32), ("pixelStride", u32), ("pixelSampler", u32), ("integrator", u32)]


class RenderStats(C.Structure):
    _fields_ = [("closestRays", u64), ("occlusionRays", u64), ("pathSamples", u64), ("kernelLaunches", u64),
                ("renderMs", f32), ("generateMs", f32), ("closestMs", f32), ("shadeMs", f32), ("shadowMs", f32),
                ("closestLaunches", u32), ("shadeLaunches", u32), ("shadowLaunches", u32),
                ("closestNodes", u64), ("closestTris", u64), ("shadowNodes", u64), ("shadowTris", u64)]


class VcmSettings(C.Structure):
    _fields_ = [("width", u32), ("height", u32), ("iterations", u32), ("firstIteration", u32), ("maxPathLength", u32),
                ("seed", u32), ("radiusFactor", f32), ("radiusAlpha", f32), ("boundingSphere", f32 * 4), ("flags", u32),
                ("reserved", u32)]


class VcmStats(C.Structure):
    _fields_ = [("lightPaths", u64), ("cameraPaths", u64), ("lightVertices", u64), ("closestRays", u64), ("occlusionRays", u64),
                ("mergeCandidates", u64), ("mergesAccepted", u64), ("connections", u64), ("kernelLaunches", u64),
                ("renderMs", f32), ("lightPassMs", f32), ("gridBuildMs", f32), ("cameraPassMs", f32), ("mergeMs", f32),
                ("connectMs", f32), ("lastRadius", f32)]


class BvhStats(C.Structure):
    _fields_ = [("triangleCount", u32), ("nodeCount", u32), ("maxDepth", u32), ("reserved", u32), ("buildMs", f32),
                ("sahCost", f32), ("deviceBytes", u64), ("sahClusters", u32), ("sahGenerations", u32)]


# numpy views of the plain-data records
RAY_DTYPE = np.dtype([("ox", "f4"), ("oy", "f4"), ("oz", "f4"), ("tnear", "f4"), ("dx", "f4"), ("dy", "f4"),
                      ("dz", "f4"), ("tfar", "f4")])
HIT_DTYPE = np.dtype([("t", "f4"), ("u", "f4"), ("v", "f4"), ("primID", "u4"), ("geomID", "u4"), ("instID0", "u4"),
                      ("instID1", "u4"), ("pad", "u4")])
SURFACE_DTYPE = np.dtype([("position", "f4", 3), ("worldToTangent", "f4", 9), ("error", "f4"), ("view", "f4", 3),
                          ("baseColor", "f4", 3), ("transmittanceColor", "f4", 3), ("sheen", "f4"), ("sheenTint", "f4"),
                          ("clearcoat", "f4"), ("clearcoatGloss", "f4"), ("metallic", "f4"), ("specTrans", "f4"),
                          ("diffTrans", "f4"), ("flatness", "f4"), ("anisotropic", "f4"), ("relativeIOR", "f4"),
                          ("specularTint", "f4"), ("roughness", "f4"), ("scatterDistance", "f4"), ("ior", "f4"),
                          ("shader", "u4"), ("materialFlags", "u4"), ("lightSetIndex", "u4")])
BSDF_SAMPLE_DTYPE = np.dtype([("flags", "u4"), ("mediumPhase", "u4"), ("mediumExtinction", "f4", 3),
                              ("reflectance", "f4", 3), ("wi", "f4", 3), ("forwardPdfW", "f4"), ("reversePdfW", "f4"),
                              ("valid", "u4"), ("randomsUsed", "u4")])
LIGHT_SAMPLE_DTYPE = np.dtype([("index", "u4"), ("direction", "f4", 3), ("radiance", "f4", 3), ("distance", "f4"),
                               ("pdfW", "f4")])
LIGHT_DTYPE = np.dtype([("type", "u4"), ("position", "f4", 3), ("direction", "f4", 3), ("x", "f4", 3), ("z", "f4", 3),
                        ("radiance", "f4", 3)])
VCM_VERTEX_DTYPE = np.dtype([("throughput", "f4", 3), ("indexAndLength", "u4"), ("dVCM", "f4"), ("dVC", "f4"), ("dVM", "f4"),
                             ("continuationProbability", "f4"), ("ray", RAY_DTYPE), ("hit", HIT_DTYPE)])
assert VCM_VERTEX_DTYPE.itemsize == 96
MATERIAL_DTYPE = np.dtype([("shader", "u4"), ("flags", "u4"), ("baseColor", "f4", 3), ("transmittanceColor", "f4", 3),
                           ("scalars", "f4", SCALAR_COUNT)])
assert RAY_DTYPE.itemsize == 32 and HIT_DTYPE.itemsize == 32 and SURFACE_DTYPE.itemsize == 156
assert BSDF_SAMPLE_DTYPE.itemsize == 60 and LIGHT_SAMPLE_DTYPE.itemsize == 36 and LIGHT_DTYPE.itemsize == 64
assert MATERIAL_DTYPE.itemsize == 88 == C.sizeof(Material) and C.sizeof(Camera) == 104

EXPORTS = [
    "selas_cuda_context_create", "selas_cuda_context_destroy", "selas_cuda_context_synchronize", "selas_cuda_last_error", "selas_cuda_device_count",
    "selas_cuda_version", "selas_cuda_scene_create", "selas_cuda_scene_destroy", "selas_cuda_scene_bvh_stats",
    "selas_cuda_intersect", "selas_cuda_occluded", "selas_cuda_intersect_device", "selas_cuda_occluded_device",
    "selas_cuda_intersect_counted", "selas_cuda_camera_init", "selas_cuda_camera_rays", "selas_cuda_cmj", "selas_cuda_bsdf_eval",
    "selas_cuda_bsdf_sample", "selas_cuda_light_sample", "selas_cuda_light_pdf", "selas_cuda_surface_params",
    "selas_cuda_ibl_lookup", "selas_cuda_ibl_sample", "selas_cuda_texture_sample", "selas_cuda_texture_filter",
    "selas_cuda_render_pathtracer", "selas_cuda_render_pathtracer_device", "selas_cuda_render_collect_stats", "selas_cuda_film_scale_device",
    "selas_cuda_reduce_film", "selas_cuda_render_vcm", "selas_cuda_render_vcm_device", "selas_cuda_vcm_light_pass", "selas_cuda_vcm_iteration_constants",
    "selas_cuda_hashgrid_build", "selas_cuda_hashgrid_destroy", "selas_cuda_hashgrid_read",
    "selas_cuda_hashgrid_query", "selas_cuda_hashgrid_build_device", "selas_cuda_hashgrid_query_device",
    "selas_cuda_ptex_load", "selas_cuda_ptex_load_memory", "selas_cuda_ptex_destroy", "selas_cuda_ptex_eval",
    "selas_cuda_ptex_eval_device",
]


class SelasError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"selas_cuda error {code}: {message}")
        self.code = code


_lib = None


def load_library() -> C.CDLL:
    """dlopen libselas_cuda.so; raises if it has not been built (there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("SELAS_CUDA_LIB", str(LIB_PATH))
    if not os.path.exists(path):
        raise ImportError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          f"or `make -C selas_b200/csrc` -- there is no CPU fallback")
    lib = C.CDLL(path)
    lib.selas_cuda_last_error.restype = C.c_char_p
    lib.selas_cuda_last_error.argtypes = [C.c_void_p]
    lib.selas_cuda_version.restype = C.c_char_p
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("selas_cuda_last_error", "selas_cuda_version"):
            fn.restype = C.c_int
    _lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


class Context:
    """rtcNewDevice / rtcReleaseDevice stand-in (main.cpp:99,:130)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self.handle = C.c_void_p()
        rc = self.lib.selas_cuda_context_create(C.c_int(device), C.byref(self.handle))
        if rc:
            raise SelasError(rc, self.lib.selas_cuda_last_error(None).decode())
        self.device = device
        self._children = weakref.WeakSet()  # scenes / grids / textures: destroyed before the context

    def adopt(self, child):
        self._children.add(child)

    def check(self, rc: int):
        if rc:
            raise SelasError(rc, self.lib.selas_cuda_last_error(self.handle).decode())

    def close(self):
        if self.handle:
            for child in list(self._children):
                child.close()
            self.lib.selas_cuda_context_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- batch operators -------------------------------------------------------------------------------------
    def camera_rays(self, camera: Camera, px, py, jitter=None):
        px, py = _f32(px), _f32(py)
        n = px.size
        jit = None if jitter is None else _f32(jitter, (n, 2))
        rays = np.zeros(n, dtype=RAY_DTYPE)
        self.check(self.lib.selas_cuda_camera_rays(self.handle, C.byref(camera), C.c_size_t(n), _ptr(px), _ptr(py),
                                                   _ptr(jit), _ptr(rays)))
        return rays

    def cmj(self, s, m, n, p):
        s = np.ascontiguousarray(s, dtype=np.int32)
        p = np.ascontiguousarray(p, dtype=np.int32)
        out = np.zeros((s.size, 2), np.float32)
        self.check(self.lib.selas_cuda_cmj(self.handle, C.c_size_t(s.size), _ptr(s), C.c_int32(m), C.c_int32(n), _ptr(p), _ptr(out)))
        return out

    def bsdf_eval(self, surfaces, v, l):
        surfaces = np.ascontiguousarray(surfaces, dtype=SURFACE_DTYPE)
        n = surfaces.size
        v, l = _f32(v, (n, 3)), _f32(l, (n, 3))
        rgb = np.zeros((n, 3), np.float32)
        fwd = np.zeros(n, np.float32)
        rev = np.zeros(n, np.float32)
        self.check(self.lib.selas_cuda_bsdf_eval(self.handle, C.c_size_t(n), _ptr(surfaces), _ptr(v), _ptr(l),
                                                 _ptr(rgb), _ptr(fwd), _ptr(rev)))
        return rgb, fwd, rev

    def bsdf_sample(self, surfaces, v, u):
        surfaces = np.ascontiguousarray(surfaces, dtype=SURFACE_DTYPE)
        n = surfaces.size
        v, u = _f32(v, (n, 3)), _f32(u, (n, SELAS_BSDF_RANDOMS))
        out = np.zeros(n, dtype=BSDF_SAMPLE_DTYPE)
        self.check(self.lib.selas_cuda_bsdf_sample(self.handle, C.c_size_t(n), _ptr(surfaces), _ptr(v), _ptr(u),
                                                   _ptr(out)))
        return out

    def texture_sample(self, texture, uv, srgb=True):
        """SampleTextureFloat3 (SurfaceParameters.cpp:97-118) = Triangle filter at level 0 (+ sRGB -> linear) over a batch of uvs."""
        a = texture_array(texture)
        td = texture_desc(a)
        uv = _f32(uv).reshape(-1, 2)
        out = np.zeros((uv.shape[0], 3), np.float32)
        self.check(self.lib.selas_cuda_texture_sample(self.handle, C.byref(td), C.c_size_t(uv.shape[0]), _ptr(uv), C.c_int(1 if srgb else 0),
                                                      _ptr(out)))
        return out

    def texture_filter(self, levels, filter, st, dst0=None, dst1=None, level=0):
        """TextureFiltering::Point / Triangle / Trilinear / EWA (TextureFiltering.h:54-234) over a batch; levels as texture_mips_desc."""
        d, keep = texture_mips_desc(levels)
        st = _f32(st).reshape(-1, 2)
        n = st.shape[0]
        dst0 = None if dst0 is None else _f32(dst0, (n, 2))
        dst1 = None if dst1 is None else _f32(dst1, (n, 2))
        out = np.zeros((n, d.format + 1), np.float32)
        self.check(self.lib.selas_cuda_texture_filter(self.handle, C.byref(d), C.c_int(filter), C.c_int(level), C.c_size_t(n), _ptr(st),
                                                      _ptr(dst0) if dst0 is not None else None, _ptr(dst1) if dst1 is not None else None, _ptr(out)))
        del keep
        return out

    def light_sample(self, lights, position, normal, u3):
        lights = np.ascontiguousarray(lights, dtype=LIGHT_DTYPE)
        position = _f32(position).reshape(-1, 3)
        n = position.shape[0]
        normal, u3 = _f32(normal, (n, 3)), _f32(u3, (n, 3))
        out = np.zeros(n, dtype=LIGHT_SAMPLE_DTYPE)
        self.check(self.lib.selas_cuda_light_sample(self.handle, C.c_size_t(n), _ptr(lights), u32(lights.size),
                                                    _ptr(position), _ptr(normal), _ptr(u3), _ptr(out)))
        return out

    def light_pdf(self, lights, index, position, wi):
        lights = np.ascontiguousarray(lights, dtype=LIGHT_DTYPE)
        index = np.ascontiguousarray(index, dtype=np.uint32)
        n = index.size
        position, wi = _f32(position, (n, 3)), _f32(wi, (n, 3))
        out = np.zeros(n, np.float32)
        self.check(self.lib.selas_cuda_light_pdf(self.handle, C.c_size_t(n), _ptr(lights), u32(lights.size),
                                                 _ptr(index), _ptr(position), _ptr(wi), _ptr(out)))
        return out


def camera_init(position, look_at, up, fov_degrees, width, height, znear=0.1, zfar=10000.0) -> Camera:
    """InitializeRayCastCamera (GeometryLib/Camera.cpp:89-115)."""
    lib = load_library()
    st = CameraSettings()
    st.position[:] = position
    st.lookAt[:] = look_at
    st.up[:] = up
    st.fovDegrees = fov_degrees
    st.znear, st.zfar = znear, zfar
    cam = Camera()
    rc = lib.selas_cuda_camera_init(C.byref(st), u32(width), u32(height), C.byref(cam))
    if rc:
        raise SelasError(rc, lib.selas_cuda_last_error(None).decode())
    return cam


SCENE_TWO_LEVEL_BVH = 1
SCENE_INSTANCE_IDS = 2


def texture_array(tex):
    a = np.ascontiguousarray(tex, dtype=np.float32)
    if a.ndim == 2:
        a = a[..., None]
    assert a.ndim == 3 and 1 <= a.shape[2] <= 4
    return np.ascontiguousarray(a)


def texture_desc(a):
    """SelasTextureDesc over a texture_array (the array must outlive the descriptor's use)."""
    td = TextureDesc()
    td.width, td.height, td.format, td.reserved = a.shape[1], a.shape[0], a.shape[2] - 1, 0
    td.texels = a.ctypes.data
    return td


class SceneData:
    """Host-side scene description (numpy) from which both the CUDA scene and the oracle scene are created.

    meshes: list of dicts {positions (V,3) f32, indices (F,3|4) u32, normals?, tangents?, uvs?, material, lightSet, ptex?}
    """

    def __init__(self, meshes, materials, lights=None, light_sets=None, background=(0, 0, 0, 0), instances=None, ibl=None, two_level=False,
                 instance_ids=None, curves=None):
        self.two_level = two_level  # SELAS_SCENE_TWO_LEVEL_BVH: per-mesh object-space trees under a top-level tree over the instances
        self.ibl = ibl  # dict: light (H,W,3), marginal (H,), conditional (H,W), miss (h,w,3), rotation, exposure_scale -- see scenes.build_ibl
        self.meshes = []
        for m in meshes:
            mm = dict(m)
            mm["positions"] = _f32(m["positions"]).reshape(-1, 3)
            idx = np.ascontiguousarray(m["indices"], dtype=np.uint32)
            assert idx.ndim == 2 and idx.shape[1] in (3, 4)
            mm["indices"] = idx
            for k, w in (("normals", 3), ("tangents", 4), ("uvs", 2)):
                mm[k] = None if m.get(k) is None else _f32(m[k]).reshape(-1, w)
            mm.setdefault("material", 0)
            mm.setdefault("lightSet", 0)
            mm.setdefault("ptex", None)
            tex = m.get("texture")  # (H, W) or (H, W, 1|2|3|4) float32: base colour texture, level 0 (SelasTextureDesc)
            mm["texture"] = None if tex is None else texture_array(tex)
            self.meshes.append(mm)
        self.materials = np.ascontiguousarray(materials, dtype=MATERIAL_DTYPE)
        self.lights = np.zeros(0, LIGHT_DTYPE) if lights is None else np.ascontiguousarray(lights, dtype=LIGHT_DTYPE)
        if light_sets is None:
            light_sets = [(0, self.lights.size)]
        self.light_sets = np.array(light_sets, dtype=np.uint32).reshape(-1, 2)
        self.background = tuple(float(x) for x in background)
        self.instances = instances  # list of (meshIndex, 4x4) or None
        # per instance (instID0, instID1, geomID): the reference's two instancing levels as hits report them (SELAS_SCENE_INSTANCE_IDS), or None
        self.instance_ids = instance_ids
        # curve geometries (RTC_GEOMETRY_TYPE_ROUND_BSPLINE_CURVE): dicts with vertices (n, 4) float32 {x, y, z, radius}, indices (first
        # control vertex per segment), material, lightSet; their geomID = len(meshes) + position in this list
        self.curves = [dict(c) for c in (curves or [])]
        for c in self.curves:
            c["vertices"] = np.ascontiguousarray(c["vertices"], np.float32).reshape(-1, 4)
            c["indices"] = np.ascontiguousarray(c["indices"], np.uint32).reshape(-1)

    def triangle_count(self):
        return sum(m["indices"].shape[0] * (2 if m["indices"].shape[1] == 4 else 1) for m in self.meshes)

    def to_desc(self):
        """Build a SelasSceneDesc; returns (desc, keepalive)."""
        keep = []
        md = (MeshDesc * max(1, len(self.meshes)))()
        for i, m in enumerate(self.meshes):
            md[i].positions = m["positions"].ctypes.data
            md[i].vertexCount = m["positions"].shape[0]
            md[i].indices = m["indices"].ctypes.data
            md[i].faceCount = m["indices"].shape[0]
            md[i].indicesPerFace = m["indices"].shape[1]
            md[i].normals = None if m["normals"] is None else m["normals"].ctypes.data
            md[i].tangents = None if m["tangents"] is None else m["tangents"].ctypes.data
            md[i].uvs = None if m["uvs"] is None else m["uvs"].ctypes.data
            md[i].materialIndex = m["material"]
            md[i].lightSetIndex = m["lightSet"]
            if m["texture"] is not None:
                td = texture_desc(m["texture"])
                keep.append(td)
                md[i].baseColorTexture = C.pointer(td)
            ptex = m["ptex"]
            md[i].ptex = None if ptex is None else (ptex.handle if hasattr(ptex, "handle") else ptex)
        desc = SceneDesc()
        desc.meshes = md
        desc.meshCount = len(self.meshes)
        if self.curves:
            cd = (CurveDesc * len(self.curves))()
            for i, c in enumerate(self.curves):
                cd[i].vertices, cd[i].vertexCount = c["vertices"].ctypes.data, c["vertices"].shape[0]
                cd[i].indices, cd[i].segmentCount = c["indices"].ctypes.data, c["indices"].size
                cd[i].materialIndex, cd[i].lightSetIndex = int(c.get("material", 0)), int(c.get("lightSet", 0))
            desc.curves, desc.curveCount = cd, len(self.curves)
            keep.append(cd)
        if self.instances:
            inst = (InstanceDesc * len(self.instances))()
            for i, (mi, mat) in enumerate(self.instances):
                inst[i].meshIndex = mi
                inst[i].localToWorld[:] = list(np.asarray(mat, dtype=np.float32).reshape(16))
                ids = self.instance_ids[i] if self.instance_ids is not None else (i, 0, mi)
                inst[i].instID0, inst[i].instID1, inst[i].geomID = (int(x) for x in ids)
            desc.instances = inst
            desc.instanceCount = len(self.instances)
            keep.append(inst)
        desc.materials = C.cast(self.materials.ctypes.data, C.POINTER(Material))
        desc.materialCount = self.materials.size
        desc.lights = C.cast(self.lights.ctypes.data, C.POINTER(Light)) if self.lights.size else None
        desc.lightCount = self.lights.size
        desc.lightSets = C.cast(self.light_sets.ctypes.data, C.POINTER(LightSetRange))
        desc.lightSetCount = self.light_sets.shape[0]
        desc.backgroundIntensity[:] = self.background
        desc.flags = (SCENE_TWO_LEVEL_BVH if getattr(self, "two_level", False) else 0) | (SCENE_INSTANCE_IDS if self.instance_ids is not None else 0)
        if self.ibl is not None:
            i = self.ibl
            for k in ("light", "marginal", "conditional", "miss"):
                i[k] = np.ascontiguousarray(i[k], dtype=np.float32)
            d = IblDesc()
            d.height, d.width = i["light"].shape[0], i["light"].shape[1]
            d.marginalDensityFunction = i["marginal"].ctypes.data
            d.conditionalDensityFunctions = i["conditional"].ctypes.data
            d.lightData = i["light"].ctypes.data
            d.missHeight, d.missWidth = i["miss"].shape[0], i["miss"].shape[1]
            d.missData = i["miss"].ctypes.data
            d.rotationRadians = i.get("rotation", 0.0)
            d.exposureScale = i.get("exposure_scale", 1.0)
            desc.ibl = C.pointer(d)
            keep.append(d)
        keep.append(md)
        return desc, keep


class Scene:
    """Device scene: replaces the Embree scene build (ModelResource.cpp:164-200,:461-468; SceneResource.cpp:218-249)."""

    def __init__(self, ctx: Context, data: SceneData):
        self.ctx = ctx
        self.lib = ctx.lib
        self.data = data
        desc, keep = data.to_desc()
        self.handle = C.c_void_p()
        ctx.check(self.lib.selas_cuda_scene_create(ctx.handle, C.byref(desc), C.byref(self.handle)))
        ctx.adopt(self)
        del keep

    def close(self):
        if self.handle:
            self.lib.selas_cuda_scene_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def bvh_stats(self) -> BvhStats:
        st = BvhStats()
        self.ctx.check(self.lib.selas_cuda_scene_bvh_stats(self.handle, C.byref(st)))
        return st

    def intersect(self, rays, want_ng=False):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.zeros(rays.size, dtype=HIT_DTYPE)
        ng = np.zeros((rays.size, 4), np.float32) if want_ng else None
        self.ctx.check(self.lib.selas_cuda_intersect(self.handle, C.c_size_t(rays.size), _ptr(rays), _ptr(hits), _ptr(ng)))
        return (hits, ng) if want_ng else hits

    def occluded(self, rays):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        occ = np.zeros(rays.size, dtype=np.uint8)
        self.ctx.check(self.lib.selas_cuda_occluded(self.handle, C.c_size_t(rays.size), _ptr(rays), _ptr(occ)))
        return occ

    def intersect_counted(self, rays, any_hit=False):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.zeros(rays.size, dtype=HIT_DTYPE)
        counts = np.zeros((rays.size, 2), dtype=np.uint32)
        self.ctx.check(self.lib.selas_cuda_intersect_counted(self.handle, C.c_size_t(rays.size), _ptr(rays), _ptr(hits),
                                                             _ptr(counts), C.c_int(1 if any_hit else 0)))
        return hits, counts

    def surface_params(self, rays, hits):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.ascontiguousarray(hits, dtype=HIT_DTYPE)
        out = np.zeros(rays.size, dtype=SURFACE_DTYPE)
        self.ctx.check(self.lib.selas_cuda_surface_params(self.handle, C.c_size_t(rays.size), _ptr(rays), _ptr(hits), _ptr(out)))
        return out

    def ibl_lookup(self, wi):
        wi = _f32(wi).reshape(-1, 3)
        n = wi.shape[0]
        rgb, pdf, miss = np.zeros((n, 3), np.float32), np.zeros(n, np.float32), np.zeros((n, 3), np.float32)
        self.ctx.check(self.lib.selas_cuda_ibl_lookup(self.handle, C.c_size_t(n), _ptr(wi), _ptr(rgb), _ptr(pdf), _ptr(miss)))
        return rgb, pdf, miss

    def ibl_sample(self, u2):
        u2 = _f32(u2).reshape(-1, 2)
        n = u2.shape[0]
        tpp, xy = np.zeros((n, 3), np.float32), np.zeros((n, 2), np.uint32)
        self.ctx.check(self.lib.selas_cuda_ibl_sample(self.handle, C.c_size_t(n), _ptr(u2), _ptr(tpp), _ptr(xy)))
        return tpp, xy

    def render(self, camera: Camera, settings: RenderSettings):
        """PathTracer::GenerateImage (PathTracer.cpp:301-350) with host film buffers; returns (layers, stats).
        layers[l] is (H, W, 3) float32 holding the SUM over samples (FrameBuffer_Scale is the caller's)."""
        w, h, layers = settings.width, settings.height, settings.layerCount
        film = [np.zeros((h, w, 3), np.float32) for _ in range(layers)]
        ptrs = (C.c_void_p * layers)(*[f.ctypes.data for f in film])
        stats = RenderStats()
        self.ctx.check(self.lib.selas_cuda_render_pathtracer(self.handle, C.byref(camera), C.byref(settings), ptrs, C.byref(stats)))
        return film, stats

    def render_device(self, camera: Camera, settings: RenderSettings, d_film_ptr: int, stream: int = 0, want_stats: bool = True):
        """want_stats=False: the call only enqueues and returns while the GPU works (collect_stats() later)."""
        stats = RenderStats() if want_stats else None
        self.ctx.check(self.lib.selas_cuda_render_pathtracer_device(self.handle, C.byref(camera), C.byref(settings), C.c_void_p(d_film_ptr),
                                                                    C.byref(stats) if want_stats else None, C.c_void_p(stream)))
        return stats

    def collect_stats(self) -> RenderStats:
        """Waits for the enqueued renders; per-kernel times, ray totals and samples since the last collect."""
        stats = RenderStats()
        self.ctx.check(self.lib.selas_cuda_render_collect_stats(self.handle, C.byref(stats)))
        return stats

    def render_vcm(self, camera: Camera, settings: VcmSettings):
        """VCM::GenerateImage (VCM.cpp:738-807) with a host film; returns ((H, W, 3) SUM over iterations, stats)."""
        film = np.zeros((settings.height, settings.width, 3), np.float32)
        stats = VcmStats()
        self.ctx.check(self.lib.selas_cuda_render_vcm(self.handle, C.byref(camera), C.byref(settings), _ptr(film), C.byref(stats)))
        return film, stats

    def render_vcm_device(self, camera: Camera, settings: VcmSettings, d_film_ptr: int, stream: int = 0):
        stats = VcmStats()
        self.ctx.check(self.lib.selas_cuda_render_vcm_device(self.handle, C.byref(camera), C.byref(settings), C.c_void_p(d_film_ptr),
                                                             C.byref(stats), C.c_void_p(stream)))
        return stats

    def vcm_light_pass(self, camera: Camera, settings: VcmSettings):
        """Light pass of iteration settings.firstIteration: (vertices in the reference's order, pathEnds)."""
        vm_count = settings.width * settings.height
        total = C.c_size_t(0)
        ends = np.zeros(vm_count, np.uint32)
        self.ctx.check(self.lib.selas_cuda_vcm_light_pass(self.handle, C.byref(camera), C.byref(settings), None, C.c_size_t(0), _ptr(ends),
                                                          C.byref(total)))
        verts = np.zeros(total.value, dtype=VCM_VERTEX_DTYPE)
        if total.value:
            self.ctx.check(self.lib.selas_cuda_vcm_light_pass(self.handle, C.byref(camera), C.byref(settings), _ptr(verts),
                                                              C.c_size_t(total.value), _ptr(ends), C.byref(total)))
        return verts, ends

    def intersect_device(self, n, d_rays, d_hits, d_ng=0, stream=0):
        self.ctx.check(self.lib.selas_cuda_intersect_device(self.handle, C.c_size_t(n), C.c_void_p(d_rays), C.c_void_p(d_hits),
                                                            C.c_void_p(d_ng) if d_ng else None, C.c_void_p(stream)))

    def occluded_device(self, n, d_rays, d_occ, stream=0):
        self.ctx.check(self.lib.selas_cuda_occluded_device(self.handle, C.c_size_t(n), C.c_void_p(d_rays), C.c_void_p(d_occ),
                                                           C.c_void_p(stream)))


RENDER_COUNT_TRAVERSAL = 1
RENDER_SKIP_NONFINITE = 2


INTEGRATOR_PATHTRACER, INTEGRATOR_DEFERRED = 0, 1


def render_settings(width, height, spp, max_path_length=8, rr_start=8, layers=1, seed=1, first_sample=0, mis_mode=0,
                    flags=0, pixel_stride=0, cmj=None, integrator=0):
    rs = RenderSettings()
    rs.integrator = integrator
    rs.flags, rs.pixelStride = flags, pixel_stride
    rs.pixelSampler = 0 if cmj is None else (cmj[0] | (cmj[1] << 16))
    rs.width, rs.height, rs.samplesPerPixel = width, height, spp
    rs.maxPathLength, rs.rrStartBounce, rs.layerCount = max_path_length, rr_start, layers
    rs.seed, rs.firstSample, rs.misMode = seed, first_sample, mis_mode
    return rs


(VCM_NO_LIGHT_TRACING, VCM_NO_DIRECT, VCM_NO_CONNECTIONS, VCM_NO_MERGING, VCM_NO_SKY, VCM_SKIP_NONFINITE) = 1, 2, 4, 8, 16, 32


def vcm_settings(width, height, iterations, bounding_sphere, first_iteration=0, max_path_length=10, seed=1, radius_factor=0.0025,
                 radius_alpha=0.75, flags=0):
    """Defaults are the reference's macros: MaxBounceCount_ 10, VcmRadiusFactor_ 0.0025, VcmRadiusAlpha_ 0.75 (VCM.cpp:28-35)."""
    vs = VcmSettings()
    vs.width, vs.height, vs.iterations, vs.firstIteration = width, height, iterations, first_iteration
    vs.maxPathLength, vs.seed, vs.radiusFactor, vs.radiusAlpha, vs.flags = max_path_length, seed, radius_factor, radius_alpha, flags
    vs.boundingSphere[:] = [float(x) for x in bounding_sphere]
    return vs


def vcm_iteration_constants(vm_count, vc_count, base_radius, alpha, iteration):
    out = np.zeros(5, np.float32)
    rc = load_library().selas_cuda_vcm_iteration_constants(u32(vm_count), u32(vc_count), f32(base_radius), f32(alpha), f32(iteration), _ptr(out))
    if rc:
        raise SelasError(rc, "selas_cuda_vcm_iteration_constants")
    return out


class HashGrid:
    """BuildHashGrid / SearchHashGrid (VCMHashGrid.cpp:48-155)."""

    def __init__(self, ctx: Context, positions, radius: float, cell_count: int):
        self.ctx = ctx
        self.lib = ctx.lib
        positions = _f32(positions).reshape(-1, 3)
        self.n = positions.shape[0]
        self.cell_count = cell_count
        self.handle = C.c_void_p()
        ctx.check(self.lib.selas_cuda_hashgrid_build(ctx.handle, C.c_size_t(self.n), _ptr(positions), f32(radius),
                                                     u32(cell_count), C.byref(self.handle)))
        ctx.adopt(self)

    def read(self):
        ends = np.zeros(self.cell_count, np.uint32)
        idx = np.zeros(self.n, np.uint32)
        aabb = np.zeros(6, np.float32)
        self.ctx.check(self.lib.selas_cuda_hashgrid_read(self.handle, _ptr(ends), _ptr(idx), _ptr(aabb)))
        return ends, idx, aabb

    def query(self, queries, max_per_query=0):
        queries = _f32(queries).reshape(-1, 3)
        m = queries.shape[0]
        counts = np.zeros(m, np.uint32)
        indices = np.full((m, max_per_query), 0xFFFFFFFF, np.uint32) if max_per_query else None
        self.ctx.check(self.lib.selas_cuda_hashgrid_query(self.handle, C.c_size_t(m), _ptr(queries), _ptr(counts),
                                                          _ptr(indices), u32(max_per_query)))
        return counts, indices

    def close(self):
        if self.handle:
            self.lib.selas_cuda_hashgrid_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Ptex:
    """PtexCache::get + PtexFilter::getFilter(f_bspline) (TextureCache.cpp:99-126, SurfaceParameters.cpp:196-208)."""

    def __init__(self, ctx: Context, path: str):
        self.ctx = ctx
        self.lib = ctx.lib
        self.handle = C.c_void_p()
        ctx.check(self.lib.selas_cuda_ptex_load(ctx.handle, str(path).encode(), C.byref(self.handle)))
        ctx.adopt(self)

    def eval(self, faceid, u, v):
        faceid = np.ascontiguousarray(faceid, dtype=np.int32)
        u, v = _f32(u), _f32(v)
        out = np.zeros((faceid.size, 3), np.float32)
        self.ctx.check(self.lib.selas_cuda_ptex_eval(self.handle, C.c_size_t(faceid.size), _ptr(faceid), _ptr(u), _ptr(v), _ptr(out)))
        return out

    def eval_device(self, n, d_faceid, d_u, d_v, d_rgb, stream=0):
        self.ctx.check(self.lib.selas_cuda_ptex_eval_device(self.handle, C.c_size_t(n), C.c_void_p(d_faceid), C.c_void_p(d_u),
                                                            C.c_void_p(d_v), C.c_void_p(d_rgb), C.c_void_p(stream)))

    def close(self):
        if self.handle:
            self.lib.selas_cuda_ptex_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
