"""ctypes bindings of the CPU checkers under oracle/ -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
  * liboracle.so            the restatement (oracle/selas_oracle.cpp, orc_shading.h, cpu_bvh.h)        prefix orc_
  * _ref/libselas_ref.so    the UNMODIFIED reference translation units + a thin harness                prefix ref_
  * _ref/libptex_ref.so     the reference's vendored Ptex v2.2 + harness                               prefix ptexref_
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

from selas_b200 import capi

ORACLE_DIR = Path(__file__).resolve().parent.parent / "oracle"


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a if shape is None else a.reshape(shape)


def build_oracle():
    """make -C oracle (also builds oracle/_ref when /root/reference is present)."""
    subprocess.run(["make", "-C", str(ORACLE_DIR), "-j8"], check=True, stdout=subprocess.DEVNULL)


class _Lib:
    def __init__(self, path, prefix):
        self.lib = C.CDLL(str(path))
        self.prefix = prefix

    def fn(self, name):
        f = getattr(self.lib, self.prefix + name)
        f.restype = C.c_int
        return f


class ShadingOracle(_Lib):
    """Batch shading operators shared by liboracle.so (orc_) and libselas_ref.so (ref_)."""

    def bsdf_eval(self, surfaces, v, l):
        surfaces = np.ascontiguousarray(surfaces, dtype=capi.SURFACE_DTYPE)
        n = surfaces.size
        v, l = _f32(v, (n, 3)), _f32(l, (n, 3))
        rgb, fwd, rev = np.zeros((n, 3), np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)
        self.fn("bsdf_eval")(C.c_size_t(n), _ptr(surfaces), _ptr(v), _ptr(l), _ptr(rgb), _ptr(fwd), _ptr(rev))
        return rgb, fwd, rev

    def bsdf_sample(self, surfaces, v, u):
        surfaces = np.ascontiguousarray(surfaces, dtype=capi.SURFACE_DTYPE)
        n = surfaces.size
        v, u = _f32(v, (n, 3)), _f32(u, (n, capi.SELAS_BSDF_RANDOMS))
        out = np.zeros(n, dtype=capi.BSDF_SAMPLE_DTYPE)
        self.fn("bsdf_sample")(C.c_size_t(n), _ptr(surfaces), _ptr(v), _ptr(u), _ptr(out))
        return out

    def light_sample(self, lights, position, normal, u3):
        lights = np.ascontiguousarray(lights, dtype=capi.LIGHT_DTYPE)
        position = _f32(position).reshape(-1, 3)
        n = position.shape[0]
        normal, u3 = _f32(normal, (n, 3)), _f32(u3, (n, 3))
        out = np.zeros(n, dtype=capi.LIGHT_SAMPLE_DTYPE)
        self.fn("light_sample")(C.c_size_t(n), _ptr(lights), C.c_uint32(lights.size), _ptr(position), _ptr(normal), _ptr(u3), _ptr(out))
        return out

    def light_pdf(self, lights, index, position, wi):
        lights = np.ascontiguousarray(lights, dtype=capi.LIGHT_DTYPE)
        index = np.ascontiguousarray(index, dtype=np.uint32)
        n = index.size
        position, wi = _f32(position, (n, 3)), _f32(wi, (n, 3))
        out = np.zeros(n, np.float32)
        self.fn("light_pdf")(C.c_size_t(n), _ptr(lights), C.c_uint32(lights.size), _ptr(index), _ptr(position), _ptr(wi), _ptr(out))
        return out

    def camera_init(self, position, look_at, up, fov, width, height, znear=0.1, zfar=10000.0):
        st = capi.CameraSettings()
        st.position[:], st.lookAt[:], st.up[:] = position, look_at, up
        st.fovDegrees, st.znear, st.zfar = fov, znear, zfar
        cam = capi.Camera()
        self.fn("camera_init")(C.byref(st), C.c_uint32(width), C.c_uint32(height), C.byref(cam))
        return cam

    def camera_rays(self, camera, px, py, jitter=None):
        px, py = _f32(px), _f32(py)
        n = px.size
        jit = None if jitter is None else _f32(jitter, (n, 2))
        rays = np.zeros(n, dtype=capi.RAY_DTYPE)
        self.fn("camera_rays")(C.byref(camera), C.c_size_t(n), _ptr(px), _ptr(py), _ptr(jit), _ptr(rays))
        return rays

    def cmj(self, s, m, n, p):
        s = np.ascontiguousarray(s, dtype=np.int32)
        p = np.ascontiguousarray(p, dtype=np.int32)
        out = np.zeros((s.size, 2), np.float32)
        self.fn("cmj")(C.c_size_t(s.size), _ptr(s), C.c_int32(m), C.c_int32(n), _ptr(p), _ptr(out))
        return out

    def medium(self, extinction, u2, dist):
        extinction = _f32(extinction).reshape(-1, 3)
        n = extinction.shape[0]
        u2, dist = _f32(u2, (n, 2)), _f32(dist, (n,))
        sampled, trans = np.zeros(n, np.float32), np.zeros((n, 3), np.float32)
        self.fn("medium")(C.c_size_t(n), _ptr(extinction), _ptr(u2), _ptr(dist), _ptr(sampled), _ptr(trans))
        return sampled, trans

    # scene-level ---------------------------------------------------------------------------------------------
    def scene_create(self, data: capi.SceneData):
        desc, keep = data.to_desc()
        h = C.c_void_p()
        rc = self.fn("scene_create")(C.byref(desc), C.byref(h))
        if rc:
            raise RuntimeError(f"{self.prefix}scene_create failed: {rc}")
        return h

    def scene_destroy(self, h):
        self.fn("scene_destroy")(h)

    def scene_set_ptex(self, h, mesh_index, ptexref, ptex_handle):
        """Bind mesh `mesh_index`'s Ptex base colour to the reference Ptex library's lookup (PtexRef.open handle)."""
        rc = self.fn("scene_set_ptex")(h, C.c_uint32(mesh_index), C.cast(ptexref.fn("eval1"), C.c_void_p), C.c_void_p(ptex_handle))
        assert rc == 0, rc

    def ibl_lookup(self, h, wi):
        wi = _f32(wi).reshape(-1, 3)
        n = wi.shape[0]
        rgb, pdf, miss = np.zeros((n, 3), np.float32), np.zeros(n, np.float32), np.zeros((n, 3), np.float32)
        rc = self.fn("ibl_lookup")(h, C.c_size_t(n), _ptr(wi), _ptr(rgb), _ptr(pdf), _ptr(miss))
        assert rc == 0, rc
        return rgb, pdf, miss

    def ibl_sample(self, h, u2):
        u2 = _f32(u2).reshape(-1, 2)
        n = u2.shape[0]
        tpp, xy = np.zeros((n, 3), np.float32), np.zeros((n, 2), np.uint32)
        rc = self.fn("ibl_sample")(h, C.c_size_t(n), _ptr(u2), _ptr(tpp), _ptr(xy))
        assert rc == 0, rc
        return tpp, xy

    def texture_sample(self, texture, uv, srgb=True):
        a = capi.texture_array(texture)
        td = capi.texture_desc(a)
        uv = _f32(uv).reshape(-1, 2)
        out = np.zeros((uv.shape[0], 3), np.float32)
        rc = self.fn("texture_sample")(C.byref(td), C.c_size_t(uv.shape[0]), _ptr(uv), C.c_int(1 if srgb else 0), _ptr(out))
        assert rc == 0, rc
        return out

    def texture_filter(self, levels, filter, st, dst0=None, dst1=None, level=0):
        """TextureFiltering::Point / Triangle / Trilinear / EWA (TextureFiltering.h:54-234); levels as capi.texture_mips_desc."""
        d, keep = capi.texture_mips_desc(levels)
        st = _f32(st).reshape(-1, 2)
        n = st.shape[0]
        dst0 = None if dst0 is None else _f32(dst0, (n, 2))
        dst1 = None if dst1 is None else _f32(dst1, (n, 2))
        out = np.zeros((n, d.format + 1), np.float32)
        rc = self.fn("texture_filter")(C.byref(d), C.c_int(filter), C.c_int(level), C.c_size_t(n), _ptr(st),
                                       _ptr(dst0) if dst0 is not None else None, _ptr(dst1) if dst1 is not None else None, _ptr(out))
        assert rc == 0, rc
        del keep
        return out

    # VCM building blocks (live code in the reference: AreaLighting.cpp:246-331, VCMCommon.cpp:37-112) ------------------
    def ibl_emit(self, h, sphere4, u4):
        u4 = _f32(u4).reshape(-1, 4)
        n = u4.shape[0]
        sphere4 = _f32(sphere4, (4,))
        pdr, pdfs = np.zeros((n, 9), np.float32), np.zeros((n, 3), np.float32)
        rc = self.fn("ibl_emit")(h, _ptr(sphere4), C.c_size_t(n), _ptr(u4), _ptr(pdr), _ptr(pdfs))
        assert rc == 0, rc
        return pdr, pdfs

    def ibl_direct(self, h, u2):
        u2 = _f32(u2).reshape(-1, 2)
        out = np.zeros(u2.shape[0], dtype=capi.LIGHT_SAMPLE_DTYPE)
        rc = self.fn("ibl_direct")(h, C.c_size_t(u2.shape[0]), _ptr(u2), _ptr(out))
        assert rc == 0, rc
        return out

    def ibl_radiance(self, h, sphere4, dirs):
        dirs = _f32(dirs).reshape(-1, 3)
        n = dirs.shape[0]
        sphere4 = _f32(sphere4, (4,))
        rgb, pdfs = np.zeros((n, 3), np.float32), np.zeros((n, 2), np.float32)
        rc = self.fn("ibl_radiance")(h, _ptr(sphere4), C.c_size_t(n), _ptr(dirs), _ptr(rgb), _ptr(pdfs))
        assert rc == 0, rc
        return rgb, pdfs

    def vcm_iteration_constants(self, vm_count, vc_count, base_radius, alpha, iteration):
        out = np.zeros(5, np.float32)
        self.fn("vcm_iteration_constants")(C.c_uint32(vm_count), C.c_uint32(vc_count), C.c_float(base_radius), C.c_float(alpha),
                                           C.c_float(iteration), _ptr(out))
        return out

    def render(self, h, camera, settings, nthreads=8):
        w, hgt, layers = settings.width, settings.height, settings.layerCount
        film = [np.zeros((hgt, w, 3), np.float32) for _ in range(layers)]
        ptrs = (C.c_void_p * layers)(*[f.ctypes.data for f in film])
        stats = capi.RenderStats()
        rc = self.fn("render_pathtracer")(h, C.byref(camera), C.byref(settings), ptrs, C.byref(stats), C.c_int(nthreads))
        if rc:
            raise RuntimeError(f"{self.prefix}render_pathtracer failed: {rc}")
        return film, stats


class Oracle(ShadingOracle):
    def __init__(self, path):
        super().__init__(path, "orc_")
        self.lib.orc_scene_triangle_count.restype = C.c_size_t

    def intersect(self, h, rays, mode=1, nthreads=8, want_ng=False):
        """mode 0: float BVH (timed baseline), 1: double BVH (hit-ID oracle), 2: double brute force."""
        rays = np.ascontiguousarray(rays, dtype=capi.RAY_DTYPE)
        hits = np.zeros(rays.size, dtype=capi.HIT_DTYPE)
        ng = np.zeros((rays.size, 4), np.float32) if want_ng else None
        self.fn("intersect")(h, C.c_size_t(rays.size), _ptr(rays), _ptr(hits), _ptr(ng), C.c_int(mode), C.c_int(nthreads))
        return (hits, ng) if want_ng else hits

    def occluded(self, h, rays, mode=1, nthreads=8):
        rays = np.ascontiguousarray(rays, dtype=capi.RAY_DTYPE)
        occ = np.zeros(rays.size, np.uint8)
        self.fn("occluded")(h, C.c_size_t(rays.size), _ptr(rays), _ptr(occ), C.c_int(mode), C.c_int(nthreads))
        return occ

    def prim_distance(self, h, rays, hits):
        rays = np.ascontiguousarray(rays, dtype=capi.RAY_DTYPE)
        hits = np.ascontiguousarray(hits, dtype=capi.HIT_DTYPE)
        t = np.zeros(rays.size, np.float64)
        self.fn("prim_distance")(h, C.c_size_t(rays.size), _ptr(rays), _ptr(hits), _ptr(t))
        return t

    def surface_params(self, h, rays, hits):
        rays = np.ascontiguousarray(rays, dtype=capi.RAY_DTYPE)
        hits = np.ascontiguousarray(hits, dtype=capi.HIT_DTYPE)
        out = np.zeros(rays.size, dtype=capi.SURFACE_DTYPE)
        self.fn("surface_params")(h, C.c_size_t(rays.size), _ptr(rays), _ptr(hits), _ptr(out))
        return out

    def hashgrid_build(self, positions, radius, cell_count):
        positions = _f32(positions).reshape(-1, 3)
        n = positions.shape[0]
        ends, idx, aabb = np.zeros(cell_count, np.uint32), np.zeros(n, np.uint32), np.zeros(6, np.float32)
        self.fn("hashgrid_build")(C.c_size_t(n), _ptr(positions), C.c_float(radius), C.c_uint32(cell_count), _ptr(ends), _ptr(idx), _ptr(aabb))
        return ends, idx, aabb

    def hashgrid_query(self, positions, radius, cell_count, grid, queries, max_per_query=0):
        positions = _f32(positions).reshape(-1, 3)
        queries = _f32(queries).reshape(-1, 3)
        ends, idx, aabb = grid
        m = queries.shape[0]
        counts = np.zeros(m, np.uint32)
        indices = np.full((m, max_per_query), 0xFFFFFFFF, np.uint32) if max_per_query else None
        self.fn("hashgrid_query")(C.c_size_t(positions.shape[0]), _ptr(positions), C.c_float(radius), C.c_uint32(cell_count), _ptr(ends),
                                  _ptr(idx), _ptr(aabb), C.c_size_t(m), _ptr(queries), _ptr(counts), _ptr(indices), C.c_uint32(max_per_query))
        return counts, indices

    def world_to_image(self, camera, points):
        points = _f32(points).reshape(-1, 3)
        out = np.zeros((points.shape[0], 2), np.int32)
        self.fn("world_to_image")(C.byref(camera), C.c_size_t(points.shape[0]), _ptr(points), _ptr(out))
        return out

    def vcm_light_pass(self, h, camera, settings, nthreads=8, film=None):
        vm_count = settings.width * settings.height
        total = C.c_size_t(0)
        ends = np.zeros(vm_count, np.uint32)
        cap = vm_count * max(1, settings.maxPathLength)
        verts = np.zeros(cap, dtype=capi.VCM_VERTEX_DTYPE)
        rc = self.fn("vcm_light_pass")(h, C.byref(camera), C.byref(settings), _ptr(verts), C.c_size_t(cap), _ptr(ends), C.byref(total),
                                       _ptr(film), C.c_int(nthreads))
        if rc:
            raise RuntimeError(f"orc_vcm_light_pass failed: {rc}")
        return verts[:total.value].copy(), ends

    def render_vcm(self, h, camera, settings, nthreads=8):
        film = np.zeros((settings.height, settings.width, 3), np.float32)
        stats = capi.VcmStats()
        rc = self.fn("render_vcm")(h, C.byref(camera), C.byref(settings), _ptr(film), C.byref(stats), C.c_int(nthreads))
        if rc:
            raise RuntimeError(f"orc_render_vcm failed: {rc}")
        return film, stats

    def rng(self, seed, pixel, sample, n):
        out = np.zeros(n, np.float32)
        self.fn("rng")(C.c_uint32(seed), C.c_uint32(pixel), C.c_uint32(sample), C.c_size_t(n), _ptr(out))
        return out


class Ref(ShadingOracle):
    def __init__(self, path):
        super().__init__(path, "ref_")

    def sizeof(self, name):
        return self.fn("sizeof")(name.encode())

    def render_reference_loop(self, h, camera, settings, nthreads=8):
        """The reference's own PathTracer.cpp EvaluatePath + SurfaceParameters.cpp CalculateSurfaceParams + Framebuffer.cpp writers,
        compiled unmodified behind the Embree-API shim of oracle/ref_pathtracer.cpp (integrator 0, misMode 0, RR after bounce 8)."""
        w, hgt, layers = settings.width, settings.height, settings.layerCount
        film = [np.zeros((hgt, w, 3), np.float32) for _ in range(layers)]
        ptrs = (C.c_void_p * layers)(*[f.ctypes.data for f in film])
        stats = capi.RenderStats()
        rc = self.lib.refpt_render(h, C.byref(camera), C.byref(settings), ptrs, C.byref(stats), C.c_int(nthreads))
        if rc:
            raise RuntimeError(f"refpt_render failed: {rc}")
        return film, stats

    def surface_params_reference(self, h, rays, hits):
        """CalculateSurfaceParams (SurfaceParameters.cpp:144-247, unmodified) on a batch of hits."""
        rays = np.ascontiguousarray(rays, dtype=capi.RAY_DTYPE)
        hits = np.ascontiguousarray(hits, dtype=capi.HIT_DTYPE)
        out = np.zeros(rays.size, dtype=capi.SURFACE_DTYPE)
        rc = self.lib.refpt_surface_params(h, C.c_size_t(rays.size), _ptr(rays), _ptr(hits), _ptr(out))
        assert rc == 0, rc
        return out

    def hashgrid_build(self, positions, radius, cell_count):
        positions = _f32(positions).reshape(-1, 3)
        n = positions.shape[0]
        ends, idx, aabb = np.zeros(cell_count, np.uint32), np.zeros(n, np.uint32), np.zeros(6, np.float32)
        self.fn("hashgrid_build")(C.c_size_t(n), _ptr(positions), C.c_float(radius), C.c_uint32(cell_count), _ptr(ends), _ptr(idx), _ptr(aabb))
        return ends, idx, aabb

    def hashgrid_query(self, positions, radius, cell_count, queries, max_per_query=0):
        positions = _f32(positions).reshape(-1, 3)
        queries = _f32(queries).reshape(-1, 3)
        m = queries.shape[0]
        counts = np.zeros(m, np.uint32)
        indices = np.full((m, max_per_query), 0xFFFFFFFF, np.uint32) if max_per_query else None
        self.fn("hashgrid_query")(C.c_size_t(positions.shape[0]), _ptr(positions), C.c_float(radius), C.c_uint32(cell_count), C.c_size_t(m),
                                  _ptr(queries), _ptr(counts), _ptr(indices), C.c_uint32(max_per_query))
        return counts, indices

    # baked blobs written by the reference's own SerializeToBinary (IoLib/BinaryStreamSerializer.h)
    def write_ibl_blob(self, ibl, path):
        desc = capi.IblDesc()
        keep = [np.ascontiguousarray(ibl[k], np.float32) for k in ("marginal", "conditional", "light", "miss")]
        desc.height, desc.width = keep[2].shape[:2]
        desc.missHeight, desc.missWidth = keep[3].shape[:2]
        desc.marginalDensityFunction, desc.conditionalDensityFunctions = keep[0].ctypes.data, keep[1].ctypes.data
        desc.lightData, desc.missData = keep[2].ctypes.data, keep[3].ctypes.data
        desc.rotationRadians, desc.exposureScale = ibl["rotation"], ibl["exposure_scale"]
        rc = self.fn("write_ibl_blob")(C.byref(desc), str(path).encode())
        assert rc == 0, rc

    def write_texture_blob(self, texture, path):
        a = capi.texture_array(texture)
        td = capi.texture_desc(a)
        rc = self.fn("write_texture_blob")(C.byref(td), str(path).encode())
        assert rc == 0, rc

    def write_hdr(self, path, rgb):
        """FrameBuffer_Save's writer: the reference's StbImageWrite(..., HDR, ...) = stbi_write_hdr, unmodified."""
        rgb = np.ascontiguousarray(rgb, np.float32)
        rc = self.lib.ref_write_hdr(str(path).encode(), C.c_uint32(rgb.shape[1]), C.c_uint32(rgb.shape[0]), _ptr(rgb))
        assert rc == 0, rc

    def asset_version(self, type_name):
        f = self.lib.ref_asset_version
        f.restype = C.c_ulonglong
        return int(f(type_name.encode()))

    def vcm_generate_light(self, h, sphere4, vc_weight, u4):
        """VCMCommon::GenerateLightSample: rows of 13 floats (position, direction, throughput, dVCM, dVC, dVM, length|area<<8)."""
        u4 = _f32(u4).reshape(-1, 4)
        sphere4 = _f32(sphere4, (4,))
        out = np.zeros((u4.shape[0], 13), np.float32)
        rc = self.fn("vcm_generate_light")(h, _ptr(sphere4), C.c_float(vc_weight), C.c_size_t(u4.shape[0]), _ptr(u4), _ptr(out))
        assert rc == 0, rc
        return out

    def vcm_generate_camera(self, camera, light_path_count, xy, u2):
        xy = np.ascontiguousarray(xy, dtype=np.uint32).reshape(-1, 2)
        u2 = _f32(u2).reshape(-1, 2)
        out, index = np.zeros((xy.shape[0], 13), np.float32), np.zeros(xy.shape[0], np.uint32)
        rc = self.fn("vcm_generate_camera")(C.byref(camera), C.c_float(light_path_count), C.c_size_t(xy.shape[0]), _ptr(xy), _ptr(u2),
                                            _ptr(out), _ptr(index))
        assert rc == 0, rc
        return out, index


class PtexRef(_Lib):
    def __init__(self, path):
        super().__init__(path, "ptexref_")

    def write_grid(self, path, nu, nv, ulog2, vlog2, res_jitter=0, dtype=0, nchan=3, seed=4, genmipmaps=True, constant_every=0):
        rc = self.fn("write_grid")(str(path).encode(), nu, nv, ulog2, vlog2, res_jitter, dtype, nchan, C.c_uint32(seed),
                                   1 if genmipmaps else 0, constant_every)
        if rc:
            raise RuntimeError(f"ptexref_write_grid failed: {rc}")

    def eval(self, path, faceid, u, v):
        faceid = np.ascontiguousarray(faceid, dtype=np.int32)
        u, v = _f32(u), _f32(v)
        out = np.zeros((faceid.size, 3), np.float32)
        rc = self.fn("eval")(str(path).encode(), C.c_size_t(faceid.size), _ptr(faceid), _ptr(u), _ptr(v), _ptr(out))
        if rc:
            raise RuntimeError(f"ptexref_eval failed: {rc}")
        return out

    def open(self, path):
        f = self.fn("open")
        f.restype = C.c_void_p
        h = f(str(path).encode())
        if not h:
            raise RuntimeError(f"ptexref_open failed: {path}")
        return h

    def close(self, handle):
        self.fn("close")(C.c_void_p(handle))

    def info(self, path):
        out = (C.c_int * 8)()
        rc = self.fn("info")(str(path).encode(), out)
        if rc:
            raise RuntimeError("ptexref_info failed")
        return list(out)


_cache = {}


def load_oracle() -> Oracle:
    if "orc" not in _cache:
        path = ORACLE_DIR / "liboracle.so"
        if not path.exists():
            build_oracle()
        _cache["orc"] = Oracle(path)
    return _cache["orc"]


def load_ref():
    if "ref" not in _cache:
        path = ORACLE_DIR / "_ref" / "libselas_ref.so"
        if not path.exists() and os.path.exists("/root/reference/Source"):
            build_oracle()
        _cache["ref"] = Ref(path) if path.exists() else None
    return _cache["ref"]


def load_ptexref():
    if "ptex" not in _cache:
        path = ORACLE_DIR / "_ref" / "libptex_ref.so"
        if not path.exists() and os.path.exists("/root/reference/Middleware/ptex"):
            build_oracle()
        _cache["ptex"] = PtexRef(path) if path.exists() else None
    return _cache["ptex"]
