"""Test helper: a small two-level scene (scene -> subscene instances -> model instances -> meshes) written as the reference's baked
blobs by the reference's own serializer (oracle/_ref, ref_write_*_blob), and the flat capi.SceneData the host loader must produce from
them.  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import zlib
from pathlib import Path

import numpy as np

from selas_b200 import capi, scenes
from tests import inputs

MESH_META_DTYPE = np.dtype([("indexCount", "u4"), ("indexOffset", "u4"), ("vertexCount", "u4"), ("vertexOffset", "u4"), ("materialHash", "u4"),
                            ("indicesPerFace", "u4"), ("nameHash", "u4"), ("name", "S64")])
MATERIAL_RES_DTYPE = np.dtype([("baseColorTexture", "S512"), ("payload", capi.MATERIAL_DTYPE)])
INSTANCE_DTYPE = np.dtype([("index", "u8"), ("localToWorld", "f4", 16)])
CAMERA_DTYPE = np.dtype([("name", "S128"), ("position", "f4", 3), ("fovDegrees", "f4"), ("lookAt", "f4", 3), ("znear", "f4"), ("up", "f4", 3),
                         ("zfar", "f4")])
assert MESH_META_DTYPE.itemsize == 92 and MATERIAL_RES_DTYPE.itemsize == 600 and INSTANCE_DTYPE.itemsize == 72 and CAMERA_DTYPE.itemsize == 176

VERSIONS = {"Textures": 1534887410, "SceneResource": 1539147471, "SubsceneResource": 1539551841, "ModelResource": 1539636447, "ModelGeometryResource": 1539636447,
            "IBL": 1538087123}


def asset_path(root, type_name, name):
    p = Path(root) / "_Assets" / type_name / str(VERSIONS[type_name]) / f"{name}.bin"
    p.parent.mkdir(parents=True, exist_ok=True)
    return p


def xform(tx, ty, tz, s, ang):
    c, sn = np.cos(ang), np.sin(ang)
    return np.array([[s * c, 0, -s * sn, 0], [0, s, 0, 0], [s * sn, 0, s * c, 0], [tx, ty, tz, 1]], np.float32)


def mul44(a, b):
    """MatrixMultiply (row-major, row vectors) in float32 with the host loader's summation order."""
    a, b = a.astype(np.float32), b.astype(np.float32)
    c = np.zeros((4, 4), np.float32)
    for i in range(4):
        for j in range(4):
            acc = np.float32(a[i, 0] * b[0, j])
            for k in (1, 2, 3):
                acc = np.float32(acc + np.float32(a[i, k] * b[k, j]))
            c[i, j] = acc
    return c


def _names(strings, width=512):
    a = np.zeros(len(strings), dtype=f"S{width}")
    a[:] = [s.encode() for s in strings]
    return a


def _material_res(materials, names=None):
    a = np.zeros(len(materials), MATERIAL_RES_DTYPE)
    for i, m in enumerate(materials):
        a[i]["payload"] = m
        if names and names[i]:
            a[i]["baseColorTexture"] = names[i].encode()
    return a


def _p(a):
    return None if a is None or a.size == 0 else a.ctypes.data_as(C.c_void_p)


def small_two_level_scene():
    """Two models (A: attributed quad sphere, two meshes sharing one vertex array; B: bare triangle sphere + floor), one subscene holding
    three model instances, two subscene instances, three quad lights, one camera, an IBL."""
    base = inputs.small_scene(with_attrs=True)
    sph, tri, floor = base.meshes
    rng = np.random.default_rng(9)
    uvs = rng.uniform(0, 1, size=(sph["positions"].shape[0], 2)).astype(np.float32)
    half = sph["indices"].shape[0] // 2
    checker = ((np.add.outer(np.arange(8), np.arange(8)) % 2) * 0.7 + 0.2).astype(np.float32)
    texture = np.stack([checker, checker[::-1], np.full((8, 8), 0.5, np.float32)], axis=2)
    model_a = dict(name="modelA", textures={0xA1: "checker"}, meshes=[dict(name="capTop", material=0xA1, indices=sph["indices"][:half]),
                                          dict(name="capBottom", material=0xB7, indices=sph["indices"][half:])],
                   positions=sph["positions"], normals=sph["normals"], tangents=sph["tangents"], uvs=uvs,
                   materials={0xA1: base.materials[0]})  # 0xB7 is only in the subscene's material list
    model_b = dict(name="modelB", meshes=[dict(name="blob", material=0xC3, indices=tri["indices"], positions=tri["positions"]),
                                          dict(name="floor", material=0x99, indices=floor["indices"], positions=floor["positions"])],
                   normals=None, tangents=None, uvs=None, materials={0xC3: base.materials[1]})  # 0x99 resolves to the default material
    sub = dict(name="sub0", light_set=0, models=[model_a, model_b], scene_materials={0xB7: base.materials[2]},
               instances=[(0, np.eye(4, dtype=np.float32)), (1, np.eye(4, dtype=np.float32)), (0, xform(-2.5, 0.4, 0.5, 0.6, 0.7))])
    ibl_img = (rng.uniform(0, 1, size=(16, 32, 3)) ** 4 * 6.0).astype(np.float32)
    ibl = scenes.build_ibl(ibl_img, miss=rng.uniform(0, 1, size=(4, 8, 3)).astype(np.float32), rotation=0.4, exposure=0.0)
    scene = dict(name="twolevel", ibl_name="sky", ibl=ibl, textures={"checker": texture}, background=(0.2, 0.3, 0.4, 0.0), subscenes=[sub],
                 instances=[(0, np.eye(4, dtype=np.float32)), (0, xform(1.0, 2.6, 1.5, 0.5, -0.4))], lights=base.lights, light_sets=[(0, 3)],
                 cameras=[dict(name="main", position=(0.0, 1.0, -6.0), lookAt=(0.3, 0.2, 0.0), up=(0.0, 1.0, 0.0), fovDegrees=45.0, znear=0.1, zfar=500.0)])
    return scene


def grass_model_scene():
    """A model with a floor mesh and two curve geometries (RTC_GEOMETRY_TYPE_ROUND_BSPLINE_CURVE, ModelResource.cpp:300-333) sharing one
    control-vertex array, instanced twice (the second instance rotated, moved and uniformly scaled): what the reference's island scene
    is full of.  The curves' material is the one named curveModelName (ModelResource.cpp:382-383)."""
    g = scenes.grass_patch(blades=60, extent=2.0)
    floor = g.meshes[0]
    cv, ci = g.curves[0]["vertices"], g.curves[0]["indices"]
    half = ci.size // 2
    model = dict(name="grassModel", meshes=[dict(name="soil", material=0x11, indices=floor["indices"], positions=floor["positions"])],
                 normals=None, tangents=None, uvs=None, materials={0x11: g.materials[0], 0x22: g.materials[1]},
                 curve_vertices=cv, curves=[ci[:half], ci[half:]], curve_material=0x22)
    sub = dict(name="grassSub", light_set=0, models=[model], scene_materials={},
               instances=[(0, np.eye(4, dtype=np.float32)), (0, xform(5.0, 0.3, 1.0, 1.4, 0.6))])
    rng = np.random.default_rng(4)
    ibl = scenes.build_ibl((rng.uniform(0, 1, size=(8, 16, 3)) * 0.5).astype(np.float32), miss=np.full((2, 4, 3), 0.3, np.float32), rotation=0.0, exposure=0.0)
    return dict(name="grass", ibl_name="dimsky", ibl=ibl, textures={}, background=(0.3, 0.45, 0.7, 0.0), subscenes=[sub],
                instances=[(0, np.eye(4, dtype=np.float32))], lights=g.lights, light_sets=[(0, 1)],
                cameras=[dict(name="main", position=(2.0, 2.5, -9.0), lookAt=(2.0, 0.5, 0.0), up=(0.0, 1.0, 0.0), fovDegrees=50.0, znear=0.1, zfar=500.0)])


def ptex_sheet_scene():
    """One model with one quad mesh ("sheet", 5 x 4 quads = the faces of tests/golden/ptex_jitter_u8.ptx) whose material uses Ptex: the
    loader must resolve Content/<material.baseColorTexture>/<mesh name>.ptx (ModelResource.cpp:358-364)."""
    flat = scenes.ptex_quad_grid(5, 4, ptex=None)
    sheet = flat.meshes[0]
    mat = scenes.material(shaderName="DisneySolid", baseColor=[0.6, 0.6, 0.6], metallic=0.0, roughness=0.8, usesPtex=True)
    model = dict(name="sheetModel", textures={0x51: "sheetTex"}, meshes=[dict(name="sheet", material=0x51, indices=sheet["indices"], positions=sheet["positions"])],
                 normals=None, tangents=None, uvs=None, materials={0x51: mat})
    sub = dict(name="sheetSub", light_set=0, models=[model], scene_materials={}, instances=[(0, np.eye(4, dtype=np.float32))])
    rng = np.random.default_rng(3)
    ibl = scenes.build_ibl((rng.uniform(0, 1, size=(8, 16, 3)) * 0.5).astype(np.float32), miss=np.full((2, 4, 3), 0.3, np.float32), rotation=0.0, exposure=0.0)
    return dict(name="ptexsheet", ibl_name="dim", ibl=ibl, textures={}, background=(0.3, 0.35, 0.4, 0.0), subscenes=[sub],
                instances=[(0, np.eye(4, dtype=np.float32))], lights=flat.lights, light_sets=[(0, 1)],
                cameras=[dict(name="below", position=(0.0, -9.0, -14.0), lookAt=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), fovDegrees=55.0, znear=0.1, zfar=500.0)])


def write_baked(ref, root, scene):
    """All blobs of `scene` under root/_Assets/... through the reference serializer."""
    lib = ref.lib
    ref.write_ibl_blob(scene["ibl"], asset_path(root, "IBL", scene["ibl_name"]))
    for name, tex in scene.get("textures", {}).items():
        ref.write_texture_blob(tex, asset_path(root, "Textures", name))
    for sub in scene["subscenes"]:
        for model in sub["models"]:
            meshes = model["meshes"]
            shared = "positions" in model
            pos, idx, meta = [], [], np.zeros(len(meshes), MESH_META_DTYPE)
            voff = ioff = 0
            for k, m in enumerate(meshes):
                ind = np.ascontiguousarray(m["indices"], np.uint32)
                if shared:  # indices are already absolute into the model-wide arrays
                    lo, nv, add = int(ind.min()), int(ind.max()) - int(ind.min()) + 1, 0
                else:
                    p = np.ascontiguousarray(m["positions"], np.float32)
                    pos.append(p)
                    lo, nv, add = voff, p.shape[0], voff
                    voff += p.shape[0]
                meta[k] = (ind.size, ioff, nv, lo, m["material"], ind.shape[1], zlib.crc32(m["name"].encode()), m["name"].encode())
                idx.append((ind + np.uint32(add)).ravel())
                ioff += ind.size
            positions = np.ascontiguousarray(model["positions"] if shared else np.concatenate(pos), np.float32)
            indices = np.ascontiguousarray(np.concatenate(idx), np.uint32)
            nrm = None if model["normals"] is None else np.ascontiguousarray(model["normals"], np.float32)
            tan = None if model["tangents"] is None else np.ascontiguousarray(model["tangents"], np.float32)
            uvs = None if model["uvs"] is None else np.ascontiguousarray(model["uvs"], np.float32)
            hashes = np.array(sorted(model["materials"]), np.uint32)
            mats = _material_res([model["materials"][h] for h in hashes], [model.get("textures", {}).get(int(h)) for h in hashes])
            aabb = np.concatenate([positions.min(axis=0), positions.max(axis=0)]).astype(np.float32)
            if model.get("curves"):  # curve geometries: model-wide control vertices, CurveMetaData {indexOffset, indexCount} per curve
                cv = np.ascontiguousarray(model["curve_vertices"], np.float32).reshape(-1, 4)
                cidx, ranges, off = [], [], 0
                for c in model["curves"]:
                    ci = np.ascontiguousarray(c, np.uint32).ravel()
                    ranges.append((off, ci.size))
                    cidx.append(ci)
                    off += ci.size
                cidx, ranges = np.ascontiguousarray(np.concatenate(cidx), np.uint32), np.array(ranges, np.uint32)
                assert lib.ref_set_next_model_curves(_p(cidx), C.c_uint64(cidx.size), _p(cv), C.c_uint32(cv.shape[0]), _p(ranges), C.c_uint32(len(ranges)),
                                                     C.c_uint32(model["curve_material"])) == 0
            rc = lib.ref_write_model_blobs(str(asset_path(root, "ModelResource", model["name"])).encode(),
                                           str(asset_path(root, "ModelGeometryResource", model["name"])).encode(), _p(aabb),
                                           C.c_uint32(positions.shape[0]), _p(indices), C.c_uint64(indices.size), _p(positions), _p(nrm), _p(tan), _p(uvs),
                                           _p(meta), C.c_uint32(len(meshes)), _p(mats), _p(hashes), C.c_uint32(hashes.size))
            assert rc == 0
        inst = np.zeros(len(sub["instances"]), INSTANCE_DTYPE)
        for k, (i, m) in enumerate(sub["instances"]):
            inst[k] = (i, m.astype(np.float32).ravel())
        sm_hashes = np.array(list(sub["scene_materials"]), np.uint32)
        sm = _material_res([sub["scene_materials"][h] for h in sm_hashes])
        names = _names([m["name"] for m in sub["models"]])
        rc = lib.ref_write_subscene_blob(str(asset_path(root, "SubsceneResource", sub["name"])).encode(), sub["name"].encode(), C.c_uint64(sub["light_set"]),
                                         _p(names), C.c_uint32(names.size), _p(inst), C.c_uint32(inst.size), _p(sm_hashes), _p(sm), C.c_uint32(sm.size))
        assert rc == 0
    inst = np.zeros(len(scene["instances"]), INSTANCE_DTYPE)
    for k, (i, m) in enumerate(scene["instances"]):
        inst[k] = (i, m.astype(np.float32).ravel())
    lights = np.ascontiguousarray(scene["lights"], capi.LIGHT_DTYPE)
    sets = np.array(scene["light_sets"], np.uint32).reshape(-1, 2)
    cams = np.zeros(len(scene["cameras"]), CAMERA_DTYPE)
    for k, c in enumerate(scene["cameras"]):
        cams[k] = (c["name"].encode(), c["position"], c["fovDegrees"], c["lookAt"], c["znear"], c["up"], c["zfar"])
    bg = np.array(scene["background"], np.float32)
    names = _names([s["name"] for s in scene["subscenes"]])
    rc = lib.ref_write_scene_blob(str(asset_path(root, "SceneResource", scene["name"])).encode(), scene["name"].encode(), scene["ibl_name"].encode(), _p(bg),
                                  _p(names), C.c_uint32(names.size), _p(inst), C.c_uint32(inst.size), _p(lights), C.c_uint32(lights.size), _p(sets),
                                  C.c_uint32(sets.shape[0]), _p(cams), C.c_uint32(cams.size))
    assert rc == 0


def expected_flat_scene(scene, with_ids=False):
    """What LoadBakedScene must build: meshes in subscene -> model -> mesh order with FindMeshMaterial-resolved materials, and the two
    instancing levels flattened (localToWorld = model * subscene)."""
    meshes, materials, ranges, curves = [], [], [], []
    default_index = None
    for sub in scene["subscenes"]:
        scene_base = len(materials)
        sm_hashes = list(sub["scene_materials"])
        materials += [sub["scene_materials"][h] for h in sm_hashes]
        sub_ranges = []
        for model in sub["models"]:
            model_base = len(materials)
            hashes = sorted(model["materials"])
            materials += [model["materials"][h] for h in hashes]
            first = len(meshes)
            for m in model["meshes"]:
                ind = np.ascontiguousarray(m["indices"], np.uint32)
                if "positions" in model:
                    lo, hi = int(ind.min()), int(ind.max()) + 1
                    d = dict(positions=model["positions"][lo:hi], indices=ind - np.uint32(lo))
                    for k in ("normals", "tangents", "uvs"):
                        d[k] = None if model[k] is None else model[k][lo:hi]
                else:
                    d = dict(positions=m["positions"], indices=ind)
                if m["material"] in model["materials"]:
                    d["material"] = model_base + hashes.index(m["material"])
                    tname = model.get("textures", {}).get(m["material"])
                    if tname and tname in scene["textures"]:  # a Ptex material names a directory of .ptx files instead
                        d["texture"] = scene["textures"][tname]
                elif m["material"] in sub["scene_materials"]:
                    d["material"] = scene_base + sm_hashes.index(m["material"])
                else:
                    if default_index is None:
                        default_index = len(materials)
                        dm = np.zeros((), capi.MATERIAL_DTYPE)
                        dm["shader"], dm["baseColor"] = capi.SHADER_DISNEY_SOLID, (0.6, 0.6, 0.6)
                        dm["scalars"][capi.IOR] = 1.5
                        materials.append(dm)
                    d["material"] = default_index
                d["lightSet"] = sub["light_set"]
                meshes.append(d)
            curve_first = len(curves)
            if model.get("curves"):  # FindMeshMaterial(curveModelName): the model's materials, the subscene's, else the default
                h = model["curve_material"]
                if h in model["materials"]:
                    cmat = model_base + hashes.index(h)
                elif h in sub["scene_materials"]:
                    cmat = scene_base + sm_hashes.index(h)
                else:
                    if default_index is None:
                        default_index = len(materials)
                        dm = np.zeros((), capi.MATERIAL_DTYPE)
                        dm["shader"], dm["baseColor"] = capi.SHADER_DISNEY_SOLID, (0.6, 0.6, 0.6)
                        dm["scalars"][capi.IOR] = 1.5
                        materials.append(dm)
                    cmat = default_index
                cv = np.ascontiguousarray(model["curve_vertices"], np.float32).reshape(-1, 4)
                for c in model["curves"]:
                    ci = np.ascontiguousarray(c, np.uint32).ravel()
                    lo, hi = int(ci.min()), int(ci.max()) + 4
                    curves.append(dict(vertices=cv[lo:hi], indices=ci - np.uint32(lo), material=cmat, lightSet=sub["light_set"]))
            sub_ranges.append((first, len(model["meshes"]), curve_first, len(model.get("curves", []))))
        ranges.append(sub_ranges)
    instances, ids = [], []
    for s_idx, (si, ms) in enumerate(scene["instances"]):
        for m_idx, (mi, mm) in enumerate(scene["subscenes"][si]["instances"]):
            first, count, cfirst, ccount = ranges[si][mi]
            m = mul44(mm, ms)
            instances += [(first + k, m) for k in range(count)]
            # what a hit reports in the reference: instID[0] = subscene instance, instID[1] = model instance, geomID = mesh inside the model
            ids += [(s_idx, m_idx, k) for k in range(count)]
            # the model's curve geometries are attached after its meshes (ModelResource.cpp:461-466); globally they follow all meshes
            instances += [(len(meshes) + cfirst + k, m) for k in range(ccount)]
            ids += [(s_idx, m_idx, count + k) for k in range(ccount)]
    return capi.SceneData(meshes, np.array(materials), scene["lights"], scene["light_sets"], background=scene["background"], instances=instances,
                          ibl=scene["ibl"], instance_ids=ids if with_ids else None, curves=curves)
