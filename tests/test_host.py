"""The C++ host surface (selas_b200/host): material JSON import + BuildMaterial against the Python restatement of the same
reference rules for the six material files the reference ships, the Radiance .hdr writer, and (GPU) the selas_render
executable = the reference's main() for the Cornell scene, whose image must equal the one rendered through the C ABI."""
import ctypes as C
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

from selas_b200 import capi, scenes

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "selas_b200" / "lib"


@pytest.fixture(scope="module")
def host():
    path = LIB / "libselas_host.so"
    if not path.exists():
        subprocess.run(["make", "-C", str(ROOT / "selas_b200" / "host")], check=True, stdout=subprocess.DEVNULL)
    capi.load_library()
    return C.CDLL(str(path))


def read_hdr(path):
    data = Path(path).read_bytes()
    assert data.startswith(b"#?RADIANCE") and b"FORMAT=32-bit_rle_rgbe" in data[:128]
    a = _decode_hdr(data).astype(np.float32)  # flat or run-length scanlines
    scale = np.where(a[..., 3] > 0, np.exp2(a[..., 3] - 136.0), 0.0).astype(np.float32)
    return a[..., :3] * scale[..., None]


def test_material_import_matches_reference_rules(host):
    for name, d in scenes.REFERENCE_MATERIALS.items():
        want = scenes.material(**d)
        got = capi.Material()
        err = C.create_string_buffer(256)
        assert host.selas_host_import_material_json(json.dumps(d).encode(), C.byref(got), err, 256) == 0, err.value
        assert bytes(got) == want.tobytes(), name
    got = capi.Material()
    assert host.selas_host_import_material_json(b" ", C.byref(got), None, 0) == 0  # Default.json is an empty object
    assert bytes(got) == scenes.material().tobytes()
    err = C.create_string_buffer(256)
    assert host.selas_host_import_material_json(b'{"roughness": ', C.byref(got), err, 256) == 4 and b"json" in err.value
    assert host.selas_host_import_material_json(b'{"shaderName":"disneythin","usesPtex":true,"alphaTesting":true}', C.byref(got), None, 0) == 0
    assert got.shader == capi.SHADER_DISNEY_THIN and got.flags == (capi.MAT_USES_PTEX | capi.MAT_ALPHA_TESTED)


def test_hdr_writer_roundtrip(host, tmp_path):
    rng = np.random.default_rng(3)
    img = (rng.uniform(0, 1, size=(7, 13, 3)) ** 4 * 50).astype(np.float32)
    img[0, 0] = 0
    path = tmp_path / "t.hdr"
    assert host.selas_host_write_hdr(str(path).encode(), 13, 7, img.ctypes.data_as(C.c_void_p)) == 0
    back = read_hdr(path)
    m = img.max(axis=2, keepdims=True)
    assert np.all(np.abs(back - img) <= m / 128.0 + 1e-6)  # 8-bit mantissa shared exponent


@pytest.mark.gpu
def test_selas_render_cornell(ctx, host, tmp_path):
    exe = LIB / "selas_render"
    out = subprocess.run([str(exe), "--width", "96", "--height", "96", "--spp", "8", "--depth", "6", "--roughness", "0.1", "--seed", "1",
                          "--name", "t"], cwd=tmp_path, capture_output=True, text=True, check=True).stdout
    assert "Scene render time" in out
    stats = json.loads(out.strip().splitlines()[-1])
    assert stats["paths"] == 96 * 96 * 8 and stats["kernel_launches"] > 0
    radiance = read_hdr(tmp_path / "_Images" / "t_1.hdr")
    albedo = read_hdr(tmp_path / "_Images" / "t_0.hdr")
    sc = capi.Scene(ctx, scenes.cornell_box(copper_roughness=0.1))
    film, _ = sc.render(scenes.cornell_camera(96, 96), capi.render_settings(96, 96, 8, max_path_length=6, rr_start=8, layers=2, seed=1))
    for got, want in ((albedo, film[0] / 8.0), (radiance, film[1] / 8.0)):
        m = want.max(axis=2, keepdims=True)
        close = np.abs(got - want) <= m / 100.0 + 1e-4
        assert close.mean() > 0.995  # RGBE quantisation + atomic float accumulation order
    sc.close()


@pytest.mark.gpu
def test_selas_render_vcm(ctx, host, tmp_path):
    """VCM::GenerateImage through the C++ host on its own C++ copy of the config 5 scene (MakeCausticScene, the IBL density
    functions of ImageBasedLight::BuildDensityFunctions) against the Python-built scene rendered through the C ABI."""
    exe = LIB / "selas_render"
    w, h, its = 128, 72, 3
    out = subprocess.run([str(exe), "--integrator", "vcm", "--width", str(w), "--height", str(h), "--iterations", str(its), "--seed", "4",
                          "--name", "v"], cwd=tmp_path, capture_output=True, text=True, check=True).stdout
    assert f"Vcm integration performed with {its} iterations" in out
    stats = json.loads(out.strip().splitlines()[-1])
    assert stats["iterations"] == its and stats["light_vertices"] > 0 and stats["merges"] > 0 and stats["connections"] > 0
    got = read_hdr(tmp_path / "_Images" / "v_0.hdr")
    data = scenes.caustic_scene()
    sc = capi.Scene(ctx, data)
    film, st = sc.render_vcm(scenes.caustic_camera(w, h), capi.vcm_settings(w, h, its, scenes.bounding_sphere(data), seed=4))
    want = film / its
    assert abs(int(st.lightVertices) - stats["light_vertices"]) <= 0.01 * st.lightVertices
    # same streams, but the two scene builders round vertex positions and CDFs independently: compare as pictures
    assert abs(got.mean() / want.mean() - 1.0) < 0.02
    gb = got.reshape(h // 4, 4, w // 4, 4, 3).mean(axis=(1, 3))
    wb = want.reshape(h // 4, 4, w // 4, 4, 3).mean(axis=(1, 3))
    assert float(np.mean((gb - wb) ** 2 / (wb ** 2 + 1e-2))) < 0.02
    sc.close()


@pytest.mark.gpu
def test_selas_render_deferred(ctx, host, tmp_path):
    """DeferredPathTracer::GenerateImage through the C++ host equals the same settings through the C ABI."""
    exe = LIB / "selas_render"
    out = subprocess.run([str(exe), "--integrator", "deferred", "--width", "80", "--height", "80", "--roughness", "0.1", "--seed", "2", "--depth", "32",
                          "--name", "d"], cwd=tmp_path, capture_output=True, text=True, check=True).stdout
    stats = json.loads(out.strip().splitlines()[-1])
    assert stats["paths"] == 80 * 80 * 16 and stats["occlusion_rays"] > stats["paths"]  # two NEE samples per vertex
    got = read_hdr(tmp_path / "_Images" / "d_0.hdr")
    sc = capi.Scene(ctx, scenes.cornell_box(copper_roughness=0.1))
    film, _ = sc.render(scenes.cornell_camera(80, 80), capi.render_settings(80, 80, 16, max_path_length=32, layers=1, seed=2, cmj=(4, 4),
                                                                             integrator=capi.INTEGRATOR_DEFERRED))
    want = film[0] / 16.0
    m = want.max(axis=2, keepdims=True)
    assert (np.abs(got - want) <= m / 100.0 + 1e-4).mean() > 0.995
    sc.close()


# ---- baked assets: blobs written by the reference's own serializer, read by the host library -------------------
GOLDEN = ROOT / "tests" / "golden"


def _read_ibl(host, path):
    info, params = (C.c_uint32 * 4)(), (C.c_float * 2)()
    err = C.create_string_buffer(256)
    assert host.selas_host_read_ibl_blob(str(path).encode(), info, params, None, None, None, None, err, 256) == 0, err.value
    w, h, mw, mh = list(info)
    marg, cond = np.zeros(h, np.float32), np.zeros((h, w), np.float32)
    light, miss = np.zeros((h, w, 3), np.float32), np.zeros((mh, mw, 3), np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert host.selas_host_read_ibl_blob(str(path).encode(), info, params, p(marg), p(cond), p(light), p(miss), err, 256) == 0, err.value
    return dict(marginal=marg, conditional=cond, light=light, miss=miss, rotation=params[0], exposure_scale=params[1])


def _read_texture(host, path):
    info = (C.c_uint32 * 4)()
    err = C.create_string_buffer(256)
    assert host.selas_host_read_texture_blob(str(path).encode(), info, None, err, 256) == 0, err.value
    mips, fmt, w, h = list(info)
    tex = np.zeros((h, w, fmt + 1), np.float32)
    assert host.selas_host_read_texture_blob(str(path).encode(), info, tex.ctypes.data_as(C.c_void_p), err, 256) == 0
    return mips, tex


def _golden_assets():
    rng = np.random.default_rng(71)
    img = (rng.uniform(0, 1, size=(12, 24, 3)) ** 4 * 30.0).astype(np.float32)
    ibl = scenes.build_ibl(img, miss=rng.uniform(0, 2, size=(3, 6, 3)).astype(np.float32), rotation=0.7, exposure=1.0)
    tex = rng.uniform(0, 1, size=(5, 7, 3)).astype(np.float32)
    return ibl, tex


def test_baked_blobs_from_the_reference_serializer(host, ref, tmp_path):
    """Written here by the reference's SerializeToBinary, read back by ReadImageBasedLightResource / ReadTextureResource."""
    ibl, tex = _golden_assets()
    ref.write_ibl_blob(ibl, tmp_path / "ibl.bin")
    ref.write_texture_blob(tex, tmp_path / "tex.bin")
    got = _read_ibl(host, tmp_path / "ibl.bin")
    for k in ("marginal", "conditional", "light", "miss"):
        np.testing.assert_array_equal(got[k], ibl[k], err_msg=k)
    assert got["rotation"] == np.float32(ibl["rotation"]) and got["exposure_scale"] == np.float32(ibl["exposure_scale"])
    mips, t = _read_texture(host, tmp_path / "tex.bin")
    assert mips == 1
    np.testing.assert_array_equal(t, tex)
    # the committed golden blobs are these very files
    assert (tmp_path / "ibl.bin").read_bytes() == (GOLDEN / "baked_ibl.bin").read_bytes()
    assert (tmp_path / "tex.bin").read_bytes() == (GOLDEN / "baked_texture.bin").read_bytes()
    assert ref.asset_version("IBL") == 1538087123 and ref.asset_version("Textures") == 1534887410


def test_baked_blobs_golden(host, tmp_path):
    """The committed blobs (made by tests/golden/make_golden.py from the reference serializer) without the reference at hand."""
    ibl, tex = _golden_assets()
    got = _read_ibl(host, GOLDEN / "baked_ibl.bin")
    for k in ("marginal", "conditional", "light", "miss"):
        np.testing.assert_array_equal(got[k], ibl[k], err_msg=k)
    mips, t = _read_texture(host, GOLDEN / "baked_texture.bin")
    np.testing.assert_array_equal(t, tex)
    # malformed files are reported, not read
    bad = tmp_path / "bad.bin"
    bad.write_bytes((GOLDEN / "baked_ibl.bin").read_bytes()[:100])
    err = C.create_string_buffer(256)
    assert host.selas_host_read_ibl_blob(str(bad).encode(), None, None, None, None, None, None, err, 256) == 4 and b"outside the file" in err.value
    assert host.selas_host_read_texture_blob(str(tmp_path / "missing.bin").encode(), None, None, err, 256) == 4 and b"Failed to open" in err.value


def _load_baked(host, root, name):
    counts = (C.c_uint64 * 8)()
    err = C.create_string_buffer(512)
    rc = host.selas_host_load_baked_scene(str(root).encode() + b"/", name.encode(), counts, None, None, None, err, 512)
    assert rc == 0, err.value
    n_mesh, n_inst, n_mat = counts[0], counts[1], counts[2]
    mats, mesh = np.zeros((n_inst, 4, 4), np.float32), np.zeros(n_inst, np.uint32)
    materials = np.zeros(n_mat, capi.MATERIAL_DTYPE)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert host.selas_host_load_baked_scene(str(root).encode() + b"/", name.encode(), counts, p(mats), p(mesh), p(materials), err, 512) == 0
    return list(counts), mats, mesh, materials


def test_baked_scene_from_the_reference_serializer(host, ref, tmp_path):
    """Scene / subscene / model / geometry / IBL blobs written by the reference's Serialize() functions (oracle/_ref) and read by
    LoadBakedScene: counts, FindMeshMaterial resolution and the flattened two-level instance transforms."""
    from tests import baked

    scene = baked.small_two_level_scene()
    baked.write_baked(ref, tmp_path, scene)
    want = baked.expected_flat_scene(scene)
    counts, mats, mesh, materials = _load_baked(host, tmp_path, scene["name"])
    assert counts[0] == len(want.meshes) == 4 and counts[1] == len(want.instances) == 12 and counts[3] == 3 and counts[4] == 1 and counts[5] == 1
    assert counts[6] == sum(want.meshes[i]["indices"].shape[0] * (2 if want.meshes[i]["indices"].shape[1] == 4 else 1) for i, _ in want.instances)
    assert (counts[7] & ((1 << 62) - 1)) == (1 | (1 << 8) | (1 << 32)) and counts[7] >> 62 == 1  # IBL, one distinct texture ("checker"), one textured mesh
    np.testing.assert_array_equal(mesh, [i for i, _ in want.instances])
    np.testing.assert_array_equal(mats, np.stack([m for _, m in want.instances]))  # bit-exact: same float32 summation order
    assert materials.tobytes() == want.materials.tobytes()
    for t in ("SceneResource", "SubsceneResource", "ModelResource", "ModelGeometryResource"):
        assert ref.asset_version(t) == baked.VERSIONS[t]
    # a missing subscene blob is an error, not an empty scene
    (baked.asset_path(tmp_path, "SubsceneResource", "sub0")).unlink()
    err = C.create_string_buffer(512)
    assert host.selas_host_load_baked_scene(str(tmp_path).encode() + b"/", b"twolevel", None, None, None, None, err, 512) == 4
    assert b"Failed to open" in err.value


@pytest.mark.gpu
def test_baked_scene_renders_like_the_flat_description(ctx, host, ref, tmp_path):
    """The baked two-level scene rendered through the host loader equals the flat SceneData rendered through the C ABI."""
    from tests import baked

    scene = baked.small_two_level_scene()
    baked.write_baked(ref, tmp_path, scene)
    want = baked.expected_flat_scene(scene)
    w, h, spp = 96, 64, 8
    c = scene["cameras"][0]
    cam = capi.camera_init(c["position"], c["lookAt"], c["up"], c["fovDegrees"], w, h, c["znear"], c["zfar"])
    rs = capi.render_settings(w, h, spp, max_path_length=6, rr_start=8, layers=1, seed=3)
    film = np.zeros((h, w, 3), np.float32)
    err = C.create_string_buffer(512)
    rc = host.selas_host_render_baked_scene(str(tmp_path).encode() + b"/", b"twolevel", 0, C.byref(cam), C.byref(rs), film.ctypes.data_as(C.c_void_p), err, 512)
    assert rc == 0, err.value
    sc = capi.Scene(ctx, want)
    ref_film = sc.render(cam, rs)[0][0]
    sc.close()
    assert ref_film.sum() > 0
    fin = np.isfinite(ref_film).all(axis=2) & np.isfinite(film).all(axis=2)
    np.testing.assert_allclose(film[fin], ref_film[fin], rtol=2e-3, atol=2e-3)  # identical inputs: only the float atomics' order differs


def _ptex_sheet(ref, tmp_path, with_file=True):
    import shutil

    from tests import baked

    scene = baked.ptex_sheet_scene()
    baked.write_baked(ref, tmp_path, scene)
    if with_file:
        d = tmp_path / "Content" / "sheetTex"
        d.mkdir(parents=True)
        shutil.copy(GOLDEN / "ptex_jitter_u8.ptx", d / "sheet.ptx")
    return scene


def test_baked_scene_resolves_ptex_files(host, ref, tmp_path):
    """A mesh whose material uses Ptex names Content/<baseColorTexture>/<mesh name>.ptx (ModelResource.cpp:358-364); a missing file is
    the reference's "Ptex file ... does not exist" (TextureCache.cpp:105-108)."""
    from tests import baked

    scene = _ptex_sheet(ref, tmp_path, with_file=False)
    want = baked.expected_flat_scene(scene)
    assert len(want.meshes) == 1 and want.meshes[0].get("texture") is None and want.to_desc() is not None
    err = C.create_string_buffer(512)
    assert host.selas_host_load_baked_scene(str(tmp_path).encode() + b"/", b"ptexsheet", None, None, None, None, err, 512) == 4
    assert b"Ptex file" in err.value and b"Content/sheetTex/sheet.ptx does not exist" in err.value
    d = tmp_path / "Content" / "sheetTex"
    d.mkdir(parents=True)
    (d / "sheet.ptx").write_bytes((GOLDEN / "ptex_jitter_u8.ptx").read_bytes())
    counts, mats, mesh, materials = _load_baked(host, tmp_path, "ptexsheet")
    assert counts[0] == 1 and counts[1] == 1 and materials[0]["flags"] & capi.MAT_USES_PTEX


@pytest.mark.gpu
def test_baked_scene_with_ptex_renders_like_the_flat_description(ctx, host, ref, tmp_path):
    """The loader attaches the Ptex file itself: the baked sheet through the host equals the flat scene with the texture passed by hand."""
    from tests import baked

    scene = _ptex_sheet(ref, tmp_path)
    want = baked.expected_flat_scene(scene)
    tex = capi.Ptex(ctx, GOLDEN / "ptex_jitter_u8.ptx")
    want.meshes[0]["ptex"] = tex
    w, h, spp = 96, 64, 8
    c = scene["cameras"][0]
    cam = capi.camera_init(c["position"], c["lookAt"], c["up"], c["fovDegrees"], w, h, c["znear"], c["zfar"])
    rs = capi.render_settings(w, h, spp, max_path_length=5, rr_start=8, layers=1, seed=4)
    film = np.zeros((h, w, 3), np.float32)
    err = C.create_string_buffer(512)
    rc = host.selas_host_render_baked_scene(str(tmp_path).encode() + b"/", b"ptexsheet", 0, C.byref(cam), C.byref(rs), film.ctypes.data_as(C.c_void_p), err, 512)
    assert rc == 0, err.value
    sc = capi.Scene(ctx, want)
    ref_film = sc.render(cam, rs)[0][0]
    want.meshes[0]["ptex"] = None
    plain = capi.Scene(ctx, want)
    plain_film = plain.render(cam, rs)[0][0]
    sc.close()
    plain.close()
    tex.close()
    assert ref_film.sum() > 0
    np.testing.assert_allclose(film, ref_film, rtol=2e-3, atol=2e-3)  # identical inputs: only the float atomics' order differs
    assert np.abs(plain_film - ref_film).mean() > 0.05 * ref_film.mean()  # and the texture is what makes the picture


def test_baked_scene_golden(host):
    """The committed baked scene (tests/golden/baked_scene, written by make_golden.py with the reference serializer) without the
    reference at hand."""
    from tests import baked

    scene = baked.small_two_level_scene()
    want = baked.expected_flat_scene(scene)
    counts, mats, mesh, materials = _load_baked(host, GOLDEN / "baked_scene", scene["name"])
    assert counts[:6] == [4, 12, len(want.materials), 3, 1, 1] and (counts[7] & ((1 << 62) - 1)) == (1 | (1 << 8) | (1 << 32)) and counts[7] >> 62 == 1
    np.testing.assert_array_equal(mesh, [i for i, _ in want.instances])
    np.testing.assert_array_equal(mats, np.stack([m for _, m in want.instances]))
    assert materials.tobytes() == want.materials.tobytes()


@pytest.mark.gpu
def test_selas_render_baked_scene(host, tmp_path):
    """The reference's main() for a baked scene: selas_render --scene/--assets with the DeferredPathTracer (main.cpp:115-126), one image
    per scene camera, named after the camera."""
    exe = LIB / "selas_render"
    root = str(GOLDEN / "baked_scene") + "/"
    out = subprocess.run([str(exe), "--scene", "twolevel", "--assets", root, "--integrator", "deferred", "--width", "96", "--height", "64", "--depth", "24"],
                         cwd=tmp_path, capture_output=True, text=True, check=True).stdout
    stats = json.loads(out.strip().splitlines()[-1])
    assert stats["image"] == "main" and stats["paths"] == 96 * 64 * 16 and stats["occlusion_rays"] > 0
    img = read_hdr(tmp_path / "_Images" / "main_0.hdr")
    assert img.shape == (64, 96, 3) and np.isfinite(img).all() and img.mean() > 0.01


def test_baked_scene_bounds_follow_the_instance_transforms(host):
    """SceneResource::aaBox / boundingSphere (CalculateSceneBoundingBox, SceneResource.cpp:194-203): every instance's object-space box
    through its localToWorld, not the raw mesh vertices -- the VCM integrator sizes its emission disc and merge radius with this sphere
    (VCM.cpp:778, AreaLighting.cpp:274-275).  The committed two-level scene instances its meshes with translations / rotations / scales."""
    from tests import baked

    scene = baked.small_two_level_scene()
    want = baked.expected_flat_scene(scene)
    aabb, sphere = np.zeros(6, np.float32), np.zeros(4, np.float32)
    err = C.create_string_buffer(512)
    rc = host.selas_host_baked_scene_bounds(str(GOLDEN / "baked_scene").encode() + b"/", scene["name"].encode(), aabb.ctypes.data_as(C.c_void_p),
                                            sphere.ctypes.data_as(C.c_void_p), err, 512)
    assert rc == 0, err.value
    # IncludeBox on every instance: the eight corners of the mesh's box through the instance matrix
    lo, hi = np.full(3, np.inf), np.full(3, -np.inf)
    raw_lo, raw_hi = np.full(3, np.inf), np.full(3, -np.inf)
    for mi, mat in want.instances:
        p = np.asarray(want.meshes[mi]["positions"], np.float64)
        blo, bhi = p.min(axis=0), p.max(axis=0)
        raw_lo, raw_hi = np.minimum(raw_lo, blo), np.maximum(raw_hi, bhi)
        corners = np.array([[bhi[0] if c & 1 else blo[0], bhi[1] if c & 2 else blo[1], bhi[2] if c & 4 else blo[2], 1.0] for c in range(8)])
        w = (corners @ np.asarray(mat, np.float64).reshape(4, 4))[:, :3]
        lo, hi = np.minimum(lo, w.min(axis=0)), np.maximum(hi, w.max(axis=0))
    np.testing.assert_allclose(aabb[:3], lo, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(aabb[3:], hi, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(sphere[:3], 0.5 * (lo + hi), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(sphere[3], 0.5 * np.linalg.norm(hi - lo), rtol=1e-5)
    assert np.abs(np.concatenate([lo - raw_lo, hi - raw_hi])).max() > 0.1  # the transforms matter in this scene
    np.testing.assert_allclose(scenes.bounding_sphere(want), sphere, rtol=1e-5, atol=1e-5)
    # every world-space vertex lies inside
    for mi, mat in want.instances:
        p = np.asarray(want.meshes[mi]["positions"], np.float64)
        w = (np.concatenate([p, np.ones((len(p), 1))], axis=1) @ np.asarray(mat, np.float64).reshape(4, 4))[:, :3]
        assert (w >= aabb[:3] - 1e-4).all() and (w <= aabb[3:] + 1e-4).all()


def test_baked_scene_keeps_both_instance_levels(host):
    """LoadBakedScene names every flattened instance by the reference's triple: instID[0] = subscene instance (SceneResource.cpp:135),
    instID[1] = model instance inside the subscene (:136), geomID = mesh inside the model -- what ModelDataFromRayIds takes (:362-378)."""
    from tests import baked

    scene = baked.small_two_level_scene()
    want = baked.expected_flat_scene(scene, with_ids=True)
    ids = np.zeros((len(want.instances), 3), np.uint32)
    err = C.create_string_buffer(512)
    rc = host.selas_host_baked_scene_instance_ids(str(GOLDEN / "baked_scene").encode() + b"/", scene["name"].encode(), ids.ctypes.data_as(C.c_void_p),
                                                  C.c_uint64(len(ids)), err, 512)
    assert rc == 0, err.value
    np.testing.assert_array_equal(ids, np.array(want.instance_ids, np.uint32))
    assert len({tuple(r) for r in ids}) == len(ids) and ids[:, 0].max() > 0 and ids[:, 1].max() > 0


@pytest.mark.gpu
def test_hits_report_both_instance_levels(ctx, oracle):
    """SELAS_SCENE_INSTANCE_IDS: the same rays against the same flattened / two-level scene with and without the id table -- geometry
    identical, ids translated (instID0 = subscene instance, instID1 = model instance, geomID = mesh in model); the hits go back into
    selas_cuda_surface_params unchanged, like HitParameters into ModelDataFromRayIds."""
    from tests import baked, inputs

    scene = baked.small_two_level_scene()
    plain, named = baked.expected_flat_scene(scene), baked.expected_flat_scene(scene, with_ids=True)
    table = np.array(named.instance_ids, np.uint32)
    rays = inputs.scene_rays(77, 40000, radius=9.0)
    for two_level in (False, True):
        plain.two_level = named.two_level = two_level
        a, b = capi.Scene(ctx, plain), capi.Scene(ctx, named)
        ha, hb = a.intersect(rays), b.intersect(rays)
        hit = ha["geomID"] != capi.SELAS_INVALID_ID
        assert hit.mean() > 0.2 and np.array_equal(hit, hb["geomID"] != capi.SELAS_INVALID_ID)
        for f in ("t", "u", "v", "primID"):
            np.testing.assert_array_equal(ha[f], hb[f])
        flat = ha["instID0"][hit]
        np.testing.assert_array_equal(hb["instID0"][hit], table[flat, 0])
        np.testing.assert_array_equal(hb["instID1"][hit], table[flat, 1])
        np.testing.assert_array_equal(hb["geomID"][hit], table[flat, 2])
        assert len(np.unique(hb["instID0"][hit])) > 1 and len(np.unique(hb["instID1"][hit])) > 1
        sa, sb = a.surface_params(rays, ha), b.surface_params(rays, hb)
        assert sa.tobytes() == sb.tobytes()
        bad = hb[hit][:1].copy()
        bad["instID1"] = 99
        with pytest.raises(capi.SelasError):
            b.surface_params(rays[hit][:1], bad)
        a.close()
        b.close()


@pytest.mark.parametrize("shape", [(5, 7), (17, 64), (3, 8), (33, 300), (2, 40000)])
def test_hdr_writer_matches_the_reference_stb_writer(host, ref, tmp_path, shape):
    """FrameBuffer_Save (Framebuffer.cpp:43-59) -> StbImageWrite(..., HDR, ...) -> stbi_write_hdr (TextureLib/StbImage.cpp:64-87 and
    Middleware/stb/stb_image_write.h, compiled unmodified into oracle/_ref): the product's writer against it.  Widths below 8 and above
    32767 take stb's flat (non-RLE) path, the others its per-scanline RLE; pixels are compared decoded and the files byte for byte."""
    h, w = shape
    rng = np.random.default_rng(h * 1000 + w)
    img = (rng.uniform(0, 1, size=(h, w, 3)) ** 4 * 50.0).astype(np.float32)
    img[0, 0] = 0.0
    img[-1, -1] = [1e-30, 0.0, 0.0]            # below the rgbe range: stb writes 0 0 0 0
    img[h // 2, : w // 2] = [0.25, 0.5, 1.0]   # a run, for the RLE encoder
    img[0, 1] = [65504.0, 1e-3, 3.0]
    ours, theirs = tmp_path / "ours.hdr", tmp_path / "theirs.hdr"
    assert host.selas_host_write_hdr(str(ours).encode(), w, h, img.ctypes.data_as(C.c_void_p)) == 0
    ref.write_hdr(theirs, img)
    a, b = ours.read_bytes(), theirs.read_bytes()
    # the decoder of this test reads flat RGBE; decode both through a reference-agnostic RLE-aware reader instead
    da, db = _decode_hdr(a), _decode_hdr(b)
    np.testing.assert_array_equal(da, db)
    assert a == b, "the two .hdr files differ in bytes although they decode alike"


def _decode_hdr(data):
    head, _, rest = data.partition(b"\n\n")
    assert head.startswith(b"#?RADIANCE")
    line, _, pix = rest.partition(b"\n")
    parts = line.split()
    h, w = int(parts[1]), int(parts[3])
    out = np.zeros((h, w, 4), np.uint8)
    pos = 0
    for y in range(h):
        if 8 <= w < 32768 and pix[pos] == 2 and pix[pos + 1] == 2 and (pix[pos + 2] << 8 | pix[pos + 3]) == w:
            pos += 4
            for c in range(4):
                x = 0
                while x < w:
                    n = pix[pos]
                    if n > 128:
                        out[y, x:x + n - 128, c] = pix[pos + 1]
                        pos += 2
                        x += n - 128
                    else:
                        out[y, x:x + n, c] = np.frombuffer(pix[pos + 1:pos + 1 + n], np.uint8)
                        pos += 1 + n
                        x += n
        else:
            out[y] = np.frombuffer(pix[pos:pos + 4 * w], np.uint8).reshape(w, 4)
            pos += 4 * w
    assert pos == len(pix)
    return out


def test_baked_scene_with_curve_geometry(host, ref, tmp_path):
    """Model blobs with curve geometry written by the reference's Serialize(ModelResourceData&) / (ModelGeometryData&): the loader keeps
    the curves (round 1 refused them), names them after the model's meshes (geomID = meshes + curve, ModelResource.cpp:461-466), gives
    them the curveModelName material and instances them with the model."""
    from tests import baked

    scene = baked.grass_model_scene()
    baked.write_baked(ref, tmp_path, scene)
    want = baked.expected_flat_scene(scene, with_ids=True)
    assert len(want.curves) == 2 and len(want.meshes) == 1 and len(want.instances) == 6
    counts, mats, mesh, materials = _load_baked(host, tmp_path, scene["name"])
    segments = sum(c["indices"].size for c in want.curves)
    assert counts[0] == 1 and counts[1] == 6 and counts[6] == 2 * (2 + segments)  # per model instance: the floor quad's two triangles + every segment
    np.testing.assert_array_equal(mesh, [i for i, _ in want.instances])           # curve geometries are numbered after the meshes: 1, 2
    np.testing.assert_array_equal(mats, np.stack([m for _, m in want.instances]))
    ids = np.zeros((6, 3), np.uint32)
    err = C.create_string_buffer(512)
    assert host.selas_host_baked_scene_instance_ids(str(tmp_path).encode() + b"/", scene["name"].encode(), ids.ctypes.data_as(C.c_void_p), C.c_uint64(6), err, 512) == 0
    np.testing.assert_array_equal(ids, np.array(want.instance_ids, np.uint32))    # geomID 0 = soil, 1 and 2 = the two curve geometries
    aabb, sphere = np.zeros(6, np.float32), np.zeros(4, np.float32)
    assert host.selas_host_baked_scene_bounds(str(tmp_path).encode() + b"/", scene["name"].encode(), aabb.ctypes.data_as(C.c_void_p),
                                              sphere.ctypes.data_as(C.c_void_p), err, 512) == 0
    top = max(float(c["vertices"][:, 1].max()) for c in want.curves)
    assert aabb[4] >= top * 1.4 * 0.99  # the scaled instance's blades are inside the scene box


@pytest.mark.gpu
def test_baked_scene_with_curves_renders_like_the_flat_description(ctx, host, ref, tmp_path):
    from tests import baked

    scene = baked.grass_model_scene()
    baked.write_baked(ref, tmp_path, scene)
    want = baked.expected_flat_scene(scene)
    want.two_level = True
    w, h, spp = 160, 90, 8
    c = scene["cameras"][0]
    cam = capi.camera_init(c["position"], c["lookAt"], c["up"], c["fovDegrees"], w, h, c["znear"], c["zfar"])
    rs = capi.render_settings(w, h, spp, max_path_length=5, rr_start=8, layers=1, seed=6)
    film = np.zeros((h, w, 3), np.float32)
    err = C.create_string_buffer(512)
    rc = host.selas_host_render_baked_scene(str(tmp_path).encode() + b"/", b"grass", 0, C.byref(cam), C.byref(rs), film.ctypes.data_as(C.c_void_p), err, 512)
    assert rc == 0, err.value
    sc = capi.Scene(ctx, want)
    ref_film = sc.render(cam, rs)[0][0]
    hits = sc.intersect(scenes.primary_rays(cam, w, h))
    sc.close()
    assert (hits["geomID"] >= 1).mean() > 0.02  # blades in view
    fin = np.isfinite(ref_film).all(axis=2) & np.isfinite(film).all(axis=2)
    np.testing.assert_allclose(film[fin], ref_film[fin], rtol=2e-3, atol=2e-3)
