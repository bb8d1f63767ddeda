"""Pins the CPU restatement (oracle/liboracle.so) against the UNMODIFIED reference translation units
(oracle/_ref/libselas_ref.so, compiled in place from /root/reference by oracle/Makefile).

The reference ships no tests or golden vectors, so "the reference itself run here" is the
pin.  These run on CPU (-m "not gpu") and are skipped on a box without oracle/_ref (the committed vectors in
tests/golden/ then take over: test_golden.py).
"""
import numpy as np
import pytest

from selas_b200 import capi
from tests import inputs

N = 20000


def test_struct_sizes(ref):
    assert ref.sizeof("SurfaceParameters") == capi.SURFACE_DTYPE.itemsize == 156
    assert ref.sizeof("SceneLight") == capi.LIGHT_DTYPE.itemsize == 64
    assert ref.sizeof("RayCastCameraSettings") == 104
    assert ref.sizeof("BsdfSample") == 52
    assert ref.sizeof("HitParameters") == 80
    assert ref.sizeof("VCMVertex") == 108


def test_bsdf_eval_bit_exact(oracle, ref):
    s, v, nrm = inputs.random_surfaces(11, N)
    l = inputs.light_directions(11, N, nrm, v)
    a = oracle.bsdf_eval(s, v, l)
    b = ref.bsdf_eval(s, v, l)
    for x, y, name in zip(a, b, ("rgb", "fwd", "rev")):
        np.testing.assert_array_equal(x.view(np.uint32), y.view(np.uint32), err_msg=name)
    assert np.count_nonzero(a[0].sum(axis=1) > 0) > N // 3


def test_bsdf_sample_bit_exact(oracle, ref):
    s, v, _ = inputs.random_surfaces(12, N)
    u = inputs.uniforms(12, N, capi.SELAS_BSDF_RANDOMS)
    a = oracle.bsdf_sample(s, v, u)
    b = ref.bsdf_sample(s, v, u)
    for f in a.dtype.names:
        np.testing.assert_array_equal(a[f].view(np.uint32), b[f].view(np.uint32), err_msg=f)
    assert a["valid"].mean() > 0.8
    # every lobe and event type is exercised
    assert set(np.unique(a["flags"][a["valid"] == 1])) >= {1, 2, 5, 6}
    assert set(np.unique(a["randomsUsed"])) >= {1, 3, 4}


def test_light_sampling_bit_exact(oracle, ref):
    rng = np.random.default_rng(13)
    lights = inputs.random_lights(13, 5)
    pos = rng.uniform(-6, 6, size=(N, 3)).astype(np.float32)
    pos[:, 1] = rng.uniform(-3, 3, size=N)
    nrm = inputs.unit_vectors(rng, N)
    u3 = inputs.uniforms(13, N, 3)
    u3[:7, 0] = 1.0  # the p0 == 1.0 overrun: both sides clamp
    a = oracle.light_sample(lights, pos, nrm, u3)
    b = ref.light_sample(lights, pos, nrm, u3)
    for f in ("index", "direction", "radiance", "distance", "pdfW"):
        np.testing.assert_array_equal(a[f].view(np.uint32), b[f].view(np.uint32), err_msg=f)
    assert np.count_nonzero(a["radiance"].sum(axis=1) > 0) > N // 4
    # LightingPdf along the sampled direction and along random ones
    wi = np.where((rng.uniform(size=N) < 0.5)[:, None], a["direction"], inputs.unit_vectors(rng, N)).astype(np.float32)
    pa = oracle.light_pdf(lights, a["index"], pos, wi)
    pb = ref.light_pdf(lights, a["index"], pos, wi)
    np.testing.assert_array_equal(pa.view(np.uint32), pb.view(np.uint32))
    assert np.count_nonzero(pa > 0) > N // 10


def test_camera_bit_exact(oracle, ref):
    rng = np.random.default_rng(14)
    for (w, h, fov) in ((512, 512, 37.0), (1920, 1080, 40.0), (1024, 429, 53.0)):
        args = ((278.0, 273.0, -800.0), (278.0, 273.0, 0.0), (0.0, 1.0, 0.0), fov, w, h)
        ca, cb = oracle.camera_init(*args), ref.camera_init(*args)
        assert bytes(ca) == bytes(cb)
        px = rng.integers(0, w, size=5000).astype(np.float32)
        py = rng.integers(0, h, size=5000).astype(np.float32)
        jit = rng.uniform(size=(5000, 2)).astype(np.float32)
        ra, rb = oracle.camera_rays(ca, px, py, jit), ref.camera_rays(cb, px, py, jit)
        assert ra.tobytes() == rb.tobytes()
        ra, rb = oracle.camera_rays(ca, px, py), ref.camera_rays(cb, px, py)
        assert ra.tobytes() == rb.tobytes()


def test_cmj_bit_exact(oracle, ref):
    rng = np.random.default_rng(15)
    s = rng.integers(0, 16, size=4096).astype(np.int32)
    p = rng.integers(0, 1 << 26, size=4096).astype(np.int32)
    a, b = oracle.cmj(s, 4, 4, p), ref.cmj(s, 4, 4, p)
    np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32))


def test_medium_bit_exact(oracle, ref):
    rng = np.random.default_rng(16)
    ext = rng.uniform(0.01, 3.0, size=(N, 3)).astype(np.float32)
    u2 = inputs.uniforms(16, N, 2)
    dist = rng.uniform(0, 50, size=N).astype(np.float32)
    a, b = oracle.medium(ext, u2, dist), ref.medium(ext, u2, dist)
    np.testing.assert_array_equal(a[0].view(np.uint32), b[0].view(np.uint32))
    np.testing.assert_array_equal(a[1].view(np.uint32), b[1].view(np.uint32))


@pytest.mark.parametrize("n,cells,radius,scale", [(5000, 5000, 0.05, 1.0), (20000, 4099, 0.02, 1.0), (3000, 100000, 0.5, 300.0),
                                                   (1, 7, 0.1, 1.0), (64, 1, 0.1, 1.0)])
def test_hashgrid_bit_exact(oracle, ref, n, cells, radius, scale):
    rng = np.random.default_rng(17 + n)
    pos = (rng.uniform(-1, 1, size=(n, 3)) * scale).astype(np.float32)
    pos[: n // 4, 1] = 0.0  # a dense "floor" => crowded cells
    ga = oracle.hashgrid_build(pos, radius, cells)
    gb = ref.hashgrid_build(pos, radius, cells)
    for x, y in zip(ga, gb):
        np.testing.assert_array_equal(x, y)
    q = np.concatenate([pos[: min(n, 2000)] + rng.normal(scale=radius * 0.5, size=(min(n, 2000), 3)).astype(np.float32),
                        (rng.uniform(-1.2, 1.2, size=(500, 3)) * scale).astype(np.float32)]).astype(np.float32)
    ca, ia = oracle.hashgrid_query(pos, radius, cells, ga, q, 32)
    cb, ib = ref.hashgrid_query(pos, radius, cells, q, 32)
    np.testing.assert_array_equal(ca, cb)
    np.testing.assert_array_equal(ia, ib)
    if n > 100:
        assert ca.max() > 0


def test_tier_b_integrator_matches_restatement(oracle, ref):
    """The restated EvaluatePath over restated shading (orc_) vs the same loop over the reference's own
    Shading objects (ref_): identical RNG streams => images agree to float accumulation noise."""
    data = inputs.small_scene(with_attrs=True)
    cam = capi.camera_init((0.0, 1.0, -5.0), (0.3, 0.0, 0.0), (0.0, 1.0, 0.0), 45.0, 48, 32)
    rs = capi.render_settings(48, 32, 8, max_path_length=6, rr_start=3, layers=2, seed=5)
    ha, hb = oracle.scene_create(data), ref.scene_create(data)
    fa, sa = oracle.render(ha, cam, rs, nthreads=4)
    fb, sb = ref.render(hb, cam, rs, nthreads=4)
    oracle.scene_destroy(ha)
    ref.scene_destroy(hb)
    assert sa.closestRays == sb.closestRays and sa.occlusionRays == sb.occlusionRays
    for l in range(2):
        np.testing.assert_allclose(fa[l], fb[l], rtol=1e-5, atol=1e-6)
    assert fa[1].sum() > 0


def test_tier_b_integrator_cmj_pixel_sampling(oracle, ref):
    """Correlated multi-jitter primary rays (the reference's own JitteredCameraRay CMJ overload vs the restatement)."""
    data = inputs.small_scene(with_attrs=False)
    cam = capi.camera_init((0.0, 1.0, -5.0), (0.3, 0.0, 0.0), (0.0, 1.0, 0.0), 45.0, 40, 30)
    rs = capi.render_settings(40, 30, 16, max_path_length=4, rr_start=8, layers=1, seed=2, cmj=(4, 4))
    ha, hb = oracle.scene_create(data), ref.scene_create(data)
    fa, sa = oracle.render(ha, cam, rs, nthreads=2)
    fb, sb = ref.render(hb, cam, rs, nthreads=2)
    oracle.scene_destroy(ha)
    ref.scene_destroy(hb)
    assert sa.closestRays == sb.closestRays
    np.testing.assert_allclose(fa[0], fb[0], rtol=1e-5, atol=1e-6)


def _ibl_scene():
    from selas_b200 import scenes

    data = inputs.small_scene(with_attrs=False)
    rng = np.random.default_rng(31)
    img = (rng.uniform(0, 1, size=(32, 64, 3)) ** 6 * 40.0).astype(np.float32)
    img[5:8, 10:14] = 900.0  # a "sun"
    img[20] = 0.0            # an all-black row: its conditional CDF stays zero except the forced 1 at the end
    miss = rng.uniform(0, 2, size=(8, 16, 3)).astype(np.float32)
    data.ibl = scenes.build_ibl(img, miss=miss, rotation=0.7, exposure=1.0)
    return data


def test_ibl_bit_exact(oracle, ref):
    """SampleIbl / SampleIblMiss / SampleIBlPdf / Ibl (ImageBasedLightResource.cpp:118-262): restatement vs reference objects."""
    data = _ibl_scene()
    ha, hb = oracle.scene_create(data), ref.scene_create(data)
    rng = np.random.default_rng(32)
    wi = inputs.unit_vectors(rng, N)
    wi[:6] = [[0, 1, 0], [0, -1, 0], [1, 0, 0], [-1, 0, 0], [0, 0, 1], [0, 0, -1]]
    a, b = oracle.ibl_lookup(ha, wi), ref.ibl_lookup(hb, wi)
    for x, y, name in zip(a, b, ("rgb", "pdf", "miss")):
        np.testing.assert_array_equal(x.view(np.uint32), y.view(np.uint32), err_msg=name)
    u2 = inputs.uniforms(33, N, 2)
    u2[:4] = [[0, 0], [1, 1], [0, 1], [1, 0]]
    (ta, xa), (tb, xb) = oracle.ibl_sample(ha, u2), ref.ibl_sample(hb, u2)
    np.testing.assert_array_equal(xa, xb)
    np.testing.assert_array_equal(ta.view(np.uint32), tb.view(np.uint32))
    assert a[1].max() > 0 and (xa[:, 1] != 20).all()
    # the integrator with an IBL: misses read the environment (EvaluateBackground / EvaluateBackgroundMiss)
    cam = capi.camera_init((0.0, 1.0, -5.0), (0.3, 0.0, 0.0), (0.0, 1.0, 0.0), 45.0, 40, 30)
    rs = capi.render_settings(40, 30, 8, max_path_length=5, rr_start=8, layers=1, seed=3)
    fa, _ = oracle.render(ha, cam, rs, nthreads=2)
    fb, _ = ref.render(hb, cam, rs, nthreads=2)
    np.testing.assert_allclose(fa[0], fb[0], rtol=1e-5, atol=1e-6)
    oracle.scene_destroy(ha)
    ref.scene_destroy(hb)


# ---- VCM building blocks that are live code in the reference ---------------------------------------------------------------
SPHERE = (0.25, 0.5, -0.75, 6.5)


def test_vcm_ibl_emission_bit_exact(oracle, ref):
    """EmitIblLightSample / DirectIblLightSample / IblCalculateRadiance (Shading/AreaLighting.cpp:246-331): restatement vs the
    reference objects, same scripted uniforms."""
    data = _ibl_scene()
    ha, hb = oracle.scene_create(data), ref.scene_create(data)
    u4 = inputs.uniforms(41, N, 4)
    u4[:6] = [[0, 0, 0, 0], [1, 1, 1, 1], [0.5, 0.5, 0.5, 0.5], [0.2, 0.9, 0.5, 0.25], [0.3, 0.3, 0.0, 1.0], [0.7, 0.1, 1.0, 0.0]]
    (pa, fa), (pb, fb) = oracle.ibl_emit(ha, SPHERE, u4), ref.ibl_emit(hb, SPHERE, u4)
    np.testing.assert_array_equal(pa.view(np.uint32), pb.view(np.uint32))
    np.testing.assert_array_equal(fa.view(np.uint32), fb.view(np.uint32))
    assert (fa[:, 0] > 0).mean() > 0.9 and np.all(fa[:, 2] == 1.0)
    u2 = inputs.uniforms(42, N, 2)
    da, db = oracle.ibl_direct(ha, u2), ref.ibl_direct(hb, u2)
    for f in ("direction", "radiance", "distance", "pdfW"):
        np.testing.assert_array_equal(da[f].view(np.uint32), db[f].view(np.uint32), err_msg=f)
    assert np.all(da["distance"] == np.float32(1e36))
    rng = np.random.default_rng(43)
    wi = inputs.unit_vectors(rng, N)
    (ra, qa), (rb, qb) = oracle.ibl_radiance(ha, SPHERE, wi), ref.ibl_radiance(hb, SPHERE, wi)
    np.testing.assert_array_equal(ra.view(np.uint32), rb.view(np.uint32))
    np.testing.assert_array_equal(qa.view(np.uint32), qb.view(np.uint32))
    oracle.scene_destroy(ha)
    ref.scene_destroy(hb)


def test_vcm_generate_samples_match_reference(oracle, ref):
    """VCMCommon::GenerateLightSample / GenerateCameraSample (VCMCommon.cpp:37-88) from the reference objects against the
    formulas the restatement uses (orc_vcm.h), evaluated in float32 on the restatement's own emission / camera outputs."""
    f32 = np.float32
    data = _ibl_scene()
    ha, hb = oracle.scene_create(data), ref.scene_create(data)
    u4 = inputs.uniforms(44, 5000, 4)
    vc_weight = f32(0.37)
    st = ref.vcm_generate_light(hb, SPHERE, vc_weight, u4)
    pdr, pdfs = oracle.ibl_emit(ha, SPHERE, u4)
    ok = pdfs[:, 0] > 0
    emission, direction_pdf, cos_light = pdfs[:, 0], pdfs[:, 1], pdfs[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        thr = pdr[:, 6:9] * (f32(1.0) / emission)[:, None]
        dvcm, dvc = direction_pdf / emission, cos_light / emission
        dvm = cos_light / emission * vc_weight
    np.testing.assert_array_equal(st[ok, 0:6].view(np.uint32), pdr[ok, 0:6].view(np.uint32))
    np.testing.assert_array_equal(st[ok, 6:9].view(np.uint32), thr[ok].view(np.uint32))
    for col, want in ((9, dvcm), (10, dvc), (11, dvm)):
        np.testing.assert_array_equal(st[ok, col].view(np.uint32), want[ok].astype(f32).view(np.uint32))
    assert np.all(st[:, 12] == 1.0)  # pathLength 1, isAreaMeasure 0

    w, h = 64, 36
    cam = capi.camera_init((0.0, 1.0, -5.0), (0.3, 0.0, 0.0), (0.0, 1.0, 0.0), 45.0, w, h)
    rng = np.random.default_rng(45)
    xy = np.stack([rng.integers(0, w, 4000), rng.integers(0, h, 4000)], axis=1).astype(np.uint32)
    u2 = inputs.uniforms(46, 4000, 2)
    cs, index = ref.vcm_generate_camera(cam, float(w * h), xy, u2)
    rays = oracle.camera_rays(cam, xy[:, 0].astype(f32), xy[:, 1].astype(f32), u2)
    d = np.stack([rays["dx"], rays["dy"], rays["dz"]], axis=1)
    fwd = np.array(cam.forward[:], f32)
    cos_cam = (fwd[0] * d[:, 0] + fwd[1] * d[:, 1]) + fwd[2] * d[:, 2]
    dist = f32(cam.virtualImagePlaneDistance) / cos_cam
    inv_solid = dist * dist / cos_cam
    dvcm = f32(w * h) * (f32(1.0) / inv_solid)
    np.testing.assert_array_equal(cs[:, 3:6].view(np.uint32), d.view(np.uint32))
    np.testing.assert_array_equal(cs[:, 9].view(np.uint32), dvcm.astype(f32).view(np.uint32))
    assert np.all(cs[:, 6:9] == 1.0) and np.all(cs[:, 10:12] == 0.0) and np.all(cs[:, 12] == 257.0)
    np.testing.assert_array_equal(index, xy[:, 1] * w + xy[:, 0])
    # WorldToImage (restated: the reference no longer has it) inverts ImageToWorldDirection
    t = rng.uniform(0.5, 30.0, size=4000).astype(f32)
    p = np.array(cam.position[:], f32)[None, :] + d * t[:, None]
    frac = u2 - 0.5
    inner = (np.abs(frac) < 0.49).all(axis=1)  # away from pixel borders, where float rounding may pick the neighbour
    got = oracle.world_to_image(cam, p)
    np.testing.assert_array_equal(got[inner], xy[inner].astype(np.int32))
    oracle.scene_destroy(ha)
    ref.scene_destroy(hb)


def test_vcm_iteration_constants_bit_exact(oracle, ref):
    for vm, it in ((2073600, 1.0), (2073600, 7.0), (4096, 100.0), (640 * 360, 3.0)):
        a = oracle.vcm_iteration_constants(vm, 1, 0.0214, 0.75, it)
        b = ref.vcm_iteration_constants(vm, 1, 0.0214, 0.75, it)
        np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("with_ibl", [False, True])
def test_deferred_integrator_matches_reference_objects(oracle, ref, with_ibl):
    """DeferredPathTracer.cpp:61-148 walked path by path: the restatement vs the same loop over the reference's own
    NextEventEstimation / SampleBackground / BackgroundLightingPdf / BalanceHeuristic / EvaluateBackground[Miss] / BSDF objects,
    with the constant background (BackgroundLightSample, pdfW = 1) and with an IBL (DirectIblLightSample / SampleIBlPdf)."""
    data = _ibl_scene() if with_ibl else inputs.small_scene(with_attrs=False)
    cam = capi.camera_init((0.0, 1.0, -5.0), (0.3, 0.0, 0.0), (0.0, 1.0, 0.0), 45.0, 40, 30)
    rs = capi.render_settings(40, 30, 16, max_path_length=64, layers=1, seed=8, cmj=(4, 4), integrator=capi.INTEGRATOR_DEFERRED)
    ha, hb = oracle.scene_create(data), ref.scene_create(data)
    fa, sa = oracle.render(ha, cam, rs, nthreads=2)
    fb, sb = ref.render(hb, cam, rs, nthreads=2)
    oracle.scene_destroy(ha)
    ref.scene_destroy(hb)
    assert sa.closestRays == sb.closestRays and sa.occlusionRays == sb.occlusionRays
    assert sa.occlusionRays > sa.closestRays * 0.5  # two NEE samples per hit
    fin = np.isfinite(fb[0]).all(axis=2)
    assert np.array_equal(fin, np.isfinite(fa[0]).all(axis=2)) and fin.mean() > 0.99
    np.testing.assert_allclose(fa[0][fin], fb[0][fin], rtol=1e-5, atol=1e-6)
    # it is a different estimator from PathTracer::EvaluatePath on the same streams
    plain, _ = oracle_render_plain(oracle, data, cam)
    assert not np.allclose(plain, fa[0], rtol=1e-3)


def oracle_render_plain(oracle, data, cam):
    h = oracle.scene_create(data)
    f, s = oracle.render(h, cam, capi.render_settings(40, 30, 16, max_path_length=64, layers=1, seed=8, cmj=(4, 4)), nthreads=2)
    oracle.scene_destroy(h)
    return f[0], s


@pytest.mark.parametrize("shape", [(16, 32, 3), (7, 13, 3), (9, 5, 1), (8, 8, 4), (1, 1, 3)])
def test_texture_lookup_bit_exact(oracle, ref, shape):
    """SampleTextureFloat3 (SurfaceParameters.cpp:97-118): the reference's TextureFiltering::Triangle template and SrgbToLinearPrecise
    against the restatement, incl. non-power-of-two sizes where the int32 % uint32 Repeat wrap of negative coordinates is not w - 1."""
    rng = np.random.default_rng(51 + shape[0])
    tex = rng.uniform(0, 1, size=shape).astype(np.float32)
    uv = rng.uniform(-2.5, 3.5, size=(N, 2)).astype(np.float32)
    uv[:8] = [[0, 0], [1, 1], [0.5, 0.5], [-0.01, 0.99], [1e-4, -1e-4], [0.999, 0.001], [2.0, -1.0], [0.25, 0.75]]
    for srgb in (True, False):
        a, b = oracle.texture_sample(tex, uv, srgb), ref.texture_sample(tex, uv, srgb)
        np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32))
    assert a.max() > 0


def test_oracle_ptex_hook_surface_params(oracle, ptexref):
    """The CPU integrator's Ptex hook (orc_scene_set_ptex bound to the reference Ptex library): CalculateSurfaceParams gives
    baseColor = pow(PtexFilter::eval(primId, u, v), 2.2) (SurfaceParameters.cpp:196-208), checked against the committed lookups."""
    from pathlib import Path

    from selas_b200 import scenes

    G = Path(__file__).resolve().parent / "golden"
    z = np.load(G / "ptex_lookups.npz")
    data = scenes.ptex_quad_grid(6, 5, ptex=None)
    data.materials["flags"] |= capi.MAT_USES_PTEX
    hs = oracle.scene_create(data)
    th = ptexref.open(G / "ptex_uniform_u8.ptx")
    oracle.scene_set_ptex(hs, 0, ptexref, th)
    face, u, v = z["ptex_uniform_u8_face"], z["ptex_uniform_u8_u"], z["ptex_uniform_u8_v"]
    rays = np.zeros(face.size, dtype=capi.RAY_DTYPE)
    rays["dy"] = 1.0
    hits = np.zeros(face.size, dtype=capi.HIT_DTYPE)
    hits["t"], hits["u"], hits["v"], hits["primID"] = 1.0, u, v, face.astype(np.uint32)
    s = oracle.surface_params(hs, rays, hits)
    want = np.power(z["ptex_uniform_u8_rgb"].astype(np.float32), np.float32(2.2))
    assert np.abs(s["baseColor"] - want).max() <= 2e-6  # powf vs numpy's float32 power
    oracle.scene_destroy(hs)
    ptexref.close(th)


# ---- the reference's own integrator loop (oracle/ref_pathtracer.cpp: PathTracer.cpp + SurfaceParameters.cpp + Framebuffer.cpp unmodified) ----
def _loop_cases():
    from selas_b200 import scenes

    yield "cornell", scenes.cornell_box(copper_roughness=0.1), scenes.cornell_camera(48, 48), (48, 48), 8
    # all eight materials of configs[2] (metal-aniso, clearcoat, sheen, solid glass with a medium, thin glass, thin diffTrans, diffuse, copper)
    yield "disney_heightfield", scenes.disney_heightfield(grid=48), scenes.heightfield_camera(64, 36), (64, 36), 16
    yield "small_scene_attrs", inputs.small_scene(with_attrs=True), capi.camera_init((0.0, 1.0, -5.0), (0.3, 0.0, 0.0), (0.0, 1.0, 0.0), 45.0, 48, 32), (48, 32), 12
    # IBL + miss image.  The clearcoat lobe is switched off here: SampleDisneyClearcoat never writes sample.flags (SURVEY appendix B
    # #4) and EvaluatePath's BsdfSample is an uninitialised local (PathTracer.cpp:197), so after a clearcoat sample the unmodified loop
    # reads whatever the stack slot held -- with an IBL that garbage decides between EvaluateBackground and EvaluateBackgroundMiss
    # (17 of 1200 pixels with this compiler).  The restatement, the oracle and the CUDA path define it as eScatterEvent.
    ibl = _ibl_scene()
    ibl.materials["scalars"][:, capi.CLEARCOAT] = 0.0
    yield "ibl", ibl, capi.camera_init((0.0, 1.0, -5.0), (0.3, 0.0, 0.0), (0.0, 1.0, 0.0), 45.0, 40, 30), (40, 30), 10


@pytest.mark.parametrize("case", list(_loop_cases()), ids=lambda c: c[0])
def test_restated_loop_is_bit_identical_to_reference_pathtracer_cpp(oracle, ref, case):
    """Level (d) pin: PathTracer.cpp's EvaluatePath / RayPick / OcclusionRay, SurfaceParameters.cpp's CalculateSurfaceParams and
    Framebuffer.cpp's writers, compiled UNMODIFIED behind the Embree-API shim, against (1) the restated loop over reference objects
    (ref_render_pathtracer: bit for bit, both film layers, same ray counts) and (2) the plain-C++ oracle (float accumulation noise)."""
    name, data, cam, (w, h), depth = case
    hb = ref.scene_create(data)
    ha = oracle.scene_create(data)
    for layers in (1, 2):
        rs = capi.render_settings(w, h, 6, max_path_length=depth, rr_start=8, layers=layers, seed=21)
        loop, sl = ref.render_reference_loop(hb, cam, rs, nthreads=1)  # one thread: see the note on sample.flags in _loop_cases
        rest, sr = ref.render(hb, cam, rs, nthreads=3)
        assert sl.closestRays == sr.closestRays and sl.occlusionRays == sr.occlusionRays and sl.pathSamples == sr.pathSamples == w * h * 6
        for l in range(layers):
            np.testing.assert_array_equal(loop[l].view(np.uint32), rest[l].view(np.uint32), err_msg=f"{name} layer {l}")
        assert np.nanmean(loop[-1]) > 0
        orc, so = oracle.render(ha, cam, rs, nthreads=4)
        assert so.closestRays == sl.closestRays and so.occlusionRays == sl.occlusionRays
        fin = np.isfinite(loop[-1]).all(axis=2)
        assert np.array_equal(fin, np.isfinite(orc[-1]).all(axis=2)) and fin.mean() > 0.99
        np.testing.assert_allclose(orc[-1][fin], loop[-1][fin], rtol=2e-5, atol=1e-6)
    # sample-pass shards add up exactly (per-pixel sums are sequential in sample order on both sides)
    rs_a = capi.render_settings(w, h, 3, max_path_length=depth, rr_start=8, layers=1, seed=21)
    rs_b = capi.render_settings(w, h, 3, max_path_length=depth, rr_start=8, layers=1, seed=21, first_sample=3)
    fa, _ = ref.render_reference_loop(hb, cam, rs_a, nthreads=2)
    fb, _ = ref.render_reference_loop(hb, cam, rs_b, nthreads=2)
    full, _ = ref.render_reference_loop(hb, cam, capi.render_settings(w, h, 6, max_path_length=depth, rr_start=8, layers=1, seed=21), nthreads=2)
    fin = np.isfinite(full[0]).all(axis=2)
    np.testing.assert_allclose((fa[0] + fb[0])[fin], full[0][fin], rtol=1e-5, atol=1e-6)
    oracle.scene_destroy(ha)
    ref.scene_destroy(hb)


def test_reference_loop_rejects_what_pathtracer_cpp_cannot_do(ref):
    from selas_b200 import scenes

    hb = ref.scene_create(scenes.cornell_box())
    cam = scenes.cornell_camera(8, 8)
    for kw in (dict(rr_start=3), dict(mis_mode=1), dict(integrator=capi.INTEGRATOR_DEFERRED), dict(cmj=(4, 4))):
        args = dict(max_path_length=12, rr_start=8, layers=1, seed=1)
        args.update(kw)
        with pytest.raises(RuntimeError):
            ref.render_reference_loop(hb, cam, capi.render_settings(8, 8, 2, **args), nthreads=1)
    ref.scene_destroy(hb)


def test_surface_params_bit_exact_vs_reference_calculate_surface_params(oracle, ref):
    """CalculateSurfaceParams (SurfaceParameters.cpp:144-247, unmodified; ModelDataFromRayIds through the reference's own scene /
    subscene / model tables; rtcInterpolate0 through the shim) against the restatement, on hits of a scene whose meshes do and do
    not carry normals / tangents (triangles and quads)."""
    data = inputs.small_scene(with_attrs=True)
    ha, hb = oracle.scene_create(data), ref.scene_create(data)
    rays = inputs.scene_rays(31, 20000)
    hits = oracle.intersect(ha, rays, mode=0)
    assert (hits["geomID"] != capi.SELAS_INVALID_ID).mean() > 0.3
    a = oracle.surface_params(ha, rays, hits)
    b = ref.surface_params_reference(hb, rays, hits)
    for f in a.dtype.names:
        np.testing.assert_array_equal(a[f].view(np.uint32), b[f].view(np.uint32), err_msg=f)
    oracle.scene_destroy(ha)
    ref.scene_destroy(hb)


@pytest.mark.parametrize("comps", [1, 2, 3, 4])
def test_texture_filters_bit_exact_vs_reference_templates(oracle, ref, comps):
    """TextureFiltering::Point / Triangle(level) / Trilinear / EWA / EWA(level) (TextureFiltering.h:54-234): the restatement
    against the reference's own templates instantiated for float / float2 / float3 / float4 over a 7-level mip chain."""
    from tests import inputs

    levels = inputs.mip_chain(64, comps, seed=10 + comps)
    st, d0, d1 = inputs.filter_queries(3000, seed=20 + comps)
    for flt, level in ((capi.FILTER_POINT, 0), (capi.FILTER_TRIANGLE, 0), (capi.FILTER_TRIANGLE, 3), (capi.FILTER_TRIANGLE, 40), (capi.FILTER_TRILINEAR, 0),
                       (capi.FILTER_EWA, 0), (capi.FILTER_EWA_LEVEL, 2), (capi.FILTER_EWA_LEVEL, 9)):
        a = oracle.texture_filter(levels, flt, st, d0, d1, level)
        b = ref.texture_filter(levels, flt, st, d0, d1, level)
        assert np.isfinite(b).all(), flt
        np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32), err_msg=f"filter {flt} level {level}")
    # the filters do different things: EWA of a wide footprint is far from the point sample
    wide = np.linalg.norm(d1, axis=1) > 0.05
    assert np.abs(oracle.texture_filter(levels, capi.FILTER_EWA, st, d0, d1)[wide] - oracle.texture_filter(levels, capi.FILTER_POINT, st)[wide]).mean() > 0.02
