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
