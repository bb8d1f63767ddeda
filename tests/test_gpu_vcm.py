"""GPU parity of the VCM integrator (selas_cuda_render_vcm, selas_b200/csrc/vcm.cu) against the CPU restatement
(oracle/orc_vcm.h) through the C ABI: same counter-keyed RNG streams => the GPU and the oracle walk the same light and camera
sub-paths, so arrays compare element by element and images far tighter than statistically.

Tolerances: integer / index data (pathEnds, index:26|pathLength:5 words, hit ids, the hash grid) exact up to the handful of paths
where a 1-ulp libm difference (sinf/cosf/powf/acosf) flips a branch; float fields 1e-4 relative; images relMSE < 2e-3 with the
mean within 0.5 %."""
import numpy as np
import pytest

from selas_b200 import capi, scenes

pytestmark = pytest.mark.gpu

W, H = 160, 90
ALL_OFF = capi.VCM_NO_LIGHT_TRACING | capi.VCM_NO_DIRECT | capi.VCM_NO_CONNECTIONS | capi.VCM_NO_MERGING | capi.VCM_NO_SKY


def _relmse(a, b):
    return float(np.mean((a - b) ** 2 / (b ** 2 + 1e-2)))


@pytest.fixture(scope="module")
def world(ctx, oracle):
    data = scenes.caustic_scene(sphere_res=(32, 64), floor_cells=8, ibl_size=(256, 128))
    scene = capi.Scene(ctx, data)
    h = oracle.scene_create(data)
    yield data, scene, h, scenes.caustic_camera(W, H), scenes.bounding_sphere(data)
    oracle.scene_destroy(h)
    scene.close()


def test_iteration_constants_exact(oracle):
    for vm, it in ((2073600, 1.0), (2073600, 9.0), (W * H, 2.0)):
        a = capi.vcm_iteration_constants(vm, 1, 0.0214, 0.75, it)
        b = oracle.vcm_iteration_constants(vm, 1, 0.0214, 0.75, it)
        np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32))


def test_light_pass_matches_oracle(world, oracle):
    data, scene, h, cam, sphere = world
    vs = capi.vcm_settings(W, H, 1, sphere, seed=11)
    gv, ge = scene.vcm_light_pass(cam, vs)
    ov, oe = oracle.vcm_light_pass(h, cam, vs, nthreads=8)
    n = W * H
    gs, os_ = np.concatenate([[0], ge[:-1]]), np.concatenate([[0], oe[:-1]])
    gc, oc = ge - gs, oe - os_
    same_len = gc == oc
    assert same_len.mean() >= 0.999, f"{(~same_len).sum()} of {n} light paths differ in length"
    assert ge[-1] == gv.size and abs(int(ge[-1]) - int(oe[-1])) <= 0.002 * oe[-1]
    # vertices of the paths with equal length, paired in order
    gi = np.concatenate([np.arange(s, s + c) for s, c in zip(gs[same_len], gc[same_len])]).astype(np.int64)
    oi = np.concatenate([np.arange(s, s + c) for s, c in zip(os_[same_len], oc[same_len])]).astype(np.int64)
    g, o = gv[gi], ov[oi]
    np.testing.assert_array_equal(g["indexAndLength"] & 0x7FFFFFFF, o["indexAndLength"] & 0x7FFFFFFF)
    same_hit = (g["hit"]["geomID"] == o["hit"]["geomID"]) & (g["hit"]["primID"] == o["hit"]["primID"])
    assert same_hit.mean() >= 0.999
    np.testing.assert_array_equal((g["indexAndLength"] >> 31)[same_hit], (o["indexAndLength"] >> 31)[same_hit])
    for f, tol in (("throughput", 2e-4), ("dVCM", 2e-4), ("dVC", 2e-4), ("dVM", 2e-4), ("continuationProbability", 1e-5)):
        a, b = g[f][same_hit].astype(np.float64), o[f][same_hit].astype(np.float64)
        a, b = a.reshape(a.shape[0], -1), b.reshape(b.shape[0], -1)
        err = np.abs(a - b).max(axis=1) / (np.abs(b).max(axis=1) + 1e-6)
        assert (err > tol).mean() <= 2e-3, f"{f}: {(err > tol).sum()} of {err.size} beyond {tol} (max {err.max():.3g})"
    gp = np.stack([g["ray"]["ox"] + g["hit"]["t"] * g["ray"]["dx"], g["ray"]["oy"] + g["hit"]["t"] * g["ray"]["dy"],
                   g["ray"]["oz"] + g["hit"]["t"] * g["ray"]["dz"]], axis=1)[same_hit]
    op = np.stack([o["ray"]["ox"] + o["hit"]["t"] * o["ray"]["dx"], o["ray"]["oy"] + o["hit"]["t"] * o["ray"]["dy"],
                   o["ray"]["oz"] + o["hit"]["t"] * o["ray"]["dz"]], axis=1)[same_hit]
    assert (np.abs(gp - op).max(axis=1) > 1e-3).mean() <= 2e-3


@pytest.mark.parametrize("name,flags", [
    ("light tracing", ALL_OFF & ~capi.VCM_NO_LIGHT_TRACING),
    ("direct", ALL_OFF & ~capi.VCM_NO_DIRECT),
    ("connections", ALL_OFF & ~capi.VCM_NO_CONNECTIONS),
    ("merging", ALL_OFF & ~capi.VCM_NO_MERGING),
    ("sky", ALL_OFF & ~capi.VCM_NO_SKY),
    ("full", 0),
])
def test_image_matches_oracle(world, oracle, name, flags):
    data, scene, h, cam, sphere = world
    vs = capi.vcm_settings(W, H, 2, sphere, seed=7, flags=flags)
    gf, gst = scene.render_vcm(cam, vs)
    of, ost = oracle.render_vcm(h, cam, vs, nthreads=8)
    assert np.isfinite(gf).all() and np.isfinite(of).all()
    assert gst.lightPaths == ost.lightPaths == 2 * W * H and gst.kernelLaunches > 0
    assert abs(int(gst.lightVertices) - int(ost.lightVertices)) <= 0.002 * ost.lightVertices
    if flags == 0:
        for f in ("closestRays", "occlusionRays", "mergeCandidates", "mergesAccepted", "connections"):
            a, b = int(getattr(gst, f)), int(getattr(ost, f))
            assert abs(a - b) <= 0.005 * b + 4, (f, a, b)
    assert of.mean() > 0
    assert abs(gf.mean() / of.mean() - 1.0) < 5e-3, (name, gf.mean(), of.mean())
    # splats of single diverged paths can be large; relMSE over a 4x4-binned image is the robust form of "same picture"
    gb = gf.reshape(H // 2, 2, W // 2, 2, 3).mean(axis=(1, 3))
    ob = of.reshape(H // 2, 2, W // 2, 2, 3).mean(axis=(1, 3))
    assert _relmse(gb, ob) < 2e-3, (name, _relmse(gb, ob))
    close = np.abs(gf - of) <= 1e-3 * (np.abs(of) + 1e-2)
    assert close.mean() > 0.97, (name, close.mean())


def test_iteration_shards_sum(world):
    """Multi-GPU split: passes [0,1) and [1,2) rendered separately add up to [0,2)."""
    data, scene, h, cam, sphere = world
    full, st = scene.render_vcm(cam, capi.vcm_settings(W, H, 2, sphere, seed=3))
    a, _ = scene.render_vcm(cam, capi.vcm_settings(W, H, 1, sphere, seed=3, first_iteration=0))
    b, sb = scene.render_vcm(cam, capi.vcm_settings(W, H, 1, sphere, seed=3, first_iteration=1))
    # float atomics add in a different order run to run: compare with an absolute floor
    np.testing.assert_allclose(a.astype(np.float64) + b, full, rtol=1e-3, atol=1e-3)
    assert sb.lastRadius < st.lastRadius * 1.0001 and st.renderMs > 0


def test_techniques_are_additive_at_1080p(ctx):
    """Size-independent property at the configs[4] frame size: the five MIS-weighted techniques, rendered one at a time, sum to
    the full image; and the light-vertex arrays are consistent (pathEnds monotone, counts bounded by maxPathLength - 3)."""
    data = scenes.caustic_scene()
    scene = capi.Scene(ctx, data)
    w, h = 1920, 1080
    cam = scenes.caustic_camera(w, h)
    sphere = scenes.bounding_sphere(data)
    full, st = scene.render_vcm(cam, capi.vcm_settings(w, h, 1, sphere, seed=9, flags=capi.VCM_SKIP_NONFINITE))
    assert np.isfinite(full).all() and st.lightVertices > w * h // 4 and st.mergesAccepted > 0 and st.connections > 0
    parts = np.zeros(full.shape, np.float64)
    for keep in (capi.VCM_NO_LIGHT_TRACING, capi.VCM_NO_DIRECT, capi.VCM_NO_CONNECTIONS, capi.VCM_NO_MERGING, capi.VCM_NO_SKY):
        f, _ = scene.render_vcm(cam, capi.vcm_settings(w, h, 1, sphere, seed=9, flags=(ALL_OFF & ~keep) | capi.VCM_SKIP_NONFINITE))
        parts += f
    assert abs(parts.mean() / full.mean() - 1.0) < 1e-4
    np.testing.assert_allclose(parts, full, rtol=2e-3, atol=2e-3)
    verts, ends = scene.vcm_light_pass(cam, capi.vcm_settings(w, h, 1, sphere, seed=9))
    counts = np.diff(np.concatenate([[0], ends]).astype(np.int64))
    assert ends[-1] == verts.size == st.lightVertices and counts.min() >= 0 and counts.max() <= 7
    length = (verts["indexAndLength"] >> 26) & 31
    index = verts["indexAndLength"] & 0x3FFFFFF
    np.testing.assert_array_equal(index, np.repeat(np.arange(w * h, dtype=np.uint32), counts))
    assert length.min() == 1 and length.max() <= 7
    scene.close()


def test_vcm_errors(ctx):
    scene = capi.Scene(ctx, scenes.cornell_box())
    cam = scenes.cornell_camera(32, 32)
    with pytest.raises(capi.SelasError) as e:
        scene.render_vcm(cam, capi.vcm_settings(32, 32, 1, (0, 0, 0, 1)))
    assert e.value.code == 2 and "image based light" in str(e.value)
    scene.close()
    data = scenes.caustic_scene(sphere_res=(8, 16), floor_cells=2, ibl_size=(64, 32))
    scene = capi.Scene(ctx, data)
    cam = scenes.caustic_camera(32, 32)
    for bad in (dict(max_path_length=32), dict(radius_factor=0.0)):
        with pytest.raises(capi.SelasError):
            scene.render_vcm(cam, capi.vcm_settings(32, 32, 1, scenes.bounding_sphere(data), **bad))
    with pytest.raises(capi.SelasError):
        scene.render_vcm(cam, capi.vcm_settings(32, 32, 1, (0, 0, 0, 0)))
    # maxPathLength <= 3: no light vertices at all, still a valid render (direct + sky only)
    f, st = scene.render_vcm(cam, capi.vcm_settings(32, 32, 1, scenes.bounding_sphere(data), max_path_length=3))
    assert st.lightVertices == 0 and np.isfinite(f).all() and f.sum() > 0
    scene.close()


def test_vcm_degenerate_inputs(ctx, oracle):
    """No geometry at all (every path escapes: the image is the IBL seen through ConnectToSkyLight with pathLength 1), a 1x1 and a
    ragged 33x7 frame (queue sizes that are not multiples of the warp / tile sizes), zero iterations."""
    ibl = scenes.sun_sky_ibl(64, 32)
    empty = capi.SceneData([], np.array([scenes.material()]), None, None, background=(0, 0, 0, 0), ibl=ibl)
    sc, hs = capi.Scene(ctx, empty), oracle.scene_create(empty)
    cam = scenes.caustic_camera(33, 7)
    vs = capi.vcm_settings(33, 7, 2, (0.0, 0.0, 0.0, 5.0), seed=2)
    g, gst = sc.render_vcm(cam, vs)
    o, ost = oracle.render_vcm(hs, cam, vs, nthreads=2)
    assert gst.lightVertices == ost.lightVertices == 0 and gst.occlusionRays == 0 and gst.mergesAccepted == 0
    np.testing.assert_allclose(g, o, rtol=1e-6, atol=1e-7)
    assert g.min() > 0  # sky everywhere
    oracle.scene_destroy(hs)
    sc.close()
    data = scenes.caustic_scene(sphere_res=(8, 16), floor_cells=2, ibl_size=(64, 32))
    sc, hs = capi.Scene(ctx, data), oracle.scene_create(data)
    sphere = scenes.bounding_sphere(data)
    for w, h in ((1, 1), (33, 7)):
        cam = scenes.caustic_camera(w, h)
        vs = capi.vcm_settings(w, h, 3, sphere, seed=5)
        g, gst = sc.render_vcm(cam, vs)
        o, ost = oracle.render_vcm(hs, cam, vs, nthreads=1)
        assert np.isfinite(g).all() and gst.lightPaths == 3 * w * h
        assert abs(int(gst.lightVertices) - int(ost.lightVertices)) <= 1
        np.testing.assert_allclose(g.sum(), o.sum(), rtol=2e-2, atol=1e-3)
    z, zst = sc.render_vcm(scenes.caustic_camera(8, 8), capi.vcm_settings(8, 8, 0, sphere))
    assert not z.any() and zst.lightPaths == 0
    oracle.scene_destroy(hs)
    sc.close()


@pytest.mark.parametrize("name,flags", [("connections", ALL_OFF & ~capi.VCM_NO_CONNECTIONS), ("full", 0)])
def test_vcm_with_ptex_matches_oracle(ctx, oracle, name, flags):
    """Ptex base colours through VCM: the camera and light shade stages look them up, and a light vertex keeps its colour for the
    surface ConnectPathVertices rebuilds (VCM.cpp:314-317).  The oracle's CalculateSurfaceParams calls the reference PtexFilter."""
    from pathlib import Path

    from tests import oracle_binding

    ptexref = oracle_binding.load_ptexref()
    if ptexref is None:
        pytest.skip("oracle/_ref/libptex_ref.so not built")
    path = Path(__file__).resolve().parent / "golden" / "ptex_jitter_u8.ptx"  # 5 x 4 faces
    tex = capi.Ptex(ctx, path)
    data = scenes.ptex_quad_grid(5, 4, ptex=tex)
    # the sheet's geometric normal is -y and the camera sits below it: a sky whose bright patch is in the lower hemisphere
    img = np.full((64, 128, 3), 0.2, np.float32)
    img[44:52, 40:48] = 300.0
    data.ibl = scenes.build_ibl(img, miss=np.full((16, 32, 3), 0.1, np.float32), rotation=0.3, exposure=0.0)
    sphere = scenes.bounding_sphere(data)
    w, h = 96, 64
    cam = scenes.ptex_grid_camera(w, h)
    sc = capi.Scene(ctx, data)
    hs = oracle.scene_create(data)
    th = ptexref.open(path)
    oracle.scene_set_ptex(hs, 0, ptexref, th)
    vs = capi.vcm_settings(w, h, 2, sphere, seed=11, flags=flags)
    gf, gst = sc.render_vcm(cam, vs)
    of, ost = oracle.render_vcm(hs, cam, vs, nthreads=8)
    # the same render with the constant base colour must be clearly different, or the comparison proves nothing
    plain = scenes.ptex_quad_grid(5, 4, ptex=None)
    plain.ibl = data.ibl
    hp = oracle.scene_create(plain)
    pf, _ = oracle.render_vcm(hp, cam, vs, nthreads=8)
    oracle.scene_destroy(hs)
    oracle.scene_destroy(hp)
    ptexref.close(th)
    assert np.isfinite(gf).all() and of.mean() > 0
    assert abs(int(gst.lightVertices) - int(ost.lightVertices)) <= 0.002 * ost.lightVertices + 2
    assert abs(gf.mean() / of.mean() - 1.0) < 5e-3, (name, gf.mean(), of.mean())
    scale = 1.0 / of.mean()  # connections alone are a small term on this nearly flat sheet: compare in units of the image mean
    gb = gf.reshape(h // 2, 2, w // 2, 2, 3).mean(axis=(1, 3)) * scale
    ob = of.reshape(h // 2, 2, w // 2, 2, 3).mean(axis=(1, 3)) * scale
    pb = pf.reshape(h // 2, 2, w // 2, 2, 3).mean(axis=(1, 3)) * scale
    assert _relmse(gb, ob) < 2e-3, (name, _relmse(gb, ob))
    assert _relmse(pb, ob) > 20 * max(_relmse(gb, ob), 1e-4), (name, _relmse(pb, ob), _relmse(gb, ob))
    assert abs(pf.mean() / of.mean() - 1.0) > 0.1
    sc.close()
    tex.close()
