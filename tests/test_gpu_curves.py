"""Round B-spline curve geometry (RTC_GEOMETRY_TYPE_ROUND_BSPLINE_CURVE, ModelResource.cpp:300-333) through the C ABI,
against the oracle's double-precision checker (oracle/orc_curves.h: 128 parts per segment, exhaustive over the segments).

Parity is UNPINNED like all ray casting (Embree is absent): the checker restates the documented sweep-surface semantics
(RTC_GEOMETRY_TYPE_CURVE.3embree3).  The CUDA intersector cuts a segment into 8 cones with spherical joints in fp32, so the two agree
up to that subdivision: silhouette rays may disagree on hit / miss, t agrees to about 1e-3 of the blade radius over the distance.
"""
import numpy as np
import pytest

from selas_b200 import capi, scenes

pytestmark = pytest.mark.gpu
W, H = 320, 180


def _unit(a):
    return a / np.maximum(np.linalg.norm(a, axis=1, keepdims=True), 1e-30)


def test_curve_hits_match_the_double_precision_checker(ctx, oracle):
    data = scenes.grass_patch(blades=400)
    cam = scenes.grass_camera(W, H)
    rays = scenes.primary_rays(cam, W, H)
    sc, hs = capi.Scene(ctx, data), oracle.scene_create(data)
    g, gng = sc.intersect(rays, want_ng=True)
    o, ong = oracle.intersect(hs, rays, mode=1, want_ng=True)
    curve_geom = len(data.meshes)
    gc, oc = g["geomID"] == curve_geom, o["geomID"] == curve_geom
    assert oc.mean() > 0.03, oc.mean()                      # the blades are in view
    same_geom = g["geomID"] == o["geomID"]
    assert same_geom.mean() > 0.995, same_geom.mean()       # silhouette rays may go either way
    both = gc & oc & (g["primID"] == o["primID"])
    assert both.sum() > 0.97 * oc.sum(), (both.sum(), oc.sum())
    # the reported record: u along the segment, v = 0, the element code in the last word, t close to the finer subdivision's
    assert np.all(g["v"][gc] == 0.0) and np.all(g["pad"][gc] > 0) and np.all(g["pad"][~gc] == 0)
    assert np.all((g["u"][gc] >= 0.0) & (g["u"][gc] <= 1.0))
    terr = np.abs(g["t"][both] - o["t"][both])
    print(f"curves: {oc.sum()} rays on blades, same segment {both.sum()}, |dt| median {np.median(terr):.2e} p99 {np.quantile(terr, 0.99):.2e} max {terr.max():.2e}")
    assert np.quantile(terr, 0.99) < 2e-3 and np.median(terr) < 1e-4
    assert np.abs(g["u"][both] - o["u"][both]).max() < 0.05 and np.median(np.abs(g["u"][both] - o["u"][both])) < 2e-3
    cosang = np.sum(_unit(gng[both][:, :3]) * _unit(ong[both][:, :3]), axis=1)
    # eight parts against 128: the normals differ by the chord angle of a part (measured: median 0.7 degrees, 2 % beyond 3 degrees)
    assert np.quantile(cosang, 0.02) > 0.997 and np.median(cosang) > 0.9999, (np.quantile(cosang, 0.02), np.median(cosang))
    # triangles of the same scene are untouched by the curve variant of the kernel
    tri = (g["geomID"] == 0) & (o["geomID"] == 0)
    np.testing.assert_array_equal(g["primID"][tri], o["primID"][tri])
    np.testing.assert_allclose(g["t"][tri], o["t"][tri], rtol=1e-5)
    # any hit: shadow rays from the floor towards the light through the grass
    floor = (o["geomID"] == 0)
    sh = np.zeros(int(floor.sum()), dtype=capi.RAY_DTYPE)
    p = np.stack([rays[k][floor] + o["t"][floor] * rays["d" + k[1]][floor] for k in ("ox", "oy", "oz")], axis=1)
    to = np.array([0.0, 6.0, 0.0], np.float32) - p
    dist = np.linalg.norm(to, axis=1)
    sh["ox"], sh["oy"], sh["oz"] = p[:, 0], p[:, 1] + 1e-4, p[:, 2]
    sh["dx"], sh["dy"], sh["dz"] = (to / dist[:, None]).T
    sh["tnear"], sh["tfar"] = 1e-4, dist
    go, oo = sc.occluded(sh), oracle.occluded(hs, sh, mode=1)
    assert 0.02 < oo.mean() < 0.9 and (go == oo).mean() > 0.99, (oo.mean(), (go == oo).mean())
    # the hits go back into CalculateSurfaceParams as they came out (the element code rebuilds the normal)
    sel = np.nonzero(both)[0][:20000]
    gs, osf = sc.surface_params(rays[sel], g[sel]), oracle.surface_params(hs, rays[sel], o[sel])
    gn, on = gs["worldToTangent"].reshape(-1, 3, 3)[:, :, 1], osf["worldToTangent"].reshape(-1, 3, 3)[:, :, 1]
    assert np.quantile(np.sum(gn * on, axis=1), 0.02) > 0.997
    np.testing.assert_allclose(gs["baseColor"], osf["baseColor"], rtol=1e-5)
    assert np.all(gs["lightSetIndex"] == osf["lightSetIndex"]) and np.all(gs["shader"] == osf["shader"])
    oracle.scene_destroy(hs)
    sc.close()


def test_curve_scene_renders_like_the_checker(ctx, oracle):
    """PathTracer over a scene with curve geometry, same RNG streams as the oracle's loop over its double-precision curve checker.  The
    blades' normals differ by the chord angle of the coarser subdivision (a fraction of a degree), which moves every blade pixel a little
    and lets silhouette paths diverge: the comparison is on the mean (measured 1.00002), on 4 x 4 binned pixels (relMSE 6e-5) and on the
    pixel majority (78 % within 2 %: the floor and sky pixels)."""
    data = scenes.grass_patch(blades=300)
    cam = scenes.grass_camera(W, H)
    spp = 16
    rs = capi.render_settings(W, H, spp, max_path_length=6, rr_start=8, layers=1, seed=13)
    sc, hs = capi.Scene(ctx, data), oracle.scene_create(data)
    g, gs = sc.render(cam, rs)
    o, os_ = oracle.render(hs, cam, rs, nthreads=16)
    g, o = g[0] / spp, o[0] / spp
    assert abs(int(gs.closestRays) - int(os_.closestRays)) < 0.01 * os_.closestRays
    assert abs(g.mean() / o.mean() - 1.0) < 2e-3, (g.mean(), o.mean())
    close = np.isclose(g, o, rtol=2e-2, atol=1e-3).all(axis=2)
    gb, ob = g.reshape(H // 4, 4, W // 4, 4, 3).mean(axis=(1, 3)), o.reshape(H // 4, 4, W // 4, 4, 3).mean(axis=(1, 3))
    relmse = float(np.mean((gb - ob) ** 2 / (ob ** 2 + 1e-2)))
    print(f"grass render vs checker: {close.mean():.4f} of the pixels within 2 %, 4 x 4 binned relMSE {relmse:.3e}, mean ratio {g.mean() / o.mean():.5f}")
    assert close.mean() > 0.7 and relmse < 5e-4
    oracle.scene_destroy(hs)
    sc.close()


def test_instanced_curves_flattened_and_two_level_agree(ctx, oracle):
    """A curve geometry instanced three times (identity, rotated + translated, uniformly scaled) together with a mesh: the flattened tree
    (control vertices transformed into the pool) and the two-level tree (ray transformed into the geometry's space) report the same
    segments, and both agree with the checker."""
    data = scenes.grass_patch(blades=120, extent=1.5)

    def xform(tx, ty, tz, s, ang):
        c, sn = np.cos(ang), np.sin(ang)
        return np.array([[s * c, 0, -s * sn, 0], [0, s, 0, 0], [s * sn, 0, s * c, 0], [tx, ty, tz, 1]], np.float32)

    curve_geom = len(data.meshes)
    inst = [(0, np.eye(4, dtype=np.float32)), (curve_geom, np.eye(4, dtype=np.float32)), (curve_geom, xform(3.5, 0.0, 1.0, 1.0, 0.8)),
            (curve_geom, xform(-3.5, 0.2, 0.5, 1.6, -0.5))]
    flat = capi.SceneData(data.meshes, data.materials, data.lights, [(0, 1)], background=(0.3, 0.45, 0.7, 0.0), instances=inst, curves=data.curves)
    two = capi.SceneData(data.meshes, data.materials, data.lights, [(0, 1)], background=(0.3, 0.45, 0.7, 0.0), instances=inst, curves=data.curves,
                         two_level=True)
    cam = scenes.grass_camera(W, H)
    rays = scenes.primary_rays(cam, W, H)
    a, b = capi.Scene(ctx, flat), capi.Scene(ctx, two)
    hs = oracle.scene_create(flat)
    ha, hb, ho = a.intersect(rays), b.intersect(rays), oracle.intersect(hs, rays, mode=1)
    on_curves = ho["geomID"] == curve_geom
    assert on_curves.mean() > 0.02 and len(np.unique(ho["instID0"][on_curves])) == 3
    for h, name in ((ha, "flattened"), (hb, "two-level")):
        same = (h["geomID"] == ho["geomID"]) & (h["instID0"] == ho["instID0"]) & (h["primID"] == ho["primID"])
        assert same.mean() > 0.99, (name, same.mean())
        m = same & on_curves
        assert np.quantile(np.abs(h["t"][m] - ho["t"][m]), 0.99) < 4e-3, name
    assert ((ha["geomID"] == hb["geomID"]) & (ha["primID"] == hb["primID"]) & (ha["instID0"] == hb["instID0"])).mean() > 0.995
    # rendering the two forms gives the same picture up to silhouette paths
    rs = capi.render_settings(W, H, 8, max_path_length=5, rr_start=8, layers=1, seed=2)
    fa, fb = a.render(cam, rs)[0][0], b.render(cam, rs)[0][0]
    assert abs(fa.mean() / fb.mean() - 1.0) < 0.01
    oracle.scene_destroy(hs)
    a.close()
    b.close()


def test_curve_descriptor_errors(ctx):
    data = scenes.grass_patch(blades=4)
    bad = scenes.grass_patch(blades=4)
    bad.curves[0]["indices"] = np.array([10 ** 6], np.uint32)
    with pytest.raises(capi.SelasError):
        capi.Scene(ctx, bad)
    bad = scenes.grass_patch(blades=4)
    bad.curves[0]["material"] = 99
    with pytest.raises(capi.SelasError):
        capi.Scene(ctx, bad)
    sc = capi.Scene(ctx, data)
    assert sc.bvh_stats().triangleCount == 2 + 4 * 3 * 8  # the floor quad's two triangles and twelve curve segments of eight parts each
    sc.close()
