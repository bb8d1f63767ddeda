"""Two-level BVH (SELAS_SCENE_TWO_LEVEL_BVH): per-mesh object-space trees under a top-level tree over the instances, the GPU form of the
reference's Embree instance geometries / subscene-instance callbacks (SubsceneResource.cpp:63-79, SceneResource.cpp:87-187).  It must
return what the flattened world-space build returns (same ids, t / u / v up to the float rounding of the ray transform) while storing the
instanced geometry once."""
import numpy as np
import pytest

from selas_b200 import capi, scenes
from tests import inputs
from tests.test_gpu_parity import _assert_rel, _hit_agreement, _relmse

pytestmark = pytest.mark.gpu


def _two_level(data):
    data.two_level = True
    return data


def test_two_level_small_scene_vs_brute_force(ctx, oracle):
    data = _two_level(inputs.small_scene(instances=True))
    sc, hs = capi.Scene(ctx, data), oracle.scene_create(data)
    rays = inputs.scene_rays(41, 20000)
    g, ng = sc.intersect(rays, want_ng=True)
    o, ong = oracle.intersect(hs, rays, mode=2, want_ng=True)
    agree, _, _ = _hit_agreement(oracle, hs, rays, g, o)
    assert agree >= 0.9999, agree
    hit = (g["geomID"] != capi.SELAS_INVALID_ID) & (g["geomID"] == o["geomID"]) & (g["primID"] == o["primID"]) & (g["instID0"] == o["instID0"])
    assert hit.sum() > 5000 and len(np.unique(g["instID0"][hit])) == 5
    _assert_rel(g["t"][hit], o["t"][hit], 2e-5, "t")
    np.testing.assert_allclose(g["u"][hit], o["u"][hit], atol=5e-4)
    np.testing.assert_allclose(g["v"][hit], o["v"][hit], atol=5e-4)
    miss = g["geomID"] == capi.SELAS_INVALID_ID
    assert np.all(g["t"][miss] == rays["tfar"][miss])
    assert (sc.occluded(rays) != oracle.occluded(hs, rays, mode=2)).mean() <= 1e-4
    seg = rays.copy()
    seg["tnear"] = np.where(hit, o["t"] * 0.5, 0).astype(np.float32)
    seg["tfar"] = np.where(hit, o["t"] * 0.999, 1.0).astype(np.float32)
    assert (sc.occluded(seg) != oracle.occluded(hs, seg, mode=2)).mean() <= 1e-4
    # the surface stage resolves (instID0, geomID) the same way
    ok = g["geomID"] != capi.SELAS_INVALID_ID
    gs, os_ = sc.surface_params(rays[ok], g[ok]), oracle.surface_params(hs, rays[ok], g[ok])
    _assert_rel(gs["position"], os_["position"], 1e-6, "position", floor=1e-3)
    _assert_rel(gs["worldToTangent"], os_["worldToTangent"], 2e-5, "frame", floor=1e-3)
    oracle.scene_destroy(hs)
    sc.close()


def test_two_level_without_instance_list(ctx):
    """No instance list: one identity instance per mesh is implied, instID0 = mesh index; everything else equals the flat build."""
    flat, two = capi.Scene(ctx, inputs.small_scene()), capi.Scene(ctx, _two_level(inputs.small_scene()))
    rays = inputs.scene_rays(42, 20000)
    a, b = flat.intersect(rays), two.intersect(rays)
    same = (a["geomID"] == b["geomID"]) & (a["primID"] == b["primID"])
    assert same.mean() >= 0.9999
    hit = same & (a["geomID"] != capi.SELAS_INVALID_ID)
    np.testing.assert_array_equal(b["instID0"][hit], b["geomID"][hit])
    np.testing.assert_allclose(a["t"][hit], b["t"][hit], rtol=1e-6)
    flat.close()
    two.close()


def _forest(n_inst, two_level):
    rng = np.random.default_rng(77)
    sph = scenes.displaced_sphere(24, 48, as_quads=True, with_normals=True)
    blob = scenes.displaced_sphere(12, 20, as_quads=False)
    blob["material"] = 1
    floor = dict(positions=np.array([[-40, -2, -40], [-40, -2, 40], [40, -2, 40], [40, -2, -40]], np.float32), indices=np.array([[0, 1, 2, 3]], np.uint32),
                 material=2, lightSet=0)
    inst = [(2, np.eye(4, dtype=np.float32))]
    for k in range(n_inst):
        ang, s = rng.uniform(0, 2 * np.pi), rng.uniform(0.3, 1.2)
        c, sn = np.cos(ang), np.sin(ang)
        sx, sy, sz = s * rng.uniform(0.7, 1.3), s * rng.uniform(0.7, 1.3), s  # non-uniform scale: the inverse transform is not a rotation
        m = np.array([[sx * c, 0, -sx * sn, 0], [0, sy, 0, 0], [sz * sn, 0, sz * c, 0],
                      [rng.uniform(-30, 30), rng.uniform(-1, 6), rng.uniform(-30, 30), 1]], np.float32)
        inst.append((k % 2, m))
    mats = np.array([scenes.material(**scenes.DISNEY_MATERIAL_SET[1]), scenes.material(**scenes.DISNEY_MATERIAL_SET[6]),
                     scenes.material(**scenes.REFERENCE_MATERIALS["RoughDiffuse"])])
    lights = np.array([scenes.quad_light((0.0, 30.0, 0.0), (20.0, 0.0, 0.0), (0.0, 0.0, 20.0), (0.0, -1.0, 0.0), (40.0, 38.0, 35.0))])
    return capi.SceneData([sph, blob, floor], mats, lights, [(0, 1)], background=(0.3, 0.4, 0.6, 0.0), instances=inst, two_level=two_level)


def test_two_level_many_instances_matches_flattened(ctx):
    n_inst = 600
    flat, two = capi.Scene(ctx, _forest(n_inst, False)), capi.Scene(ctx, _forest(n_inst, True))
    fs, ts = flat.bvh_stats(), two.bvh_stats()
    per = [24 * 48 * 2, 12 * 20 * 2]
    assert fs.triangleCount == 2 + sum(per[k % 2] for k in range(n_inst))
    assert ts.triangleCount == 2 + sum(per) and ts.reserved == fs.triangleCount  # geometry stored once
    assert ts.deviceBytes * 20 < fs.deviceBytes
    cam = capi.camera_init((0.0, 14.0, -48.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 50.0, 320, 180)
    rays = scenes.primary_rays(cam, 320, 180)
    a, b = flat.intersect(rays), two.intersect(rays)
    same = (a["geomID"] == b["geomID"]) & (a["primID"] == b["primID"]) & (a["instID0"] == b["instID0"])
    assert same.mean() >= 0.999, same.mean()  # silhouette / shared-edge rays may pick the neighbour after the ray transform
    hit = same & (a["geomID"] != capi.SELAS_INVALID_ID)
    assert hit.mean() > 0.5 and len(np.unique(b["instID0"][hit])) > 200
    np.testing.assert_allclose(a["t"][hit], b["t"][hit], rtol=3e-5)
    assert (flat.occluded(rays) != two.occluded(rays)).mean() <= 1e-3
    # the integrator on top: same streams, same picture
    rs = capi.render_settings(320, 180, 8, max_path_length=6, rr_start=8, layers=1, seed=5)
    fa, fst = flat.render(cam, rs)
    fb, tst = two.render(cam, rs)
    assert abs(int(fst.closestRays) - int(tst.closestRays)) <= 0.01 * fst.closestRays
    x, y = fa[0] / 8.0, fb[0] / 8.0
    assert abs(x.mean() / y.mean() - 1.0) < 5e-3 and _relmse(y, x) < 5e-3
    flat.close()
    two.close()


def test_two_level_degenerate(ctx):
    """Instances of an empty mesh, a scene with nothing to hit, and a singular transform (rejected)."""
    sph = scenes.displaced_sphere(8, 16, as_quads=True)
    empty = dict(positions=np.zeros((3, 3), np.float32), indices=np.zeros((0, 3), np.uint32), material=0, lightSet=0)
    mats = np.array([scenes.material()])
    eye = np.eye(4, dtype=np.float32)
    data = capi.SceneData([sph, empty], mats, None, None, instances=[(1, eye), (0, eye), (1, eye)], two_level=True)
    sc = capi.Scene(ctx, data)
    rays = inputs.scene_rays(43, 4000)
    g = sc.intersect(rays)
    hit = g["geomID"] != capi.SELAS_INVALID_ID
    assert hit.any() and np.all(g["instID0"][hit] == 1) and np.all(g["geomID"][hit] == 0)
    sc.close()
    nothing = capi.Scene(ctx, capi.SceneData([empty], mats, None, None, instances=[(0, eye)], two_level=True))
    assert np.all(nothing.intersect(rays)["geomID"] == capi.SELAS_INVALID_ID) and not nothing.occluded(rays).any()
    nothing.close()
    flatm = eye.copy()
    flatm[1, 1] = 0.0
    with pytest.raises(capi.SelasError) as e:
        capi.Scene(ctx, capi.SceneData([sph], mats, None, None, instances=[(0, flatm)], two_level=True))
    assert "singular" in str(e.value)
