"""CPU checks of the restated VCM integrator (oracle/orc_vcm.h; VCM.cpp is dead code in the reference, so the orchestration is
pinned only through its live building blocks -- test_oracle_vs_ref.py -- and through the properties below)."""
import numpy as np

from selas_b200 import capi, scenes

W, H = 64, 36
ALL_OFF = capi.VCM_NO_LIGHT_TRACING | capi.VCM_NO_DIRECT | capi.VCM_NO_CONNECTIONS | capi.VCM_NO_MERGING | capi.VCM_NO_SKY


def _scene():
    return scenes.caustic_scene(sphere_res=(16, 32), floor_cells=4, ibl_size=(128, 64))


def test_light_pass_arrays(oracle):
    data = _scene()
    h = oracle.scene_create(data)
    cam = scenes.caustic_camera(W, H)
    vs = capi.vcm_settings(W, H, 1, scenes.bounding_sphere(data), seed=3)
    verts, ends = oracle.vcm_light_pass(h, cam, vs, nthreads=4)
    v1, e1 = oracle.vcm_light_pass(h, cam, vs, nthreads=1)
    oracle.scene_destroy(h)
    # thread count must not change the reference-ordered arrays
    np.testing.assert_array_equal(ends, e1)
    assert verts.tobytes() == v1.tobytes()
    assert ends[-1] == verts.size and verts.size > 0 and np.all(np.diff(ends.astype(np.int64)) >= 0)
    starts = np.concatenate([[0], ends[:-1]])
    counts = ends - starts
    assert counts.max() <= vs.maxPathLength - 3  # while(pathLength + 2 < maxPathLength), VCM.cpp:509
    index = verts["indexAndLength"] & 0x3FFFFFF
    length = (verts["indexAndLength"] >> 26) & 31
    # path by path, pathLength 1, 2, ... inside a path
    np.testing.assert_array_equal(index, np.repeat(np.arange(W * H, dtype=np.uint32), counts))
    np.testing.assert_array_equal(length, np.concatenate([np.arange(1, c + 1) for c in counts]).astype(np.uint32))
    assert np.all(verts["hit"]["geomID"] != capi.SELAS_INVALID_ID)


def test_techniques_are_additive_and_shards_sum(oracle):
    data = _scene()
    h = oracle.scene_create(data)
    cam = scenes.caustic_camera(W, H)
    sphere = scenes.bounding_sphere(data)
    full, st = oracle.render_vcm(h, cam, capi.vcm_settings(W, H, 2, sphere, seed=5), nthreads=4)
    assert np.isfinite(full).all() and full.mean() > 0
    assert st.lightPaths == st.cameraPaths == 2 * W * H and st.lightVertices > 0 and st.mergesAccepted > 0 and st.connections > 0
    parts = np.zeros_like(full, dtype=np.float64)
    for keep in (capi.VCM_NO_LIGHT_TRACING, capi.VCM_NO_DIRECT, capi.VCM_NO_CONNECTIONS, capi.VCM_NO_MERGING, capi.VCM_NO_SKY):
        f, _ = oracle.render_vcm(h, cam, capi.vcm_settings(W, H, 2, sphere, seed=5, flags=ALL_OFF & ~keep), nthreads=4)
        assert f.sum() > 0
        parts += f
    np.testing.assert_allclose(parts, full, rtol=2e-4, atol=1e-4)
    # iteration shards (the multi-GPU split): passes [0,1) + [1,2) == [0,2)
    a, _ = oracle.render_vcm(h, cam, capi.vcm_settings(W, H, 1, sphere, seed=5, first_iteration=0), nthreads=4)
    b, sb = oracle.render_vcm(h, cam, capi.vcm_settings(W, H, 1, sphere, seed=5, first_iteration=1), nthreads=4)
    np.testing.assert_allclose(a.astype(np.float64) + b, full, rtol=2e-4, atol=1e-4)
    assert sb.lastRadius < capi.vcm_iteration_constants(W * H, 1, 0.0025 * sphere[3], 0.75, 1.0)[0]
    oracle.scene_destroy(h)


def test_settings_are_checked(oracle):
    data = _scene()
    cam = scenes.caustic_camera(W, H)
    no_ibl = scenes.cornell_box()
    h = oracle.scene_create(no_ibl)
    try:
        oracle.render_vcm(h, cam, capi.vcm_settings(W, H, 1, (0, 0, 0, 1)))
        raise AssertionError("VCM without an IBL must fail")
    except RuntimeError:
        pass
    oracle.scene_destroy(h)
