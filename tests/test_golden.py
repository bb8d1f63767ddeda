"""The CPU restatement (oracle/liboracle.so) against the committed golden vectors in tests/golden/, which were produced by
running the reference's own translation units (tests/golden/make_golden.py).  Runs everywhere (no GPU, no reference tree)."""
from pathlib import Path

import numpy as np

from selas_b200 import capi

G = Path(__file__).resolve().parent / "golden"


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def test_shading_golden(oracle):
    z = np.load(G / "shading.npz")
    rgb, fwd, rev = oracle.bsdf_eval(z["surfaces"], z["v"], z["l"])
    np.testing.assert_array_equal(_bits(rgb), _bits(z["eval_rgb"]))
    np.testing.assert_array_equal(_bits(fwd), _bits(z["eval_fwd"]))
    np.testing.assert_array_equal(_bits(rev), _bits(z["eval_rev"]))
    smp = oracle.bsdf_sample(z["surfaces"], z["v"], z["u"])
    assert smp.tobytes() == z["sample"].tobytes()
    ls = oracle.light_sample(z["lights"], z["light_pos"], z["light_nrm"], z["light_u3"])
    assert ls.tobytes() == z["light_sample"].tobytes()
    lp = oracle.light_pdf(z["lights"], z["light_sample"]["index"], z["light_pos"], z["light_sample"]["direction"])
    np.testing.assert_array_equal(_bits(lp), _bits(z["light_pdf"]))


def test_camera_cmj_medium_golden(oracle):
    z = np.load(G / "shading.npz")
    cam = oracle.camera_init((278.0, 273.0, -800.0), (278.0, 273.0, 0.0), (0.0, 1.0, 0.0), 37.0, 512, 512)
    assert bytes(cam) == z["camera"].tobytes()
    rays = oracle.camera_rays(cam, z["cam_px"], z["cam_py"], z["cam_jitter"])
    assert rays.tobytes() == z["cam_rays"].tobytes()
    np.testing.assert_array_equal(_bits(oracle.cmj(z["cmj_s"], 4, 4, z["cmj_p"])), _bits(z["cmj"]))
    sampled, trans = oracle.medium(z["med_ext"], z["med_u"], z["med_dist"])
    np.testing.assert_array_equal(_bits(sampled), _bits(z["med_sampled"]))
    np.testing.assert_array_equal(_bits(trans), _bits(z["med_trans"]))


def test_ibl_golden(oracle):
    from tests import inputs

    z = np.load(G / "ibl.npz")
    data = inputs.small_scene(with_attrs=False)
    data.ibl = dict(light=z["light"], marginal=z["marginal"], conditional=z["conditional"], miss=z["miss"], rotation=float(z["rotation"]),
                    exposure_scale=float(z["exposure_scale"]))
    hs = oracle.scene_create(data)
    rgb, pdf, miss = oracle.ibl_lookup(hs, z["wi"])
    tpp, xy = oracle.ibl_sample(hs, z["u2"])
    oracle.scene_destroy(hs)
    np.testing.assert_array_equal(_bits(rgb), _bits(z["rgb"]))
    np.testing.assert_array_equal(_bits(pdf), _bits(z["pdf"]))
    np.testing.assert_array_equal(_bits(miss), _bits(z["miss_rgb"]))
    np.testing.assert_array_equal(_bits(tpp), _bits(z["theta_phi_pdf"]))
    np.testing.assert_array_equal(xy, z["xy"])


def test_hashgrid_golden(oracle):
    z = np.load(G / "hashgrid.npz")
    ends, idx, aabb = oracle.hashgrid_build(z["positions"], float(z["radius"]), int(z["cells"]))
    np.testing.assert_array_equal(ends, z["ends"])
    np.testing.assert_array_equal(idx, z["indices"])
    np.testing.assert_array_equal(aabb, z["aabb"])
    counts, found = oracle.hashgrid_query(z["positions"], float(z["radius"]), int(z["cells"]), (ends, idx, aabb), z["queries"], 16)
    np.testing.assert_array_equal(counts, z["counts"])
    np.testing.assert_array_equal(found, z["found"])
    assert counts.max() > 1


def test_ptex_fixtures_match_reference(ptexref):
    """When the reference Ptex library is available, the committed lookups are reproduced exactly by it."""
    z = np.load(G / "ptex_lookups.npz")
    names = sorted({k[:-5] for k in z.files if k.endswith("_face")})
    assert len(names) >= 6
    for name in names:
        got = ptexref.eval(G / f"{name}.ptx", z[f"{name}_face"], z[f"{name}_u"], z[f"{name}_v"])
        np.testing.assert_array_equal(got[:, : z[f"{name}_rgb"].shape[1]], z[f"{name}_rgb"])


def test_record_layouts():
    assert capi.C.sizeof(capi.RenderSettings) == 52 and capi.C.sizeof(capi.RenderStats) == 96
    assert capi.C.sizeof(capi.VcmSettings) == 56 and capi.C.sizeof(capi.VcmStats) == 104 and capi.VCM_VERTEX_DTYPE.itemsize == 96
    assert capi.C.sizeof(capi.MeshDesc) == 80 and capi.C.sizeof(capi.TextureDesc) == 24 and capi.C.sizeof(capi.SceneDesc) == 120 and capi.C.sizeof(capi.IblDesc) == 56
    assert capi.C.sizeof(capi.CurveDesc) == 40 and capi.C.sizeof(capi.InstanceDesc) == 80 and capi.C.sizeof(capi.BvhStats) == 40
    assert capi.C.sizeof(capi.TextureMipsDesc) == 8 + 64 + 64 + 128 + 8 + 8


def test_vcm_texture_deferred_golden(oracle):
    """Round-1 additions (make_golden.py extra): EmitIblLightSample / DirectIblLightSample / IblCalculateRadiance, the texture lookup
    and a DeferredPathTracer image, all produced by the reference objects."""
    from tests import inputs

    z = np.load(G / "vcm_texture.npz")
    data = inputs.small_scene(with_attrs=False)
    data.ibl = dict(light=z["light"], marginal=z["marginal"], conditional=z["conditional"], miss=z["miss"], rotation=float(z["rotation"]),
                    exposure_scale=float(z["exposure_scale"]))
    hs = oracle.scene_create(data)
    pdr, pdfs = oracle.ibl_emit(hs, z["sphere"], z["u4"])
    np.testing.assert_array_equal(_bits(pdr), _bits(z["emit_pdr"]))
    np.testing.assert_array_equal(_bits(pdfs), _bits(z["emit_pdfs"]))
    assert oracle.ibl_direct(hs, z["u2"]).tobytes() == z["direct"].tobytes()
    rgb, rp = oracle.ibl_radiance(hs, z["sphere"], z["wi"])
    np.testing.assert_array_equal(_bits(rgb), _bits(z["rad_rgb"]))
    np.testing.assert_array_equal(_bits(rp), _bits(z["rad_pdfs"]))
    cam = oracle.camera_init((0.0, 1.0, -5.0), (0.3, 0.0, 0.0), (0.0, 1.0, 0.0), 45.0, 40, 30)
    rs = capi.render_settings(40, 30, 16, max_path_length=64, layers=1, seed=8, cmj=(4, 4), integrator=capi.INTEGRATOR_DEFERRED)
    film, st = oracle.render(hs, cam, rs, nthreads=1)
    oracle.scene_destroy(hs)
    assert [st.closestRays, st.occlusionRays] == list(z["deferred_rays"])
    fin = np.isfinite(z["deferred_film"]).all(axis=2)
    np.testing.assert_allclose(film[0][fin], z["deferred_film"][fin], rtol=1e-5, atol=1e-6)
    np.testing.assert_array_equal(_bits(oracle.texture_sample(z["tex"], z["tex_uv"], True)), _bits(z["tex_srgb"]))
    np.testing.assert_array_equal(_bits(oracle.texture_sample(z["tex"], z["tex_uv"], False)), _bits(z["tex_linear"]))


def test_texture_filter_golden(oracle):
    """Point / Triangle(level) / Trilinear / EWA / EWA(level) against vectors written by the reference's own TextureFiltering templates
    (tests/golden/make_texture_filter.py).  Trilinear and EWA go through log2f / expf of the C library that produced the fixture: they are
    compared to 2e-6, everything else bit for bit."""
    from tests import inputs
    from tests.golden.make_texture_filter import CASES

    z = np.load(G / "texture_filter.npz")
    for comps in (1, 3, 4):
        levels = inputs.mip_chain(32, comps, seed=30 + comps)
        st, d0, d1 = inputs.filter_queries(600, seed=40 + comps)
        for flt, level in CASES:
            got, want = oracle.texture_filter(levels, flt, st, d0, d1, level), z[f"c{comps}_f{flt}_l{level}"]
            if flt in (capi.FILTER_POINT, capi.FILTER_TRIANGLE):
                np.testing.assert_array_equal(_bits(got), _bits(want))
            else:
                np.testing.assert_allclose(got, want, rtol=2e-6, atol=2e-7)
