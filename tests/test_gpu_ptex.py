"""Parity level (c): the CUDA Ptex lookup against the reference's own PtexFilter(f_bspline)::eval outputs, within 1e-4
(BASELINE.json north_star).  The expected values are the committed golden vectors (tests/golden/ptex_lookups.npz, written by
the reference's vendored Ptex library); when oracle/_ref/libptex_ref.so is present the live library is consulted as well."""
from pathlib import Path

import numpy as np
import pytest

from selas_b200 import capi, scenes
from tests import oracle_binding

pytestmark = pytest.mark.gpu
G = Path(__file__).resolve().parent / "golden"
TOL = 1e-4
NAMES = ["ptex_uniform_u8", "ptex_jitter_u8", "ptex_jitter_nomip_u16", "ptex_const_float", "ptex_lowres_u8", "ptex_onechan_u8",
         "ptex_fourchan_float"]


@pytest.mark.parametrize("name", NAMES)
def test_ptex_lookup_golden(ctx, name):
    z = np.load(G / "ptex_lookups.npz")
    tex = capi.Ptex(ctx, G / f"{name}.ptx")
    got = tex.eval(z[f"{name}_face"], z[f"{name}_u"], z[f"{name}_v"])
    want = z[f"{name}_rgb"]
    err = np.abs(got[:, : want.shape[1]] - want)
    assert err.max() <= TOL, (name, float(err.max()), int((err > TOL).sum()))
    if want.shape[1] < 3:
        assert not got[:, want.shape[1]:].any()  # channels the texture lacks come back as 0
    tex.close()


def test_ptex_lookup_live_reference_large(ctx, tmp_path):
    ptexref = oracle_binding.load_ptexref()
    if ptexref is None:
        pytest.skip("oracle/_ref/libptex_ref.so not present")
    path = tmp_path / "grid.ptx"
    ptexref.write_grid(path, 40, 30, 3, 3, 1, 0, 3, 11, True, 7)  # 1200 faces, mixed 8/16 resolutions, some constant faces
    rng = np.random.default_rng(5)
    n = 200_000
    face = rng.integers(0, 1200, size=n).astype(np.int32)
    u = rng.uniform(0, 1, size=n).astype(np.float32)
    v = rng.uniform(0, 1, size=n).astype(np.float32)
    tex = capi.Ptex(ctx, path)
    got = tex.eval(face, u, v)
    want = ptexref.eval(path, face, u, v)
    err = np.abs(got - want)
    assert err.max() <= TOL, (float(err.max()), int((err > TOL).sum()))
    # out-of-range ids and coordinates: the library returns without touching the result / clamps to the face
    edge = tex.eval(np.array([-1, 1200, 5, 5], np.int32), np.array([0.5, 0.5, -3.0, 7.0], np.float32), np.array([0.5, 0.5, 0.5, 0.5], np.float32))
    assert not edge[:2].any()
    want_edge = ptexref.eval(path, np.array([5, 5], np.int32), np.array([-3.0, 7.0], np.float32), np.array([0.5, 0.5], np.float32))
    assert np.abs(edge[2:] - want_edge).max() <= TOL
    tex.close()


@pytest.mark.parametrize("log2,jitter,dtype,mips", [(3, 2, 0, True), (3, 1, 1, False), (4, 1, 3, True), (3, 3, 2, True)])
def test_ptex_mixed_resolution_neighbours_vs_live_reference(ctx, tmp_path, log2, jitter, dtype, mips):
    """Kernels that leave their face onto a neighbour of another resolution (since round 2 the register path reduces
    the piece, PtexSeparableKernel::downresU/V, or reads the neighbour's reduced copy, PtexReader::getData(faceid, res)): lookups hugging
    the edges and corners of faces whose neighbours are 2^jitter times finer / coarser along u or v, with and without stored mip levels,
    in all four data types, against the reference library itself."""
    ptexref = oracle_binding.load_ptexref()
    if ptexref is None:
        pytest.skip("oracle/_ref/libptex_ref.so not present")
    path = tmp_path / "mixed.ptx"
    nu, nv = 24, 20
    ptexref.write_grid(path, nu, nv, log2, log2, jitter, dtype, 3, 21 + jitter, mips, 11)
    rng = np.random.default_rng(17 + log2)
    n = 120_000
    face = rng.integers(0, nu * nv, size=n).astype(np.int32)
    # two thirds of the lookups within a texel and a half of an edge (both sides, both axes), the rest anywhere
    near = rng.uniform(0.0, 1.5 / (1 << log2), size=(n, 2)).astype(np.float32)
    side = rng.integers(0, 2, size=(n, 2))
    uv = np.where(side == 1, 1.0 - near, near).astype(np.float32)
    anywhere = rng.uniform(0, 1, size=(n, 2)).astype(np.float32)
    pick = rng.integers(0, 3, size=(n, 2))
    uv = np.where(pick == 0, anywhere, uv).astype(np.float32)
    tex = capi.Ptex(ctx, path)
    got = tex.eval(face, uv[:, 0].copy(), uv[:, 1].copy())
    want = ptexref.eval(path, face, uv[:, 0].copy(), uv[:, 1].copy())
    err = np.abs(got - want)
    assert err.max() <= TOL, (float(err.max()), int((err > TOL).sum()))
    tex.close()


def test_ptex_errors(ctx, tmp_path):
    bad = tmp_path / "bad.ptx"
    bad.write_bytes(b"not a ptex file at all, just sixty-four bytes of text padding.........")
    with pytest.raises(capi.SelasError) as e:
        capi.Ptex(ctx, bad)
    assert e.value.code == 4
    with pytest.raises(capi.SelasError):
        capi.Ptex(ctx, tmp_path / "missing.ptx")
    data = (G / "ptex_uniform_u8.ptx").read_bytes()
    trunc = tmp_path / "trunc.ptx"
    trunc.write_bytes(data[: len(data) // 2])
    with pytest.raises(capi.SelasError):
        capi.Ptex(ctx, trunc)


def test_ptex_in_surface_params(ctx):
    """CalculateSurfaceParams with a Ptex material (SurfaceParameters.cpp:196-208): baseColor = pow(eval(primId, u, v), 2.2)."""
    z = np.load(G / "ptex_lookups.npz")
    tex = capi.Ptex(ctx, G / "ptex_uniform_u8.ptx")
    data = scenes.ptex_quad_grid(6, 5, ptex=tex)
    sc = capi.Scene(ctx, data)
    face, u, v = z["ptex_uniform_u8_face"], z["ptex_uniform_u8_u"], z["ptex_uniform_u8_v"]
    rays = np.zeros(face.size, dtype=capi.RAY_DTYPE)
    rays["dy"] = 1.0
    hits = np.zeros(face.size, dtype=capi.HIT_DTYPE)
    hits["t"], hits["u"], hits["v"], hits["primID"] = 1.0, u, v, face.astype(np.uint32)
    s = sc.surface_params(rays, hits)
    want = np.power(z["ptex_uniform_u8_rgb"].astype(np.float32), np.float32(2.2))
    assert np.abs(s["baseColor"] - want).max() <= 2e-4
    sc.close()
    tex.close()


def test_ptex_render_matches_oracle(ctx, ptexref):
    """End to end: the wavefront PathTracer with the Ptex stage between closest-hit and shade, against the CPU integrator whose
    Ptex hook calls the reference's own PtexFilter (same RNG streams, so the images differ only by float rounding and the
    1e-4 lookup tolerance)."""
    orc = oracle_binding.load_oracle()
    path = G / "ptex_jitter_u8.ptx"
    nu, nv = 5, 4  # the fixture's face grid (tests/golden/make_golden.py)
    assert ptexref.info(path)[0] == nu * nv
    tex = capi.Ptex(ctx, path)
    data = scenes.ptex_quad_grid(nu, nv, ptex=tex)
    cam = scenes.ptex_grid_camera(96, 64)
    rs = capi.render_settings(96, 64, 16, max_path_length=5, rr_start=8, layers=1, seed=5)
    sc = capi.Scene(ctx, data)
    film, st = sc.render(cam, rs)
    hs = orc.scene_create(data)
    th = ptexref.open(path)
    orc.scene_set_ptex(hs, 0, ptexref, th)
    ofilm, ost = orc.render(hs, cam, rs, nthreads=8)
    flat = scenes.ptex_quad_grid(nu, nv, ptex=None)
    hs2 = orc.scene_create(flat)
    pfilm, _ = orc.render(hs2, cam, rs, nthreads=8)
    orc.scene_destroy(hs)
    orc.scene_destroy(hs2)
    ptexref.close(th)
    g, o, p = film[0] / 16.0, ofilm[0] / 16.0, pfilm[0] / 16.0
    relmse = float(np.mean((g - o) ** 2 / (o ** 2 + 1e-2)))
    textured = float(np.mean((p - o) ** 2 / (o ** 2 + 1e-2)))
    assert st.closestRays == ost.closestRays or abs(int(st.closestRays) - int(ost.closestRays)) <= 1e-3 * ost.closestRays
    assert abs(g.mean() / o.mean() - 1.0) < 5e-3 and relmse < 5e-3, (g.mean(), o.mean(), relmse)
    assert textured > 20 * max(relmse, 1e-6), (textured, relmse)  # the texture is what the comparison is sensitive to
    sc.close()
    tex.close()
