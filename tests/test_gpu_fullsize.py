"""BASELINE.json's full sizes through size-independent properties (the oracle only finishes subsets in seconds):
  configs[1]  1M-triangle displaced sphere, 1920x1080 primary rays: closest-hit / any-hit consistency, the reported primitive really
              is hit at the reported t (float64 recomputation from the mesh), run-to-run determinism, and a 200k-ray subset against the
              oracle's BVH;
  configs[2]  ~10M-triangle Disney scene, 1080p, depth 16: sample-pass shards add up, ray accounting, finiteness."""
import numpy as np
import pytest

from selas_b200 import capi, scenes

pytestmark = pytest.mark.gpu


def test_config2_raycast_properties(ctx, oracle):
    data = scenes.raycast_scene()
    assert data.triangle_count() == 1_000_000
    sc = capi.Scene(ctx, data)
    cam = scenes.raycast_camera(1920, 1080)
    rays = scenes.primary_rays(cam, 1920, 1080)
    assert rays.size == 2_073_600
    hits = sc.intersect(rays)
    again = sc.intersect(rays)
    assert hits.tobytes() == again.tobytes()  # persistent threads + dynamic fetch, but every ray's result is a pure function of the ray
    hit = hits["geomID"] != capi.SELAS_INVALID_ID
    assert 0.8 < hit.mean() < 0.95
    # any-hit agrees with closest-hit: nothing in front of the reported hit, something at it
    seg = rays.copy()
    seg["tfar"] = np.where(hit, hits["t"] * np.float32(0.999), rays["tfar"])
    assert not sc.occluded(seg)[hit].any()
    seg["tfar"] = np.where(hit, np.where(hit, hits["t"], 1.0) * np.float32(1.001), rays["tfar"])
    occ = sc.occluded(seg)
    assert occ[hit].all() and not occ[~hit].any()
    # the reported primitive is hit at the reported (t, u, v): float64 recomputation on a random subset
    rng = np.random.default_rng(1)
    pick = rng.choice(np.nonzero(hit)[0], size=100_000, replace=False)
    mesh = data.meshes[0]
    quad = mesh["indices"][hits["primID"][pick]]
    P = mesh["positions"].astype(np.float64)
    u, v = hits["u"][pick].astype(np.float64), hits["v"][pick].astype(np.float64)
    p0, p1, p2, p3 = P[quad[:, 0]], P[quad[:, 1]], P[quad[:, 2]], P[quad[:, 3]]
    first = (u + v) <= 1.0  # RTC_GEOMETRY_TYPE_QUAD.3embree3:38-52: triangles (v0,v1,v3) and (v2,v3,v1), u/v flipped on the second
    surf = np.where(first[:, None], p0 + u[:, None] * (p1 - p0) + v[:, None] * (p3 - p0),
                    p2 + (1 - u)[:, None] * (p3 - p2) + (1 - v)[:, None] * (p1 - p2))
    o = np.stack([rays["ox"], rays["oy"], rays["oz"]], axis=1)[pick].astype(np.float64)
    d = np.stack([rays["dx"], rays["dy"], rays["dz"]], axis=1)[pick].astype(np.float64)
    at = o + hits["t"][pick].astype(np.float64)[:, None] * d
    err = np.linalg.norm(at - surf, axis=1)
    assert np.quantile(err, 0.999) < 2e-5 and err.max() < 2e-4, (np.quantile(err, 0.999), err.max())
    # a subset against the oracle's BVH (double precision traversal)
    sub = rng.choice(rays.size, size=200_000, replace=False)
    hs = oracle.scene_create(data)
    o_hits = oracle.intersect(hs, rays[sub], mode=1, nthreads=16)
    same = (hits["geomID"][sub] == o_hits["geomID"]) & (hits["primID"][sub] == o_hits["primID"])
    mism = np.nonzero(~same)[0]
    if mism.size:
        tg = oracle.prim_distance(hs, rays[sub][mism], hits[sub][mism])
        to = oracle.prim_distance(hs, rays[sub][mism], o_hits[mism])
        tie = np.count_nonzero((tg >= 0) & (to >= 0) & (np.abs(tg - to) <= 1e-5 * np.maximum(np.abs(to), 1e-3)))
    else:
        tie = 0
    assert (same.sum() + tie) / sub.size >= 0.9999
    oracle.scene_destroy(hs)
    sc.close()


def test_config3_render_properties(ctx):
    data = scenes.disney_heightfield()
    assert data.triangle_count() > 10_000_000
    sc = capi.Scene(ctx, data)
    cam = scenes.heightfield_camera(1920, 1080)
    w, h = 1920, 1080
    full, st = sc.render(cam, capi.render_settings(w, h, 2, max_path_length=16, rr_start=8, layers=1, seed=1, flags=capi.RENDER_SKIP_NONFINITE))
    a, sa = sc.render(cam, capi.render_settings(w, h, 1, 16, 8, 1, seed=1, first_sample=0, flags=capi.RENDER_SKIP_NONFINITE))
    b, sb = sc.render(cam, capi.render_settings(w, h, 1, 16, 8, 1, seed=1, first_sample=1, flags=capi.RENDER_SKIP_NONFINITE))
    assert st.pathSamples == 2 * w * h and sa.pathSamples == sb.pathSamples == w * h
    # the same paths are walked whichever way the samples are sharded: ray accounting is exact, the film up to float atomics' order
    assert st.closestRays == sa.closestRays + sb.closestRays and st.occlusionRays == sa.occlusionRays + sb.occlusionRays
    assert st.closestRays >= st.pathSamples and st.occlusionRays <= st.closestRays
    assert np.isfinite(full[0]).all()
    np.testing.assert_allclose(a[0].astype(np.float64) + b[0], full[0], rtol=1e-3, atol=1e-3)
    assert 0.2 < full[0].mean() / 2 < 5.0
    again, st2 = sc.render(cam, capi.render_settings(w, h, 2, 16, 8, 1, seed=1, flags=capi.RENDER_SKIP_NONFINITE))
    assert (st2.closestRays, st2.occlusionRays) == (st.closestRays, st.occlusionRays)
    other, st3 = sc.render(cam, capi.render_settings(w, h, 2, 16, 8, 1, seed=2, flags=capi.RENDER_SKIP_NONFINITE))
    # another seed walks other paths to the same picture; the mean is firefly-dominated (NEE without MIS, PathTracer.cpp:215), so compare
    # the median pixel
    assert st3.closestRays != st.closestRays
    assert abs(np.median(other[0].sum(axis=2)) / np.median(full[0].sum(axis=2)) - 1.0) < 0.03
    sc.close()


def test_config5_hashgrid_properties(ctx, oracle):
    """8M light vertices / 8M cells (the configs[4] microbench size): the arrays are a counting sort of the points by hashed cell --
    checked as invariants in numpy -- and the range query of a vertex always returns the vertex itself; a slice of queries is compared
    with the oracle's SearchHashGrid over the GPU-built arrays."""
    n = cells = 8_000_000
    rng = np.random.default_rng(4)
    pos = np.empty((n, 3), np.float32)
    pos[: n // 2] = np.stack([rng.uniform(-6, 6, n // 2), np.zeros(n // 2), rng.uniform(-6, 6, n // 2)], axis=1)  # the floor
    d = rng.normal(size=(n - n // 2, 3))
    pos[n // 2:] = (d / np.linalg.norm(d, axis=1, keepdims=True) * 1.2 + [0, 1.2, 0]).astype(np.float32)          # the sphere
    radius = np.float32(0.0025 * 8.57)
    grid = capi.HashGrid(ctx, pos, float(radius), cells)
    ends, idx, aabb = grid.read()
    np.testing.assert_array_equal(aabb, np.concatenate([pos.min(axis=0), pos.max(axis=0)]))
    assert ends[-1] == n and np.all(np.diff(ends.astype(np.int64)) >= 0)
    assert np.array_equal(np.sort(idx), np.arange(n, dtype=np.uint32))  # a permutation
    # every point sits in the cell its position hashes to (GetCellIndex with the reference's 64-bit arithmetic), ascending inside a cell
    inv = np.float32(1.0) / (np.float32(2.0) * radius)
    c = np.floor(inv * (pos - aabb[:3])).astype(np.int64)
    h = ((c[:, 0] * 73856093) ^ (c[:, 1] * 19349663) ^ (c[:, 2] * 83492791)).astype(np.uint64) % np.uint64(cells)
    cell_of_slot = np.searchsorted(ends, np.arange(n), side="right")
    np.testing.assert_array_equal(h[idx], cell_of_slot.astype(np.uint64))
    same_cell = cell_of_slot[1:] == cell_of_slot[:-1]
    assert np.all(idx[1:][same_cell] > idx[:-1][same_cell])
    # a vertex is always within its own radius; compare a slice of queries with the oracle walking the GPU-built arrays
    pick = rng.choice(n, size=20000, replace=False)
    counts, found = grid.query(pos[pick], 64)
    assert counts.min() >= 1 and np.all((found == pick[:, None].astype(np.uint32)).any(axis=1) | (counts > 64))
    oc, of = oracle.hashgrid_query(pos, float(radius), cells, (ends, idx, aabb), pos[pick], 64)
    np.testing.assert_array_equal(counts, oc)
    np.testing.assert_array_equal(found, of)
    grid.close()
