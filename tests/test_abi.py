"""Host-side checks that need no GPU: the C-ABI library loads and exports every symbol include/selas_cuda.h declares,
fails loudly (no fallback) without a CUDA device, and the host logic (material import, scene generators, sample sharding)."""
import ctypes as C
import json
import re
from pathlib import Path

import numpy as np
import pytest

from selas_b200 import capi, distributed, scenes

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "selas_cuda.h").read_text()
    return sorted(set(re.findall(r"\b(selas_cuda_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"libselas_cuda.so does not export {n}"
    assert set(names) == set(capi.EXPORTS)
    assert b"sm_100a" in lib.selas_cuda_version()


def test_no_silent_fallback_without_gpu():
    lib = capi.load_library()
    if lib.selas_cuda_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.SelasError) as e:
        capi.Context(0)
    assert e.value.code == 1 and "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    for p in (ROOT / "selas_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h", ".cpp") and p.is_file():
            text = p.read_text()
            assert "oracle_binding" not in text and "liboracle" not in text and "oracle/" not in text.replace("the oracle", ""), p


def test_material_import_defaults():
    m = scenes.material_from_json("{}")  # Content/Materials/Default.json
    assert m["shader"] == capi.SHADER_DISNEY_SOLID and m["flags"] == 0
    np.testing.assert_allclose(m["baseColor"], [0.6, 0.6, 0.6])
    s = m["scalars"]
    assert s[capi.METALLIC] == np.float32(0.3) and s[capi.ROUGHNESS] == 0.5 and s[capi.IOR] == 1.5 and s[capi.SPEC_TRANS] == 0
    glass = scenes.material_from_json(json.dumps(scenes.REFERENCE_MATERIALS["BSDF"]))
    assert glass["flags"] & capi.MAT_TRANSPARENT and glass["scalars"][capi.SCATTER_DISTANCE] == 15.0
    thin = scenes.material(**scenes.REFERENCE_MATERIALS["DiffuseTransmissionThin"])
    assert thin["shader"] == capi.SHADER_DISNEY_THIN and not (thin["flags"] & capi.MAT_TRANSPARENT)  # eTransparent is solid-only
    paper = scenes.material(**scenes.REFERENCE_MATERIALS["Paper"])  # texture name: constant falls back to 0.6
    np.testing.assert_allclose(paper["baseColor"], [0.6, 0.6, 0.6])
    assert scenes.material(shaderName="deltatransparent")["shader"] == capi.SHADER_DIRAC_TRANSPARENT
    with pytest.raises(KeyError):
        scenes.material(bogus=1.0)


def test_scene_generators():
    c = scenes.cornell_box()
    assert c.triangle_count() == 2 * (3 + 5 + 1 + 1 + 5) and c.lights.size == 1
    s = scenes.raycast_scene(20, 40)
    assert s.triangle_count() == 2 * 20 * 40
    pos = s.meshes[0]["positions"]
    r = np.linalg.norm(pos, axis=1)
    assert 0.9 < r.min() and r.max() < 1.1
    h = scenes.disney_heightfield(grid=64, spheres=2, sphere_res=(6, 8))
    assert h.triangle_count() == 2 * 64 * 64 + 2 * 2 * 6 * 8 and len(h.materials) == 8
    cam = scenes.raycast_camera(64, 36)
    rays = scenes.primary_rays(cam, 64, 36)
    d = np.stack([rays["dx"], rays["dy"], rays["dz"]], 1)
    np.testing.assert_allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-6)


def test_shard_samples_partition():
    for total in (0, 1, 7, 256):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                first, count = distributed.shard_samples(total, r, world)
                seen += list(range(first, first + count))
            assert seen == list(range(total))
    with pytest.raises(ValueError):
        distributed.shard_samples(8, 2, 2)
