"""Authors the converged-image pins for parity level (d) by running the REFERENCE's own integrator loop in this container:
oracle/_ref/libselas_ref.so :: refpt_render = PathTracer.cpp's EvaluatePath + SurfaceParameters.cpp + Framebuffer.cpp compiled
unmodified behind the Embree-API shim of oracle/ref_pathtracer.cpp (ray casting by the oracle's CPU BVH: Embree is absent).

    python tests/golden/make_converged.py            # both images (about 20 minutes on 8 cores)

Outputs (float32 means of the radiance layer, plus the noise floor an independent reference render of the test's spp has against them):
  tests/golden/converged_cornell_512.npz      BASELINE configs[0]: Cornell box 512x512, depth 8, 4096 spp, stored as
                                              2 x 2 binned means (256 x 256) to keep the fixture below 1 MB
  tests/golden/converged_heightfield_320.npz  configs[2] downsized: disney_heightfield(grid=128), all eight materials, four lights,
                                              depth 16, RR after 8, 320x180, 2048 spp
SURVEY 8(d)(v): relMSE = mean((img - ref)^2 / (ref^2 + 1e-2)); a render passes when its relMSE against the converged image is within
1.5 x that of an independent reference render at the same spp, and its mean is within 1 % per channel.
"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))

from selas_b200 import capi, scenes  # noqa: E402
from tests import oracle_binding  # noqa: E402

OUT = Path(__file__).resolve().parent
CONVERGED_SEED = 1001   # streams of the converged image
FLOOR_SEED = 2002       # streams of the independent same-spp render that defines the noise floor


def bin2(img):
    """2 x 2 box filter over the finite pixels (the reference integrator emits a few non-finite samples: a pixel that got one is left out)."""
    a = np.where(np.isfinite(img).all(axis=2, keepdims=True), img, np.nan).astype(np.float64)
    h, w = a.shape[0] // 2, a.shape[1] // 2
    with np.errstate(all="ignore"):
        return np.nanmean(a[: 2 * h, : 2 * w].reshape(h, 2, w, 2, 3), axis=(1, 3))


def relmse(img, ref, keep=0.99):
    """SURVEY 8(d)(v) relMSE over the pixels finite on both sides, without the `1 - keep` worst ones (fireflies of the unclamped estimator)."""
    img, ref = np.asarray(img, np.float64), np.asarray(ref, np.float64)
    ok = np.isfinite(img).all(axis=2) & np.isfinite(ref).all(axis=2)
    e = np.sort(((img - ref) ** 2 / (ref ** 2 + 1e-2)).mean(axis=2)[ok])
    return float(e[: int(len(e) * keep)].mean())


def accumulate(ref, h, cam, w, hgt, depth, spp, seed, batch, threads):
    acc = np.zeros((hgt, w, 3), np.float64)
    done = 0
    while done < spp:
        k = min(batch, spp - done)
        rs = capi.render_settings(w, hgt, k, max_path_length=depth, rr_start=8, layers=1, seed=seed, first_sample=done)
        film, _ = ref.render_reference_loop(h, cam, rs, nthreads=threads)
        acc += film[0]
        done += k
    return (acc / spp).astype(np.float32)


def make(name, data, cam, w, hgt, depth, spp, test_spp, threads=8, binned=False):
    ref = oracle_binding.load_ref()
    assert ref is not None, "build oracle/_ref first (make -C oracle) with /root/reference mounted"
    h = ref.scene_create(data)
    t0 = time.time()
    conv = accumulate(ref, h, cam, w, hgt, depth, spp, CONVERGED_SEED, 64, threads)
    floor_img = accumulate(ref, h, cam, w, hgt, depth, test_spp, FLOOR_SEED, test_spp, threads)
    ref.scene_destroy(h)
    if binned:  # keep the fixture small: 2 x 2 binned means; the test bins its render the same way
        conv, floor_img = bin2(conv).astype(np.float32), bin2(floor_img).astype(np.float32)
    floor = relmse(floor_img, conv)
    ok = np.isfinite(conv).all(axis=2) & np.isfinite(floor_img).all(axis=2)
    np.savez_compressed(OUT / f"{name}.npz", image=conv.astype(np.float32), spp=np.int32(spp), test_spp=np.int32(test_spp), binned=np.int32(binned),
                        noise_floor_relmse=np.float64(floor), channel_means=conv[ok].mean(axis=0).astype(np.float64),
                        floor_channel_means=floor_img[ok].mean(axis=0).astype(np.float64), nonfinite_pixels=np.int32((~ok).sum()),
                        converged_seed=np.int32(CONVERGED_SEED), floor_seed=np.int32(FLOOR_SEED), depth=np.int32(depth))
    print(f"{name}: {w}x{hgt} x {spp} spp in {time.time() - t0:.0f} s; noise floor of {test_spp} spp = {floor:.3e}; "
          f"means {conv[ok].mean(axis=0)}, non-finite pixels {(~ok).sum()}", flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["cornell", "heightfield"]
    if "heightfield" in which:
        make("converged_heightfield_320", scenes.disney_heightfield(grid=128), scenes.heightfield_camera(320, 180), 320, 180, 16, 2048, 64)
    if "cornell" in which:
        make("converged_cornell_512", scenes.cornell_box(copper_roughness=0.1), scenes.cornell_camera(512, 512), 512, 512, 8, 4096, 64, binned=True)
