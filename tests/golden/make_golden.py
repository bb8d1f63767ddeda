"""Authors the committed golden vectors by running the REFERENCE itself in this container:
  * oracle/_ref/libselas_ref.so  = the unmodified reference translation units (Disney BSDF, lights, camera, CMJ, medium,
    VCM hash grid, VCM iteration constants), compiled in place from /root/reference by oracle/Makefile;
  * oracle/_ref/libptex_ref.so   = the reference's vendored Ptex v2.2 (PtexWriter for the .ptx fixtures, PtexFilter(f_bspline)
    for the expected lookups).
The reference ships no test vectors of its own, so these are the pins.  Run from the repo root:
    python tests/golden/make_golden.py
Outputs: tests/golden/shading.npz, tests/golden/ibl.npz, tests/golden/hashgrid.npz, tests/golden/ptex_*.ptx, tests/golden/ptex_lookups.npz
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))

from selas_b200 import capi  # noqa: E402
from tests import inputs, oracle_binding  # noqa: E402

OUT = Path(__file__).resolve().parent

PTEX_FILES = {
    # name: (nu, nv, ulog2, vlog2, res_jitter, dtype, nchan, seed, genmipmaps, constant_every)
    "ptex_uniform_u8": (6, 5, 3, 3, 0, 0, 3, 4, True, 0),
    "ptex_jitter_u8": (5, 4, 3, 2, 1, 0, 3, 5, True, 0),
    "ptex_jitter_nomip_u16": (4, 4, 2, 3, 2, 1, 3, 6, False, 0),
    "ptex_const_float": (4, 3, 3, 3, 0, 3, 3, 7, True, 3),
    "ptex_lowres_u8": (3, 3, 1, 1, 1, 0, 3, 8, True, 0),
    "ptex_onechan_u8": (3, 2, 3, 3, 0, 0, 1, 9, True, 0),
    "ptex_fourchan_float": (3, 3, 2, 2, 1, 3, 4, 10, True, 0),
}


def ptex_lookups(rng, nfaces, n):
    face = rng.integers(0, nfaces, size=n).astype(np.int32)
    u = rng.uniform(0, 1, size=n).astype(np.float32)
    v = rng.uniform(0, 1, size=n).astype(np.float32)
    k = n // 4  # a quarter of the lookups sit on / next to edges and corners
    edge = rng.integers(0, 6, size=k)
    vals = np.array([0.0, 1.0, 0.001, 0.999, 0.0625, 0.9375], np.float32)
    u[:k] = np.where(rng.uniform(size=k) < 0.7, vals[edge], u[:k])
    v[:k] = np.where(rng.uniform(size=k) < 0.7, vals[rng.integers(0, 6, size=k)], v[:k])
    return face, u, v


def extra():
    """Round-1 additions: VCM emission building blocks, texture lookups, a deferred-integrator image and two baked blobs, all from the
    reference objects.  `python tests/golden/make_golden.py extra` writes only these."""
    from selas_b200 import scenes
    ref = oracle_binding.load_ref()
    assert ref is not None, "build oracle/_ref first (make -C oracle) with /root/reference mounted"
    rng = np.random.default_rng(301)
    img = (rng.uniform(0, 1, size=(32, 64, 3)) ** 6 * 40.0).astype(np.float32)
    img[5:8, 10:14] = 900.0
    ibl = scenes.build_ibl(img, miss=rng.uniform(0, 2, size=(8, 16, 3)).astype(np.float32), rotation=0.7, exposure=1.0)
    data = inputs.small_scene(with_attrs=False)
    data.ibl = ibl
    hs = ref.scene_create(data)
    n = 2048
    sphere = np.array([0.25, 0.5, -0.75, 6.5], np.float32)
    u4, u2 = inputs.uniforms(302, n, 4), inputs.uniforms(303, n, 2)
    wi = inputs.unit_vectors(rng, n)
    emit_pdr, emit_pdfs = ref.ibl_emit(hs, sphere, u4)
    direct = ref.ibl_direct(hs, u2)
    rad_rgb, rad_pdfs = ref.ibl_radiance(hs, sphere, wi)
    light_state = ref.vcm_generate_light(hs, sphere, 0.37, u4)
    cam = ref.camera_init((0.0, 1.0, -5.0), (0.3, 0.0, 0.0), (0.0, 1.0, 0.0), 45.0, 40, 30)
    rs = capi.render_settings(40, 30, 16, max_path_length=64, layers=1, seed=8, cmj=(4, 4), integrator=capi.INTEGRATOR_DEFERRED)
    deferred_film, dst = ref.render(hs, cam, rs, nthreads=1)
    ref.scene_destroy(hs)
    tex = rng.uniform(0, 1, size=(7, 13, 3)).astype(np.float32)
    tex_uv = rng.uniform(-2.5, 3.5, size=(n, 2)).astype(np.float32)
    np.savez_compressed(OUT / "vcm_texture.npz", light=ibl["light"], marginal=ibl["marginal"], conditional=ibl["conditional"], miss=ibl["miss"],
                        rotation=np.float32(ibl["rotation"]), exposure_scale=np.float32(ibl["exposure_scale"]), sphere=sphere, u4=u4, u2=u2,
                        wi=wi, emit_pdr=emit_pdr, emit_pdfs=emit_pdfs, direct=direct, rad_rgb=rad_rgb, rad_pdfs=rad_pdfs,
                        light_state=light_state, deferred_film=deferred_film[0], deferred_rays=np.array([dst.closestRays, dst.occlusionRays]),
                        tex=tex, tex_uv=tex_uv, tex_srgb=ref.texture_sample(tex, tex_uv, True), tex_linear=ref.texture_sample(tex, tex_uv, False))
    # baked blobs (tests/test_host.py::_golden_assets)
    brng = np.random.default_rng(71)
    bimg = (brng.uniform(0, 1, size=(12, 24, 3)) ** 4 * 30.0).astype(np.float32)
    bibl = scenes.build_ibl(bimg, miss=brng.uniform(0, 2, size=(3, 6, 3)).astype(np.float32), rotation=0.7, exposure=1.0)
    btex = brng.uniform(0, 1, size=(5, 7, 3)).astype(np.float32)
    ref.write_ibl_blob(bibl, OUT / "baked_ibl.bin")
    ref.write_texture_blob(btex, OUT / "baked_texture.bin")
    # a whole baked scene (scene / subscene / two models + geometry / IBL), tests/baked.py
    import shutil
    from tests import baked
    shutil.rmtree(OUT / "baked_scene", ignore_errors=True)
    baked.write_baked(ref, OUT / "baked_scene", baked.small_two_level_scene())
    print("extra golden vectors written")


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "extra":
        return extra()
    ref = oracle_binding.load_ref()
    ptex = oracle_binding.load_ptexref()
    assert ref is not None and ptex is not None, "build oracle/_ref first (make -C oracle) with /root/reference mounted"

    # ---- shading vectors -------------------------------------------------------------------------------------------
    n = 4096
    s, v, nrm = inputs.random_surfaces(101, n)
    l = inputs.light_directions(101, n, nrm, v)
    u = inputs.uniforms(101, n, capi.SELAS_BSDF_RANDOMS)
    rgb, fwd, rev = ref.bsdf_eval(s, v, l)
    smp = ref.bsdf_sample(s, v, u)
    rng = np.random.default_rng(102)
    lights = inputs.random_lights(102, 4)
    pos = rng.uniform(-6, 6, size=(n, 3)).astype(np.float32)
    pos[:, 1] = rng.uniform(-3, 3, size=n)
    ln = inputs.unit_vectors(rng, n)
    u3 = inputs.uniforms(102, n, 3)
    ls = ref.light_sample(lights, pos, ln, u3)
    lpdf = ref.light_pdf(lights, ls["index"], pos, ls["direction"])
    cam = ref.camera_init((278.0, 273.0, -800.0), (278.0, 273.0, 0.0), (0.0, 1.0, 0.0), 37.0, 512, 512)
    px = rng.integers(0, 512, size=1024).astype(np.float32)
    py = rng.integers(0, 512, size=1024).astype(np.float32)
    jit = rng.uniform(size=(1024, 2)).astype(np.float32)
    rays = ref.camera_rays(cam, px, py, jit)
    cs = rng.integers(0, 16, size=1024).astype(np.int32)
    cp = rng.integers(0, 1 << 26, size=1024).astype(np.int32)
    cmj = ref.cmj(cs, 4, 4, cp)
    ext = rng.uniform(0.01, 3.0, size=(1024, 3)).astype(np.float32)
    mu = inputs.uniforms(103, 1024, 2)
    md = rng.uniform(0, 50, size=1024).astype(np.float32)
    msamp, mtrans = ref.medium(ext, mu, md)
    vcm = np.stack([ref.vcm_iteration_constants(2073600, 1, 0.05, 0.75, float(i)) for i in range(1, 9)])
    # image based light (ImageBasedLightResource.cpp) on a small synthetic environment
    from selas_b200 import scenes
    irng = np.random.default_rng(105)
    img = (irng.uniform(0, 1, size=(32, 64, 3)) ** 6 * 40.0).astype(np.float32)
    img[5:8, 10:14] = 900.0
    img[20] = 0.0
    ibl = scenes.build_ibl(img, miss=irng.uniform(0, 2, size=(8, 16, 3)).astype(np.float32), rotation=0.7, exposure=1.0)
    idata = inputs.small_scene(with_attrs=False)
    idata.ibl = ibl
    hs = ref.scene_create(idata)
    iwi = inputs.unit_vectors(irng, n)
    iu2 = inputs.uniforms(105, n, 2)
    irgb, ipdf, imiss = ref.ibl_lookup(hs, iwi)
    itpp, ixy = ref.ibl_sample(hs, iu2)
    ref.scene_destroy(hs)
    np.savez_compressed(OUT / "ibl.npz", light=ibl["light"], marginal=ibl["marginal"], conditional=ibl["conditional"], miss=ibl["miss"],
                        rotation=np.float32(ibl["rotation"]), exposure_scale=np.float32(ibl["exposure_scale"]), wi=iwi, u2=iu2,
                        rgb=irgb, pdf=ipdf, miss_rgb=imiss, theta_phi_pdf=itpp, xy=ixy)
    np.savez_compressed(OUT / "shading.npz", surfaces=s, v=v, l=l, u=u, eval_rgb=rgb, eval_fwd=fwd, eval_rev=rev, sample=smp,
                        lights=lights, light_pos=pos, light_nrm=ln, light_u3=u3, light_sample=ls, light_pdf=lpdf,
                        camera=np.frombuffer(bytes(cam), np.uint8), cam_px=px, cam_py=py, cam_jitter=jit, cam_rays=rays,
                        cmj_s=cs, cmj_p=cp, cmj=cmj, med_ext=ext, med_u=mu, med_dist=md, med_sampled=msamp, med_trans=mtrans, vcm_constants=vcm)

    # ---- hash grid ---------------------------------------------------------------------------------------------------
    rng = np.random.default_rng(104)
    hp = (rng.uniform(-1, 1, size=(6000, 3)) * 40.0).astype(np.float32)
    hp[:1500, 1] = 0.0
    radius, cells = 0.8, 6000
    ends, idx, aabb = ref.hashgrid_build(hp, radius, cells)
    hq = np.concatenate([hp[:2000] + rng.normal(scale=0.4, size=(2000, 3)).astype(np.float32),
                         (rng.uniform(-1.1, 1.1, size=(500, 3)) * 40.0).astype(np.float32)]).astype(np.float32)
    counts, found = ref.hashgrid_query(hp, radius, cells, hq, 16)
    np.savez_compressed(OUT / "hashgrid.npz", positions=hp, radius=np.float32(radius), cells=np.uint32(cells), ends=ends, indices=idx,
                        aabb=aabb, queries=hq, counts=counts, found=found)

    # ---- Ptex --------------------------------------------------------------------------------------------------------
    look = {}
    for name, (nu, nv, ul, vl, jit_, dt, nchan, seed, mip, ce) in PTEX_FILES.items():
        path = OUT / f"{name}.ptx"
        ptex.write_grid(path, nu, nv, ul, vl, jit_, dt, nchan, seed, mip, ce)
        rng = np.random.default_rng(200 + seed)
        face, uu, vv = ptex_lookups(rng, nu * nv, 3000)
        look[f"{name}_face"], look[f"{name}_u"], look[f"{name}_v"] = face, uu, vv
        look[f"{name}_rgb"] = ptex.eval(path, face, uu, vv)[:, :min(3, nchan)]
        print(name, path.stat().st_size, "bytes", ptex.info(path))
    np.savez_compressed(OUT / "ptex_lookups.npz", **look)
    extra()


if __name__ == "__main__":
    main()
