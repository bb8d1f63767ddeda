// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Thin extern "C" wrappers (prefix ref_) around the UNMODIFIED reference translation units, which
// oracle/Makefile compiles in place from /root/reference into oracle/_ref/libselas_ref.so.  Nothing from the
// reference is copied into this repository; this file only calls it.  It provides
//   * the few link-time symbols the reference's platform layer would supply (allocator, asset path, file
//     read) and a CSampler whose uniform stream can be scripted, so that SampleBsdfFunction /
//     NextEventEstimation can be driven with the same numbers as the CUDA kernels;
//   * ref_render_pathtracer: the reference's own EvaluateBsdf / SampleBsdfFunction / NextEventEstimation /
//     medium functions driven by a restated EvaluatePath loop (PathTracer.cpp:142-258) over the oracle's
//     CPU BVH, because Embree (the reference's ray caster) is not available.  This is BASELINE.md's tier B
//     "reference shading + restated ray cast".
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

#include <xmmintrin.h>
#include <pmmintrin.h>

#include "Shading/SurfaceScattering.h"
#include "Shading/SurfaceParameters.h"
#include "Shading/AreaLighting.h"
#include "Shading/IntegratorContexts.h"
#include "Shading/VolumetricScattering.h"
#include "Shading/Scattering.h"
#include "SceneLib/SceneResource.h"
#include "SceneLib/SubsceneResource.h"
#include "SceneLib/ModelResource.h"
#include "SceneLib/ImageBasedLightResource.h"
#include "GeometryLib/Camera.h"
#include "GeometryLib/Ray.h"
#include "GeometryLib/CoordinateSystem.h"
#include "MathLib/Sampler.h"
#include "MathLib/CorrelatedMultiJitter.h"
#include "MathLib/FloatFuncs.h"
#include "MathLib/ColorSpace.h"
#include "TextureLib/TextureFiltering.h"
#include "TextureLib/TextureResource.h"
#include "TextureLib/StbImage.h"
#include "IoLib/Directory.h"
#include "MathLib/ImportanceSampling.h"
#include "IoLib/File.h"
#include "IoLib/BinaryStreamSerializer.h"
#include "Assets/AssetFileUtils.h"
#include "VCMHashGrid.h"
#include "VCMCommon.h"

#include "../include/selas_cuda.h"
#include "cpu_bvh.h"

static_assert(sizeof(SelasSurface) == sizeof(Selas::SurfaceParameters), "SurfaceParameters layout");
static_assert(sizeof(SelasLight) == sizeof(Selas::SceneLight), "SceneLight layout");
static_assert(sizeof(SelasCamera) == sizeof(Selas::RayCastCameraSettings), "RayCastCameraSettings layout");
static_assert(sizeof(Selas::BsdfSample) == 52, "BsdfSample layout");
static_assert(sizeof(Selas::LightDirectSample) == sizeof(SelasLightSample), "LightDirectSample layout");

// ---- scripted uniform source ------------------------------------------------------------------------------
#include "ref_harness.h"
using namespace refh;
thread_local refh::DrawSource refh::g_src = {nullptr, 0, 0, 0, false};
thread_local std::mt19937* refh::g_mt = nullptr;

namespace Selas {
// platform shims -------------------------------------------------------------------------------------------
void* SelasMalloc(uint64 size, const char*, const char*, int) { return malloc(size); }
void* SelasRealloc(void* p, uint64 size, const char*, const char*, int) { return realloc(p, size); }
void SelasFree(void* p) { free(p); }
void* SelasAlignedMalloc(uint64 size, uint64 alignment, const char*, const char*, int) {
    void* p = nullptr;
    if (posix_memalign(&p, alignment < sizeof(void*) ? sizeof(void*) : alignment, size)) return nullptr;
    return p;
}
void SelasAlignedFree(void* p) { free(p); }
namespace AssetFileUtils {
void AssetFilePath(cpointer, uint64, cpointer, FilePathString& filepath) { filepath.Clear(); }
}
namespace File {
Error ReadWholeFile(cpointer, void**, uint64*) { return Error_("oracle harness: no file IO"); }
}
// TextureLib/TextureResource.cpp is linked for its Serialize(TextureResourceData&); its debug image dump is never called here
namespace Directory {
void EnsureDirectoryExists(const char*) {}
}
// CSampler with a scriptable stream (MathLib/Sampler.cpp is not linked).
void CSampler::Initialize(uint32 seed) {
    if (!g_mt) g_mt = new std::mt19937();
    g_mt->seed(seed);
}
void CSampler::Shutdown() {}
void CSampler::Reseed(uint32 seed) { Initialize(seed); }
float CSampler::UniformFloat() {
    DrawSource& s = g_src;
    if (s.usePcg) return pcgNext(s.pcg);
    if (s.script) {
        float v = s.pos < s.count ? s.script[s.pos] : 0.5f;
        ++s.pos;
        return v;
    }
    if (!g_mt) g_mt = new std::mt19937(0);
    return (*g_mt)() / (float)(0xFFFFFFFF);  // MathLib/Random.cpp:47-50
}
uint32 CSampler::UniformUInt32() {
    if (!g_mt) g_mt = new std::mt19937(0);
    return (uint32)(*g_mt)();
}
float3 CSampler::UniformSphere() {  // MathLib/Sampler.cpp:44-54, restated because that TU is not linked
    float r0 = UniformFloat();
    float r1 = UniformFloat();
    float u = 2.0f * r1 - 1.0f;
    float norm = Math::Sqrtf(Max(0.0f, 1.0f - u * u));
    float theta = Math::TwoPi_ * r0;
    return float3(norm * Math::Cosf(theta), u, norm * Math::Sinf(theta));
}
float CSampler::UniformSpherePdf() { return Math::Inv4Pi_; }
}  // namespace Selas

using namespace Selas;

namespace {
void copySample(const BsdfSample& b, bool ok, uint32_t used, SelasBsdfSample& o) {
    o.flags = b.flags;
    o.mediumPhase = (uint32_t)b.medium.phaseFunction;
    st3(o.mediumExtinction, b.medium.extinction);
    st3(o.reflectance, b.reflectance);
    st3(o.wi, b.wi);
    o.forwardPdfW = b.forwardPdfW;
    o.reversePdfW = b.reversePdfW;
    o.valid = ok ? 1u : 0u;
    o.randomsUsed = used;
}

}  // namespace

extern "C" {

int ref_sizeof(const char* name) {
    if (!strcmp(name, "SurfaceParameters")) return (int)sizeof(SurfaceParameters);
    if (!strcmp(name, "BsdfSample")) return (int)sizeof(BsdfSample);
    if (!strcmp(name, "SceneLight")) return (int)sizeof(SceneLight);
    if (!strcmp(name, "HitParameters")) return (int)sizeof(HitParameters);
    if (!strcmp(name, "RayCastCameraSettings")) return (int)sizeof(RayCastCameraSettings);
    if (!strcmp(name, "VCMVertex")) return (int)sizeof(VCMVertex);
    if (!strcmp(name, "MaterialResourceData")) return (int)sizeof(MaterialResourceData);
    return -1;
}

int ref_bsdf_eval(size_t n, const SelasSurface* s, const float* v, const float* l, float* rgb, float* fwd, float* rev) {
    setFtz();
    for (size_t i = 0; i < n; ++i) {
        SurfaceParameters sp;
        memcpy(&sp, &s[i], sizeof(sp));
        float f = 0.f, r = 0.f;
        float3 c = EvaluateBsdf(sp, f3(v + 3 * i), f3(l + 3 * i), f, r);
        st3(rgb + 3 * i, c);
        fwd[i] = f;
        rev[i] = r;
    }
    return 0;
}

int ref_bsdf_sample(size_t n, const SelasSurface* s, const float* v, const float* u, SelasBsdfSample* out) {
    setFtz();
    CSampler sampler;
    for (size_t i = 0; i < n; ++i) {
        SurfaceParameters sp;
        memcpy(&sp, &s[i], sizeof(sp));
        g_src = DrawSource{u + SELAS_BSDF_RANDOMS * i, 0, SELAS_BSDF_RANDOMS, 0, false};
        BsdfSample bs;
        bs.flags = eScatterEvent;  // no default in Scattering.h:36; see SURVEY.md appendix B #4
        bool ok = SampleBsdfFunction(&sampler, sp, f3(v + 3 * i), bs);
        copySample(bs, ok, g_src.pos, out[i]);
    }
    g_src = DrawSource{nullptr, 0, 0, 0, false};
    return 0;
}

int ref_light_sample(size_t n, const SelasLight* lights, uint32_t count, const float* pos, const float* nrm, const float* u3,
                     SelasLightSample* out) {
    setFtz();
    FakeContext fc;
    SelasLightSetRange range = {0, count};
    fc.init(lights, count, &range, 1, nullptr);
    for (size_t i = 0; i < n; ++i) {
        // clamp the p0 == 1.0 overrun the same way the restatement does (appendix B #20)
        float uu[3] = {u3[3 * i], u3[3 * i + 1], u3[3 * i + 2]};
        if (count && (uint32_t)(uu[0] * count) >= count) uu[0] = nextafterf(1.0f, 0.0f);
        g_src = DrawSource{uu, 0, 3, 0, false};
        LightDirectSample ls;
        memset(&ls, 0, sizeof(ls));
        NextEventEstimation(fc.ctx(), 0, f3(pos + 3 * i), f3(nrm + 3 * i), ls);
        memcpy(&out[i], &ls, sizeof(ls));
    }
    g_src = DrawSource{nullptr, 0, 0, 0, false};
    return 0;
}

int ref_light_pdf(size_t n, const SelasLight* lights, uint32_t count, const uint32_t* index, const float* pos, const float* wi,
                  float* pdf) {
    setFtz();
    FakeContext fc;
    SelasLightSetRange range = {0, count};
    fc.init(lights, count, &range, 1, nullptr);
    for (size_t i = 0; i < n; ++i) {
        LightDirectSample ls;
        memset(&ls, 0, sizeof(ls));
        ls.index = index[i];
        pdf[i] = LightingPdf(fc.ctx(), 0, ls, f3(pos + 3 * i), f3(wi + 3 * i));
    }
    return 0;
}

int ref_camera_init(const SelasCameraSettings* st, uint32_t w, uint32_t h, SelasCamera* out) {
    CameraSettings cs;
    memset(&cs, 0, sizeof(cs));
    cs.position = f3(st->position);
    cs.fovDegrees = st->fovDegrees;
    cs.lookAt = f3(st->lookAt);
    cs.znear = st->znear;
    cs.up = f3(st->up);
    cs.zfar = st->zfar;
    RayCastCameraSettings rc;
    memset(&rc, 0, sizeof(rc));
    InitializeRayCastCamera(cs, w, h, rc);
    memcpy(out, &rc, sizeof(rc));
    return 0;
}

int ref_camera_rays(const SelasCamera* cam, size_t n, const float* px, const float* py, const float* jitter, SelasRay* rays) {
    RayCastCameraSettings rc;
    memcpy(&rc, cam, sizeof(rc));
    CSampler sampler;
    for (size_t i = 0; i < n; ++i) {
        float jj[2] = {jitter ? jitter[2 * i] : 0.5f, jitter ? jitter[2 * i + 1] : 0.5f};
        g_src = DrawSource{jj, 0, 2, 0, false};
        Ray r = JitteredCameraRay(&rc, &sampler, px[i], py[i]);
        rays[i].ox = r.origin.x; rays[i].oy = r.origin.y; rays[i].oz = r.origin.z;
        rays[i].tnear = 0.0f;
        rays[i].dx = r.direction.x; rays[i].dy = r.direction.y; rays[i].dz = r.direction.z;
        rays[i].tfar = FloatMax_;
    }
    g_src = DrawSource{nullptr, 0, 0, 0, false};
    return 0;
}

int ref_cmj(size_t n, const int32_t* s, int32_t m, int32_t nn, const int32_t* p, float* out2) {
    for (size_t i = 0; i < n; ++i) {
        float2 r = CorrelatedMultiJitter(s[i], m, nn, p[i]);
        out2[2 * i] = r.x;
        out2[2 * i + 1] = r.y;
    }
    return 0;
}

int ref_medium(size_t n, const float* extinction3, const float* u2, const float* dist, float* sampled, float* trans3) {
    CSampler sampler;
    for (size_t i = 0; i < n; ++i) {
        MediumParameters m;
        m.phaseFunction = eIsotropic;
        m.extinction = f3(extinction3 + 3 * i);
        g_src = DrawSource{u2 + 2 * i, 0, 2, 0, false};
        float pdf;
        sampled[i] = SampleDistance(&sampler, m, &pdf);
        st3(trans3 + 3 * i, Transmission(m, dist[i]));
    }
    g_src = DrawSource{nullptr, 0, 0, 0, false};
    return 0;
}

// BuildHashGrid + readback of its arrays (VCMHashGrid.cpp:48-97)
int ref_hashgrid_build(size_t n, const float* pos, float radius, uint32_t cellCount, uint32_t* cellRangeEnds,
                       uint32_t* cellIndices, float* aabb6) {
    CArray<VCMVertex> verts;
    verts.Resize(n);
    for (size_t i = 0; i < n; ++i) {
        memset(&verts[i], 0, sizeof(VCMVertex));
        verts[i].hit.position = f3(pos + 3 * i);
    }
    VCMHashGrid grid;
    BuildHashGrid(&grid, cellCount, radius, verts);
    memcpy(cellRangeEnds, grid.cellRangeEnds.DataPointer(), sizeof(uint32_t) * cellCount);
    memcpy(cellIndices, grid.cellIndices.DataPointer(), sizeof(uint32_t) * n);
    st3(aabb6, grid.aaBox.min);
    st3(aabb6 + 3, grid.aaBox.max);
    ShutdownHashGrid(&grid);
    verts.Shutdown();
    return 0;
}

struct QueryCollect {
    const VCMVertex* base;
    uint32_t* indices;
    uint32_t count, maxCount;
};
static void collectCallback(const VCMVertex& v, void* user) {
    QueryCollect* q = (QueryCollect*)user;
    if (q->indices && q->count < q->maxCount) q->indices[q->count] = (uint32_t)(&v - q->base);
    ++q->count;
}
// BuildHashGrid then SearchHashGrid for every query (VCMHashGrid.cpp:107-155); returns build and query seconds
int ref_hashgrid_query(size_t n, const float* pos, float radius, uint32_t cellCount, size_t m, const float* q, uint32_t* counts,
                       uint32_t* indices, uint32_t maxPerQuery) {
    CArray<VCMVertex> verts;
    verts.Resize(n);
    for (size_t i = 0; i < n; ++i) {
        memset(&verts[i], 0, sizeof(VCMVertex));
        verts[i].hit.position = f3(pos + 3 * i);
    }
    VCMHashGrid grid;
    BuildHashGrid(&grid, cellCount, radius, verts);
    for (size_t i = 0; i < m; ++i) {
        QueryCollect qc = {verts.DataPointer(), indices ? indices + i * maxPerQuery : nullptr, 0, maxPerQuery};
        SearchHashGrid(&grid, verts, f3(q + 3 * i), &qc, collectCallback);
        counts[i] = qc.count;
    }
    ShutdownHashGrid(&grid);
    verts.Shutdown();
    return 0;
}

int ref_vcm_iteration_constants(uint32_t vmCount, uint32_t vcCount, float baseRadius, float alpha, float iteration, float* out5) {
    VCMIterationConstants c = VCMCommon::CalculateIterationConstants(vmCount, vcCount, baseRadius, alpha, iteration);
    out5[0] = c.vmSearchRadius;
    out5[1] = c.vmSearchRadiusSqr;
    out5[2] = c.vmNormalization;
    out5[3] = c.vmWeight;
    out5[4] = c.vcWeight;
    return 0;
}

// ImageBasedLightResource.cpp: SampleIbl / SampleIblMiss / Ibl on a scene created with an IBL (ref_scene_create)
int ref_ibl_lookup(void* s, size_t n, const float* wi3, float* rgb3, float* pdf, float* miss3);
int ref_ibl_sample(void* s, size_t n, const float* u2, float* thetaPhiPdf3, uint32_t* xy2);

// ---- tier-B CPU integrator: reference shading + restated ray cast --------------------------------------------
int ref_scene_create(const SelasSceneDesc* d, void** out) {
    if (d->instanceCount || d->curveCount) return 5;  // the tier-B integrator handles identity single-instance triangle / quad scenes only
    for (uint32_t m = 0; m < d->meshCount; ++m)
        if (d->meshes[m].baseColorTexture) return 5;  // refSurface has no texture fetch (the lookup itself is pinned by ref_texture_sample)
    RefScene* sc = new RefScene();
    sc->meshes.resize(d->meshCount);
    for (uint32_t m = 0; m < d->meshCount; ++m) {
        const SelasMeshDesc& md = d->meshes[m];
        RefScene::Mesh& M = sc->meshes[m];
        M.positions.assign(md.positions, md.positions + 3 * (size_t)md.vertexCount);
        if (md.normals) M.normals.assign(md.normals, md.normals + 3 * (size_t)md.vertexCount);
        if (md.tangents) M.tangents.assign(md.tangents, md.tangents + 4 * (size_t)md.vertexCount);
        if (md.uvs) M.uvs.assign(md.uvs, md.uvs + 2 * (size_t)md.vertexCount);
        M.indices.assign(md.indices, md.indices + (size_t)md.faceCount * md.indicesPerFace);
        M.indicesPerFace = md.indicesPerFace;
        M.materialIndex = md.materialIndex;
        M.lightSetIndex = md.lightSetIndex;
        for (uint32_t f = 0; f < md.faceCount; ++f) {
            const uint32_t* ix = &M.indices[(size_t)f * M.indicesPerFace];
            orc::Tri T;
            T.primID = f; T.geomID = m; T.instID = 0; T.second = 0;
            auto P = [&](uint32_t vi, float* o) { memcpy(o, &M.positions[3 * (size_t)vi], 12); };
            if (M.indicesPerFace == 3) {
                P(ix[0], T.v0); P(ix[1], T.v1); P(ix[2], T.v2);
                sc->bvh.tris.push_back(T);
            } else {
                P(ix[0], T.v0); P(ix[1], T.v1); P(ix[3], T.v2);
                sc->bvh.tris.push_back(T);
                T.second = 1;
                P(ix[2], T.v0); P(ix[3], T.v1); P(ix[1], T.v2);
                sc->bvh.tris.push_back(T);
            }
        }
    }
    sc->materials.assign(d->materials, d->materials + d->materialCount);
    sc->lights.assign(d->lights, d->lights + d->lightCount);
    sc->sets.assign(d->lightSets, d->lightSets + d->lightSetCount);
    memcpy(sc->background, d->backgroundIntensity, sizeof(sc->background));
    sc->ibl.set(d->ibl);
    sc->bvh.build();
    *out = sc;
    return 0;
}
int ref_scene_destroy(void* s) {
    delete (RefScene*)s;
    return 0;
}

// CalculateSurfaceParams (SurfaceParameters.cpp:144-247) with the rtcInterpolate0 gathers and the
// material/cache lookups replaced by direct array reads; everything downstream is reference code.
static void refSurface(const RefScene& sc, const Ray& ray, float t, float u, float v, uint32_t geomID, uint32_t primID,
                       SurfaceParameters& surface) {
    const RefScene::Mesh& M = sc.meshes[geomID];
    const SelasMaterial& mat = sc.materials[M.materialIndex];
    const uint32_t* ix = &M.indices[(size_t)primID * M.indicesPerFace];
    auto interp = [&](const std::vector<float>& a, int comps, float* o) {
        for (int c = 0; c < comps; ++c) {
            if (M.indicesPerFace == 3)
                o[c] = (1.0f - u - v) * a[ix[0] * comps + c] + u * a[ix[1] * comps + c] + v * a[ix[2] * comps + c];
            else
                o[c] = (1.0f - v) * ((1.0f - u) * a[ix[0] * comps + c] + u * a[ix[1] * comps + c]) +
                       v * ((1.0f - u) * a[ix[3] * comps + c] + u * a[ix[2] * comps + c]);
        }
    };
    float3 normal;
    if (!M.normals.empty()) {
        float nn[3];
        interp(M.normals, 3, nn);
        normal = f3(nn);
    } else {
        const float* P = M.positions.data();
        uint32_t i0, i1, i2;
        if (M.indicesPerFace == 3) { i0 = ix[0]; i1 = ix[1]; i2 = ix[2]; }
        else if (u + v <= 1.0f) { i0 = ix[0]; i1 = ix[1]; i2 = ix[3]; }
        else { i0 = ix[2]; i1 = ix[3]; i2 = ix[1]; }
        normal = Cross(f3(P + 3 * i1) - f3(P + 3 * i0), f3(P + 3 * i2) - f3(P + 3 * i0));
    }
    float3 n = Normalize(normal);
    float3 tt, b;
    if (!M.tangents.empty()) {
        float tg[4];
        interp(M.tangents, 4, tg);
        tt = f3(tg);
        b = Cross(n, tt) * tg[3];
    } else {
        MakeOrthogonalCoordinateSystem(n, &tt, &b);
    }
    surface.baseColor = Pow(f3(mat.baseColor), 2.2f);
    float3x3 tangentToWorld = MakeFloat3x3(tt, n, b);
    surface.worldToTangent = MatrixTranspose(tangentToWorld);
    float3 pos = float3(ray.origin.x + t * ray.direction.x, ray.origin.y + t * ray.direction.y, ray.origin.z + t * ray.direction.z);
    surface.position = pos;
    const float kErr = 32.0f * 1.19209e-07f;
    surface.error = kErr * Max(Max(Math::Absf(pos.x), Math::Absf(pos.y)), Max(Math::Absf(pos.z), t));
    surface.materialFlags = mat.flags;
    surface.transmittanceColor = f3(mat.transmittanceColor);
    surface.sheen = mat.scalars[eSheen];
    surface.sheenTint = mat.scalars[eSheenTint];
    surface.clearcoat = mat.scalars[eClearcoat];
    surface.clearcoatGloss = mat.scalars[eClearcoatGloss];
    surface.specTrans = Saturate(mat.scalars[eSpecTrans]);
    surface.diffTrans = mat.scalars[eDiffuseTrans] * 0.5f;
    surface.flatness = mat.scalars[eFlatness];
    surface.anisotropic = mat.scalars[eAnisotropic];
    surface.specularTint = mat.scalars[eSpecularTint];
    surface.roughness = mat.scalars[eRoughness];
    surface.metallic = Saturate(mat.scalars[eMetallic]);
    surface.scatterDistance = mat.scalars[eScatterDistance];
    surface.ior = mat.scalars[eIor];
    surface.lightSetIndex = M.lightSetIndex;
    surface.shader = (ShaderType)mat.shader;
    surface.view = -ray.direction;
    surface.relativeIOR = ((mat.flags & eTransparent) && Dot(surface.view, n) < 0.0f) ? surface.ior : 1.0f / surface.ior;
}

int ref_render_pathtracer(void* s, const SelasCamera* cam, const SelasRenderSettings* rs, float* const* film,
                          SelasRenderStats* stats, int nthreads) {
    const RefScene& sc = *(RefScene*)s;
    size_t pixels = (size_t)rs->width * rs->height;
    for (uint32_t l = 0; l < rs->layerCount; ++l) memset(film[l], 0, pixels * 3 * sizeof(float));
    RayCastCameraSettings rc;
    memcpy(&rc, cam, sizeof(rc));
    std::atomic<size_t> next(0);
    std::atomic<uint64_t> closest(0), shadow(0);
    const size_t stride = rs->pixelStride > 1 ? rs->pixelStride : 1;
    auto worker = [&]() {
        setFtz();
        FakeContext fc;
        fc.init(sc.lights.data(), (uint32_t)sc.lights.size(), sc.sets.data(), (uint32_t)sc.sets.size(), sc.background);
        if (sc.ibl.present) fc.setIbl(&sc.ibl.data);
        GIIntegratorContext* ctx = fc.ctx();
        ctx->camera = &rc;
        ctx->maxPathLength = rs->maxPathLength;
        uint64_t c = 0, sh = 0;
        for (;;) {
            size_t b = next.fetch_add(1024);
            if (b >= pixels) break;
            size_t e = std::min(pixels, b + 1024);
            for (size_t p = b; p < e; ++p) {
                if (stride > 1 && (p % stride) != 0) continue;  // bounded sample (SelasRenderSettings.pixelStride)
                uint32_t y = (uint32_t)(p / rs->width), x = (uint32_t)(p - (size_t)y * rs->width);
                for (uint32_t k = 0; k < rs->samplesPerPixel; ++k) {
                    g_src = DrawSource{nullptr, 0, 0, mix32(rs->seed ^ mix32((uint32_t)p ^ mix32(rs->firstSample + k + 0x9e3779b9u))), true};
                    Ray ray = rs->pixelSampler == 0
                                  ? JitteredCameraRay(&rc, &ctx->sampler, (float)x, (float)y)
                                  : JitteredCameraRay(&rc, (int32)x, (int32)y, (int32)(rs->firstSample + k), (int32)(rs->pixelSampler & 0xffffu),
                                                      (int32)(rs->pixelSampler >> 16), (int32)p);
                    if (rs->integrator == SELAS_INTEGRATOR_DEFERRED) {
                        // ---- DeferredPathTracer.cpp:61-148, :151-290 walked path by path, with the reference's own NextEventEstimation,
                        // SampleBackground, BackgroundLightingPdf, BalanceHeuristic, EvaluateBackground[Miss], EvaluateBsdf, SampleBsdfFunction
                        float3 L = float3::Zero_;
                        float3 throughput = float3::One_;
                        bool diracScatterOnly = true;
                        uint32_t trackedBounces = 0, bounce = 0;
                        auto occludedFrom = [&](const SurfaceParameters& surface, float3 dir, float distance) {
                            float3 org = OffsetRayOrigin(surface, dir, 0.1f);
                            float so[3] = {org.x, org.y, org.z}, sd[3] = {dir.x, dir.y, dir.z};
                            orc::RayHit shh;
                            ++sh;
                            return sc.bvh.trace<float>(so, sd, 0.0f, distance, true, shh);
                        };
                        while (bounce < ctx->maxPathLength) {
                            orc::RayHit rh;
                            float o[3] = {ray.origin.x, ray.origin.y, ray.origin.z}, d[3] = {ray.direction.x, ray.direction.y, ray.direction.z};
                            bool hit = sc.bvh.trace<float>(o, d, 0.0f, FloatMax_, false, rh);
                            ++c;
                            if (!hit) {
                                float3 smp = diracScatterOnly ? EvaluateBackgroundMiss(ctx, ray.direction) : EvaluateBackground(ctx, ray.direction);
                                L += smp * throughput;
                                break;
                            }
                            SurfaceParameters surface;
                            const orc::Tri& T = sc.bvh.tris[rh.tri];
                            float hu = T.second ? (float)(1.0 - rh.u) : (float)rh.u, hv = T.second ? (float)(1.0 - rh.v) : (float)rh.v;
                            refSurface(sc, ray, (float)rh.t, hu, hv, T.geomID, T.primID, surface);
                            float3 view = -ray.direction;
                            LightDirectSample lightSample;
                            memset(&lightSample, 0, sizeof(lightSample));
                            NextEventEstimation(ctx, surface.lightSetIndex, surface.position, GeometricNormal(surface), lightSample);
                            if (Dot(lightSample.radiance, float3::One_) > 0) {
                                float fw = 0, rv = 0;
                                float3 reflectance = EvaluateBsdf(surface, view, lightSample.direction, fw, rv);
                                float weight = 1.0f;
                                float3 smp = weight * reflectance * lightSample.radiance * (1.0f / lightSample.pdfW);
                                if (Dot(smp, float3::One_) > 0 && !occludedFrom(surface, lightSample.direction, lightSample.distance)) L += smp * throughput;
                            }
                            LightDirectSample skySample;
                            memset(&skySample, 0, sizeof(skySample));
                            SampleBackground(ctx, skySample);
                            if (Dot(skySample.radiance, float3::One_) > 0) {
                                float fw = 0, rv = 0;
                                float3 reflectance = EvaluateBsdf(surface, view, skySample.direction, fw, rv);
                                float misWeight = ImportanceSampling::BalanceHeuristic(1, skySample.pdfW, 1, fw);
                                float3 smp = misWeight * reflectance * skySample.radiance * (1.0f / skySample.pdfW);
                                if (Dot(smp, float3::One_) > 0 && !occludedFrom(surface, skySample.direction, skySample.distance)) L += smp * throughput;
                            }
                            BsdfSample bsdfSample;
                            bsdfSample.flags = eScatterEvent;
                            if (SampleBsdfFunction(&ctx->sampler, surface, view, bsdfSample) == false) break;
                            float skyPdfW = BackgroundLightingPdf(ctx, bsdfSample.wi);
                            float misWeight = ImportanceSampling::BalanceHeuristic(1, bsdfSample.forwardPdfW, 1, skyPdfW);
                            throughput = misWeight * throughput * bsdfSample.reflectance;
                            if (LengthSquared(throughput) == 0.0f) break;
                            if (trackedBounces >= MaxTrackedBounces_) {
                                float continuationProb = Max<float>(Max<float>(throughput.x, throughput.y), throughput.z);
                                if (ctx->sampler.UniformFloat() >= continuationProb) break;
                                throughput = throughput * (1.0f / continuationProb);
                            }
                            float3 offsetOrigin = OffsetRayOrigin(surface, bsdfSample.wi, 1.0f);
                            ray = MakeRay(offsetOrigin, bsdfSample.wi);
                            diracScatterOnly = diracScatterOnly && (bsdfSample.flags & SurfaceEventFlags::eDiracEvent);
                            trackedBounces = Min<uint32>(MaxTrackedBounces_, trackedBounces + 1);
                            ++bounce;
                        }
                        float* out = film[rs->layerCount == 2 ? 1 : 0];
                        out[3 * p] += L.x; out[3 * p + 1] += L.y; out[3 * p + 2] += L.z;
                        continue;
                    }
                    // ---- EvaluatePath, PathTracer.cpp:142-258 ----
                    float3 Ld[2] = {float3::Zero_, float3::Zero_};
                    float3 throughput = float3::One_;
                    MediumParameters vacuum;
                    MediumParameters currentMedium = vacuum;
                    bool isDeltaOnly = true;
                    uint32_t bounceCount = 0;
                    while (bounceCount < ctx->maxPathLength) {
                        float pdf;
                        float rayDistance = SampleDistance(&ctx->sampler, currentMedium, &pdf);
                        orc::RayHit rh;
                        float o[3] = {ray.origin.x, ray.origin.y, ray.origin.z}, d[3] = {ray.direction.x, ray.direction.y, ray.direction.z};
                        bool hit = sc.bvh.trace<float>(o, d, 0.0f, rayDistance, false, rh);
                        ++c;
                        SurfaceParameters surface;
                        if (hit) {
                            const orc::Tri& T = sc.bvh.tris[rh.tri];
                            float hu = T.second ? (float)(1.0 - rh.u) : (float)rh.u, hv = T.second ? (float)(1.0 - rh.v) : (float)rh.v;
                            refSurface(sc, ray, (float)rh.t, hu, hv, T.geomID, T.primID, surface);
                            rayDistance = Length(surface.position - ray.origin);
                        }
                        throughput = throughput * Transmission(currentMedium, rayDistance);
                        if (hit) {
                            LightDirectSample lightSample;
                            memset(&lightSample, 0, sizeof(lightSample));
                            NextEventEstimation(ctx, surface.lightSetIndex, surface.position, GeometricNormal(surface), lightSample);
                            if (Dot(lightSample.radiance, float3::One_) > 0) {
                                float fw = 0, rv = 0;
                                float3 reflectance = EvaluateBsdf(surface, -ray.direction, lightSample.direction, fw, rv);
                                if (Dot(reflectance, float3::One_) > 0) {
                                    float3 org = OffsetRayOrigin(surface, lightSample.direction, 0.1f);
                                    float so[3] = {org.x, org.y, org.z}, sd[3] = {lightSample.direction.x, lightSample.direction.y, lightSample.direction.z};
                                    orc::RayHit shh;
                                    ++sh;
                                    if (!sc.bvh.trace<float>(so, sd, surface.error, lightSample.distance, true, shh)) {
                                        float3 smp = reflectance * lightSample.radiance * (1.0f / lightSample.pdfW);
                                        Ld[1] += smp * throughput;
                                    }
                                }
                            }
                            BsdfSample bsdfSample;
                            bsdfSample.flags = eScatterEvent;
                            if (SampleBsdfFunction(&ctx->sampler, surface, -ray.direction, bsdfSample) == false) break;
                            isDeltaOnly = isDeltaOnly && (bsdfSample.flags & SurfaceEventFlags::eDiracEvent);
                            if (bsdfSample.flags == SurfaceEventFlags::eTransmissionEvent) {
                                if (currentMedium.phaseFunction == eVacuum) currentMedium = bsdfSample.medium;
                                else currentMedium = vacuum;
                            }
                            float weight = 1.0f;
                            if (rs->misMode == 1) {
                                float lightPdfW = LightingPdf(ctx, surface.lightSetIndex, lightSample, surface.position, bsdfSample.wi);
                                weight = bsdfSample.forwardPdfW / (bsdfSample.forwardPdfW + lightPdfW);
                            }
                            throughput = weight * throughput * bsdfSample.reflectance;
                            float3 offsetOrigin = OffsetRayOrigin(surface, bsdfSample.wi, 1.0f);
                            ray = MakeRay(offsetOrigin, bsdfSample.wi);
                            if (bounceCount == 0) Ld[0] = bsdfSample.reflectance;
                        } else if (currentMedium.phaseFunction != MediumPhaseFunction::eVacuum) {
                            float3 origin = ray.origin + rayDistance * ray.direction;
                            float mediumPdf;
                            float3 direction = SampleScatterDirection(&ctx->sampler, currentMedium, ray.direction, &mediumPdf);
                            ray = MakeRay(origin, direction);
                        } else {
                            float3 smp = isDeltaOnly ? EvaluateBackgroundMiss(ctx, ray.direction) : EvaluateBackground(ctx, ray.direction);
                            Ld[1] += smp * throughput;
                            break;
                        }
                        ++bounceCount;
                        if (bounceCount > rs->rrStartBounce) {
                            float continuationProb = Max<float>(Max<float>(throughput.x, throughput.y), throughput.z);
                            if (ctx->sampler.UniformFloat() > continuationProb) break;
                            throughput = throughput * (1.0f / continuationProb);
                        }
                    }
                    if (rs->layerCount == 1) {
                        film[0][3 * p] += Ld[1].x; film[0][3 * p + 1] += Ld[1].y; film[0][3 * p + 2] += Ld[1].z;
                    } else {
                        for (int l = 0; l < 2; ++l) { film[l][3 * p] += Ld[l].x; film[l][3 * p + 1] += Ld[l].y; film[l][3 * p + 2] += Ld[l].z; }
                    }
                }
            }
        }
        closest += c;
        shadow += sh;
        g_src = DrawSource{nullptr, 0, 0, 0, false};
    };
    std::vector<std::thread> th;
    for (int t = 0; t < std::max(1, nthreads); ++t) th.emplace_back(worker);
    for (auto& x : th) x.join();
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->closestRays = closest;
        stats->occlusionRays = shadow;
        stats->pathSamples = ((pixels + stride - 1) / stride) * rs->samplesPerPixel;
    }
    return 0;
}

int ref_ibl_lookup(void* s, size_t n, const float* wi3, float* rgb3, float* pdf, float* miss3) {
    setFtz();
    const RefScene& sc = *(RefScene*)s;
    if (!sc.ibl.present) return 2;
    for (size_t i = 0; i < n; ++i) {
        float p = 0.f;
        float3 c = SampleIbl(&sc.ibl.data, f3(wi3 + 3 * i), p);
        if (rgb3) st3(rgb3 + 3 * i, c);
        if (pdf) pdf[i] = p;
        if (miss3) st3(miss3 + 3 * i, SampleIblMiss(&sc.ibl.data, f3(wi3 + 3 * i), p));
    }
    return 0;
}
int ref_ibl_sample(void* s, size_t n, const float* u2, float* thetaPhiPdf3, uint32_t* xy2) {
    setFtz();
    const RefScene& sc = *(RefScene*)s;
    if (!sc.ibl.present) return 2;
    for (size_t i = 0; i < n; ++i) {
        float theta, phi, pdf;
        Selas::uint x, y;
        Ibl(&sc.ibl.data, u2[2 * i], u2[2 * i + 1], theta, phi, x, y, pdf);
        thetaPhiPdf3[3 * i] = theta; thetaPhiPdf3[3 * i + 1] = phi; thetaPhiPdf3[3 * i + 2] = pdf;
        xy2[2 * i] = (uint32_t)x; xy2[2 * i + 1] = (uint32_t)y;
    }
    return 0;
}

// ---- VCM building blocks that are still live code in the reference ---------------------------------------------------
// EmitIblLightSample / DirectIblLightSample / IblCalculateRadiance (Shading/AreaLighting.cpp:246-331) and
// VCMCommon::GenerateLightSample / GenerateCameraSample (VCMCommon.cpp:37-88), driven with a scripted CSampler.
namespace {
struct VcmContext {
    // SceneResource / ImageBasedLightResource have out-of-line constructors in translation units that are not linked
    // (they need Embree and the asset IO); raw zeroed storage of the right size stands in for them.
    alignas(16) unsigned char sceneStorage[sizeof(SceneResource)];
    alignas(16) unsigned char iblStorage[sizeof(ImageBasedLightResource)];
    alignas(16) unsigned char ctxStorage[sizeof(GIIntegratorContext)];
    GIIntegratorContext* ctx;
    VcmContext(const RefScene& sc, const float* sphere4, const RayCastCameraSettings* cam) {
        memset(sceneStorage, 0, sizeof(sceneStorage));
        memset(iblStorage, 0, sizeof(iblStorage));
        memset(ctxStorage, 0, sizeof(ctxStorage));
        SceneResource* scene = reinterpret_cast<SceneResource*>(sceneStorage);
        ImageBasedLightResource* ibl = reinterpret_cast<ImageBasedLightResource*>(iblStorage);
        ibl->data = const_cast<ImageBasedLightResourceData*>(&sc.ibl.data);
        scene->iblResource = ibl;
        if (sphere4) scene->boundingSphere = float4(sphere4[0], sphere4[1], sphere4[2], sphere4[3]);
        ctx = reinterpret_cast<GIIntegratorContext*>(ctxStorage);
        ctx->scene = scene;
        ctx->camera = cam;
    }
};
}  // namespace

int ref_ibl_emit(void* s, const float* sphere4, size_t n, const float* u4, float* posDirRad9, float* pdfs3) {
    setFtz();
    const RefScene& sc = *(RefScene*)s;
    if (!sc.ibl.present) return 2;
    VcmContext vc(sc, sphere4, nullptr);
    for (size_t i = 0; i < n; ++i) {
        g_src = DrawSource{u4 + 4 * i, 0, 4, 0, false};
        LightEmissionSample e;
        EmitIblLightSample(vc.ctx, e);
        st3(posDirRad9 + 9 * i, e.position);
        st3(posDirRad9 + 9 * i + 3, e.direction);
        st3(posDirRad9 + 9 * i + 6, e.radiance);
        pdfs3[3 * i] = e.emissionPdfW; pdfs3[3 * i + 1] = e.directionPdfA; pdfs3[3 * i + 2] = e.cosThetaLight;
    }
    g_src = DrawSource{nullptr, 0, 0, 0, false};
    return 0;
}
int ref_ibl_direct(void* s, size_t n, const float* u2, SelasLightSample* out) {
    setFtz();
    const RefScene& sc = *(RefScene*)s;
    if (!sc.ibl.present) return 2;
    VcmContext vc(sc, nullptr, nullptr);
    for (size_t i = 0; i < n; ++i) {
        g_src = DrawSource{u2 + 2 * i, 0, 2, 0, false};
        LightDirectSample d;
        d.index = 0;
        DirectIblLightSample(vc.ctx, d);
        out[i].index = 0;
        st3(out[i].direction, d.direction);
        st3(out[i].radiance, d.radiance);
        out[i].distance = d.distance;
        out[i].pdfW = d.pdfW;
    }
    g_src = DrawSource{nullptr, 0, 0, 0, false};
    return 0;
}
int ref_ibl_radiance(void* s, const float* sphere4, size_t n, const float* dir3, float* rgb3, float* pdfs2) {
    setFtz();
    const RefScene& sc = *(RefScene*)s;
    if (!sc.ibl.present) return 2;
    VcmContext vc(sc, sphere4, nullptr);
    for (size_t i = 0; i < n; ++i) {
        float a = 0.f, b = 0.f;
        st3(rgb3 + 3 * i, IblCalculateRadiance(vc.ctx, f3(dir3 + 3 * i), a, b));
        pdfs2[2 * i] = a; pdfs2[2 * i + 1] = b;
    }
    return 0;
}
// out13 per sample: position, direction, throughput, dVCM, dVC, dVM, (pathLength | isAreaMeasure << 8 as float)
static void storeState(const PathState& st, float* o) {
    st3(o, st.position); st3(o + 3, st.direction); st3(o + 6, st.throughput);
    o[9] = st.dVCM; o[10] = st.dVC; o[11] = st.dVM;
    o[12] = (float)(st.pathLength | (st.isAreaMeasure << 8));
}
int ref_vcm_generate_light(void* s, const float* sphere4, float vcWeight, size_t n, const float* u4, float* out13) {
    setFtz();
    const RefScene& sc = *(RefScene*)s;
    if (!sc.ibl.present) return 2;
    VcmContext vc(sc, sphere4, nullptr);
    for (size_t i = 0; i < n; ++i) {
        g_src = DrawSource{u4 + 4 * i, 0, 4, 0, false};
        PathState st;
        VCMCommon::GenerateLightSample(vc.ctx, vcWeight, (Selas::uint)i, st);
        storeState(st, out13 + 13 * i);
    }
    g_src = DrawSource{nullptr, 0, 0, 0, false};
    return 0;
}
int ref_vcm_generate_camera(const SelasCamera* cam, float lightPathCount, size_t n, const uint32_t* xy2, const float* u2, float* out13,
                            uint32_t* index) {
    setFtz();
    RefScene dummy;
    VcmContext vc(dummy, nullptr, reinterpret_cast<const RayCastCameraSettings*>(cam));
    for (size_t i = 0; i < n; ++i) {
        g_src = DrawSource{u2 + 2 * i, 0, 2, 0, false};
        PathState st;
        VCMCommon::GenerateCameraSample(vc.ctx, xy2[2 * i], xy2[2 * i + 1], lightPathCount, st);
        storeState(st, out13 + 13 * i);
        index[i] = st.index;
    }
    g_src = DrawSource{nullptr, 0, 0, 0, false};
    return 0;
}

// TextureFiltering::Point / Triangle / Trilinear / EWA (TextureLib/TextureFiltering.h:54-234), the reference's own templates over a
// TextureResourceData built from the descriptor.  EWAFilterLut is a `static` array in that header, i.e. one copy per translation unit,
// and InitializeEWAFilterWeights (TextureFiltering.cpp:14-21) only fills the copy of its own unit: any other caller of EWA reads zeros
// and gets 0 / 0.  This unit therefore fills ITS copy with the same formula (Math::Expf) before the first lookup.
}  // extern "C"
template <typename T>
static void refFilterOne(TextureResourceData* data, int filter, int level, float2 st, float2 d0, float2 d1, float* out, int comps) {
    T r = T(0.0f);
    switch (filter) {
        case SELAS_FILTER_POINT: TextureFiltering::Point<T>(data, st, r); break;
        case SELAS_FILTER_TRIANGLE: TextureFiltering::Triangle<T>(data, level, st, r); break;
        case SELAS_FILTER_TRILINEAR: TextureFiltering::Trilinear<T>(data, st, d0, d1, r); break;
        case SELAS_FILTER_EWA: TextureFiltering::EWA<T>(data, st, d0, d1, r); break;
        default: TextureFiltering::EWA<T>(data, (int32)level, st, d0, d1, r); break;
    }
    memcpy(out, &r, sizeof(float) * comps);
}
extern "C" {
int ref_texture_filter(const SelasTextureMipsDesc* t, int filter, int level, size_t n, const float* st2, const float* dst0, const float* dst1, float* out) {
    setFtz();
    static bool lutReady = false;
    if (!lutReady) {
        for (uint32_t i = 0; i < (uint32_t)EwaLutSize; ++i) {
            float alpha = 2.0f;
            float r2 = float(i) / (EwaLutSize - 1.0f);
            EWAFilterLut[i] = Math::Expf(-alpha * r2) - Math::Expf(-alpha);
        }
        lutReady = true;
    }
    TextureResourceData data;
    memset(&data, 0, sizeof(data));
    data.mipCount = t->mipCount;
    data.dataSize = (uint32)t->dataSize;
    for (uint32_t l = 0; l < t->mipCount; ++l) {
        data.mipWidths[l] = t->mipWidths[l];
        data.mipHeights[l] = t->mipHeights[l];
        data.mipOffsets[l] = t->mipOffsets[l];
    }
    data.format = (TextureResourceData::TextureDataType)t->format;
    data.texture = reinterpret_cast<uint8*>(const_cast<void*>(t->texels));
    const int comps = (int)t->format + 1;
    for (size_t i = 0; i < n; ++i) {
        float2 st(st2[2 * i], st2[2 * i + 1]);
        float2 d0 = dst0 ? float2(dst0[2 * i], dst0[2 * i + 1]) : float2(0.0f, 0.0f);
        float2 d1 = dst1 ? float2(dst1[2 * i], dst1[2 * i + 1]) : float2(0.0f, 0.0f);
        float* o = out + i * comps;
        switch (comps) {
            case 1: refFilterOne<float>(&data, filter, level, st, d0, d1, o, 1); break;
            case 2: refFilterOne<float2>(&data, filter, level, st, d0, d1, o, 2); break;
            case 3: refFilterOne<float3>(&data, filter, level, st, d0, d1, o, 3); break;
            default: refFilterOne<float4>(&data, filter, level, st, d0, d1, o, 4); break;
        }
    }
    return 0;
}

// SampleTextureFloat3(texture, uvs, sRGB, default) of Shading/SurfaceParameters.cpp:97-118 (static there, and that translation unit
// is not linked): the reference's own TextureFiltering::Triangle template (TextureLib/TextureFiltering.h:77-104) and
// Math::SrgbToLinearPrecise (MathLib/ColorSpace.cpp) under the same format dispatch.
int ref_texture_sample(const SelasTextureDesc* t, size_t n, const float* uv2, int sRGB, float* rgb3) {
    setFtz();
    TextureResourceData data;
    memset(&data, 0, sizeof(data));
    data.mipCount = 1;
    data.mipWidths[0] = t->width;
    data.mipHeights[0] = t->height;
    data.mipOffsets[0] = 0;
    data.format = (TextureResourceData::TextureDataType)t->format;
    data.texture = reinterpret_cast<uint8*>(const_cast<float*>(t->texels));
    for (size_t i = 0; i < n; ++i) {
        float2 uvs(uv2[2 * i], uv2[2 * i + 1]);
        float3 out = float3::Zero_;
        if (data.format == TextureResourceData::Float) {
            float v;
            TextureFiltering::Triangle(&data, 0, uvs, v);
            if (sRGB) v = Math::SrgbToLinearPrecise(v);
            out = float3(v, v, v);
        } else if (data.format == TextureResourceData::Float3) {
            float3 v;
            TextureFiltering::Triangle(&data, 0, uvs, v);
            if (sRGB) v = Math::SrgbToLinearPrecise(v);
            out = v;
        } else if (data.format == TextureResourceData::Float4) {
            float4 v;
            TextureFiltering::Triangle(&data, 0, uvs, v);
            if (sRGB) v = Math::SrgbToLinearPrecise(v);
            out = v.XYZ();
        }
        st3(rgb3 + 3 * i, out);
    }
    return 0;
}

// ---- baked blobs written by the reference's own serializer (IoLib/BinaryStreamSerializer.h: SerializeToBinary) -----------------
// The product's host library reads these files (selas_b200/host: ReadImageBasedLightResource / ReadTextureResource); the baker itself
// (BuildCore) is not buildable here, so the blobs are produced from in-memory resource data by the same Serialize() functions it calls.
static int writeFile(const char* path, const uint8* data, uint64 size) {
    FILE* f = fopen(path, "wb");
    if (!f) return 1;
    size_t n = fwrite(data, 1, size, f);
    fclose(f);
    return n == size ? 0 : 1;
}
int ref_write_ibl_blob(const SelasIblDesc* I, const char* path) {
    ImageBasedLightResourceData data;
    memset(&data, 0, sizeof(data));
    data.densityfunctions.width = I->width;
    data.densityfunctions.height = I->height;
    data.densityfunctions.marginalDensityFunction = const_cast<float*>(I->marginalDensityFunction);
    data.densityfunctions.conditionalDensityFunctions = const_cast<float*>(I->conditionalDensityFunctions);
    data.missWidth = I->missWidth;
    data.missHeight = I->missHeight;
    data.rotationRadians = I->rotationRadians;
    data.exposureScale = I->exposureScale;
    data.lightData = reinterpret_cast<float3*>(const_cast<float*>(I->lightData));
    data.missData = reinterpret_cast<float3*>(const_cast<float*>(I->missData));
    uint8* blob = nullptr;
    Selas::uint blobSize = 0;
    SerializeToBinary(data, blob, blobSize);
    int rc = writeFile(path, blob, blobSize);
    SelasAlignedFree(blob);
    return rc;
}
int ref_write_texture_blob(const SelasTextureDesc* t, const char* path) {
    TextureResourceData data;
    memset(&data, 0, sizeof(data));
    data.mipCount = 1;
    data.dataSize = (uint32)(sizeof(float) * (t->format + 1) * t->width * t->height);
    data.mipWidths[0] = t->width;
    data.mipHeights[0] = t->height;
    data.mipOffsets[0] = 0;
    data.format = (TextureResourceData::TextureDataType)t->format;
    data.texture = reinterpret_cast<uint8*>(const_cast<float*>(t->texels));
    uint8* blob = nullptr;
    Selas::uint blobSize = 0;
    SerializeToBinary(data, blob, blobSize);
    int rc = writeFile(path, blob, blobSize);
    SelasAlignedFree(blob);
    return rc;
}
// FrameBuffer_Save's writer: StbImageWrite(..., HDR, ...) = stbi_write_hdr (TextureLib/StbImage.cpp:64-87 + Middleware/stb, unmodified)
int ref_write_hdr(const char* path, uint32_t width, uint32_t height, const float* rgb) {
    Error e = StbImageWrite(path, width, height, 3, HDR, const_cast<float*>(rgb));
    return e.Failed() ? 1 : 0;
}
unsigned long long ref_asset_version(const char* type) {
    if (!strcmp(type, "IBL")) return ImageBasedLightResource::kDataVersion;
    if (!strcmp(type, "Textures")) return TextureResource::kDataVersion;
    if (!strcmp(type, "SceneResource")) return SceneResource::kDataVersion;
    if (!strcmp(type, "SubsceneResource")) return SubsceneResource::kDataVersion;
    if (!strcmp(type, "ModelResource") || !strcmp(type, "ModelGeometryResource")) return ModelResource::kDataVersion;
    return 0;
}

// ---- baked scene / subscene / model blobs -----------------------------------------------------------------------------------
// Inputs are raw arrays in the reference's own element layouts (MeshMetaData 92 B, MaterialResourceData 600 B, Instance 72 B,
// CameraSettings 176 B, SceneLight 64 B, SceneLightSetRange 8 B, FilePathString 512 B); they are copied into the reference's CArrays and
// written with SerializeToBinary through the reference's Serialize(SceneResourceData&) / (SubsceneResourceData&) / (ModelResourceData&)
// / (ModelGeometryData&) (SceneLib/*.cpp, linked unmodified).
}  // extern "C"
template <typename T>
static void fillArray(CArray<T>& a, const void* src, uint64 count) {
    a.Resize(count);
    if (count) memcpy(a.DataPointer(), src, count * sizeof(T));
}
template <int N>
static void copyFixed(FixedString<N>& dst, const char* src) {  // FixedString::Copy goes through StringUtil, which is not built here
    memset(dst.str, 0, N);
    strncpy(dst.str, src, N - 1);
}
template <typename T>
static int writeBlob(T& data, const char* path) {
    uint8* blob = nullptr;
    Selas::uint blobSize = 0;
    SerializeToBinary(data, blob, blobSize);
    int rc = writeFile(path, blob, blobSize);
    SelasAlignedFree(blob);
    return rc;
}
static_assert(sizeof(MeshMetaData) == 92 && sizeof(MaterialResourceData) == 600 && sizeof(Instance) == 72 && sizeof(CameraSettings) == 176,
              "reference element layouts");
extern "C" {

// Curve geometry of the NEXT ref_write_model_blobs call (ModelResourceData::curves / curveModelName / totalCurveVertexCount and
// ModelGeometryData::curveIndices / curveVertices): borrowed pointers, consumed and cleared by that call.
static struct {
    const uint32_t* indices = nullptr;
    uint64_t indexCount = 0;
    const float* vertices4 = nullptr;
    uint32_t vertexCount = 0;
    const uint32_t* ranges2 = nullptr;  // CurveMetaData {indexOffset, indexCount}
    uint32_t curveCount = 0;
    uint32_t modelName = 0;
} g_nextCurves;
int ref_set_next_model_curves(const uint32_t* indices, uint64_t indexCount, const float* vertices4, uint32_t vertexCount, const uint32_t* ranges2,
                              uint32_t curveCount, uint32_t curveModelName) {
    g_nextCurves.indices = indices; g_nextCurves.indexCount = indexCount; g_nextCurves.vertices4 = vertices4; g_nextCurves.vertexCount = vertexCount;
    g_nextCurves.ranges2 = ranges2; g_nextCurves.curveCount = curveCount; g_nextCurves.modelName = curveModelName;
    return 0;
}

int ref_write_model_blobs(const char* dataPath, const char* geomPath, const float* aabb6, uint32_t totalVertexCount, const uint32_t* indices,
                          uint64_t indexCount, const float* positions, const float* normals, const float* tangents, const float* uvs,
                          const void* meshes92, uint32_t meshCount, const void* materials600, const uint32_t* materialHashes,
                          uint32_t materialCount) {
    ModelResourceData* data = new ModelResourceData();
    memset(&data->aaBox, 0, sizeof(data->aaBox));
    data->aaBox.min = float3(aabb6[0], aabb6[1], aabb6[2]);
    data->aaBox.max = float3(aabb6[3], aabb6[4], aabb6[5]);
    data->totalVertexCount = totalVertexCount;
    data->totalCurveVertexCount = g_nextCurves.vertexCount;
    data->curveModelName = g_nextCurves.modelName;
    data->pad = 0;
    data->indexSize = indexCount * sizeof(uint32);
    data->faceIndexSize = 0;
    data->positionSize = (uint64)totalVertexCount * sizeof(float3);
    data->normalsSize = normals ? (uint64)totalVertexCount * sizeof(float3) : 0;
    data->tangentsSize = tangents ? (uint64)totalVertexCount * sizeof(float4) : 0;
    data->uvsSize = uvs ? (uint64)totalVertexCount * sizeof(float2) : 0;
    data->curveIndexSize = g_nextCurves.indexCount * sizeof(uint32);
    data->curveVertexSize = (uint64)g_nextCurves.vertexCount * sizeof(float4);
    static_assert(sizeof(CurveMetaData) == 8, "CurveMetaData layout");
    fillArray(data->curves, g_nextCurves.ranges2, g_nextCurves.curveCount);
    fillArray(data->materials, materials600, materialCount);
    fillArray(data->materialHashes, materialHashes, materialCount);
    fillArray(data->meshes, meshes92, meshCount);
    int rc = writeBlob(*data, dataPath);
    data->materials.Shutdown(); data->materialHashes.Shutdown(); data->meshes.Shutdown(); data->curves.Shutdown();
    delete data;
    ModelGeometryData geom;
    memset(&geom, 0, sizeof(geom));
    geom.indexSize = indexCount * sizeof(uint32);
    geom.positionSize = (uint64)totalVertexCount * sizeof(float3);
    geom.normalsSize = normals ? (uint64)totalVertexCount * sizeof(float3) : 0;
    geom.tangentsSize = tangents ? (uint64)totalVertexCount * sizeof(float4) : 0;
    geom.uvsSize = uvs ? (uint64)totalVertexCount * sizeof(float2) : 0;
    geom.indices = const_cast<uint32*>(indices);
    geom.positions = reinterpret_cast<float3*>(const_cast<float*>(positions));
    geom.normals = reinterpret_cast<float3*>(const_cast<float*>(normals));
    geom.tangents = reinterpret_cast<float4*>(const_cast<float*>(tangents));
    geom.uvs = reinterpret_cast<float2*>(const_cast<float*>(uvs));
    geom.curveIndexSize = g_nextCurves.indexCount * sizeof(uint32);
    geom.curveVertexSize = (uint64)g_nextCurves.vertexCount * sizeof(float4);
    geom.curveIndices = const_cast<uint32*>(g_nextCurves.indices);
    geom.curveVertices = reinterpret_cast<float4*>(const_cast<float*>(g_nextCurves.vertices4));
    if (!rc) rc = writeBlob(geom, geomPath);
    g_nextCurves = {};
    return rc;
}
int ref_write_subscene_blob(const char* path, const char* name, uint64_t lightSetIndex, const void* modelNames512, uint32_t modelCount,
                            const void* instances72, uint32_t instanceCount, const uint32_t* sceneMaterialHashes, const void* sceneMaterials600,
                            uint32_t sceneMaterialCount) {
    SubsceneResourceData* data = new SubsceneResourceData();
    copyFixed(data->name, name);
    data->lightSetIndex = lightSetIndex;
    fillArray(data->modelNames, modelNames512, modelCount);
    fillArray(data->modelInstances, instances72, instanceCount);
    fillArray(data->sceneMaterialNames, sceneMaterialHashes, sceneMaterialCount);
    fillArray(data->sceneMaterials, sceneMaterials600, sceneMaterialCount);
    int rc = writeBlob(*data, path);
    data->modelNames.Shutdown(); data->modelInstances.Shutdown(); data->sceneMaterialNames.Shutdown(); data->sceneMaterials.Shutdown();
    delete data;
    return rc;
}
int ref_write_scene_blob(const char* path, const char* name, const char* iblName, const float* background4, const void* subsceneNames512,
                         uint32_t subsceneCount, const void* instances72, uint32_t instanceCount, const void* lights64, uint32_t lightCount,
                         const void* lightSets8, uint32_t lightSetCount, const void* cameras176, uint32_t cameraCount) {
    SceneResourceData* data = new SceneResourceData();
    copyFixed(data->name, name);
    copyFixed(data->iblName, iblName ? iblName : "");
    data->backgroundIntensity = float4(background4[0], background4[1], background4[2], background4[3]);
    fillArray(data->subsceneNames, subsceneNames512, subsceneCount);
    fillArray(data->subsceneInstances, instances72, instanceCount);
    fillArray(data->lights, lights64, lightCount);
    fillArray(data->lightSetRanges, lightSets8, lightSetCount);
    fillArray(data->cameras, cameras176, cameraCount);
    int rc = writeBlob(*data, path);
    data->subsceneNames.Shutdown(); data->subsceneInstances.Shutdown(); data->lights.Shutdown(); data->lightSetRanges.Shutdown();
    data->cameras.Shutdown();
    delete data;
    return rc;
}

}  // extern "C"
