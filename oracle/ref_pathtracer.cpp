// oracle/ref_pathtracer.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// The reference's own integrator loop as the oracle for parity level (d).  Two reference translation units are used UNMODIFIED
// and compiled where they lie under /root/reference (oracle/Makefile):
//   * Source/Applications/Selas/Source/PathTracer.cpp -- #included below, so that the file-static EvaluatePath (:142-258), RayPick
//     (:95-139) and OcclusionRay (:71-92) are reachable from this harness; nothing of it is copied or edited;
//   * Source/Core/Shading/SurfaceParameters.cpp       -- CalculateSurfaceParams (:144-247), GeometricNormal, OffsetRayOrigin;
//   * Source/Core/TextureLib/Framebuffer.cpp          -- FramebufferWriter_Write / Flush (:85-135), FrameBuffer_Initialize / Scale.
// What they call that the reference gets from Embree and its caches is supplied here as a shim with Embree's documented semantics
// (share/man/man3/rtcIntersect1.3embree3, rtcOccluded1.3embree3, rtcInterpolate.3embree3, RTC_GEOMETRY_TYPE_TRIANGLE / _QUAD):
//   rtcIntersect1 / rtcOccluded1   over the oracle's CPU BVH (oracle/cpu_bvh.h) -- Embree itself is absent, so the ray-cast
//                                  arithmetic stays "parity unpinned"; everything downstream of the hit is reference code;
//   rtcInterpolate                 barycentric (triangles) / bilinear (quads) attribute interpolation;
//   GeometryCache / TextureCache   the geometry is always resident; FetchTexture returns the mesh's base colour texture or null;
//   spin locks                     SystemLib/OSThreading_Osx.cpp does not compile here (dispatch/dispatch.h).
// Random numbers: CSampler::UniformFloat (ref_harness.cpp) replays the product's counter-keyed PCG stream per (seed, pixel, sample),
// so this loop, the restated loop in ref_harness.cpp and the CUDA wavefront walk the same paths.
//
// The scene graph handed to the reference code is the real one: SceneResource -> subsceneInstances[0] (identity) ->
// SubsceneResource -> modelInstances[0] (identity) -> ModelResource::userDatas[geomId] (ModelDataFromRayIds, SceneResource.cpp:362-378,
// SubsceneResource.cpp:186-195, both linked unmodified); hits report instID[0] = instID[1] = 0 as Embree's instance stack would.
#include "PathTracer.cpp"  // the unmodified reference translation unit (include path: -I$(REF)/Source/Applications/Selas/Source)

#include <atomic>
#include <thread>

#include "SceneLib/GeometryCache.h"
#include "TextureLib/TextureCache.h"
#include "TextureLib/Framebuffer.h"
#include "ref_harness.h"

using namespace Selas;
using namespace refh;

// ---- Embree shim -----------------------------------------------------------------------------------------------------------------
namespace {
struct ShimGeometry {  // what an RTCGeometry handle points at
    const RefScene::Mesh* mesh;
};
struct ShimScene {  // what an RTCScene handle points at
    const RefScene* scene;
    std::atomic<uint64_t> closest{0}, occluded{0};
};
}  // namespace

extern "C" {
RTC_API void rtcIntersect1(RTCScene scene, struct RTCIntersectContext*, struct RTCRayHit* rh) {
    ShimScene* s = reinterpret_cast<ShimScene*>(scene);
    float o[3] = {rh->ray.org_x, rh->ray.org_y, rh->ray.org_z}, d[3] = {rh->ray.dir_x, rh->ray.dir_y, rh->ray.dir_z};
    orc::RayHit hit;
    s->closest.fetch_add(1, std::memory_order_relaxed);
    if (!s->scene->bvh.trace<float>(o, d, rh->ray.tnear, rh->ray.tfar, false, hit)) return;
    const orc::Tri& T = s->scene->bvh.tris[hit.tri];
    rh->ray.tfar = (float)hit.t;
    // quads: (u, v) of the second half are flipped (RTC_GEOMETRY_TYPE_QUAD.3embree3:38-41)
    rh->hit.u = T.second ? (float)(1.0 - hit.u) : (float)hit.u;
    rh->hit.v = T.second ? (float)(1.0 - hit.v) : (float)hit.v;
    const float e1[3] = {T.v1[0] - T.v0[0], T.v1[1] - T.v0[1], T.v1[2] - T.v0[2]}, e2[3] = {T.v2[0] - T.v0[0], T.v2[1] - T.v0[1], T.v2[2] - T.v0[2]};
    rh->hit.Ng_x = e1[1] * e2[2] - e1[2] * e2[1];  // unnormalised object-space normal (rtcIntersect1.3embree3:53)
    rh->hit.Ng_y = e1[2] * e2[0] - e1[0] * e2[2];
    rh->hit.Ng_z = e1[0] * e2[1] - e1[1] * e2[0];
    rh->hit.primID = T.primID;
    rh->hit.geomID = T.geomID;
    rh->hit.instID[0] = 0;  // subscene instance 0 ...
    rh->hit.instID[1] = 0;  // ... model instance 0 (SceneResource.cpp:135-136)
}

RTC_API void rtcOccluded1(RTCScene scene, struct RTCIntersectContext*, struct RTCRay* ray) {
    ShimScene* s = reinterpret_cast<ShimScene*>(scene);
    float o[3] = {ray->org_x, ray->org_y, ray->org_z}, d[3] = {ray->dir_x, ray->dir_y, ray->dir_z};
    orc::RayHit hit;
    s->occluded.fetch_add(1, std::memory_order_relaxed);
    if (s->scene->bvh.trace<float>(o, d, ray->tnear, ray->tfar, true, hit)) ray->tfar = -INFINITY;  // rtcOccluded1.3embree3:43
}

RTC_API void rtcInterpolate(const struct RTCInterpolateArguments* a) {
    const RefScene::Mesh& M = *reinterpret_cast<const ShimGeometry*>(a->geometry)->mesh;
    const std::vector<float>* src = nullptr;
    if (a->bufferType == RTC_BUFFER_TYPE_VERTEX) src = &M.positions;
    else if (a->bufferSlot == 0) src = &M.normals;
    else if (a->bufferSlot == 1) src = &M.tangents;
    else src = &M.uvs;
    const unsigned comps = a->bufferType == RTC_BUFFER_TYPE_VERTEX ? 3u : (a->bufferSlot == 0 ? 3u : (a->bufferSlot == 1 ? 4u : 2u));
    const uint32_t* ix = &M.indices[(size_t)a->primID * M.indicesPerFace];
    const float u = a->u, v = a->v;
    for (unsigned c = 0; c < a->valueCount && c < comps; ++c) {
        const float* p = src->data();
        if (M.indicesPerFace == 3)  // RTC_GEOMETRY_TYPE_TRIANGLE.3embree3:47
            a->P[c] = (1.0f - u - v) * p[ix[0] * comps + c] + u * p[ix[1] * comps + c] + v * p[ix[2] * comps + c];
        else  // RTC_GEOMETRY_TYPE_QUAD.3embree3:48-52
            a->P[c] = (1.0f - v) * ((1.0f - u) * p[ix[0] * comps + c] + u * p[ix[1] * comps + c]) +
                      v * ((1.0f - u) * p[ix[3] * comps + c] + u * p[ix[2] * comps + c]);
    }
}
}  // extern "C"

// ---- caches and platform layer -----------------------------------------------------------------------------------------------------
namespace Selas {
struct TextureCacheData {
    std::vector<TextureResource*> textures;  // handle hash h -> textures[h - 1]
};
TextureCache::TextureCache() : cacheData(nullptr) {}
TextureCache::~TextureCache() {}
const TextureResource* TextureCache::FetchTexture(TextureHandle handle) {
    if (handle.Invalid() || !cacheData) return nullptr;
    return cacheData->textures[handle.hash - 1];
}
void TextureCache::ReleaseTexture(TextureHandle) {}

void GeometryCache::EnsureSubsceneGeometryLoaded(SubsceneResource*) {}
void GeometryCache::FinishUsingSubceneGeometry(SubsceneResource*) {}

// SystemLib/OSThreading_Osx.cpp:89-130 (does not compile here); Framebuffer.cpp's writers need these four
void CreateSpinLock(uint8 spin[CacheLineSize_]) { memset(spin, 0, CacheLineSize_); }
bool TryEnterSpinLock(void* spinlock) { return __sync_val_compare_and_swap(reinterpret_cast<volatile int32*>(spinlock), 0, 1) == 0; }
void EnterSpinLock(void* spinlock) {
    while (__sync_val_compare_and_swap(reinterpret_cast<volatile int32*>(spinlock), 0, 1) == 1) {}
}
void LeaveSpinLock(void* spinlock) {
    __sync_synchronize();
    *reinterpret_cast<volatile int32*>(spinlock) = 0;
}
}  // namespace Selas

// ---- the reference scene graph around a RefScene ------------------------------------------------------------------------------------
namespace {

template <typename T>
T* zeroed() {  // resource structs whose constructors / destructors assert on a lifecycle this harness does not go through
    void* p = calloc(1, sizeof(T));
    return reinterpret_cast<T*>(p);
}

struct RefGraph {
    ShimScene shim;
    std::vector<ShimGeometry> geometries;
    std::vector<MaterialResourceData*> materials;
    SceneResource* scene = nullptr;
    SceneResourceData* sceneData = nullptr;
    SubsceneResource* subscene = nullptr;
    SubsceneResourceData* subsceneData = nullptr;
    ModelResource* model = nullptr;
    SubsceneResource* subsceneTable[1];
    ModelResource* modelTable[1];
    FakeContext lightsAndSky;  // reuses the light-set / background / IBL wiring of the operator harness
    TextureCache textureCache;
    GeometryCache* geometryCache = nullptr;

    explicit RefGraph(const RefScene& sc) {
        shim.scene = &sc;
        lightsAndSky.init(sc.lights.data(), (uint32_t)sc.lights.size(), sc.sets.data(), (uint32_t)sc.sets.size(), sc.background);
        if (sc.ibl.present) lightsAndSky.setIbl(&sc.ibl.data);
        scene = lightsAndSky.scene();
        sceneData = lightsAndSky.data();
        scene->rtcScene = reinterpret_cast<RTCScene>(&shim);
        Instance identity;  // Instance() sets localToWorld = identity (SubsceneResource.h:26-35)
        identity.index = 0;
        sceneData->subsceneInstances.Add(identity);
        subscene = zeroed<SubsceneResource>();
        subsceneData = new SubsceneResourceData();
        subsceneData->modelInstances.Add(identity);
        subscene->data = subsceneData;
        subscene->rtcScene = scene->rtcScene;
        model = zeroed<ModelResource>();
        modelTable[0] = model;
        subscene->models = modelTable;
        subsceneTable[0] = subscene;
        scene->subscenes = subsceneTable;
        geometryCache = zeroed<GeometryCache>();

        geometries.resize(sc.meshes.size());
        materials.resize(sc.materials.size());
        for (size_t m = 0; m < sc.materials.size(); ++m) {
            MaterialResourceData* md = new MaterialResourceData();
            const SelasMaterial& src = sc.materials[m];
            md->shader = (ShaderType)src.shader;
            md->flags = src.flags;
            md->baseColor = float3(src.baseColor[0], src.baseColor[1], src.baseColor[2]);
            md->transmittanceColor = float3(src.transmittanceColor[0], src.transmittanceColor[1], src.transmittanceColor[2]);
            for (int k = 0; k < (int)eMaterialPropertyCount && k < (int)(sizeof(src.scalars) / sizeof(float)); ++k) md->scalarAttributeValues[k] = src.scalars[k];
            materials[m] = md;
        }
        model->userDatas.Resize((uint64)sc.meshes.size());
        for (size_t g = 0; g < sc.meshes.size(); ++g) {
            const RefScene::Mesh& M = sc.meshes[g];
            geometries[g].mesh = &M;
            ModelGeometryUserData& ud = model->userDatas[g];
            memset(&ud, 0, sizeof(ud));
            ud.material = materials[M.materialIndex];
            ud.subscene = subscene;
            ud.rtcGeometry = reinterpret_cast<RTCGeometry>(&geometries[g]);
            ud.flags = (uint16)((M.normals.empty() ? 0 : HasNormals) | (M.tangents.empty() ? 0 : HasTangents) | (M.uvs.empty() ? 0 : HasUvs));
            ud.lightSetIndex = (uint16)M.lightSetIndex;
        }
    }
    ~RefGraph() {
        sceneData->subsceneInstances.Shutdown();
        subsceneData->modelInstances.Shutdown();
        delete subsceneData;
        model->userDatas.Shutdown();
        free(model);
        free(subscene);
        free(geometryCache);
        for (MaterialResourceData* m : materials) delete m;
        scene->subscenes = nullptr;
    }
};

}  // namespace

extern "C" {

// PathTracer::EvaluatePath (PathTracer.cpp:142-258, unmodified) for every (pixel, sample) of the settings, through the reference's
// FramebufferWriter into a two-layer Framebuffer (LayerCount_ = 2: Ld[0] = first-bounce BSDF weight, Ld[1] = radiance).  The caller's
// film gets layer 1 when it asks for one layer, both otherwise; sums are unscaled (FrameBuffer_Scale is the caller's 1 / spp).
// Limits of the unmodified loop: Russian roulette starts after bounce 8 (PathTracer.cpp:249) and the MIS weight is 1 (:215), so
// rrStartBounce must be 8 (or >= maxPathLength) and misMode 0; integrator 0 only.
int refpt_render(void* s, const SelasCamera* cam, const SelasRenderSettings* rs, float* const* film, SelasRenderStats* stats, int nthreads) {
    const RefScene& sc = *(RefScene*)s;
    if (rs->integrator != SELAS_INTEGRATOR_PATHTRACER || rs->misMode != 0 || rs->pixelSampler != 0) return 5;
    if (rs->rrStartBounce != 8 && rs->rrStartBounce < rs->maxPathLength) return 5;
    if (rs->layerCount < 1 || rs->layerCount > 2) return 5;
    RefGraph graph(sc);
    RayCastCameraSettings rc;
    memcpy(&rc, cam, sizeof(rc));
    Framebuffer frame;
    FrameBuffer_Initialize(&frame, rs->width, rs->height, LayerCount_);
    const size_t pixels = (size_t)rs->width * rs->height;
    const size_t stride = rs->pixelStride > 1 ? rs->pixelStride : 1;
    std::atomic<size_t> next(0);
    auto worker = [&]() {
        setFtz();
        // PathTracerKernel, PathTracer.cpp:261-300, with the atomic pixel claiming replaced by blocks of 1024 pixels
        alignas(16) unsigned char storage[sizeof(GIIntegratorContext)];
        memset(storage, 0, sizeof(storage));
        GIIntegratorContext& context = *reinterpret_cast<GIIntegratorContext*>(storage);
        context.geometryCache = graph.geometryCache;
        context.textureCache = &graph.textureCache;
        context.rtcScene = graph.scene->rtcScene;
        context.scene = graph.scene;
        context.camera = &rc;
        context.maxPathLength = rs->maxPathLength;
        FramebufferWriter_Initialize(&context.frameWriter, &frame);
        for (;;) {
            size_t b = next.fetch_add(1024);
            if (b >= pixels) break;
            size_t e = std::min(pixels, b + 1024);
            for (size_t p = b; p < e; ++p) {
                if (stride > 1 && (p % stride) != 0) continue;
                uint32_t y = (uint32_t)(p / rs->width), x = (uint32_t)(p - (size_t)y * rs->width);
                for (uint32_t k = 0; k < rs->samplesPerPixel; ++k) {
                    g_src = DrawSource{nullptr, 0, 0, mix32(rs->seed ^ mix32((uint32_t)p ^ mix32(rs->firstSample + k + 0x9e3779b9u))), true};
                    Ray ray = JitteredCameraRay(context.camera, &context.sampler, (float)x, (float)y);
                    PathTracer::EvaluatePath(&context, ray, x, y);
                }
            }
        }
        FramebufferWriter_Shutdown(&context.frameWriter);
        g_src = DrawSource{nullptr, 0, 0, 0, false};
    };
    std::vector<std::thread> th;
    for (int t = 0; t < std::max(1, nthreads); ++t) th.emplace_back(worker);
    for (auto& x : th) x.join();
    if (rs->layerCount == 1) memcpy(film[0], frame.buffers[1], pixels * sizeof(float3));
    else
        for (uint32_t l = 0; l < 2; ++l) memcpy(film[l], frame.buffers[l], pixels * sizeof(float3));
    FrameBuffer_Shutdown(&frame);
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->closestRays = graph.shim.closest.load();
        stats->occlusionRays = graph.shim.occluded.load();
        stats->pathSamples = ((pixels + stride - 1) / stride) * rs->samplesPerPixel;
    }
    return 0;
}

// CalculateSurfaceParams (SurfaceParameters.cpp:144-247, unmodified) for a batch of hits: the surface-parameter parity oracle.
int refpt_surface_params(void* s, size_t n, const SelasRay* rays, const SelasHit* hits, SelasSurface* out) {
    const RefScene& sc = *(RefScene*)s;
    RefGraph graph(sc);
    setFtz();
    alignas(16) unsigned char storage[sizeof(GIIntegratorContext)];
    memset(storage, 0, sizeof(storage));
    GIIntegratorContext& context = *reinterpret_cast<GIIntegratorContext*>(storage);
    context.geometryCache = graph.geometryCache;
    context.textureCache = &graph.textureCache;
    context.rtcScene = graph.scene->rtcScene;
    context.scene = graph.scene;
    for (size_t i = 0; i < n; ++i) {
        memset(&out[i], 0, sizeof(out[i]));
        if (hits[i].geomID == SELAS_INVALID_ID) continue;
        const SelasRay& r = rays[i];
        // HitParameters as RayPick fills them (PathTracer.cpp:120-136)
        HitParameters hit;
        memset(&hit, 0, sizeof(hit));
        const float t = hits[i].t;
        hit.position = float3(r.ox + t * r.dx, r.oy + t * r.dy, r.oz + t * r.dz);
        const RefScene::Mesh& M = sc.meshes[hits[i].geomID];
        {  // Ng of the hit primitive half, as rtcIntersect1 above reports it
            const uint32_t* ix = &M.indices[(size_t)hits[i].primID * M.indicesPerFace];
            uint32_t i0, i1, i2;
            if (M.indicesPerFace == 3) { i0 = ix[0]; i1 = ix[1]; i2 = ix[2]; }
            else if (hits[i].u + hits[i].v <= 1.0f) { i0 = ix[0]; i1 = ix[1]; i2 = ix[3]; }
            else { i0 = ix[2]; i1 = ix[3]; i2 = ix[1]; }
            const float* P = M.positions.data();
            float3 a(P[3 * i0], P[3 * i0 + 1], P[3 * i0 + 2]), b(P[3 * i1], P[3 * i1 + 1], P[3 * i1 + 2]), c(P[3 * i2], P[3 * i2 + 1], P[3 * i2 + 2]);
            hit.normal = Cross(b - a, c - a);
        }
        hit.baryCoords = {hits[i].u, hits[i].v};
        hit.geomId = (int32)hits[i].geomID;
        hit.primId = (int32)hits[i].primID;
        hit.instId[0] = 0;
        hit.instId[1] = 0;
        hit.view = -float3(r.dx, r.dy, r.dz);
        const float kErr = 32.0f * 1.19209e-07f;
        hit.error = kErr * Max(Max(Math::Absf(hit.position.x), Math::Absf(hit.position.y)), Max(Math::Absf(hit.position.z), t));
        SurfaceParameters surface;
        memset(&surface, 0, sizeof(surface));
        if (!CalculateSurfaceParams(&context, &hit, surface)) continue;
        memcpy(&out[i], &surface, sizeof(surface));
    }
    return 0;
}

}  // extern "C"
