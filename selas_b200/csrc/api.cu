// api.cu -- the C ABI of libselas_cuda.so (include/selas_cuda.h): contexts, scene upload + BVH build, and
// the host-buffer entry points (copy in, launch, copy out, synchronise).  No CPU fallback anywhere: every
// entry point either runs its CUDA kernels or fails with SELAS_ERR_CUDA.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <new>

#include "internal.h"
#include "vec.cuh"

namespace selas {

static thread_local std::string g_threadError;

void setError(SelasContext* ctx, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_threadError = buf;
    if (ctx) ctx->lastError = buf;
}

namespace {

// Device scratch that frees itself; used by the host-buffer entry points.
struct Scratch {
    SelasContext* ctx;
    std::vector<void*> ptrs;
    explicit Scratch(SelasContext* c) : ctx(c) {}
    ~Scratch() {
        for (void* p : ptrs) cudaFree(p);
    }
    template <typename T>
    int alloc(T** out, size_t count) {
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, count * sizeof(T) + 16);
        if (e != cudaSuccess) {
            setError(ctx, "cudaMalloc(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(e));
            return e == cudaErrorMemoryAllocation ? SELAS_ERR_MEMORY : SELAS_ERR_CUDA;
        }
        ptrs.push_back(p);
        *out = (T*)p;
        return SELAS_OK;
    }
    template <typename T>
    int upload(T** out, const T* host, size_t count) {
        int rc = alloc(out, count);
        if (rc) return rc;
        if (count) SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(*out, host, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
        return SELAS_OK;
    }
};

#define SELAS_TRY(expr)          \
    do {                         \
        int rc__ = (expr);       \
        if (rc__) return rc__;   \
    } while (0)

int bind(SelasContext* ctx) {
    if (!ctx) {
        setError(nullptr, "null context");
        return SELAS_ERR_ARGUMENT;
    }
    SELAS_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    return SELAS_OK;
}

template <typename T>
int sceneUpload(SelasScene* scene, const T** out, const T* host, size_t count) {
    void* p = nullptr;
    size_t bytes = count * sizeof(T);
    SelasContext* ctx = scene->ctx;
    cudaError_t e = cudaMalloc(&p, bytes + 16);
    if (e != cudaSuccess) {
        setError(ctx, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        return SELAS_ERR_MEMORY;
    }
    scene->allocations.push_back(p);
    if (bytes) SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(p, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    *out = (const T*)p;
    return SELAS_OK;
}

}  // namespace
}  // namespace selas

using namespace selas;

extern "C" {

// ---- context -------------------------------------------------------------------------------------------------
int selas_cuda_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

const char* selas_cuda_version(void) { return "selas-b200 0.1 (sm_100a)"; }

const char* selas_cuda_last_error(const SelasContext* ctx) {
    if (ctx) return ctx->lastError.c_str();
    return g_threadError.c_str();
}

int selas_cuda_context_create(int device, SelasContext** out) {
    if (!out) {
        setError(nullptr, "selas_cuda_context_create: out is null");
        return SELAS_ERR_ARGUMENT;
    }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        setError(nullptr, "no CUDA device available (%s); this library has no CPU fallback",
                 e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return SELAS_ERR_CUDA;
    }
    if (device < 0 || device >= count) {
        setError(nullptr, "device %d out of range (0..%d)", device, count - 1);
        return SELAS_ERR_ARGUMENT;
    }
    SelasContext* ctx = new (std::nothrow) SelasContext();
    if (!ctx) return SELAS_ERR_MEMORY;
    ctx->device = device;
    SELAS_CUDA_CHECK(nullptr, cudaSetDevice(device));
    cudaDeviceProp prop;
    SELAS_CUDA_CHECK(nullptr, cudaGetDeviceProperties(&prop, device));
    ctx->smCount = prop.multiProcessorCount;
    SELAS_CUDA_CHECK(nullptr, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    SELAS_CUDA_CHECK(nullptr, cudaMalloc((void**)&ctx->d_counters, 256 * sizeof(uint32_t)));
    SELAS_CUDA_CHECK(nullptr, cudaMemset(ctx->d_counters, 0, 256 * sizeof(uint32_t)));
    *out = ctx;
    return SELAS_OK;
}

int selas_cuda_context_destroy(SelasContext* ctx) {
    if (!ctx) return SELAS_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) {
        cudaStreamSynchronize(ctx->stream);
        cudaStreamDestroy(ctx->stream);
    }
    cudaFree(ctx->d_counters);
    cudaFree(ctx->d_rayStage);
    for (cudaStream_t st : ctx->stageStream)
        if (st) cudaStreamDestroy(st);
    delete ctx;
    return SELAS_OK;
}

// ---- scene ---------------------------------------------------------------------------------------------------
int selas_cuda_scene_create(SelasContext* ctx, const SelasSceneDesc* d, SelasScene** out) {
    SELAS_TRY(bind(ctx));
    if (!d || !out) {
        setError(ctx, "selas_cuda_scene_create: null argument");
        return SELAS_ERR_ARGUMENT;
    }
    *out = nullptr;
    if (d->meshCount && !d->meshes) {
        setError(ctx, "meshCount > 0 but meshes is null");
        return SELAS_ERR_ARGUMENT;
    }
    for (uint32_t m = 0; m < d->meshCount; ++m) {
        const SelasMeshDesc& md = d->meshes[m];
        if (md.indicesPerFace != 3 && md.indicesPerFace != 4) {
            setError(ctx, "mesh %u: indicesPerFace must be 3 or 4 (got %u)", m, md.indicesPerFace);
            return SELAS_ERR_ARGUMENT;
        }
        if (md.faceCount && (!md.positions || !md.indices)) {
            setError(ctx, "mesh %u: positions/indices missing", m);
            return SELAS_ERR_ARGUMENT;
        }
        if (md.materialIndex >= d->materialCount) {
            setError(ctx, "mesh %u: materialIndex %u out of range (%u materials)", m, md.materialIndex, d->materialCount);
            return SELAS_ERR_ARGUMENT;
        }
        if (md.baseColorTexture) {
            const SelasTextureDesc& t = *md.baseColorTexture;
            if (t.width == 0 || t.height == 0 || t.format > SELAS_TEXTURE_FLOAT4 || !t.texels) {
                setError(ctx, "mesh %u: baseColorTexture needs width, height > 0, a TextureDataType format and texels", m);
                return SELAS_ERR_ARGUMENT;
            }
        }
        const size_t nidx = (size_t)md.faceCount * md.indicesPerFace;
        for (size_t k = 0; k < nidx; ++k)
            if (md.indices[k] >= md.vertexCount) {
                setError(ctx, "mesh %u: index %u at %zu out of range (%u vertices)", m, md.indices[k], k, md.vertexCount);
                return SELAS_ERR_ARGUMENT;
            }
    }
    for (uint32_t i = 0; i < d->instanceCount; ++i)
        if (d->instances[i].meshIndex >= d->meshCount + d->curveCount) {
            setError(ctx, "instance %u: meshIndex %u out of range (%u meshes + %u curve geometries)", i, d->instances[i].meshIndex, d->meshCount, d->curveCount);
            return SELAS_ERR_ARGUMENT;
        }
    if (d->curveCount && !d->curves) {
        setError(ctx, "selas_cuda_scene_create: curveCount %u with a null curves pointer", d->curveCount);
        return SELAS_ERR_ARGUMENT;
    }
    for (uint32_t c = 0; c < d->curveCount; ++c) {
        const SelasCurveDesc& cd = d->curves[c];
        if (cd.materialIndex >= d->materialCount || (cd.segmentCount && (!cd.vertices || !cd.indices))) {
            setError(ctx, "curve geometry %u: material %u out of range or null buffers", c, cd.materialIndex);
            return SELAS_ERR_ARGUMENT;
        }
        for (uint32_t k = 0; k < cd.segmentCount; ++k)
            if ((uint64_t)cd.indices[k] + 3 >= cd.vertexCount) {
                setError(ctx, "curve geometry %u: segment %u starts at control vertex %u of %u (it needs four)", c, k, cd.indices[k], cd.vertexCount);
                return SELAS_ERR_ARGUMENT;
            }
    }
    for (uint32_t i = 0; i < d->lightSetCount; ++i)
        if ((uint64_t)d->lightSets[i].start + d->lightSets[i].count > d->lightCount) {
            setError(ctx, "light set %u: range [%u, +%u) exceeds %u lights", i, d->lightSets[i].start, d->lightSets[i].count, d->lightCount);
            return SELAS_ERR_ARGUMENT;
        }

    SelasScene* scene = new (std::nothrow) SelasScene();
    if (!scene) return SELAS_ERR_MEMORY;
    scene->ctx = ctx;
    struct Guard {
        SelasScene* s;
        ~Guard() {
            if (s) selas_cuda_scene_destroy(s);
        }
    } guard{scene};

    // base colour textures: one device copy per distinct descriptor
    std::vector<const SelasTextureDesc*> texDescs;
    std::vector<TextureGpu> texGpu;
    for (uint32_t m = 0; m < d->meshCount; ++m) {
        const SelasTextureDesc* t = d->meshes[m].baseColorTexture;
        if (!t || std::find(texDescs.begin(), texDescs.end(), t) != texDescs.end()) continue;
        TextureGpu g;
        memset(&g, 0, sizeof(g));
        SELAS_TRY(sceneUpload(scene, &g.texels, t->texels, (size_t)t->width * t->height * (t->format + 1)));
        g.width = t->width;
        g.height = t->height;
        g.format = t->format;
        texDescs.push_back(t);
        texGpu.push_back(g);
    }
    const TextureGpu* d_textures = nullptr;
    if (!texGpu.empty()) SELAS_TRY(sceneUpload(scene, &d_textures, texGpu.data(), texGpu.size()));

    std::vector<MeshGpu> meshes(d->meshCount);
    for (uint32_t m = 0; m < d->meshCount; ++m) {
        const SelasMeshDesc& md = d->meshes[m];
        MeshGpu& g = meshes[m];
        memset(&g, 0, sizeof(g));
        SELAS_TRY(sceneUpload(scene, &g.positions, md.positions, 3 * (size_t)md.vertexCount));
        SELAS_TRY(sceneUpload(scene, &g.indices, md.indices, (size_t)md.faceCount * md.indicesPerFace));
        if (md.normals) SELAS_TRY(sceneUpload(scene, &g.normals, md.normals, 3 * (size_t)md.vertexCount));
        if (md.tangents) SELAS_TRY(sceneUpload(scene, &g.tangents, md.tangents, 4 * (size_t)md.vertexCount));
        if (md.uvs) SELAS_TRY(sceneUpload(scene, &g.uvs, md.uvs, 2 * (size_t)md.vertexCount));
        g.ptex = md.ptex ? ptexDevicePointer((SelasPtex*)md.ptex) : nullptr;
        if (g.ptex) scene->usesPtex = true;
        if (g.ptex && ptexMixedResolutions((SelasPtex*)md.ptex)) scene->ptexMixedRes = true;
        if (md.baseColorTexture)
            g.texture = d_textures + (std::find(texDescs.begin(), texDescs.end(), md.baseColorTexture) - texDescs.begin());
        g.faceCount = md.faceCount;
        g.indicesPerFace = md.indicesPerFace;
        g.materialIndex = md.materialIndex;
        g.lightSetIndex = md.lightSetIndex;
    }
    // curve geometries ride in the mesh table after the meshes (indicesPerFace = 1: positions = float4 control vertices, indices = first
    // control vertex per segment), so instances, materials, light sets and the hit ids treat them like any other geometry
    for (uint32_t c = 0; c < d->curveCount; ++c) {
        const SelasCurveDesc& cd = d->curves[c];
        MeshGpu g;
        memset(&g, 0, sizeof(g));
        SELAS_TRY(sceneUpload(scene, &g.positions, cd.vertices, 4 * (size_t)cd.vertexCount));
        SELAS_TRY(sceneUpload(scene, &g.indices, cd.indices, (size_t)cd.segmentCount));
        g.faceCount = cd.segmentCount;
        g.indicesPerFace = 1;
        g.materialIndex = cd.materialIndex;
        g.lightSetIndex = cd.lightSetIndex;
        meshes.push_back(g);
    }
    const uint32_t geometryCount = (uint32_t)meshes.size();
    std::vector<InstanceGpu> inst;
    if (d->instanceCount) {
        inst.resize(d->instanceCount);
        for (uint32_t i = 0; i < d->instanceCount; ++i) {
            memcpy(inst[i].m, d->instances[i].localToWorld, sizeof(inst[i].m));
            inst[i].mesh = d->instances[i].meshIndex;
            static const float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
            inst[i].identity = memcmp(inst[i].m, I, sizeof(I)) == 0 ? 1u : 0u;
            inst[i].pad[0] = inst[i].pad[1] = 0;
        }
        if (d->flags & SELAS_SCENE_INSTANCE_IDS) {  // the reference's two instancing levels as the ids public hits report
            std::vector<uint4> ids(d->instanceCount);
            for (uint32_t i = 0; i < d->instanceCount; ++i) {
                const SelasInstanceDesc& in = d->instances[i];
                ids[i] = make_uint4(in.instID0, in.instID1, in.geomID, 0u);
                if (!scene->instanceIdLookup.emplace(std::array<uint32_t, 3>{in.instID0, in.instID1, in.geomID}, i).second) {
                    setError(ctx, "instance %u repeats the ids (instID0 %u, instID1 %u, geomID %u) of an earlier instance", i, in.instID0, in.instID1, in.geomID);
                    return SELAS_ERR_ARGUMENT;  // (the guard above releases the scene)
                }
            }
            SELAS_TRY(sceneUpload(scene, &scene->gpu.instanceIds, ids.data(), ids.size()));
        }
    } else if (d->flags & SELAS_SCENE_TWO_LEVEL_BVH) {  // one identity instance per mesh, instID0 = mesh index
        inst.resize(geometryCount ? geometryCount : 1);
        for (uint32_t i = 0; i < inst.size(); ++i) {
            memset(&inst[i], 0, sizeof(InstanceGpu));
            inst[i].m[0] = inst[i].m[5] = inst[i].m[10] = inst[i].m[15] = 1.0f;
            inst[i].mesh = geometryCount ? i : 0xFFFFFFFFu;
            inst[i].identity = 1;
        }
    } else {  // every mesh once, identity, instID0 = 0
        inst.resize(1);
        memset(&inst[0], 0, sizeof(InstanceGpu));
        inst[0].m[0] = inst[0].m[5] = inst[0].m[10] = inst[0].m[15] = 1.0f;
        inst[0].mesh = 0xFFFFFFFFu;
        inst[0].identity = 1;
    }
    SceneGpu& g = scene->gpu;
    SELAS_TRY(sceneUpload(scene, &g.meshes, meshes.data(), meshes.size()));
    SELAS_TRY(sceneUpload(scene, &g.instances, inst.data(), inst.size()));
    SELAS_TRY(sceneUpload(scene, &g.materials, d->materials, d->materialCount));
    {
        void* p = nullptr;
        if (cudaMalloc(&p, sizeof(float4) * (size_t)d->materialCount + 16) != cudaSuccess) {
            setError(ctx, "cudaMalloc failed for the material table");
            return SELAS_ERR_MEMORY;
        }
        scene->allocations.push_back(p);
        g.materialBasePow = (const float4*)p;
        SELAS_TRY(launchMaterialBasePow(ctx, g.materials, d->materialCount, (float4*)p));
    }
    SELAS_TRY(sceneUpload(scene, &g.lights, d->lights, d->lightCount));
    SELAS_TRY(sceneUpload(scene, &g.lightSets, d->lightSets, d->lightSetCount));
    g.meshCount = geometryCount;  // meshes, then curve geometries
    g.instanceCount = (uint32_t)inst.size();
    g.materialCount = d->materialCount;
    g.lightCount = d->lightCount;
    g.lightSetCount = d->lightSetCount;
    memcpy(g.background, d->backgroundIntensity, sizeof(g.background));
    memset(&g.ibl, 0, sizeof(g.ibl));
    if (d->ibl) {
        const SelasIblDesc& I = *d->ibl;
        if (!I.width || !I.height || !I.missWidth || !I.missHeight || !I.marginalDensityFunction || !I.conditionalDensityFunctions ||
            !I.lightData || !I.missData) {
            setError(ctx, "ibl: width/height and all four arrays are required");
            return SELAS_ERR_ARGUMENT;
        }
        SELAS_TRY(sceneUpload(scene, &g.ibl.marginal, I.marginalDensityFunction, I.height));
        SELAS_TRY(sceneUpload(scene, &g.ibl.conditional, I.conditionalDensityFunctions, (size_t)I.width * I.height));
        SELAS_TRY(sceneUpload(scene, &g.ibl.light, I.lightData, 3 * (size_t)I.width * I.height));
        SELAS_TRY(sceneUpload(scene, &g.ibl.miss, I.missData, 3 * (size_t)I.missWidth * I.missHeight));
        g.ibl.width = I.width; g.ibl.height = I.height; g.ibl.missWidth = I.missWidth; g.ibl.missHeight = I.missHeight;
        g.ibl.rotation = I.rotationRadians;
        g.ibl.exposureScale = I.exposureScale;
        g.ibl.present = 1;
    }

    if ((d->flags & SELAS_SCENE_TWO_LEVEL_BVH) && geometryCount)
        SELAS_TRY(buildBvhTwoLevel(scene, meshes, inst));
    else
        SELAS_TRY(buildBvh(scene, meshes, inst, g.instances));
    uint64_t attrBytes = 0;
    for (uint32_t m = 0; m < d->meshCount; ++m) {
        const SelasMeshDesc& md = d->meshes[m];
        attrBytes += 12ull * md.vertexCount + 4ull * md.faceCount * md.indicesPerFace;
        if (md.normals) attrBytes += 12ull * md.vertexCount;
        if (md.tangents) attrBytes += 16ull * md.vertexCount;
        if (md.uvs) attrBytes += 8ull * md.vertexCount;
    }
    scene->stats.deviceBytes += attrBytes;
    guard.s = nullptr;
    *out = scene;
    return SELAS_OK;
}

int selas_cuda_scene_destroy(SelasScene* scene) {
    if (!scene) return SELAS_OK;
    if (scene->ctx) {
        cudaSetDevice(scene->ctx->device);
        cudaStreamSynchronize(scene->ctx->stream);
    }
    freeWavefront(scene);
    freeVcm(scene);
    cudaFree(scene->d_filmCache);
    for (void* p : scene->allocations) cudaFree(p);
    delete scene;
    return SELAS_OK;
}

int selas_cuda_scene_bvh_stats(const SelasScene* scene, SelasBvhStats* out) {
    if (!scene || !out) {
        setError(nullptr, "selas_cuda_scene_bvh_stats: null argument");
        return SELAS_ERR_ARGUMENT;
    }
    *out = scene->stats;
    return SELAS_OK;
}

// ---- ray cast ------------------------------------------------------------------------------------------------
static int checkRayCount(SelasContext* ctx, size_t n) {
    if (n > 0xFFFFFFF0ull) {
        setError(ctx, "ray batch of %zu exceeds the 32-bit queue index", n);
        return SELAS_ERR_ARGUMENT;
    }
    return SELAS_OK;
}

static uint32_t* nextRayCounter(SelasContext* ctx) { return ctx->d_counters + (ctx->counterNext++ & 255u); }

int selas_cuda_context_synchronize(SelasContext* ctx) {
    SELAS_TRY(bind(ctx));
    SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return SELAS_OK;
}

int selas_cuda_intersect_device(SelasScene* scene, size_t n, const SelasRay* d_rays, SelasHit* d_hits, float* d_Ng4, void* stream) {
    if (!scene) return SELAS_ERR_ARGUMENT;
    SELAS_TRY(bind(scene->ctx));
    SELAS_TRY(checkRayCount(scene->ctx, n));
    cudaStream_t st = stream ? (cudaStream_t)stream : scene->ctx->stream;
    return launchIntersect(scene, (uint32_t)n, d_rays, d_hits, d_Ng4, nullptr, nextRayCounter(scene->ctx), st);
}

int selas_cuda_occluded_device(SelasScene* scene, size_t n, const SelasRay* d_rays, uint8_t* d_occluded, void* stream) {
    if (!scene) return SELAS_ERR_ARGUMENT;
    SELAS_TRY(bind(scene->ctx));
    SELAS_TRY(checkRayCount(scene->ctx, n));
    cudaStream_t st = stream ? (cudaStream_t)stream : scene->ctx->stream;
    return launchOccluded(scene, (uint32_t)n, d_rays, d_occluded, nullptr, nextRayCounter(scene->ctx), st);
}

// Host-buffer ray casts (the seam a maintainer binds first: RayPick / OcclusionRay over arrays).  Round 1 paid a cudaMalloc + cudaFree
// and one monolithic pageable copy each way per call (50 Mrays/s, below the CPU on incoherent rays).  Now: device staging kept in the
// context, the batch cut into chunks that alternate between two streams (upload -> trace -> download per chunk), so downloads overlap the
// next upload; with pinned caller buffers the call is PCIe-bound, with pageable ones it is bound by the driver's staging copies.
namespace {
constexpr size_t kStageChunk = 1u << 19;        // rays per chunk (16 MB up, 16 MB down)
constexpr size_t kStageMaxRays = 1u << 23;      // staging capacity: 8 M rays (about 0.7 GB); larger batches go through it in rounds
constexpr size_t kStageBytesPerRay = sizeof(SelasRay) + sizeof(SelasHit) + 4 * sizeof(float) + 16;

int ensureRayStage(SelasContext* ctx, size_t rays) {
    for (cudaStream_t& st : ctx->stageStream)
        if (!st) SELAS_CUDA_CHECK(ctx, cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    if (ctx->rayStageCapacity >= rays) return SELAS_OK;
    cudaFree(ctx->d_rayStage);
    ctx->d_rayStage = nullptr;
    ctx->rayStageCapacity = 0;
    cudaError_t e = cudaMalloc(&ctx->d_rayStage, rays * kStageBytesPerRay);
    if (e != cudaSuccess) {
        setError(ctx, "cudaMalloc(%zu bytes) failed for the ray staging buffers: %s", rays * kStageBytesPerRay, cudaGetErrorString(e));
        cudaGetLastError();
        return e == cudaErrorMemoryAllocation ? SELAS_ERR_MEMORY : SELAS_ERR_CUDA;
    }
    ctx->rayStageCapacity = rays;
    return SELAS_OK;
}

// any == false: hits (+ Ng4) out; any == true: occluded bytes out
int castHostRays(SelasScene* scene, size_t n, const SelasRay* rays, SelasHit* hits, float* Ng4, uint8_t* occluded, bool any) {
    SelasContext* ctx = scene->ctx;
    const size_t cap = std::min(n, kStageMaxRays);
    SELAS_TRY(ensureRayStage(ctx, cap));
    SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));  // order after what the caller enqueued on the context's stream (scene build)
    char* base = (char*)ctx->d_rayStage;
    SelasRay* d_rays = (SelasRay*)base;
    SelasHit* d_hits = (SelasHit*)(base + ctx->rayStageCapacity * sizeof(SelasRay));
    float* d_ng = (float*)(base + ctx->rayStageCapacity * (sizeof(SelasRay) + sizeof(SelasHit)));
    uint8_t* d_occ = (uint8_t*)(base + ctx->rayStageCapacity * (sizeof(SelasRay) + sizeof(SelasHit) + 4 * sizeof(float)));
    for (size_t round = 0; round < n; round += cap) {
        const size_t m = std::min(cap, n - round);
        int k = 0;
        for (size_t off = 0; off < m; off += kStageChunk, k ^= 1) {
            const size_t c = std::min(kStageChunk, m - off);
            cudaStream_t st = ctx->stageStream[k];
            SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(d_rays + off, rays + round + off, c * sizeof(SelasRay), cudaMemcpyHostToDevice, st));
            if (any) {
                SELAS_TRY(launchOccluded(scene, (uint32_t)c, d_rays + off, d_occ + off, nullptr, nextRayCounter(ctx), st));
                SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(occluded + round + off, d_occ + off, c, cudaMemcpyDeviceToHost, st));
            } else {
                SELAS_TRY(launchIntersect(scene, (uint32_t)c, d_rays + off, d_hits + off, Ng4 ? d_ng + 4 * off : nullptr, nullptr, nextRayCounter(ctx), st));
                SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(hits + round + off, d_hits + off, c * sizeof(SelasHit), cudaMemcpyDeviceToHost, st));
                if (Ng4) SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(Ng4 + 4 * (round + off), d_ng + 4 * off, 4 * c * sizeof(float), cudaMemcpyDeviceToHost, st));
            }
        }
        SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stageStream[0]));
        SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stageStream[1]));
    }
    return SELAS_OK;
}
}  // namespace

int selas_cuda_intersect(SelasScene* scene, size_t n, const SelasRay* rays, SelasHit* hits, float* Ng4) {
    if (!scene) return SELAS_ERR_ARGUMENT;
    SelasContext* ctx = scene->ctx;
    SELAS_TRY(bind(ctx));
    SELAS_TRY(checkRayCount(ctx, n));
    if (n == 0) return SELAS_OK;
    if (!rays || !hits) {
        setError(ctx, "selas_cuda_intersect: null rays/hits");
        return SELAS_ERR_ARGUMENT;
    }
    return castHostRays(scene, n, rays, hits, Ng4, nullptr, false);
}

int selas_cuda_occluded(SelasScene* scene, size_t n, const SelasRay* rays, uint8_t* occluded) {
    if (!scene) return SELAS_ERR_ARGUMENT;
    SelasContext* ctx = scene->ctx;
    SELAS_TRY(bind(ctx));
    SELAS_TRY(checkRayCount(ctx, n));
    if (n == 0) return SELAS_OK;
    if (!rays || !occluded) {
        setError(ctx, "selas_cuda_occluded: null rays/occluded");
        return SELAS_ERR_ARGUMENT;
    }
    return castHostRays(scene, n, rays, nullptr, nullptr, occluded, true);
}

int selas_cuda_intersect_counted(SelasScene* scene, size_t n, const SelasRay* rays, SelasHit* hits, uint32_t* counts, int anyHit) {
    if (!scene) return SELAS_ERR_ARGUMENT;
    SelasContext* ctx = scene->ctx;
    SELAS_TRY(bind(ctx));
    SELAS_TRY(checkRayCount(ctx, n));
    if (n == 0) return SELAS_OK;
    if (!rays || !counts) {
        setError(ctx, "selas_cuda_intersect_counted: null rays/counts");
        return SELAS_ERR_ARGUMENT;
    }
    Scratch s(ctx);
    SelasRay* d_rays;
    SelasHit* d_hits;
    uint8_t* d_occ;
    uint32_t* d_counts;
    SELAS_TRY(s.upload(&d_rays, rays, n));
    SELAS_TRY(s.alloc(&d_hits, n));
    SELAS_TRY(s.alloc(&d_occ, n));
    SELAS_TRY(s.alloc(&d_counts, 2 * n));
    SELAS_CUDA_CHECK(ctx, cudaMemsetAsync(d_counts, 0, 2 * n * sizeof(uint32_t), ctx->stream));
    if (anyHit) {
        SELAS_TRY(launchOccluded(scene, (uint32_t)n, d_rays, d_occ, d_counts, nextRayCounter(ctx), ctx->stream));
        if (hits) {  // any-hit mode reports only geomID != INVALID for occluded rays
            std::vector<uint8_t> occ(n);
            SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(occ.data(), d_occ, n, cudaMemcpyDeviceToHost, ctx->stream));
            SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
            for (size_t i = 0; i < n; ++i) {
                memset(&hits[i], 0, sizeof(SelasHit));
                hits[i].t = occ[i] ? -INFINITY : rays[i].tfar;
                hits[i].geomID = hits[i].primID = hits[i].instID0 = hits[i].instID1 = occ[i] ? 0u : SELAS_INVALID_ID;
            }
        }
    } else {
        SELAS_TRY(launchIntersect(scene, (uint32_t)n, d_rays, d_hits, nullptr, d_counts, nextRayCounter(ctx), ctx->stream));
        if (hits) SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(hits, d_hits, n * sizeof(SelasHit), cudaMemcpyDeviceToHost, ctx->stream));
    }
    SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(counts, d_counts, 2 * n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return SELAS_OK;
}

// ---- camera --------------------------------------------------------------------------------------------------
// InitializeRayCastCamera, GeometryLib/Camera.cpp:89-115.  Load-time host arithmetic in the reference too.
int selas_cuda_camera_init(const SelasCameraSettings* st, uint32_t width, uint32_t height, SelasCamera* cam) {
    if (!st || !cam || width == 0 || height == 0) {
        setError(nullptr, "selas_cuda_camera_init: bad argument");
        return SELAS_ERR_ARGUMENT;
    }
    memset(cam, 0, sizeof(*cam));
    const float pi = 3.141592653589793f;
    float widthf = (float)width, heightf = (float)height;
    float aspect = widthf / heightf;
    float hfov = st->fovDegrees * (pi / 180.f);
    float hLen = tanf(hfov * 0.5f);
    float vfov = 2 * atanf(hLen / aspect);
    float vLen = tanf(vfov * 0.5f);
    V3 fwd = normalize(v3(st->lookAt) - v3(st->position));
    V3 right = normalize(cross(v3(st->up), fwd));
    store3(cam->cameraX, hLen * right);
    store3(cam->cameraY, vLen * cross(fwd, right));
    store3(cam->cameraZ, fwd);
    store3(cam->forward, fwd);
    store3(cam->position, v3(st->position));
    cam->viewportWidth = widthf;
    cam->viewportHeight = heightf;
    cam->znear = st->znear;
    cam->zfar = st->zfar;
    cam->virtualImagePlaneDistance = widthf / (2.0f * tanf(hfov));
    cam->width = width;
    cam->height = height;
    cam->aspect = aspect;
    return SELAS_OK;
}

int selas_cuda_camera_rays(SelasContext* ctx, const SelasCamera* camera, size_t n, const float* px, const float* py, const float* jitter,
                           SelasRay* rays) {
    SELAS_TRY(bind(ctx));
    if (n == 0) return SELAS_OK;
    if (!camera || !px || !py || !rays) {
        setError(ctx, "selas_cuda_camera_rays: null argument");
        return SELAS_ERR_ARGUMENT;
    }
    Scratch s(ctx);
    float *d_px, *d_py, *d_j = nullptr;
    SelasRay* d_rays;
    SELAS_TRY(s.upload(&d_px, px, n));
    SELAS_TRY(s.upload(&d_py, py, n));
    if (jitter) SELAS_TRY(s.upload(&d_j, jitter, 2 * n));
    SELAS_TRY(s.alloc(&d_rays, n));
    SELAS_TRY(launchCameraRays(ctx, *camera, n, d_px, d_py, d_j, d_rays));
    SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(rays, d_rays, n * sizeof(SelasRay), cudaMemcpyDeviceToHost, ctx->stream));
    SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return SELAS_OK;
}

int selas_cuda_ibl_lookup(SelasScene* scene, size_t n, const float* wi3, float* rgb3, float* pdf, float* missRgb3) {
    if (!scene) return SELAS_ERR_ARGUMENT;
    SelasContext* ctx = scene->ctx;
    SELAS_TRY(bind(ctx));
    if (!scene->gpu.ibl.present) {
        setError(ctx, "selas_cuda_ibl_lookup: the scene has no image based light");
        return SELAS_ERR_ARGUMENT;
    }
    if (n == 0) return SELAS_OK;
    if (!wi3) {
        setError(ctx, "selas_cuda_ibl_lookup: wi3 is null");
        return SELAS_ERR_ARGUMENT;
    }
    Scratch s(ctx);
    float *d_wi, *d_rgb = nullptr, *d_pdf = nullptr, *d_miss = nullptr;
    SELAS_TRY(s.upload(&d_wi, wi3, 3 * n));
    if (rgb3) SELAS_TRY(s.alloc(&d_rgb, 3 * n));
    if (pdf) SELAS_TRY(s.alloc(&d_pdf, n));
    if (missRgb3) SELAS_TRY(s.alloc(&d_miss, 3 * n));
    SELAS_TRY(launchIblLookup(scene, n, d_wi, d_rgb, d_pdf, d_miss));
    if (rgb3) SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(rgb3, d_rgb, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (pdf) SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(pdf, d_pdf, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (missRgb3) SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(missRgb3, d_miss, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return SELAS_OK;
}

int selas_cuda_ibl_sample(SelasScene* scene, size_t n, const float* u2, float* thetaPhiPdf3, uint32_t* xy2) {
    if (!scene) return SELAS_ERR_ARGUMENT;
    SelasContext* ctx = scene->ctx;
    SELAS_TRY(bind(ctx));
    if (!scene->gpu.ibl.present) {
        setError(ctx, "selas_cuda_ibl_sample: the scene has no image based light");
        return SELAS_ERR_ARGUMENT;
    }
    if (n == 0) return SELAS_OK;
    if (!u2 || !thetaPhiPdf3 || !xy2) {
        setError(ctx, "selas_cuda_ibl_sample: null argument");
        return SELAS_ERR_ARGUMENT;
    }
    Scratch s(ctx);
    float *d_u, *d_t;
    uint32_t* d_xy;
    SELAS_TRY(s.upload(&d_u, u2, 2 * n));
    SELAS_TRY(s.alloc(&d_t, 3 * n));
    SELAS_TRY(s.alloc(&d_xy, 2 * n));
    SELAS_TRY(launchIblSample(scene, n, d_u, d_t, d_xy));
    SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(thetaPhiPdf3, d_t, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(xy2, d_xy, 2 * n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return SELAS_OK;
}

int selas_cuda_cmj(SelasContext* ctx, size_t count, const int32_t* sIdx, int32_t m, int32_t n, const int32_t* p, float* out2) {
    SELAS_TRY(bind(ctx));
    if (count == 0) return SELAS_OK;
    if (!sIdx || !p || !out2 || m <= 0 || n <= 0) {
        setError(ctx, "selas_cuda_cmj: bad argument");
        return SELAS_ERR_ARGUMENT;
    }
    Scratch s(ctx);
    int32_t *d_s, *d_p;
    float* d_o;
    SELAS_TRY(s.upload(&d_s, sIdx, count));
    SELAS_TRY(s.upload(&d_p, p, count));
    SELAS_TRY(s.alloc(&d_o, 2 * count));
    SELAS_TRY(launchCmj(ctx, count, d_s, m, n, d_p, d_o));
    SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(out2, d_o, 2 * count * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return SELAS_OK;
}

// ---- BSDF / light / surface operators ----------------------------------------------------------------------------
int selas_cuda_bsdf_eval(SelasContext* ctx, size_t n, const SelasSurface* surfaces, const float* v3in, const float* l3in, float* rgb3,
                         float* forwardPdfW, float* reversePdfW) {
    SELAS_TRY(bind(ctx));
    if (n == 0) return SELAS_OK;
    if (!surfaces || !v3in || !l3in || !rgb3 || !forwardPdfW || !reversePdfW) {
        setError(ctx, "selas_cuda_bsdf_eval: null argument");
        return SELAS_ERR_ARGUMENT;
    }
    Scratch s(ctx);
    SelasSurface* d_s;
    float *d_v, *d_l, *d_rgb, *d_f, *d_r;
    SELAS_TRY(s.upload(&d_s, surfaces, n));
    SELAS_TRY(s.upload(&d_v, v3in, 3 * n));
    SELAS_TRY(s.upload(&d_l, l3in, 3 * n));
    SELAS_TRY(s.alloc(&d_rgb, 3 * n));
    SELAS_TRY(s.alloc(&d_f, n));
    SELAS_TRY(s.alloc(&d_r, n));
    SELAS_TRY(launchBsdfEval(ctx, n, d_s, d_v, d_l, d_rgb, d_f, d_r));
    SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(rgb3, d_rgb, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(forwardPdfW, d_f, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(reversePdfW, d_r, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return SELAS_OK;
}

int selas_cuda_bsdf_sample(SelasContext* ctx, size_t n, const SelasSurface* surfaces, const float* v3in, const float* u, SelasBsdfSample* out) {
    SELAS_TRY(bind(ctx));
    if (n == 0) return SELAS_OK;
    if (!surfaces || !v3in || !u || !out) {
        setError(ctx, "selas_cuda_bsdf_sample: null argument");
        return SELAS_ERR_ARGUMENT;
    }
    Scratch s(ctx);
    SelasSurface* d_s;
    float *d_v, *d_u;
    SelasBsdfSample* d_o;
    SELAS_TRY(s.upload(&d_s, surfaces, n));
    SELAS_TRY(s.upload(&d_v, v3in, 3 * n));
    SELAS_TRY(s.upload(&d_u, u, SELAS_BSDF_RANDOMS * n));
    SELAS_TRY(s.alloc(&d_o, n));
    SELAS_TRY(launchBsdfSample(ctx, n, d_s, d_v, d_u, d_o));
    SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(out, d_o, n * sizeof(SelasBsdfSample), cudaMemcpyDeviceToHost, ctx->stream));
    SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return SELAS_OK;
}

int selas_cuda_light_sample(SelasContext* ctx, size_t n, const SelasLight* lights, uint32_t lightCount, const float* position3,
                            const float* normal3, const float* u3, SelasLightSample* out) {
    SELAS_TRY(bind(ctx));
    if (n == 0) return SELAS_OK;
    if ((lightCount && !lights) || !position3 || !normal3 || !u3 || !out) {
        setError(ctx, "selas_cuda_light_sample: null argument");
        return SELAS_ERR_ARGUMENT;
    }
    Scratch s(ctx);
    SelasLight* d_l;
    float *d_p, *d_n, *d_u;
    SelasLightSample* d_o;
    SELAS_TRY(s.upload(&d_l, lights, lightCount));
    SELAS_TRY(s.upload(&d_p, position3, 3 * n));
    SELAS_TRY(s.upload(&d_n, normal3, 3 * n));
    SELAS_TRY(s.upload(&d_u, u3, 3 * n));
    SELAS_TRY(s.alloc(&d_o, n));
    SELAS_TRY(launchLightSample(ctx, n, d_l, lightCount, d_p, d_n, d_u, d_o));
    SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(out, d_o, n * sizeof(SelasLightSample), cudaMemcpyDeviceToHost, ctx->stream));
    SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return SELAS_OK;
}

int selas_cuda_light_pdf(SelasContext* ctx, size_t n, const SelasLight* lights, uint32_t lightCount, const uint32_t* lightIndex,
                         const float* position3, const float* wi3, float* pdf) {
    SELAS_TRY(bind(ctx));
    if (n == 0) return SELAS_OK;
    if ((lightCount && !lights) || !lightIndex || !position3 || !wi3 || !pdf) {
        setError(ctx, "selas_cuda_light_pdf: null argument");
        return SELAS_ERR_ARGUMENT;
    }
    for (size_t i = 0; i < n; ++i)
        if (lightCount && lightIndex[i] >= lightCount) {
            setError(ctx, "selas_cuda_light_pdf: lightIndex[%zu] = %u out of range", i, lightIndex[i]);
            return SELAS_ERR_ARGUMENT;
        }
    Scratch s(ctx);
    SelasLight* d_l;
    uint32_t* d_i;
    float *d_p, *d_w, *d_o;
    SELAS_TRY(s.upload(&d_l, lights, lightCount));
    SELAS_TRY(s.upload(&d_i, lightIndex, n));
    SELAS_TRY(s.upload(&d_p, position3, 3 * n));
    SELAS_TRY(s.upload(&d_w, wi3, 3 * n));
    SELAS_TRY(s.alloc(&d_o, n));
    SELAS_TRY(launchLightPdf(ctx, n, d_l, lightCount, d_i, d_p, d_w, d_o));
    SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(pdf, d_o, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return SELAS_OK;
}

int selas_cuda_surface_params(SelasScene* scene, size_t n, const SelasRay* rays, const SelasHit* hits, SelasSurface* out) {
    if (!scene) return SELAS_ERR_ARGUMENT;
    SelasContext* ctx = scene->ctx;
    SELAS_TRY(bind(ctx));
    if (n == 0) return SELAS_OK;
    if (!rays || !hits || !out) {
        setError(ctx, "selas_cuda_surface_params: null argument");
        return SELAS_ERR_ARGUMENT;
    }
    // ModelDataFromRayIds (SceneResource.cpp:362-378): the hit's ids name the instance and the mesh.  A SELAS_SCENE_INSTANCE_IDS scene
    // reports the reference's (instID[0], instID[1], geomID); they are resolved to the flat instance and its mesh here.
    std::vector<SelasHit> resolved;
    const bool publicIds = !scene->instanceIdLookup.empty();
    if (publicIds) resolved.assign(hits, hits + n);
    std::vector<uint32_t> meshOfInstance;
    if (publicIds) {
        meshOfInstance.resize(scene->gpu.instanceCount);
        std::vector<InstanceGpu> tmp(scene->gpu.instanceCount);
        SELAS_CUDA_CHECK(ctx, cudaMemcpy(tmp.data(), scene->gpu.instances, sizeof(InstanceGpu) * tmp.size(), cudaMemcpyDeviceToHost));
        for (size_t k = 0; k < tmp.size(); ++k) meshOfInstance[k] = tmp[k].mesh;
    }
    for (size_t i = 0; i < n; ++i) {
        const SelasHit& h = hits[i];
        if (h.geomID == SELAS_INVALID_ID) continue;
        if (publicIds) {
            auto it = scene->instanceIdLookup.find(std::array<uint32_t, 3>{h.instID0, h.instID1, h.geomID});
            if (it == scene->instanceIdLookup.end()) {
                setError(ctx, "selas_cuda_surface_params: hit %zu names (instID0 %u, instID1 %u, geomID %u), which no instance of the scene reports", i,
                         h.instID0, h.instID1, h.geomID);
                return SELAS_ERR_ARGUMENT;
            }
            resolved[i].instID0 = it->second;
            resolved[i].instID1 = 0;
            resolved[i].geomID = meshOfInstance[it->second];
        } else if (h.geomID >= scene->gpu.meshCount || h.instID0 >= scene->gpu.instanceCount) {
            setError(ctx, "selas_cuda_surface_params: hit %zu names geomID %u / instID %u outside the scene", i, h.geomID, h.instID0);
            return SELAS_ERR_ARGUMENT;
        }
    }
    Scratch s(ctx);
    SelasRay* d_r;
    SelasHit* d_h;
    SelasSurface* d_o;
    SELAS_TRY(s.upload(&d_r, rays, n));
    SELAS_TRY(s.upload(&d_h, publicIds ? resolved.data() : hits, n));
    SELAS_TRY(s.alloc(&d_o, n));
    SELAS_TRY(launchSurfaceParams(scene, n, d_r, d_h, d_o));
    SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(out, d_o, n * sizeof(SelasSurface), cudaMemcpyDeviceToHost, ctx->stream));
    SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return SELAS_OK;
}

// ---- textures ------------------------------------------------------------------------------------------------------
int selas_cuda_texture_sample(SelasContext* ctx, const SelasTextureDesc* t, size_t n, const float* uv2, int sRGB, float* rgb3) {
    SELAS_TRY(bind(ctx));
    if (!t || t->width == 0 || t->height == 0 || t->format > SELAS_TEXTURE_FLOAT4 || !t->texels || (n && (!uv2 || !rgb3))) {
        setError(ctx, "selas_cuda_texture_sample: bad texture descriptor or null buffers");
        return SELAS_ERR_ARGUMENT;
    }
    if (!n) return SELAS_OK;
    const size_t floats = (size_t)t->width * t->height * (t->format + 1);
    float *d_tex = nullptr, *d_uv = nullptr, *d_out = nullptr;
    int rc = SELAS_OK;
    auto fail = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && !rc) {
            setError(ctx, "%s failed: %s", what, cudaGetErrorString(e));
            rc = SELAS_ERR_CUDA;
        }
    };
    fail(cudaMalloc((void**)&d_tex, sizeof(float) * floats), "cudaMalloc(texels)");
    fail(cudaMalloc((void**)&d_uv, sizeof(float) * 2 * n), "cudaMalloc(uvs)");
    fail(cudaMalloc((void**)&d_out, sizeof(float) * 3 * n), "cudaMalloc(out)");
    if (!rc) fail(cudaMemcpyAsync(d_tex, t->texels, sizeof(float) * floats, cudaMemcpyHostToDevice, ctx->stream), "upload texels");
    if (!rc) fail(cudaMemcpyAsync(d_uv, uv2, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, ctx->stream), "upload uvs");
    if (!rc) {
        TextureGpu g;
        g.texels = d_tex; g.width = t->width; g.height = t->height; g.format = t->format; g.pad = 0;
        rc = launchTextureSample(ctx, g, n, d_uv, sRGB, d_out);
    }
    if (!rc) fail(cudaMemcpyAsync(rgb3, d_out, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, ctx->stream), "download");
    if (!rc) fail(cudaStreamSynchronize(ctx->stream), "synchronize");
    cudaFree(d_tex);
    cudaFree(d_uv);
    cudaFree(d_out);
    return rc;
}

int selas_cuda_texture_filter(SelasContext* ctx, const SelasTextureMipsDesc* t, int filter, int level, size_t n, const float* st2, const float* dst0,
                              const float* dst1, float* out) {
    SELAS_TRY(bind(ctx));
    const bool needsDerivatives = filter == SELAS_FILTER_TRILINEAR || filter == SELAS_FILTER_EWA || filter == SELAS_FILTER_EWA_LEVEL;
    bool ok = t && t->mipCount >= 1 && t->mipCount <= SELAS_TEXTURE_MAX_MIPS && t->format <= SELAS_TEXTURE_FLOAT4 && t->texels &&
              filter >= SELAS_FILTER_POINT && filter <= SELAS_FILTER_EWA_LEVEL && (!n || (st2 && out && (!needsDerivatives || (dst0 && dst1)))) &&
              !(filter == SELAS_FILTER_EWA_LEVEL && level < 0);
    for (uint32_t l = 0; ok && l < t->mipCount; ++l) {
        const uint64_t bytes = (uint64_t)t->mipWidths[l] * t->mipHeights[l] * (t->format + 1u) * sizeof(float);
        ok = t->mipWidths[l] > 0 && t->mipHeights[l] > 0 && (t->mipOffsets[l] % sizeof(float)) == 0 && t->mipOffsets[l] + bytes <= t->dataSize;
    }
    if (!ok) {
        setError(ctx, "selas_cuda_texture_filter: bad texture descriptor (mip chain outside dataSize?), filter, level or null buffers");
        return SELAS_ERR_ARGUMENT;
    }
    if (!n) return SELAS_OK;
    // InitializeEWAFilterWeights (TextureFiltering.cpp:14-21), with the C library's expf as in the reference
    float lut[128];
    for (int i = 0; i < 128; ++i) {
        const float alpha = 2.0f;
        const float r2 = float(i) / (128 - 1.0f);
        lut[i] = expf(-alpha * r2) - expf(-alpha);
    }
    Scratch s(ctx);
    uint8_t* d_tex;
    float *d_st, *d_d0 = nullptr, *d_d1 = nullptr, *d_out;
    SELAS_TRY(s.upload(&d_tex, static_cast<const uint8_t*>(t->texels), (size_t)t->dataSize));
    SELAS_TRY(s.upload(&d_st, st2, 2 * n));
    if (needsDerivatives) {
        SELAS_TRY(s.upload(&d_d0, dst0, 2 * n));
        SELAS_TRY(s.upload(&d_d1, dst1, 2 * n));
    }
    SELAS_TRY(s.alloc(&d_out, n * (t->format + 1u)));
    SELAS_TRY(launchTextureFilter(ctx, *t, d_tex, lut, filter, level, n, d_st, d_d0, d_d1, d_out));
    SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(out, d_out, n * (t->format + 1u) * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return SELAS_OK;
}

// ---- integrator ------------------------------------------------------------------------------------------------
int selas_cuda_render_pathtracer_device(SelasScene* scene, const SelasCamera* camera, const SelasRenderSettings* settings, float* d_film,
                                        SelasRenderStats* stats, void* stream) {
    if (!scene) return SELAS_ERR_ARGUMENT;
    SelasContext* ctx = scene->ctx;
    SELAS_TRY(bind(ctx));
    if (!camera || !settings || !d_film) {
        setError(ctx, "selas_cuda_render_pathtracer_device: null argument");
        return SELAS_ERR_ARGUMENT;
    }
    return renderPathTracer(scene, *camera, *settings, d_film, stats, stream ? (cudaStream_t)stream : ctx->stream);
}

int selas_cuda_render_collect_stats(SelasScene* scene, SelasRenderStats* stats) {
    if (!scene) return SELAS_ERR_ARGUMENT;
    SELAS_TRY(bind(scene->ctx));
    return collectRenderStats(scene, stats);
}

int selas_cuda_render_pathtracer(SelasScene* scene, const SelasCamera* camera, const SelasRenderSettings* settings,
                                 float* const* filmLayers, SelasRenderStats* stats) {
    if (!scene) return SELAS_ERR_ARGUMENT;
    SelasContext* ctx = scene->ctx;
    SELAS_TRY(bind(ctx));
    if (!camera || !settings || !filmLayers) {
        setError(ctx, "selas_cuda_render_pathtracer: null argument");
        return SELAS_ERR_ARGUMENT;
    }
    if (settings->layerCount < 1 || settings->layerCount > 2) {
        setError(ctx, "layerCount must be 1 or 2");
        return SELAS_ERR_ARGUMENT;
    }
    const size_t plane = (size_t)settings->width * settings->height * 3;
    if (scene->filmCacheFloats < plane * settings->layerCount) {
        cudaFree(scene->d_filmCache);
        scene->d_filmCache = nullptr;
        scene->filmCacheFloats = 0;
        SELAS_CUDA_CHECK(ctx, cudaMalloc((void**)&scene->d_filmCache, plane * settings->layerCount * sizeof(float)));
        scene->filmCacheFloats = plane * settings->layerCount;
    }
    float* d_film = scene->d_filmCache;
    SELAS_CUDA_CHECK(ctx, cudaMemsetAsync(d_film, 0, plane * settings->layerCount * sizeof(float), ctx->stream));
    // everything is enqueued before the one wait: render, the film copies, and (when asked for) the stats readback
    SELAS_TRY(renderPathTracer(scene, *camera, *settings, d_film, nullptr, ctx->stream));
    for (uint32_t l = 0; l < settings->layerCount; ++l)
        SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(filmLayers[l], d_film + plane * l, plane * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (stats) SELAS_TRY(collectRenderStats(scene, stats));
    else SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return SELAS_OK;
}

// ---- VCM integrator: VCM::GenerateImage, VCM.cpp:738-807 ---------------------------------------------------------------
int selas_cuda_render_vcm_device(SelasScene* scene, const SelasCamera* camera, const SelasVcmSettings* settings, float* d_film,
                                 SelasVcmStats* stats, void* stream) {
    if (!scene) return SELAS_ERR_ARGUMENT;
    SelasContext* ctx = scene->ctx;
    SELAS_TRY(bind(ctx));
    if (!camera || !settings || !d_film) {
        setError(ctx, "selas_cuda_render_vcm_device: null argument");
        return SELAS_ERR_ARGUMENT;
    }
    return renderVcm(scene, *camera, *settings, d_film, stats, stream ? (cudaStream_t)stream : ctx->stream);
}

int selas_cuda_render_vcm(SelasScene* scene, const SelasCamera* camera, const SelasVcmSettings* settings, float* film, SelasVcmStats* stats) {
    if (!scene) return SELAS_ERR_ARGUMENT;
    SelasContext* ctx = scene->ctx;
    SELAS_TRY(bind(ctx));
    if (!camera || !settings || !film) {
        setError(ctx, "selas_cuda_render_vcm: null argument");
        return SELAS_ERR_ARGUMENT;
    }
    const size_t plane = (size_t)settings->width * settings->height * 3;
    if (scene->filmCacheFloats < plane) {
        cudaFree(scene->d_filmCache);
        scene->d_filmCache = nullptr;
        scene->filmCacheFloats = 0;
        SELAS_CUDA_CHECK(ctx, cudaMalloc((void**)&scene->d_filmCache, plane * sizeof(float)));
        scene->filmCacheFloats = plane;
    }
    float* d_film = scene->d_filmCache;
    SELAS_CUDA_CHECK(ctx, cudaMemsetAsync(d_film, 0, plane * sizeof(float), ctx->stream));
    SELAS_TRY(renderVcm(scene, *camera, *settings, d_film, stats, ctx->stream));
    SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(film, d_film, plane * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return SELAS_OK;
}

int selas_cuda_vcm_light_pass(SelasScene* scene, const SelasCamera* camera, const SelasVcmSettings* settings, SelasVcmVertex* vertices,
                              size_t maxVertices, uint32_t* pathEnds, size_t* vertexCount) {
    if (!scene) return SELAS_ERR_ARGUMENT;
    SelasContext* ctx = scene->ctx;
    SELAS_TRY(bind(ctx));
    if (!camera || !settings) {
        setError(ctx, "selas_cuda_vcm_light_pass: null argument");
        return SELAS_ERR_ARGUMENT;
    }
    return vcmLightPassInspect(scene, *camera, *settings, vertices, maxVertices, pathEnds, vertexCount);
}

// ---- the film reduce over NCCL, bound at run time -----------------------------------------------------------------------
namespace {
typedef int (*NcclReduceFn)(const void*, void*, size_t, int, int, int, void*, cudaStream_t);
typedef const char* (*NcclErrorFn)(int);
struct NcclApi {
    void* handle = nullptr;
    NcclReduceFn reduce = nullptr;
    NcclErrorFn errorString = nullptr;
    bool tried = false;
};
NcclApi& ncclApi() {
    static NcclApi api;
    if (!api.tried) {
        api.tried = true;
        // RTLD_NOLOAD first: reuse the NCCL the process already has (e.g. the one torch.distributed loaded)
        api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!api.handle) api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) {
            api.reduce = (NcclReduceFn)dlsym(api.handle, "ncclReduce");
            api.errorString = (NcclErrorFn)dlsym(api.handle, "ncclGetErrorString");
        }
    }
    return api;
}
}  // namespace

int selas_cuda_reduce_film(SelasContext* ctx, void* ncclComm, float* d_accum, size_t count, int root, void* stream) {
    SELAS_TRY(bind(ctx));
    if (!ncclComm || (!d_accum && count)) {
        setError(ctx, "selas_cuda_reduce_film: null communicator or buffer");
        return SELAS_ERR_ARGUMENT;
    }
    NcclApi& api = ncclApi();
    if (!api.reduce) {
        setError(ctx, "selas_cuda_reduce_film: libnccl.so.2 (ncclReduce) is not available on this host");
        return SELAS_ERR_UNSUPPORTED;
    }
    const int ncclFloat32 = 7, ncclSum = 0;  // nccl.h: ncclDataType_t / ncclRedOp_t
    int rc = api.reduce(d_accum, d_accum, count, ncclFloat32, ncclSum, root, ncclComm, stream ? (cudaStream_t)stream : ctx->stream);
    if (rc != 0) {
        setError(ctx, "ncclReduce failed: %s", api.errorString ? api.errorString(rc) : "unknown NCCL error");
        return SELAS_ERR_CUDA;
    }
    return SELAS_OK;
}

int selas_cuda_film_scale_device(SelasContext* ctx, float* d_film, size_t floatCount, float scale, void* stream) {
    SELAS_TRY(bind(ctx));
    return launchScale(ctx, d_film, floatCount, scale, stream ? (cudaStream_t)stream : ctx->stream);
}

}  // extern "C"
