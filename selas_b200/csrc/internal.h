// internal.h -- host-side structures shared by the translation units of libselas_cuda.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <array>
#include <map>
#include <string>
#include <vector>

#include "../../include/selas_cuda.h"

namespace selas {

struct Bvh8;

// Device view of one mesh (= one Embree geometry: ModelGeometryUserData + its shared buffers,
// ModelResource.h:97-105, ModelResource.cpp:164-200).
// TextureResourceData level 0 on the device
struct TextureGpu {
    const float* texels;
    uint32_t width, height, format, pad;
};

struct MeshGpu {
    const float* positions;
    const uint32_t* indices;
    const float* normals;   // float3 or null
    const float* tangents;  // float4 or null
    const float* uvs;       // float2 or null
    const void* ptex;       // device PtexGpu* or null
    const TextureGpu* texture;  // base colour texture (non-Ptex materials) or null
    uint32_t faceCount;
    uint32_t indicesPerFace;  // 3 triangles, 4 quads, 1: a curve geometry (positions = float4 control vertices, indices = first vertex per segment)
    uint32_t materialIndex;
    uint32_t lightSetIndex;
};

struct InstanceGpu {
    float m[16];  // localToWorld, row-major row-vector
    uint32_t mesh;  // 0xFFFFFFFF: identity over all meshes
    uint32_t identity;
    uint32_t pad[2];
};

// ImageBasedLightResourceData on the device (ImageBasedLightResource.h:16-35)
struct IblGpu {
    const float* marginal;     // height
    const float* conditional;  // width*height
    const float* light;        // float3 per texel
    const float* miss;         // float3 per texel
    uint32_t width, height, missWidth, missHeight;
    float rotation, exposureScale;
    uint32_t present, pad;
};

// Two-level BVH (SELAS_SCENE_TWO_LEVEL_BVH): per flat instance, the world -> object transform the ray goes through when it enters the
// instance (SceneInstanceIntersectFunction, SceneResource.cpp:87-142: MatrixMultiplyPoint / MatrixMultiplyVector with worldToLocal) and
// where its mesh's BVH8 lives in the shared node / triangle arrays.
struct InstanceTrace {
    float w2l[12];  // row-vector affine: p_local = p_world * M, rows 0-2 linear part, row 3 translation
    uint32_t nodeBase, triBase, mesh, pad;
};

// Everything the shade kernel needs, passed by value.
struct SceneGpu {
    const float4* nodes;
    const float4* tris;
    uint32_t nodeCount, triCount;
    const InstanceTrace* traceInstances;  // null: one flattened world-space BVH8
    const float4* curvePool;              // four control vertices per curve segment of the tree (bvh8.cuh), or null: no curve geometry
    const uint4* instanceIds;             // SELAS_SCENE_INSTANCE_IDS: per flat instance {instID0, instID1, geomID, -} the public hits report
    const MeshGpu* meshes;
    const InstanceGpu* instances;
    const SelasMaterial* materials;
    // pow(material baseColor, 2.2) (SurfaceParameters.cpp:204,:211-212) of every material, made once at upload by the same powf the surface
    // stage would run per hit (materialBasePowKernel, shade.cu); used for meshes without a texture
    const float4* materialBasePow;
    const SelasLight* lights;
    const SelasLightSetRange* lightSets;
    uint32_t meshCount, instanceCount, materialCount, lightCount, lightSetCount;
    float background[4];
    IblGpu ibl;
};

struct DeviceBuffer {
    void* ptr = nullptr;
    size_t bytes = 0;
};

}  // namespace selas

struct SelasContext {
    int device = 0;
    cudaStream_t stream = nullptr;
    int smCount = 0;
    std::string lastError;
    // Ray-fetch counters of the persistent trace kernels: a ring of 256 words; every launch takes the next one and zeroes it on its own
    // stream, so launches in flight on different streams (or scenes) never share a counter.
    uint32_t* d_counters = nullptr;
    uint32_t counterNext = 0;
    // Staging of the host-buffer ray-cast entry points (selas_cuda_intersect / _occluded), kept between calls: no allocation per call,
    // two streams so that a chunk's download overlaps the next chunk's upload and trace.
    void* d_rayStage = nullptr;      // rays | hits | Ng | occluded for up to rayStageCapacity rays
    size_t rayStageCapacity = 0;
    cudaStream_t stageStream[2] = {nullptr, nullptr};
};

struct SelasScene {
    SelasContext* ctx = nullptr;
    selas::SceneGpu gpu{};
    std::vector<void*> allocations;  // everything cudaMalloc'ed for this scene
    SelasBvhStats stats{};
    bool usesPtex = false;  // any mesh carries a Ptex texture: selects the shade kernel built with the lookup
    bool ptexMixedRes = false;  // one of them has faces of different resolutions: the Ptex stage runs the instantiation with the reduction logic
    float boundsCenter[3] = {0.f, 0.f, 0.f};  // centre of the scene's box (ray binning quadrants)
    uint32_t topNodeCount = 0;  // leading BVH8 nodes (whole levels) the trace kernels stage in shared memory
    // SELAS_SCENE_INSTANCE_IDS: reported (instID0, instID1, geomID) -> flat instance index, for hits handed back to selas_cuda_surface_params
    std::map<std::array<uint32_t, 3>, uint32_t> instanceIdLookup;
    // wavefront scratch, grown on demand
    void* wf = nullptr;
    void* vcm = nullptr;  // VCM scratch (vcm.cu), grown on demand
    // device film of the host-buffer render entry point, kept between calls
    float* d_filmCache = nullptr;
    size_t filmCacheFloats = 0;
};

namespace selas {

// error plumbing
void setError(SelasContext* ctx, const char* fmt, ...);
#define SELAS_CUDA_CHECK(ctx, expr)                                                              \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            selas::setError(ctx, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return SELAS_ERR_CUDA;                                                               \
        }                                                                                        \
    } while (0)

// trace.cu
int buildBvh(SelasScene* scene, const std::vector<MeshGpu>& meshes, const std::vector<InstanceGpu>& hostInstances,
             const InstanceGpu* d_instances);
int buildBvhTwoLevel(SelasScene* scene, const std::vector<MeshGpu>& meshes, const std::vector<InstanceGpu>& hostInstances);
int launchIntersect(SelasScene* scene, uint32_t n, const SelasRay* d_rays, SelasHit* d_hits, float* d_ng4, uint32_t* d_counts,
                    uint32_t* d_rayCounter, cudaStream_t stream);
int launchOccluded(SelasScene* scene, uint32_t n, const SelasRay* d_rays, uint8_t* d_occ, uint32_t* d_counts, uint32_t* d_rayCounter,
                   cudaStream_t stream);
// shadow rays with film accumulation: payload = {contribution rgb, pixel}; film[pixel] += rgb when unoccluded
int launchShadowAccumulate(SelasScene* scene, const uint32_t* d_n, uint32_t maxN, const SelasRay* d_rays, const float4* d_payload,
                           float* d_film, uint32_t* d_rayCounter, unsigned long long* d_totals, cudaStream_t stream, int coop = -1, uint32_t expected = 0,
                           const uint32_t* d_order = nullptr);
int launchIntersectIndirect(SelasScene* scene, const uint32_t* d_n, uint32_t maxN, const SelasRay* d_rays, SelasHit* d_hits,
                            uint32_t* d_rayCounter, unsigned long long* d_totals, cudaStream_t stream, const uint32_t* d_order = nullptr, int coop = -1, uint32_t expected = 0);
// expected: an upper estimate of the queue length when the host has one (0 = unknown): the persistent grid is not launched wider than that many rays need
// coop: -1 = process default (SELAS_TRACE_COOP), 0 / 1 = traceKernel / traceCoopKernel (trace_coop.cuh) for this launch

// shade.cu
constexpr uint32_t kMaterialBinWords = 64;  // = kBinCount of the pre-shade counting sort
int launchMaterialOrder(SelasScene* scene, uint32_t n, const SelasHit* d_hits, uint32_t* d_bins, uint32_t* d_order, cudaStream_t stream);
int launchBsdfEval(SelasContext* ctx, size_t n, const SelasSurface* d_s, const float* d_v, const float* d_l, float* d_rgb, float* d_f,
                   float* d_r);
int launchBsdfSample(SelasContext* ctx, size_t n, const SelasSurface* d_s, const float* d_v, const float* d_u, SelasBsdfSample* d_out);
int launchLightSample(SelasContext* ctx, size_t n, const SelasLight* d_lights, uint32_t count, const float* d_pos, const float* d_nrm,
                      const float* d_u3, SelasLightSample* d_out);
int launchLightPdf(SelasContext* ctx, size_t n, const SelasLight* d_lights, uint32_t count, const uint32_t* d_idx, const float* d_pos,
                   const float* d_wi, float* d_pdf);
int launchCameraRays(SelasContext* ctx, const SelasCamera& cam, size_t n, const float* d_px, const float* d_py, const float* d_jitter,
                     SelasRay* d_rays);
int launchSurfaceParams(SelasScene* scene, size_t n, const SelasRay* d_rays, const SelasHit* d_hits, SelasSurface* d_out);
int renderPathTracer(SelasScene* scene, const SelasCamera& cam, const SelasRenderSettings& rs, float* d_film, SelasRenderStats* stats,
                     cudaStream_t stream);
int collectRenderStats(SelasScene* scene, SelasRenderStats* out);
int launchScale(SelasContext* ctx, float* d, size_t n, float s, cudaStream_t stream);
int launchIblLookup(SelasScene* scene, size_t n, const float* d_wi3, float* d_rgb3, float* d_pdf, float* d_miss3);
int launchIblSample(SelasScene* scene, size_t n, const float* d_u2, float* d_tpp3, uint32_t* d_xy2);
int launchMaterialBasePow(SelasContext* ctx, const SelasMaterial* d_materials, uint32_t count, float4* d_out);
int launchTextureSample(SelasContext* ctx, const TextureGpu& tex, size_t n, const float* d_uv2, int sRGB, float* d_rgb3);
int launchTextureFilter(SelasContext* ctx, const SelasTextureMipsDesc& desc, const void* d_texels, const float* ewaLut128, int filter, int level, size_t n,
                        const float* d_st2, const float* d_dst0, const float* d_dst1, float* d_out);
int launchCmj(SelasContext* ctx, size_t count, const int32_t* d_s, int32_t m, int32_t n, const int32_t* d_p, float* d_out2);
void freeWavefront(SelasScene* scene);

// hashgrid.cu
struct HashGridView;
HashGridView hashGridView(const SelasHashGrid* grid);
int hashGridReserve(SelasContext* ctx, SelasHashGrid** inout, size_t n, uint32_t cellCount);
uint32_t scanScratchWords(uint32_t n);
int launchScanU32(SelasContext* ctx, const uint32_t* d_in, uint32_t n, uint32_t* d_sums, uint32_t* d_inclusive, uint32_t* d_exclusive,
                  cudaStream_t stream);

// vcm.cu
int renderVcm(SelasScene* scene, const SelasCamera& cam, const SelasVcmSettings& vs, float* d_film, SelasVcmStats* stats, cudaStream_t stream);
int vcmLightPassInspect(SelasScene* scene, const SelasCamera& cam, const SelasVcmSettings& vs, SelasVcmVertex* vertices, size_t maxVertices,
                        uint32_t* pathEnds, size_t* vertexCount);
void freeVcm(SelasScene* scene);

// ptex.cu
const void* ptexDevicePointer(SelasPtex* tex);
bool ptexMixedResolutions(SelasPtex* tex);  // some face differs in resolution from the others: the lookup kernels need the reduction logic

}  // namespace selas
