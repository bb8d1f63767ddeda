/*
 * selas_cuda.h -- C ABI of the B200-native Selas path-tracing hot path.
 *
 * The reference (schuttejoe/Selas) has no plugin/FFI layer; the replaceable seams are ordinary C++
 * calls.  Each entry point below names the reference call it stands in for
 * (file:line relative to the reference tree).  Conventions:
 *   - every function returns int: 0 = ok, otherwise a SelasStatus; selas_cuda_last_error() gives the
 *     message (mirrors Error::Message(), Source/Core/SystemLib/Error.h:33).
 *   - handles are opaque; host pointers are borrowed for the call; device memory is owned by a handle.
 *   - plain "host" entry points copy host buffers to the GPU, run, copy back and synchronise;
 *     "_device" entry points take device pointers and a CUDA stream and do not synchronise: they
 *     only enqueue.  stream = NULL means the context's own stream (created non-blocking: it is NOT
 *     ordered against the legacy default stream -- pass your stream, or synchronise with
 *     selas_cuda_context_synchronize, when producers / consumers run elsewhere).  A render call that
 *     is handed a stats pointer has to wait for the device to fill it; pass NULL and collect later
 *     (selas_cuda_render_collect_stats) to stay asynchronous.
 *   - there is NO CPU fallback: if no CUDA device is usable every call fails with SELAS_ERR_CUDA.
 *   - vectors are row vectors, matrices row-major with translation in row 3
 *     (Source/Core/MathLib/FloatFuncs.cpp:383-396); tangent space is y-up
 *     (Source/Core/Shading/SurfaceParameters.cpp:217-219).
 */
#ifndef SELAS_CUDA_H
#define SELAS_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SELAS_INVALID_ID 0xFFFFFFFFu /* RTC_INVALID_GEOMETRY_ID, embree3/rtcore_common.h */

typedef enum SelasStatus {
    SELAS_OK            = 0,
    SELAS_ERR_CUDA      = 1, /* CUDA runtime error / no device */
    SELAS_ERR_ARGUMENT  = 2,
    SELAS_ERR_MEMORY    = 3,
    SELAS_ERR_FORMAT    = 4, /* malformed asset (ptex, material json, ...) */
    SELAS_ERR_UNSUPPORTED = 5
} SelasStatus;

/* ShaderType, Source/Core/SceneLib/ModelResource.h:27-33 */
enum { SELAS_SHADER_DISNEY_THIN = 0, SELAS_SHADER_DISNEY_SOLID = 1, SELAS_SHADER_DIRAC_TRANSPARENT = 2 };
/* MaterialFlags, ModelResource.h:35-42 */
enum { SELAS_MAT_TRANSPARENT = 1, SELAS_MAT_ALPHA_TESTED = 2, SELAS_MAT_DISPLACEMENT = 4,
       SELAS_MAT_INVERT_DISPLACEMENT = 8, SELAS_MAT_USES_PTEX = 16 };
/* ScalarMaterialProperties, ModelResource.h:44-64 */
enum { SELAS_METALLIC = 0, SELAS_SPECULAR_TINT, SELAS_ROUGHNESS, SELAS_ANISOTROPIC, SELAS_SHEEN,
       SELAS_SHEEN_TINT, SELAS_CLEARCOAT, SELAS_CLEARCOAT_GLOSS, SELAS_SPEC_TRANS, SELAS_DIFFUSE_TRANS,
       SELAS_FLATNESS, SELAS_IOR, SELAS_SCATTER_DISTANCE, SELAS_DISPLACEMENT, SELAS_SCALAR_COUNT };
/* SurfaceEventFlags / MediumPhaseFunction, Source/Core/Shading/Scattering.h:13-25 */
enum { SELAS_EVENT_SCATTER = 1, SELAS_EVENT_TRANSMISSION = 2, SELAS_EVENT_DIRAC = 4 };
enum { SELAS_MEDIUM_VACUUM = 0, SELAS_MEDIUM_ISOTROPIC = 1 };

/* 88-byte payload of MaterialResourceData (ModelResource.h:68-88, minus the 512-byte texture path). */
typedef struct SelasMaterial {
    uint32_t shader;
    uint32_t flags;
    float    baseColor[3];
    float    transmittanceColor[3];
    float    scalars[SELAS_SCALAR_COUNT];
} SelasMaterial;

/* SceneLight, Source/Core/SceneLib/SceneResource.h:32-40 (64 bytes). position is the quad centre. */
typedef struct SelasLight {
    uint32_t type;
    float    position[3];
    float    direction[3];
    float    x[3];
    float    z[3];
    float    radiance[3];
} SelasLight;

/* SceneLightSetRange, SceneResource.h:42-46 */
typedef struct SelasLightSetRange { uint32_t start, count; } SelasLightSetRange;

/* RayCastCameraSettings, Source/Core/GeometryLib/Camera.h:29-47 (104 bytes). */
typedef struct SelasCamera {
    float    cameraX[3];
    float    cameraY[3];
    float    cameraZ[3];
    float    position[3];
    float    forward[3];
    float    viewportWidth;
    float    viewportHeight;
    float    znear;
    float    zfar;
    uint32_t pad0;   /* the reference's `uint` is 64-bit (SystemLib/BasicTypes.h:47): width sits at offset 80 */
    uint64_t width;
    uint64_t height;
    float    aspect;
    float    virtualImagePlaneDistance;
} SelasCamera;

/* CameraSettings numeric part, Camera.h:17-26 */
typedef struct SelasCameraSettings {
    float position[3];
    float fovDegrees; /* horizontal */
    float lookAt[3];
    float znear;
    float up[3];
    float zfar;
} SelasCameraSettings;

/* SurfaceParameters, Source/Core/Shading/SurfaceParameters.h:18-49 (156 bytes), same field order. */
typedef struct SelasSurface {
    float    position[3];
    float    worldToTangent[9]; /* rows r0,r1,r2 */
    float    error;
    float    view[3];
    float    baseColor[3];
    float    transmittanceColor[3];
    float    sheen, sheenTint, clearcoat, clearcoatGloss, metallic, specTrans, diffTrans, flatness,
             anisotropic, relativeIOR, specularTint, roughness, scatterDistance, ior;
    uint32_t shader;
    uint32_t materialFlags;
    uint32_t lightSetIndex;
} SelasSurface;

/* BsdfSample, Source/Core/Shading/Scattering.h:34-43 (52 bytes) + the sampler-consumption count. */
typedef struct SelasBsdfSample {
    uint32_t flags;
    uint32_t mediumPhase;
    float    mediumExtinction[3];
    float    reflectance[3];
    float    wi[3];
    float    forwardPdfW;
    float    reversePdfW;
    uint32_t valid;      /* return value of SampleBsdfFunction (SurfaceScattering.cpp:18) */
    uint32_t randomsUsed;/* how many uniforms the call drew from its stream */
} SelasBsdfSample;

/* LightDirectSample, Source/Core/Shading/AreaLighting.h:47-54 */
typedef struct SelasLightSample {
    uint32_t index;
    float    direction[3];
    float    radiance[3];
    float    distance;
    float    pdfW;
} SelasLightSample;

/* Ray as the traversal kernels read it: two 16-byte vectors {o,tnear},{d,tfar}
 * (subset of RTCRay, embree3/rtcore_ray.h:26-41). */
typedef struct SelasRay {
    float ox, oy, oz, tnear;
    float dx, dy, dz, tfar;
} SelasRay;

/* Closest-hit record: two 16-byte vectors (RTCHit fields, rtcore_ray.h:44-55; Ng is optional, separate).
 * Miss: geomID == SELAS_INVALID_ID and t == ray.tfar. */
typedef struct SelasHit {
    float    t, u, v;
    uint32_t primID;
    uint32_t geomID;
    uint32_t instID0;
    uint32_t instID1;
    uint32_t pad;   /* 0 for triangles and quads; for a hit on a curve geometry the piece of the segment that was hit (csrc/curve.cuh), which
                       selas_cuda_surface_params needs to rebuild the surface normal: hand hits back as they were returned */
} SelasHit;

/* TextureResourceData (TextureLib/TextureResource.h:15-39), level 0 only: the path samples mip 0 with the bilinear "Triangle" filter
 * (ForceNoMips_, SurfaceParameters.cpp:21; TextureFiltering::Triangle(data, 0, uvs, ...), TextureFiltering.h:77-104, Repeat wrap). */
enum { SELAS_TEXTURE_FLOAT = 0, SELAS_TEXTURE_FLOAT2 = 1, SELAS_TEXTURE_FLOAT3 = 2, SELAS_TEXTURE_FLOAT4 = 3 }; /* TextureDataType */
typedef struct SelasTextureDesc {
    uint32_t     width, height;   /* mipWidths[0], mipHeights[0] */
    uint32_t     format;          /* SELAS_TEXTURE_*: floats per texel = format + 1 */
    uint32_t     reserved;
    const float* texels;          /* row-major, texel (s, t) at [t * width + s] */
} SelasTextureDesc;

/* One triangle/quad mesh = one Embree geometry (ModelResource.cpp:164-200, :236-297). */
typedef struct SelasMeshDesc {
    const float*    positions;      /* float3, stride 12 */
    uint32_t        vertexCount;
    const uint32_t* indices;        /* indicesPerFace per face */
    uint32_t        faceCount;
    uint32_t        indicesPerFace; /* 3 = RTC_GEOMETRY_TYPE_TRIANGLE, 4 = RTC_GEOMETRY_TYPE_QUAD */
    const float*    normals;        /* float3 per vertex or NULL (attribute slot 0) */
    const float*    tangents;       /* float4 per vertex or NULL (slot 1) */
    const float*    uvs;            /* float2 per vertex or NULL (slot 2) */
    uint32_t        materialIndex;
    uint32_t        lightSetIndex;
    void*           ptex;           /* SelasPtex* for SELAS_MAT_USES_PTEX materials, else NULL */
    const SelasTextureDesc* baseColorTexture; /* ModelGeometryUserData::baseColorTextureHandle for non-Ptex materials, or NULL: the
                                                 material's constant baseColor (SurfaceParameters.cpp:209-214); sampled at the
                                                 interpolated uvs (0,0 without a uv buffer), sRGB -> linear, then pow 2.2 */
} SelasMeshDesc;

/* One curve geometry = one Embree RTC_GEOMETRY_TYPE_ROUND_BSPLINE_CURVE (ModelResource.cpp:300-333): control vertices {x, y, z, radius}
 * (RTC_FORMAT_FLOAT4) and, per segment, the index of its first control vertex (four consecutive vertices make a uniform cubic B-spline
 * segment).  The surface is the sweep of a circle of the interpolated radius along the curve; hits report u = parameter along the
 * segment, v = 0, primID = segment, and Ng = the unnormalised surface normal (RTC_GEOMETRY_TYPE_CURVE.3embree3).  Curve geometries
 * are numbered after the meshes: geomID = meshCount + index (the reference attaches a model's curves after its meshes,
 * ModelResource.cpp:461-466), and SelasInstanceDesc.meshIndex may name one.  No attribute buffers, no textures (ModelResource.cpp:371-389). */
typedef struct SelasCurveDesc {
    const float*    vertices;      /* float4 per control vertex */
    uint32_t        vertexCount;
    const uint32_t* indices;       /* first control vertex of each segment; indices[i] + 3 < vertexCount */
    uint32_t        segmentCount;
    uint32_t        materialIndex;
    uint32_t        lightSetIndex;
} SelasCurveDesc;

/* Instance of a mesh; localToWorld is a row-major row-vector 4x4 (translation in row 3). */
typedef struct SelasInstanceDesc {
    uint32_t meshIndex;
    float    localToWorld[16];
    /* The ids a hit on this instance reports, read only when SelasSceneDesc.flags has SELAS_SCENE_INSTANCE_IDS: the reference's two
     * instancing levels (SceneResource.cpp:87-142, :362-378; SubsceneResource.cpp:186-195).  instID0 = index of the subscene instance in
     * SceneResourceData::subsceneInstances, instID1 = index of the model instance in SubsceneResourceData::modelInstances, geomID = index
     * of the mesh inside its model (ModelResource::userDatas), so that ModelDataFromRayIds(scene, hit.instId, hit.geomId, ...) of the
     * reference resolves a hit of this library unchanged.  localToWorld is then model.localToWorld * subscene.localToWorld.
     * Without the flag a hit reports instID0 = index of this record, instID1 = 0, geomID = meshIndex. */
    uint32_t instID0, instID1, geomID;
} SelasInstanceDesc;

/* ImageBasedLightResourceData, Source/Core/SceneLib/ImageBasedLightResource.h:16-35: lat-long environment with the marginal /
 * conditional CDFs the bake step computes (BuildCommon/BuildImageBasedLight.cpp:44-105) and the separate "miss" image shown to
 * paths that have only seen delta events (PathTracer.cpp:236-239). */
typedef struct SelasIblDesc {
    uint32_t     width, height;          /* of lightData and the density functions */
    const float* marginalDensityFunction;     /* height entries, inclusive CDF over rows */
    const float* conditionalDensityFunctions; /* width*height entries, inclusive CDF per row */
    const float* lightData;              /* float3 per texel */
    uint32_t     missWidth, missHeight;
    const float* missData;               /* float3 per texel */
    float        rotationRadians;
    float        exposureScale;          /* 2^exposure */
} SelasIblDesc;

typedef struct SelasSceneDesc {
    const SelasMeshDesc*      meshes;
    uint32_t                  meshCount;
    const SelasInstanceDesc*  instances;     /* NULL/0 => every mesh once, identity, instID0 = 0 */
    uint32_t                  instanceCount;
    const SelasMaterial*      materials;
    uint32_t                  materialCount;
    const SelasLight*         lights;
    uint32_t                  lightCount;
    const SelasLightSetRange* lightSets;
    uint32_t                  lightSetCount;
    float                     backgroundIntensity[4]; /* SceneResourceData::backgroundIntensity */
    const SelasIblDesc*       ibl;           /* SceneResource::iblResource, or NULL: constant background (AreaLighting.cpp:415-437) */
    uint32_t                  flags;         /* SELAS_SCENE_* */
    uint32_t                  curveCount;
    const SelasCurveDesc*     curves;        /* curve geometries, geomID = meshCount + index; NULL/0: none */
} SelasSceneDesc;
/* SelasSceneDesc.flags.  SELAS_SCENE_TWO_LEVEL_BVH: keep instanced geometry once -- one object-space BVH8 per mesh under a top-level BVH8
 * over the instances, rays transformed into object space on entry (what the reference gets from Embree instance geometries and its
 * subscene-instance callbacks, SubsceneResource.cpp:63-79, SceneResource.cpp:87-187) -- instead of flattening every instance into
 * world-space triangles.  Same hits and ids as the flattened form up to float rounding of the ray transform; without an instance list
 * one identity instance per mesh is implied (instID0 = mesh index). */
enum { SELAS_SCENE_TWO_LEVEL_BVH = 1, SELAS_SCENE_INSTANCE_IDS = 2 /* SelasInstanceDesc.instID0 / instID1 / geomID are what hits report */ };

/* Compile-time macros of the reference integrator made runtime (PathTracer.cpp:42-46, :248). */
/* SelasRenderSettings.flags */
enum { SELAS_RENDER_COUNT_TRAVERSAL = 1, /* use the counting build of the trace kernels and fill the node/triangle totals */
       SELAS_RENDER_SKIP_NONFINITE = 2   /* drop NaN/Inf radiance contributions instead of adding them to the film; off = reference behaviour
                                            (the reference checks for NaNs only in debug builds, BasicTypes.h:28) */ };

typedef struct SelasRenderSettings {
    uint32_t width, height;
    uint32_t samplesPerPixel; /* PathsPerPixel_ */
    uint32_t maxPathLength;   /* MaxBounceCount_ : loop bound bounceCount < maxPathLength */
    uint32_t rrStartBounce;   /* Russian roulette when bounceCount > rrStartBounce (reference: 8) */
    uint32_t layerCount;      /* 1 (radiance) or 2 (albedo-like Ld[0], radiance Ld[1]); LayerCount_ */
    uint32_t seed;
    uint32_t firstSample;     /* sample-pass shard: this call renders samples [firstSample, firstSample+samplesPerPixel) */
    uint32_t misMode;         /* 0 = reference (weight 1.0, PathTracer.cpp:215); 1 = balance heuristic */
    uint32_t flags;           /* SELAS_RENDER_* */
    uint32_t pixelStride;     /* 0/1 = every pixel; k = only pixels p with p % k == 0 (bounded samples of a workload) */
    uint32_t pixelSampler;    /* 0 = two uniforms from the path's stream (PathTracer, Camera.cpp:63-74); (m | n << 16) = correlated
                                 multi-jitter over an m x n grid with s = sample index, p = pixel index (the DeferredPathTracer's
                                 primary rays, DeferredPathTracer.cpp:349-380 / Camera.cpp:77-86) */
    uint32_t integrator;      /* SELAS_INTEGRATOR_*: 0 = PathTracer::EvaluatePath (PathTracer.cpp:142-258); 1 = the DeferredPathTracer's
                                 ShadeHitPosition (DeferredPathTracer.cpp:61-148): area-light NEE + sky/IBL NEE with balance-heuristic MIS,
                                 BSDF sampling MIS-weighted against the sky pdf, Russian roulette (u >= p terminates) once 7 bounces are
                                 tracked, no media, shadow rays from tnear 0, one layer; rrStartBounce / misMode are ignored and
                                 maxPathLength is only a safety bound (the reference has none) */
} SelasRenderSettings;
enum { SELAS_INTEGRATOR_PATHTRACER = 0, SELAS_INTEGRATOR_DEFERRED = 1 };

typedef struct SelasRenderStats {
    uint64_t closestRays;
    uint64_t occlusionRays;
    uint64_t pathSamples;
    uint64_t kernelLaunches;
    float    renderMs;   /* device time of the whole wavefront loop (CUDA events on the render stream) */
    float    generateMs; /* per-kernel device time, summed over launches (CUDA events around every launch) */
    float    closestMs;
    float    shadeMs;
    float    shadowMs;
    uint32_t closestLaunches, shadeLaunches, shadowLaunches;
    /* filled only with SELAS_RENDER_COUNT_TRAVERSAL: BVH8 nodes visited / triangles tested, summed over rays */
    uint64_t closestNodes, closestTris, shadowNodes, shadowTris;
} SelasRenderStats;

typedef struct SelasBvhStats {
    uint32_t triangleCount;  /* leaf primitives stored: triangles (quads count twice) and curve parts (8 per curve segment, csrc/curve.cuh) */
    uint32_t nodeCount;      /* 80-byte BVH8 nodes */
    uint32_t maxDepth;
    uint32_t reserved;
    float    buildMs;
    float    sahCost;
    uint64_t deviceBytes;
    uint32_t sahClusters;    /* binned-SAH build: Morton clusters under the SAH top tree (0 = radix tree, SELAS_BVH_BUILD=lbvh) */
    uint32_t sahGenerations; /* sahSplitKernel launches of the build */
} SelasBvhStats;

/* ---- VCM (vertex connection and merging) ---------------------------------------------------------------------------
 * Compile-time macros of the reference's VCM integrator made runtime (Applications/Selas/Source/VCM.cpp:28-35, :716-735,
 * :765-779).  One "iteration" = vmCount = width*height light paths, the hash grid over their vertices, then one camera
 * path per pixel (VertexConnectionAndMerging, VCM.cpp:664-697); vcCount = 1.  The reference only emits from the IBL
 * (VCMCommon.cpp:37-63), so the scene must have been created with one. */
enum { SELAS_VCM_NO_LIGHT_TRACING = 1, /* skip ConnectLightPathToCamera  (VCM.cpp:186-239) */
       SELAS_VCM_NO_DIRECT        = 2, /* skip ConnectCameraPathToLight  (VCM.cpp:242-292) */
       SELAS_VCM_NO_CONNECTIONS   = 4, /* skip ConnectPathVertices       (VCM.cpp:295-359) */
       SELAS_VCM_NO_MERGING       = 8, /* skip SearchHashGrid/MergeVertices (VCM.cpp:362-409) */
       SELAS_VCM_NO_SKY           = 16,/* skip ConnectToSkyLight         (VCM.cpp:168-183) */
       SELAS_VCM_SKIP_NONFINITE   = 32 /* drop NaN/Inf contributions instead of adding them to the film */ };

typedef struct SelasVcmSettings {
    uint32_t width, height;
    uint32_t iterations;      /* passes rendered by this call */
    uint32_t firstIteration;  /* iteration shard: this call renders passes [firstIteration, firstIteration + iterations);
                                 a pass's iterationIndex is pass + 1 (VCM.cpp:727-728) and its RNG streams are keyed by the pass */
    uint32_t maxPathLength;   /* MaxBounceCount_ (VCM.cpp:28: 10); at most 31 (PathStatePathLengthBitCount_, VCMCommon.h:17) */
    uint32_t seed;
    float    radiusFactor;    /* VcmRadiusFactor_ 0.0025: base radius = radiusFactor * boundingSphere.w (VCM.cpp:778) */
    float    radiusAlpha;     /* VcmRadiusAlpha_ 0.75 */
    float    boundingSphere[4]; /* SceneResource::boundingSphere (centre xyz, radius) */
    uint32_t flags;           /* SELAS_VCM_* */
    uint32_t reserved;
} SelasVcmSettings;

/* VCMVertex (VCMCommon.h:33-42) as the device stores it: HitParameters is kept as the ray + hit record it is recomputed
 * from (RayPick, VCM.cpp:123-166). */
typedef struct SelasVcmVertex {
    float    throughput[3];
    uint32_t indexAndLength;  /* index : 26, pathLength : 5 (VCMCommon.h:36-37), bit 31 = material is eTransparent */
    float    dVCM, dVC, dVM;
    float    continuationProbability; /* ContinuationProbability(lightSurface), SurfaceParameters.cpp:420-424 */
    SelasRay ray;
    SelasHit hit;
} SelasVcmVertex;

typedef struct SelasVcmStats {
    uint64_t lightPaths, cameraPaths;
    uint64_t lightVertices;       /* stored over all iterations */
    uint64_t closestRays, occlusionRays;
    uint64_t mergeCandidates;     /* hash-grid candidates distance-tested */
    uint64_t mergesAccepted;      /* MergeVertices calls (dist^2 <= r^2) */
    uint64_t connections;         /* ConnectPathVertices evaluations */
    uint64_t kernelLaunches;
    float    renderMs;            /* device time of all iterations */
    float    lightPassMs, gridBuildMs, cameraPassMs, mergeMs, connectMs;
    float    lastRadius;          /* vmSearchRadius of the last iteration */
} SelasVcmStats;

typedef struct SelasContext  SelasContext;
typedef struct SelasScene    SelasScene;
typedef struct SelasHashGrid SelasHashGrid;
typedef struct SelasPtex     SelasPtex;

/* ---- context --------------------------------------------------------------------------------- */
/* rtcNewDevice / rtcReleaseDevice at Source/Applications/Selas/Source/main.cpp:99,:130 */
int selas_cuda_context_create(int device, SelasContext** out);
int selas_cuda_context_destroy(SelasContext* ctx);
/* Waits for everything enqueued on the context's own stream (what _device calls use when their stream argument is NULL). */
int selas_cuda_context_synchronize(SelasContext* ctx);
const char* selas_cuda_last_error(const SelasContext* ctx); /* Error::Message(), SystemLib/Error.h:33; ctx may be NULL: thread-local message */
int selas_cuda_device_count(void);      /* no reference counterpart: the CPU renderer has one device */
const char* selas_cuda_version(void);

/* ---- scene: replaces the Embree scene construction ---------------------------------------------
 * rtcSetSharedGeometryBuffer/rtcCommitGeometry/rtcCommitScene at ModelResource.cpp:164-200,:461-468,
 * SubsceneResource.cpp:141-156, SceneResource.cpp:218-249.  Builds the compressed BVH8 on device. */
int selas_cuda_scene_create(SelasContext* ctx, const SelasSceneDesc* desc, SelasScene** out);
int selas_cuda_scene_destroy(SelasScene* scene); /* ShutdownSceneResource, SceneResource.cpp:318 (rtcReleaseScene) */
int selas_cuda_scene_bvh_stats(const SelasScene* scene, SelasBvhStats* out);

/* ---- ray cast: replaces rtcIntersect1 / rtcOccluded1 ---------------------------------------------
 * PathTracer.cpp:116 (RayPick) and :88 (OcclusionRay); packet form DeferredPathTracer.cpp:191,:267.
 * Closest hit in [tnear, tfar]; Ng (optional, float4 per ray, may be NULL) is the unnormalised
 * geometric normal cross(v1-v0, v2-v0).  occluded[i] = 1 when any hit exists (Embree: tfar = -inf). */
int selas_cuda_intersect(SelasScene* scene, size_t n, const SelasRay* rays, SelasHit* hits, float* Ng4);
int selas_cuda_occluded(SelasScene* scene, size_t n, const SelasRay* rays, uint8_t* occluded);
int selas_cuda_intersect_device(SelasScene* scene, size_t n, const SelasRay* d_rays, SelasHit* d_hits,
                                float* d_Ng4, void* stream);
int selas_cuda_occluded_device(SelasScene* scene, size_t n, const SelasRay* d_rays, uint8_t* d_occluded,
                               void* stream);
/* Per-ray traversal counters (nodes visited, triangles tested) from the counting build of the same
 * kernels; used for the algorithmic-bytes roofline.  counts: 2 uint32 per ray. */
int selas_cuda_intersect_counted(SelasScene* scene, size_t n, const SelasRay* rays, SelasHit* hits,
                                 uint32_t* counts, int anyHit);

/* ---- camera: InitializeRayCastCamera Camera.cpp:89-115; JitteredCameraRay Camera.cpp:63-74 ------- */
int selas_cuda_camera_init(const SelasCameraSettings* settings, uint32_t width, uint32_t height,
                           SelasCamera* out);
/* rays[i] from pixel (px[i] + jitter[2i], py[i] + jitter[2i+1]); jitter NULL => pixel centre +0.5 */
int selas_cuda_camera_rays(SelasContext* ctx, const SelasCamera* camera, size_t n, const float* px,
                           const float* py, const float* jitter, SelasRay* rays);

/* CorrelatedMultiJitter (MathLib/CorrelatedMultiJitter.cpp:52-62): out2[2i..] = sample s[i] of pattern p[i] on an m x n grid */
int selas_cuda_cmj(SelasContext* ctx, size_t count, const int32_t* s, int32_t m, int32_t n, const int32_t* p, float* out2);

/* ---- BSDF operators: SurfaceScattering.h:15-16 ----------------------------------------------------
 * EvaluateBsdf (SurfaceScattering.cpp:41): f*|cos| in rgb[3n], forward/reverse pdf.
 * SampleBsdfFunction (SurfaceScattering.cpp:18): u holds SELAS_BSDF_RANDOMS uniforms per sample,
 * consumed in the reference's order (lobe pick first; Disney.cpp:628). */
#define SELAS_BSDF_RANDOMS 8
int selas_cuda_bsdf_eval(SelasContext* ctx, size_t n, const SelasSurface* surfaces, const float* v3,
                         const float* l3, float* rgb3, float* forwardPdfW, float* reversePdfW);
int selas_cuda_bsdf_sample(SelasContext* ctx, size_t n, const SelasSurface* surfaces, const float* v3,
                           const float* u, SelasBsdfSample* out);

/* ---- lights: NextEventEstimation AreaLighting.cpp:353-373 (3 uniforms per sample) ----------------
 * and LightingPdf AreaLighting.cpp:376-390. */
int selas_cuda_light_sample(SelasContext* ctx, size_t n, const SelasLight* lights, uint32_t lightCount,
                            const float* position3, const float* normal3, const float* u3,
                            SelasLightSample* out);
int selas_cuda_light_pdf(SelasContext* ctx, size_t n, const SelasLight* lights, uint32_t lightCount,
                         const uint32_t* lightIndex, const float* position3, const float* wi3, float* pdf);

/* ---- image based light: ImageBasedLightResource.cpp:118-262 (the scene must have been created with an IBL) -------------
 * lookup: SampleIbl(ibl, wi, pdf) -> rgb3/pdf and SampleIblMiss(ibl, wi, pdf) -> missRgb3 (any output may be NULL);
 * sample: Ibl(ibl, r0, r1, theta, phi, x, y, pdf): u2 = 2 uniforms per sample; thetaPhiPdf3 = 3 floats, xy2 = 2 uint32 per sample */
int selas_cuda_ibl_lookup(SelasScene* scene, size_t n, const float* wi3, float* rgb3, float* pdf, float* missRgb3);
int selas_cuda_ibl_sample(SelasScene* scene, size_t n, const float* u2, float* thetaPhiPdf3, uint32_t* xy2);

/* ---- textures: SampleTextureFloat3(texture, uvs, sRGB, default) (SurfaceParameters.cpp:97-118) over a batch of uvs ------------------
 * = TextureFiltering::Triangle at level 0 (+ Math::SrgbToLinearPrecise when sRGB != 0), for parity checks of the lookup itself. */
int selas_cuda_texture_sample(SelasContext* ctx, const SelasTextureDesc* texture, size_t n, const float* uv2, int sRGB, float* rgb3);

/* ---- texture filters: TextureLib/TextureFiltering.h:54-234 over a batch of lookups -----------------------------------
 * The reference's path only ever calls Triangle at level 0 (above); Point, Triangle at any level, Trilinear and EWA are the rest of that
 * header, built here with the same arithmetic so that a caller that switches ForceNoMips_ off has them.  The texture is a
 * TextureResourceData with its whole mip chain: level l is mipWidths[l] x mipHeights[l] texels of (format + 1) floats at byte offset
 * mipOffsets[l] of `texels` (dataSize bytes).  Wrap mode Repeat, as hard-wired in the reference (including its unsigned modulo of
 * negative coordinates).  st2 = 2 floats per lookup; dst0_2 / dst1_2 = the two texture-space derivative vectors (TRILINEAR, EWA, EWA_LEVEL;
 * may be NULL for POINT and TRIANGLE); `level` is used by TRIANGLE and EWA_LEVEL (the per-level EWA the full EWA blends).  out =
 * (format + 1) floats per lookup.  EWA weights: the 128-entry table of InitializeEWAFilterWeights (TextureFiltering.cpp:14-21), computed on the host
 * with the C library's expf like the reference does.  (As written the reference never fills the table its EWA template reads -- the
 * table is a per-translation-unit static in the header and only TextureFiltering.cpp's own copy is initialised -- so its EWA would return
 * 0 / 0; the product uses the initialised table.) */
enum { SELAS_FILTER_POINT = 0, SELAS_FILTER_TRIANGLE = 1, SELAS_FILTER_TRILINEAR = 2, SELAS_FILTER_EWA = 3, SELAS_FILTER_EWA_LEVEL = 4 };
#define SELAS_TEXTURE_MAX_MIPS 16 /* TextureResourceData::MaxMipCount */
typedef struct SelasTextureMipsDesc {
    uint32_t     mipCount;
    uint32_t     format;          /* SELAS_TEXTURE_* */
    uint32_t     mipWidths[SELAS_TEXTURE_MAX_MIPS];
    uint32_t     mipHeights[SELAS_TEXTURE_MAX_MIPS];
    uint64_t     mipOffsets[SELAS_TEXTURE_MAX_MIPS];  /* bytes */
    uint64_t     dataSize;        /* bytes */
    const void*  texels;
} SelasTextureMipsDesc;
int selas_cuda_texture_filter(SelasContext* ctx, const SelasTextureMipsDesc* texture, int filter, int level, size_t n, const float* st2,
                              const float* dst0_2, const float* dst1_2, float* out);

/* ---- surface: CalculateSurfaceParams SurfaceParameters.cpp:144-247 over a batch of hits ---------- */
int selas_cuda_surface_params(SelasScene* scene, size_t n, const SelasRay* rays, const SelasHit* hits,
                              SelasSurface* out);

/* ---- integrator: PathTracer::GenerateImage PathTracer.cpp:301-350 --------------------------------
 * film[layer] is width*height float3, receives the SUM over this call's samples (not yet scaled by
 * 1/spp: FrameBuffer_Scale, Framebuffer.cpp:62-70, is the caller's / selas_cuda_film_scale's job). */
int selas_cuda_render_pathtracer(SelasScene* scene, const SelasCamera* camera,
                                 const SelasRenderSettings* settings, float* const* filmLayers,
                                 SelasRenderStats* stats);
/* Device-resident variant: d_film = layerCount consecutive width*height*3 float planes, accumulated into.
 * With stats == NULL the call only enqueues (queue lengths and launch shapes of every bounce stay on the device) and returns while
 * the GPU works; with stats != NULL it waits and fills them (= render, then selas_cuda_render_collect_stats). */
int selas_cuda_render_pathtracer_device(SelasScene* scene, const SelasCamera* camera,
                                        const SelasRenderSettings* settings, float* d_film,
                                        SelasRenderStats* stats, void* stream);
/* Waits for the scene's enqueued renders and returns what was measured since the last collect, summed over those calls: per-kernel
 * device times (CUDA events recorded around every launch), ray totals (summed on the device), path samples, launches; renderMs spans
 * from the first launch to the last.  The reference has no counterpart (its integrators print a wall-clock time, PathTracer.cpp:344). */
int selas_cuda_render_collect_stats(SelasScene* scene, SelasRenderStats* stats);
/* The path's only collective: ncclReduce(accum, accum, count, ncclFloat32, ncclSum, root, comm, stream) of the
 * per-GPU accumulation buffers after the last pass.  `ncclComm` is the caller's ncclComm_t (one process per GPU); NCCL is bound
 * at run time (dlopen of libnccl.so.2), so the library loads on hosts without it and this call then fails with
 * SELAS_ERR_UNSUPPORTED.  Does not synchronise. */
int selas_cuda_reduce_film(SelasContext* ctx, void* ncclComm, float* d_accum, size_t count, int root, void* stream);
/* FrameBuffer_Scale Framebuffer.cpp:62-70 on a device film. */
int selas_cuda_film_scale_device(SelasContext* ctx, float* d_film, size_t floatCount, float scale, void* stream);

/* ---- VCM integrator: VCM::GenerateImage VCM.cpp:738-807 (iterations instead of IntegrationSeconds_) --------------------
 * film: width*height float3, receives the SUM over this call's iterations (FrameBuffer_Normalize by 1/iterations,
 * VCM.cpp:801, is the caller's job: selas_cuda_film_scale_device).  Iterations are independent given the scene, so N GPUs
 * render disjoint [firstIteration, ...) shards and reduce the films. */
int selas_cuda_render_vcm(SelasScene* scene, const SelasCamera* camera, const SelasVcmSettings* settings, float* film,
                          SelasVcmStats* stats);
int selas_cuda_render_vcm_device(SelasScene* scene, const SelasCamera* camera, const SelasVcmSettings* settings, float* d_film,
                                 SelasVcmStats* stats, void* stream);
/* Light pass of iteration `settings->firstIteration` only (GenerateLightSamples + EvaluateLightPaths, VCM.cpp:492-561), for
 * inspection: lightVertices in the reference's order (path by path), pathEnds[vmCount] (LightPathSet, VCM.cpp:57-62).
 * vertices may be NULL (count only); at most maxVertices are written; *vertexCount receives the total. */
int selas_cuda_vcm_light_pass(SelasScene* scene, const SelasCamera* camera, const SelasVcmSettings* settings,
                              SelasVcmVertex* vertices, size_t maxVertices, uint32_t* pathEnds, size_t* vertexCount);
/* CalculateIterationConstants, VCMCommon.cpp:97-112: out5 = {vmSearchRadius, vmSearchRadiusSqr, vmNormalization, vmWeight, vcWeight} */
int selas_cuda_vcm_iteration_constants(uint32_t vmCount, uint32_t vcCount, float baseRadius, float radiusAlpha, float iterationIndex,
                                       float* out5);

/* ---- VCM hash grid: BuildHashGrid VCMHashGrid.cpp:48-97, SearchHashGrid :107-155 ------------------
 * positions: float3 per vertex.  Query returns, per query point, the indices of vertices with
 * dist^2 <= r^2 in the 2x2x2 neighbourhood, in the reference's visiting order. */
int selas_cuda_hashgrid_build(SelasContext* ctx, size_t n, const float* positions3, float radius,
                              uint32_t cellCount, SelasHashGrid** out);
int selas_cuda_hashgrid_destroy(SelasHashGrid* grid);
/* copies cellRangeEnds[cellCount] and cellIndices[n] (reference layout) to the host */
int selas_cuda_hashgrid_read(const SelasHashGrid* grid, uint32_t* cellRangeEnds, uint32_t* cellIndices,
                             float* aabbMinMax6);
/* counts[m]; if indices != NULL, up to maxPerQuery indices per query are written at indices[q*maxPerQuery..] */
int selas_cuda_hashgrid_query(const SelasHashGrid* grid, size_t m, const float* queries3,
                              uint32_t* counts, uint32_t* indices, uint32_t maxPerQuery);
int selas_cuda_hashgrid_build_device(SelasContext* ctx, size_t n, const float* d_positions3, float radius,
                                     uint32_t cellCount, SelasHashGrid** inout, void* stream);
int selas_cuda_hashgrid_query_device(const SelasHashGrid* grid, size_t m, const float* d_queries3,
                                     uint32_t* d_counts, uint32_t* d_indices, uint32_t maxPerQuery,
                                     void* stream);

/* ---- Ptex: PtexFilter::eval(f_bspline) as called at SurfaceParameters.cpp:203 ---------------------
 * (Ptexture.h:1010-1012; widths 0, width 1, blur 0).  rgb3: 3 floats per lookup. */
int selas_cuda_ptex_load(SelasContext* ctx, const char* path, SelasPtex** out);
int selas_cuda_ptex_load_memory(SelasContext* ctx, const void* data, size_t size, SelasPtex** out);
int selas_cuda_ptex_destroy(SelasPtex* tex);
int selas_cuda_ptex_eval(SelasPtex* tex, size_t n, const int32_t* faceid, const float* u, const float* v,
                         float* rgb3);
int selas_cuda_ptex_eval_device(SelasPtex* tex, size_t n, const int32_t* d_faceid, const float* d_u,
                                const float* d_v, float* d_rgb3, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SELAS_CUDA_H */
