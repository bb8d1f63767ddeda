// shade.cu -- shading-side kernels: the batch BSDF / light / camera / surface operators used for parity, and
// the wavefront PathTracer (generate -> [trace -> shade -> shadow-trace+accumulate]*).
//
// This translation unit is compiled with -fmad=false -ftz=true -prec-div=true -prec-sqrt=true so that its
// fp32 arithmetic rounds like the reference's SSE code (FTZ/DAZ on, main.cpp:84-85); only libm calls differ.
//
// Wavefront = device restatement of PathTracer::EvaluatePath (Applications/Selas/Source/PathTracer.cpp:142-258)
// as a state machine over SoA queues:
//   rays      32 B  {o,tnear,d,tfar}        hits   32 B  {t,u,v,prim,geom,inst0,inst1,-}
//   thrPix    16 B  {throughput rgb, pixel} rngFlags 8 B {pcg state, bounce | deltaOnly<<16 | inMedium<<17}
//   medium    16 B  {extinction rgb, -}  (touched only by paths inside a medium)
//   shadow    32 B ray + 16 B {contribution rgb, pixel}
// Paths that survive a bounce are appended to the next queue with one atomicAdd per warp (ballot-ranked),
// so every launch works on a dense queue.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "disney.cuh"
#include "internal.h"
#include "lights.cuh"
#include "surface.cuh"

namespace selas {

namespace {
inline uint32_t divUp(size_t a, uint32_t b) { return (uint32_t)((a + b - 1) / b); }
}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// batch operators (parity levels (b): SURVEY.md 8)
// ---------------------------------------------------------------------------------------------------------------
__global__ void bsdfEvalKernel(size_t n, const SelasSurface* __restrict__ surfaces, const float* __restrict__ v3in,
                               const float* __restrict__ l3in, float* __restrict__ rgb, float* __restrict__ fwd, float* __restrict__ rev) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Surface s;
    loadSurface(surfaces[i], s);
    DisneyFrame f;
    makeDisneyFrame(s, s.shader == SELAS_SHADER_DISNEY_THIN, f);
    float fp = 0.f, rp = 0.f;
    V3 c = evaluateBsdf(s, f, v3(v3in + 3 * i), v3(l3in + 3 * i), fp, rp);
    store3(rgb + 3 * i, c);
    fwd[i] = fp;
    rev[i] = rp;
}

__global__ void bsdfSampleKernel(size_t n, const SelasSurface* __restrict__ surfaces, const float* __restrict__ v3in,
                                 const float* __restrict__ u, SelasBsdfSample* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Surface s;
    loadSurface(surfaces[i], s);
    DisneyFrame f;
    makeDisneyFrame(s, s.shader == SELAS_SHADER_DISNEY_THIN, f);
    ArrayRng rng{u + SELAS_BSDF_RANDOMS * i, 0};
    BsdfResult r;
    bool ok = sampleBsdf(rng, s, f, v3(v3in + 3 * i), r);
    SelasBsdfSample o;
    o.flags = r.flags;
    o.mediumPhase = r.mediumPhase;
    store3(o.mediumExtinction, r.mediumExtinction);
    store3(o.reflectance, r.reflectance);
    store3(o.wi, r.wi);
    o.forwardPdfW = r.forwardPdfW;
    o.reversePdfW = r.reversePdfW;
    o.valid = ok ? 1u : 0u;
    o.randomsUsed = rng.pos;
    out[i] = o;
}

__global__ void lightSampleKernel(size_t n, const SelasLight* __restrict__ lights, uint32_t count, const float* __restrict__ pos,
                                  const float* __restrict__ nrm, const float* __restrict__ u3, SelasLightSample* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ArrayRng rng{u3 + 3 * i, 0};
    LightSample ls;
    nextEventEstimation(rng, lights, count, v3(pos + 3 * i), v3(nrm + 3 * i), ls);
    SelasLightSample o;
    o.index = ls.index;
    store3(o.direction, ls.direction);
    store3(o.radiance, ls.radiance);
    o.distance = ls.distance;
    o.pdfW = ls.pdfW;
    out[i] = o;
}

__global__ void lightPdfKernel(size_t n, const SelasLight* __restrict__ lights, uint32_t count, const uint32_t* __restrict__ idx,
                               const float* __restrict__ pos, const float* __restrict__ wi, float* __restrict__ pdf) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    pdf[i] = lightingPdf(lights, count, idx[i], v3(pos + 3 * i), v3(wi + 3 * i));
}

__global__ void cameraRaysKernel(SelasCamera cam, size_t n, const float* __restrict__ px, const float* __restrict__ py,
                                 const float* __restrict__ jitter, SelasRay* __restrict__ rays) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float jx = jitter ? jitter[2 * i] : 0.5f, jy = jitter ? jitter[2 * i + 1] : 0.5f;
    V3 d = imageToWorldDirection(cam, px[i] + jx, py[i] + jy);
    float4* o = reinterpret_cast<float4*>(rays + i);
    o[0] = make_float4(cam.position[0], cam.position[1], cam.position[2], 0.0f);
    o[1] = make_float4(d.x, d.y, d.z, SELAS_FLOAT_MAX);
}

__global__ void iblLookupKernel(IblGpu ibl, size_t n, const float* __restrict__ wi3, float* __restrict__ rgb3, float* __restrict__ pdf,
                                float* __restrict__ miss3) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    V3 wi = v3(wi3 + 3 * i);
    if (rgb3) store3(rgb3 + 3 * i, iblLookup(ibl, wi));
    if (pdf) pdf[i] = iblPdf(ibl, wi);
    if (miss3) store3(miss3 + 3 * i, iblLookupMiss(ibl, wi));
}
__global__ void iblSampleKernel(IblGpu ibl, size_t n, const float* __restrict__ u2, float* __restrict__ tpp3, uint32_t* __restrict__ xy2) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float theta, phi, pdf;
    uint32_t x, y;
    iblSample(ibl, u2[2 * i], u2[2 * i + 1], theta, phi, x, y, pdf);
    tpp3[3 * i] = theta; tpp3[3 * i + 1] = phi; tpp3[3 * i + 2] = pdf;
    xy2[2 * i] = x; xy2[2 * i + 1] = y;
}

__global__ void cmjKernel(size_t count, const int32_t* __restrict__ sIdx, int32_t m, int32_t n, const int32_t* __restrict__ p, float* __restrict__ out2) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    float jx, jy;
    correlatedMultiJitter(sIdx[i], m, n, p[i], jx, jy);
    out2[2 * i] = jx;
    out2[2 * i + 1] = jy;
}

__global__ void __launch_bounds__(128) surfaceParamsKernel(SceneGpu sc, size_t n, const SelasRay* __restrict__ rays,
                                                           const SelasHit* __restrict__ hits, const float* __restrict__ ptexRgb,
                                                           SelasSurface* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    SelasHit h = hits[i];
    SelasSurface o;
    memset(&o, 0, sizeof(o));
    if (h.geomID != SELAS_INVALID_ID) {
        SelasRay r = rays[i];
        Surface s;
        surfaceFromHit(sc, v3(r.ox, r.oy, r.oz), v3(r.dx, r.dy, r.dz), h.t, h.u, h.v, h.primID, h.geomID, h.instID0,
                       ptexRgb ? ptexRgb + 3 * i : nullptr, s, h.pad);
        storeSurface(s, v3(-r.dx, -r.dy, -r.dz), o);
    }
    out[i] = o;
}

__global__ void textureSampleKernel(TextureGpu tex, size_t n, const float* __restrict__ uv2, int sRGB, float* __restrict__ rgb3) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    store3(rgb3 + 3 * i, sampleTextureFloat3(tex, uv2[2 * i], uv2[2 * i + 1], sRGB != 0));
}

__global__ void materialBasePowKernel(const SelasMaterial* __restrict__ materials, uint32_t n, float4* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const V3 base = v3(materials[i].baseColor);
    out[i] = make_float4(powf(base.x, 2.2f), powf(base.y, 2.2f), powf(base.z, 2.2f), 0.f);
}

__global__ void scaleKernel(float* __restrict__ d, size_t n, float s) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] *= s;
}

int launchBsdfEval(SelasContext* ctx, size_t n, const SelasSurface* d_s, const float* d_v, const float* d_l, float* d_rgb, float* d_f, float* d_r) {
    if (!n) return SELAS_OK;
    bsdfEvalKernel<<<divUp(n, 128), 128, 0, ctx->stream>>>(n, d_s, d_v, d_l, d_rgb, d_f, d_r);
    SELAS_CUDA_CHECK(ctx, cudaGetLastError());
    return SELAS_OK;
}
int launchBsdfSample(SelasContext* ctx, size_t n, const SelasSurface* d_s, const float* d_v, const float* d_u, SelasBsdfSample* d_out) {
    if (!n) return SELAS_OK;
    bsdfSampleKernel<<<divUp(n, 128), 128, 0, ctx->stream>>>(n, d_s, d_v, d_u, d_out);
    SELAS_CUDA_CHECK(ctx, cudaGetLastError());
    return SELAS_OK;
}
int launchLightSample(SelasContext* ctx, size_t n, const SelasLight* d_lights, uint32_t count, const float* d_pos, const float* d_nrm,
                      const float* d_u3, SelasLightSample* d_out) {
    if (!n) return SELAS_OK;
    lightSampleKernel<<<divUp(n, 128), 128, 0, ctx->stream>>>(n, d_lights, count, d_pos, d_nrm, d_u3, d_out);
    SELAS_CUDA_CHECK(ctx, cudaGetLastError());
    return SELAS_OK;
}
int launchLightPdf(SelasContext* ctx, size_t n, const SelasLight* d_lights, uint32_t count, const uint32_t* d_idx, const float* d_pos,
                   const float* d_wi, float* d_pdf) {
    if (!n) return SELAS_OK;
    lightPdfKernel<<<divUp(n, 128), 128, 0, ctx->stream>>>(n, d_lights, count, d_idx, d_pos, d_wi, d_pdf);
    SELAS_CUDA_CHECK(ctx, cudaGetLastError());
    return SELAS_OK;
}
int launchCameraRays(SelasContext* ctx, const SelasCamera& cam, size_t n, const float* d_px, const float* d_py, const float* d_jitter,
                     SelasRay* d_rays) {
    if (!n) return SELAS_OK;
    cameraRaysKernel<<<divUp(n, 256), 256, 0, ctx->stream>>>(cam, n, d_px, d_py, d_jitter, d_rays);
    SELAS_CUDA_CHECK(ctx, cudaGetLastError());
    return SELAS_OK;
}
int launchSurfaceParams(SelasScene* scene, size_t n, const SelasRay* d_rays, const SelasHit* d_hits, SelasSurface* d_out) {
    if (!n) return SELAS_OK;
    float* d_rgb = nullptr;
    if (scene->usesPtex) {
        SELAS_CUDA_CHECK(scene->ctx, cudaMalloc((void**)&d_rgb, sizeof(float) * 3 * n));
        unsigned grid = (unsigned)std::min<size_t>((n + 1023) / 1024, (size_t)scene->ctx->smCount * 8);
        if (scene->ptexMixedRes) ptexStageKernel<true><<<grid, 128, 0, scene->ctx->stream>>>(scene->gpu, (uint32_t)n, d_hits, d_rgb);
        else ptexStageKernel<false><<<grid, 128, 0, scene->ctx->stream>>>(scene->gpu, (uint32_t)n, d_hits, d_rgb);
    }
    surfaceParamsKernel<<<divUp(n, 128), 128, 0, scene->ctx->stream>>>(scene->gpu, n, d_rays, d_hits, d_rgb, d_out);
    if (d_rgb) {
        cudaStreamSynchronize(scene->ctx->stream);
        cudaFree(d_rgb);
    }
    SELAS_CUDA_CHECK(scene->ctx, cudaGetLastError());
    return SELAS_OK;
}
int launchIblLookup(SelasScene* scene, size_t n, const float* d_wi3, float* d_rgb3, float* d_pdf, float* d_miss3) {
    if (!n) return SELAS_OK;
    iblLookupKernel<<<divUp(n, 128), 128, 0, scene->ctx->stream>>>(scene->gpu.ibl, n, d_wi3, d_rgb3, d_pdf, d_miss3);
    SELAS_CUDA_CHECK(scene->ctx, cudaGetLastError());
    return SELAS_OK;
}
int launchIblSample(SelasScene* scene, size_t n, const float* d_u2, float* d_tpp3, uint32_t* d_xy2) {
    if (!n) return SELAS_OK;
    iblSampleKernel<<<divUp(n, 128), 128, 0, scene->ctx->stream>>>(scene->gpu.ibl, n, d_u2, d_tpp3, d_xy2);
    SELAS_CUDA_CHECK(scene->ctx, cudaGetLastError());
    return SELAS_OK;
}
int launchMaterialBasePow(SelasContext* ctx, const SelasMaterial* d_materials, uint32_t count, float4* d_out) {
    if (!count) return SELAS_OK;
    materialBasePowKernel<<<divUp(count, 128), 128, 0, ctx->stream>>>(d_materials, count, d_out);
    SELAS_CUDA_CHECK(ctx, cudaGetLastError());
    return SELAS_OK;
}
int launchTextureSample(SelasContext* ctx, const TextureGpu& tex, size_t n, const float* d_uv2, int sRGB, float* d_rgb3) {
    if (!n) return SELAS_OK;
    textureSampleKernel<<<divUp(n, 256), 256, 0, ctx->stream>>>(tex, n, d_uv2, sRGB, d_rgb3);
    SELAS_CUDA_CHECK(ctx, cudaGetLastError());
    return SELAS_OK;
}
int launchCmj(SelasContext* ctx, size_t count, const int32_t* d_s, int32_t m, int32_t n, const int32_t* d_p, float* d_out2) {
    if (!count) return SELAS_OK;
    cmjKernel<<<divUp(count, 256), 256, 0, ctx->stream>>>(count, d_s, m, n, d_p, d_out2);
    SELAS_CUDA_CHECK(ctx, cudaGetLastError());
    return SELAS_OK;
}
int launchScale(SelasContext* ctx, float* d, size_t n, float s, cudaStream_t stream) {
    if (!n) return SELAS_OK;
    scaleKernel<<<divUp(n, 256), 256, 0, stream>>>(d, n, s);
    SELAS_CUDA_CHECK(ctx, cudaGetLastError());
    return SELAS_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// wavefront path tracer
// ---------------------------------------------------------------------------------------------------------------
constexpr uint32_t kBinCount = 64;  // material bins of the pre-shade counting sort (63 materials + miss)
constexpr uint32_t kCounterStride = 8;  // words per bounce in WavefrontBuffers::counters
constexpr uint32_t kEstimateBounces = 64;  // bounces whose queue lengths are read back for the shade launches' grid estimates
#define FLAG_DELTA_ONLY (1u << 16)
#define FLAG_IN_MEDIUM (1u << 17)

// CUDA events around every launch of the loop: per-kernel device time for the roofline report.
struct KernelTimers {
    std::vector<cudaEvent_t> pool;
    struct Span { int kind; size_t a, b; };
    std::vector<Span> spans;
    size_t used = 0;
    cudaEvent_t get() {
        if (used == pool.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            pool.push_back(e);
        }
        return pool[used++];
    }
    size_t begin(cudaStream_t st) {
        cudaEvent_t e = get();
        cudaEventRecord(e, st);
        return used - 1;
    }
    void end(int kind, size_t a, cudaStream_t st) {
        cudaEvent_t e = get();
        cudaEventRecord(e, st);
        spans.push_back({kind, a, used - 1});
    }
    void collect(SelasRenderStats& st) {
        for (const Span& sp : spans) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, pool[sp.a], pool[sp.b]);
            if (sp.kind == 0) st.generateMs += ms;
            else if (sp.kind == 1) { st.closestMs += ms; ++st.closestLaunches; }
            else if (sp.kind == 2) { st.shadeMs += ms; ++st.shadeLaunches; }
            else { st.shadowMs += ms; ++st.shadowLaunches; }
        }
        spans.clear();
        used = 0;
    }
    ~KernelTimers() {
        for (cudaEvent_t e : pool) cudaEventDestroy(e);
    }
};

struct WavefrontBuffers {
    uint32_t capacity = 0;
    SelasRay* rays[2] = {nullptr, nullptr};
    SelasHit* hits = nullptr;
    float4* thrPix[2] = {nullptr, nullptr};
    uint2* rngFlags[2] = {nullptr, nullptr};
    float4* medium[2] = {nullptr, nullptr};
    SelasRay* shadowRays = nullptr;
    float4* shadowPayload = nullptr;
    float* ptexRgb = nullptr;      // Ptex stage output, 3 floats per path (scenes with Ptex materials only)
    uint32_t* order = nullptr;     // material-sorted visiting order of the hit queue
    uint32_t* rayOrder = nullptr;  // direction-sorted visiting order of the ray queue (SELAS_RAY_SORT)
    uint32_t* rayBins = nullptr;   // per iteration: 64 counts + 64 cursors
    uint32_t* bins = nullptr;      // per iteration: 64 counts + 64 cursors
    // kCounterStride per iteration: [0] queue length of this iteration, [1] shadow rays made by the previous one, [2] closest-hit fetch
    // counter, [3] shadow fetch counter, [4] shade chunk counter
    uint32_t* counters = nullptr;
    uint32_t counterSlots = 0;
    uint32_t shadowPerPath = 1;    // shadow-ray slots per queued path (2 for the deferred integrator: area light + sky)
    bool hasPtex = false;
    // totals since the last collect: closest rays, occlusion rays, then the counting build's nodes / triangles (closest, shadow)
    unsigned long long* totals = nullptr;
    unsigned long long* h_totals = nullptr;  // pinned readback
    uint32_t* h_counts = nullptr;            // pinned readback of one queue length (long paths: periodic early-out check)
    // Queue lengths of an earlier pass, read back asynchronously at its end: the shade launches of later passes of the same shape size
    // their one-shot grids with them (never waited for: used when the copy has arrived, else the persistent launch shape is taken).
    uint32_t* h_passCounts = nullptr;        // pinned, kEstimateBounces * kCounterStride words
    cudaEvent_t evPassCounts = nullptr;
    bool passCountsPending = false;
    std::vector<uint32_t> estimate;          // per bounce, 0 = unknown
    uint32_t estimateCount = 0;              // pass size the estimates belong to
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool spanOpen = false;         // ev0 recorded at the first launch since the last collect
    uint64_t pendingSamples = 0, pendingLaunches = 0;
    cudaStream_t lastStream = nullptr;
    KernelTimers timers;  // event pool reused across calls
};

static void destroyWavefront(WavefrontBuffers* w) {
    if (!w) return;
    for (int k = 0; k < 2; ++k) {
        cudaFree(w->rays[k]);
        cudaFree(w->thrPix[k]);
        cudaFree(w->rngFlags[k]);
        cudaFree(w->medium[k]);
    }
    cudaFree(w->hits);
    cudaFree(w->shadowRays);
    cudaFree(w->shadowPayload);
    cudaFree(w->counters);
    cudaFree(w->order);
    cudaFree(w->rayOrder);
    cudaFree(w->rayBins);
    cudaFree(w->ptexRgb);
    cudaFree(w->bins);
    cudaFree(w->totals);
    if (w->h_totals) cudaFreeHost(w->h_totals);
    if (w->h_counts) cudaFreeHost(w->h_counts);
    if (w->h_passCounts) cudaFreeHost(w->h_passCounts);
    if (w->evPassCounts) cudaEventDestroy(w->evPassCounts);
    if (w->ev0) cudaEventDestroy(w->ev0);
    if (w->ev1) cudaEventDestroy(w->ev1);
    delete w;
}

void freeWavefront(SelasScene* scene) {
    destroyWavefront((WavefrontBuffers*)scene->wf);
    scene->wf = nullptr;
}

// The buffer set is built aside and published to the scene only when every allocation succeeded: a failed cudaMalloc (8 M paths in
// flight is 1.2 GB) must not leave a half-null set behind that the next call would take for "already large enough".
static int ensureWavefront(SelasScene* scene, uint32_t capacity, uint32_t iterations, uint32_t shadowPerPath, cudaStream_t stream) {
    SelasContext* ctx = scene->ctx;
    WavefrontBuffers* old = (WavefrontBuffers*)scene->wf;
    const uint32_t slots = (iterations + 2) * kCounterStride;
    if (old && old->capacity >= capacity && old->counterSlots >= slots && old->shadowPerPath >= shadowPerPath && old->hasPtex == scene->usesPtex)
        return SELAS_OK;
    if (old) {  // work of earlier calls may still be running on the old buffers
        if (old->lastStream) cudaStreamSynchronize(old->lastStream);
        capacity = std::max(capacity, old->capacity);
        shadowPerPath = std::max(shadowPerPath, old->shadowPerPath);
    }
    WavefrontBuffers* w = new (std::nothrow) WavefrontBuffers();
    if (!w) return SELAS_ERR_MEMORY;
    w->capacity = capacity;
    w->counterSlots = std::max(slots, old ? old->counterSlots : 0u);
    w->shadowPerPath = shadowPerPath;
    w->hasPtex = scene->usesPtex;
    const uint32_t binSets = w->counterSlots / kCounterStride;
    cudaError_t e = cudaSuccess;
    auto alloc = [&](void** p, size_t bytes) {
        if (e == cudaSuccess) e = cudaMalloc(p, bytes);
    };
    for (int k = 0; k < 2; ++k) {
        alloc((void**)&w->rays[k], sizeof(SelasRay) * (size_t)capacity);
        alloc((void**)&w->thrPix[k], sizeof(float4) * (size_t)capacity);
        alloc((void**)&w->rngFlags[k], sizeof(uint2) * (size_t)capacity);
        alloc((void**)&w->medium[k], sizeof(float4) * (size_t)capacity);
    }
    alloc((void**)&w->hits, sizeof(SelasHit) * (size_t)capacity);
    alloc((void**)&w->shadowRays, sizeof(SelasRay) * (size_t)capacity * shadowPerPath);
    alloc((void**)&w->shadowPayload, sizeof(float4) * (size_t)capacity * shadowPerPath);
    alloc((void**)&w->counters, sizeof(uint32_t) * w->counterSlots);
    alloc((void**)&w->order, sizeof(uint32_t) * (size_t)capacity);
    alloc((void**)&w->rayOrder, sizeof(uint32_t) * (size_t)capacity);
    alloc((void**)&w->rayBins, sizeof(uint32_t) * 2 * kBinCount * (size_t)binSets);
    if (scene->usesPtex) alloc((void**)&w->ptexRgb, sizeof(float) * 3 * (size_t)capacity);
    alloc((void**)&w->bins, sizeof(uint32_t) * 2 * kBinCount * (size_t)binSets);
    alloc((void**)&w->totals, sizeof(unsigned long long) * 8);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&w->h_totals, sizeof(unsigned long long) * 8);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&w->h_counts, sizeof(uint32_t) * 4);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&w->h_passCounts, sizeof(uint32_t) * kEstimateBounces * kCounterStride);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&w->evPassCounts, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreate(&w->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&w->ev1);
    if (e == cudaSuccess) e = cudaMemsetAsync(w->totals, 0, sizeof(unsigned long long) * 8, stream);
    if (e != cudaSuccess) {
        setError(ctx, "wavefront buffers for %u paths in flight: %s", capacity, cudaGetErrorString(e));
        destroyWavefront(w);
        cudaGetLastError();
        return e == cudaErrorMemoryAllocation ? SELAS_ERR_MEMORY : SELAS_ERR_CUDA;
    }
    if (old) {  // keep what has not been collected yet
        w->pendingSamples = old->pendingSamples;
        w->pendingLaunches = old->pendingLaunches;
        if (old->totals) cudaMemcpyAsync(w->totals, old->totals, sizeof(unsigned long long) * 8, cudaMemcpyDeviceToDevice, stream);
        cudaStreamSynchronize(stream);
        destroyWavefront(old);
    }
    scene->wf = w;
    return SELAS_OK;
}

// JitteredCameraRay (Camera.cpp:63-74) for path ids [first, first+count): pixel = id % pixels, sample = id / pixels.
__global__ void generateKernel(SelasCamera cam, uint32_t width, uint32_t height, uint32_t seed, uint32_t firstSample, uint64_t firstPath,
                               uint32_t count, uint32_t activePixels, uint32_t pixelStride, uint32_t pixelSampler, SelasRay* __restrict__ rays,
                               float4* __restrict__ thrPix, uint2* __restrict__ rngFlags, uint32_t* __restrict__ queueLength) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    if (i == 0) *queueLength = count;  // every later launch of the pass reads its queue length from device memory
    uint64_t id = firstPath + i;
    uint32_t pixel = (uint32_t)(id % activePixels) * pixelStride, sample = firstSample + (uint32_t)(id / activePixels);
    uint32_t y = pixel / width, x = pixel - y * width;
    PathRng rng;
    rng.init(seed, pixel, sample);
    float vx, vy;
    if (pixelSampler == 0) {  // JitteredCameraRay with the sampler, Camera.cpp:63-74
        vx = (float)x + rng.next();
        vy = (float)y + rng.next();
    } else {  // JitteredCameraRay with correlated multi-jitter, Camera.cpp:77-86 (s = sample, p = pixel index)
        float jx, jy;
        correlatedMultiJitter((int32_t)sample, (int32_t)(pixelSampler & 0xffffu), (int32_t)(pixelSampler >> 16), (int32_t)pixel, jx, jy);
        vx = (float)x + jx;
        vy = (float)y + jy;
    }
    V3 d = imageToWorldDirection(cam, vx, vy);
    float4* o = reinterpret_cast<float4*>(rays + i);
    o[0] = make_float4(cam.position[0], cam.position[1], cam.position[2], 0.0f);
    o[1] = make_float4(d.x, d.y, d.z, SELAS_FLOAT_MAX);
    thrPix[i] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(pixel));
    rngFlags[i] = make_uint2(rng.state, FLAG_DELTA_ONLY);
}

// ---- material binning: counting sort of the hit queue by material (misses last) -------------------------------------
// The fused shade kernel is large; with paths of many materials in one warp it ran at 6 of 32 lanes on bounce >= 1 and
// stalled on instruction fetch (profiles/r1a).  Sorting the queue by material makes warps (and neighbouring warps) walk
// the same code.  Keys: materialIndex (< kBinCount-1), last bin = miss.

SELAS_D uint32_t binOf(const SceneGpu& sc, const SelasHit* hits, uint32_t i) {
    uint32_t geomID = __ldg(reinterpret_cast<const uint32_t*>(hits + i) + 4);
    if (geomID == SELAS_INVALID_ID) return kBinCount - 1;
    uint32_t m = sc.meshes[geomID].materialIndex;
    return m < kBinCount - 1 ? m : kBinCount - 2;
}

__global__ void __launch_bounds__(256) binCountKernel(SceneGpu sc, uint32_t n, const uint32_t* __restrict__ nPtr, const SelasHit* __restrict__ hits,
                                                      uint32_t* __restrict__ bins) {
    __shared__ uint32_t local[kBinCount];
    if (nPtr) n = min(n, *nPtr);
    if (threadIdx.x < kBinCount) local[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) atomicAdd(&local[binOf(sc, hits, i)], 1u);
    __syncthreads();
    if (threadIdx.x < kBinCount && local[threadIdx.x]) atomicAdd(bins + threadIdx.x, local[threadIdx.x]);
}

// one warp: exclusive scan of the 64 bin counts into the cursors
__global__ void binScanKernel(const uint32_t* __restrict__ bins, uint32_t* __restrict__ cursor) {
    uint32_t lane = threadIdx.x;
    uint32_t a = bins[lane], b = bins[lane + 32];
    uint32_t xa = a, xb = b;
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t ya = __shfl_up_sync(0xffffffffu, xa, o), yb = __shfl_up_sync(0xffffffffu, xb, o);
        if (lane >= (uint32_t)o) { xa += ya; xb += yb; }
    }
    uint32_t totalA = __shfl_sync(0xffffffffu, xa, 31);
    cursor[lane] = xa - a;
    cursor[lane + 32] = totalA + xb - b;
}

__global__ void __launch_bounds__(256) binScatterKernel(SceneGpu sc, uint32_t n, const uint32_t* __restrict__ nPtr, const SelasHit* __restrict__ hits,
                                                        uint32_t* __restrict__ cursor, uint32_t* __restrict__ order) {
    if (nPtr) n = min(n, *nPtr);
    // per-CTA: local histogram -> one global atomic per bin reserves a contiguous range -> ranks inside the CTA
    __shared__ uint32_t local[kBinCount];
    __shared__ uint32_t base[kBinCount];
    const uint32_t tile = blockDim.x * 8;
    for (uint32_t start = blockIdx.x * tile; start < n; start += gridDim.x * tile) {
        if (threadIdx.x < kBinCount) local[threadIdx.x] = 0;
        __syncthreads();
        uint32_t key[8], rank[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint32_t i = start + k * blockDim.x + threadIdx.x;
            key[k] = 0xFFFFFFFFu;
            if (i < n) {
                key[k] = binOf(sc, hits, i);
                rank[k] = atomicAdd(&local[key[k]], 1u);
            }
        }
        __syncthreads();
        if (threadIdx.x < kBinCount) base[threadIdx.x] = local[threadIdx.x] ? atomicAdd(cursor + threadIdx.x, local[threadIdx.x]) : 0u;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint32_t i = start + k * blockDim.x + threadIdx.x;
            if (i < n) order[base[key[k]] + rank[k]] = i;
        }
        __syncthreads();
    }
}

// ---- ray binning: counting sort of a ray queue by direction octant and origin quadrant -------------------------------------------
// The reference sorts the rays of a bounce by direction category before it traces them (PathTracingBatcher.cpp:162 DetermineRayCategory,
// DeferredPathTracer.cpp:151-232).  Here: 8 octants x 4 quadrants of the scene's (x, z) extent = 32 bins; the trace kernel visits the
// queue through the resulting order (trace.cuh).  Off by default: see SELAS_RAY_SORT in renderPathTracer.
SELAS_D uint32_t rayBinOf(const SelasRay* rays, uint32_t i, float cx, float cz) {
    const float4* rp = reinterpret_cast<const float4*>(rays + i);
    const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
    return (r1.x < 0.f ? 4u : 0u) | (r1.y < 0.f ? 2u : 0u) | (r1.z < 0.f ? 1u : 0u) | (r0.x < cx ? 8u : 0u) | (r0.z < cz ? 16u : 0u);
}
__global__ void __launch_bounds__(256) rayBinCountKernel(uint32_t n, const uint32_t* __restrict__ nPtr, const SelasRay* __restrict__ rays, float cx, float cz,
                                                         uint32_t* __restrict__ bins) {
    __shared__ uint32_t local[kBinCount];
    if (nPtr) n = min(n, *nPtr);
    if (threadIdx.x < kBinCount) local[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) atomicAdd(&local[rayBinOf(rays, i, cx, cz)], 1u);
    __syncthreads();
    if (threadIdx.x < kBinCount && local[threadIdx.x]) atomicAdd(bins + threadIdx.x, local[threadIdx.x]);
}
__global__ void __launch_bounds__(256) rayBinScatterKernel(uint32_t n, const uint32_t* __restrict__ nPtr, const SelasRay* __restrict__ rays, float cx, float cz,
                                                           uint32_t* __restrict__ cursor, uint32_t* __restrict__ order) {
    if (nPtr) n = min(n, *nPtr);
    __shared__ uint32_t local[kBinCount];
    __shared__ uint32_t base[kBinCount];
    const uint32_t tile = blockDim.x * 8;
    for (uint32_t start = blockIdx.x * tile; start < n; start += gridDim.x * tile) {
        if (threadIdx.x < kBinCount) local[threadIdx.x] = 0;
        __syncthreads();
        uint32_t key[8], rank[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint32_t i = start + k * blockDim.x + threadIdx.x;
            key[k] = 0xFFFFFFFFu;
            if (i < n) {
                key[k] = rayBinOf(rays, i, cx, cz);
                rank[k] = atomicAdd(&local[key[k]], 1u);
            }
        }
        __syncthreads();
        if (threadIdx.x < kBinCount) base[threadIdx.x] = local[threadIdx.x] ? atomicAdd(cursor + threadIdx.x, local[threadIdx.x]) : 0u;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint32_t i = start + k * blockDim.x + threadIdx.x;
            if (i < n) order[base[key[k]] + rank[k]] = i;
        }
        __syncthreads();
    }
}

// Material-sorted visiting order of a hit queue for callers outside this file (the VCM light pass).  bins: 2 * kMaterialBinWords
// zeroed words of scratch; order: n words.
int launchMaterialOrder(SelasScene* scene, uint32_t n, const SelasHit* d_hits, uint32_t* d_bins, uint32_t* d_order, cudaStream_t stream) {
    SelasContext* ctx = scene->ctx;
    int grid = (int)std::min<uint32_t>(divUp(n, 2048), (uint32_t)ctx->smCount * 8);
    binCountKernel<<<grid, 256, 0, stream>>>(scene->gpu, n, nullptr, d_hits, d_bins);
    binScanKernel<<<1, 32, 0, stream>>>(d_bins, d_bins + kBinCount);
    binScatterKernel<<<grid, 256, 0, stream>>>(scene->gpu, n, nullptr, d_hits, d_bins + kBinCount, d_order);
    SELAS_CUDA_CHECK(ctx, cudaGetLastError());
    return SELAS_OK;
}

struct ShadeParams {
    uint32_t maxPathLength, rrStartBounce, layerCount, pixels, misMode, skipNonFinite;
};

// INTEGRATOR 0: PathTracer::EvaluatePath (PathTracer.cpp:142-258).  INTEGRATOR 1: the DeferredPathTracer's ShadeHitPosition
// (DeferredPathTracer.cpp:61-148) -- a second NEE sample towards the sky / IBL with balance-heuristic MIS, the BSDF sample weighted
// against the sky pdf, Russian roulette (u >= p) once MaxTrackedBounces_ = 7 bounces are tracked, no media, shadow rays from tnear 0.
#ifndef SELAS_SHADE_MIN_CTAS
#define SELAS_SHADE_MIN_CTAS 4
#endif
template <int INTEGRATOR, bool ONESHOT>
__global__ void __launch_bounds__(128, SELAS_SHADE_MIN_CTAS) shadeKernel(SceneGpu sc, ShadeParams prm, const uint32_t* __restrict__ order, uint32_t nMax,
                                                   const SelasRay* __restrict__ rays, const SelasHit* __restrict__ hits,
                                                   const float4* __restrict__ thrPix, const uint2* __restrict__ rngFlags,
                                                   const float4* __restrict__ mediumIn, SelasRay* __restrict__ raysOut,
                                                   float4* __restrict__ thrPixOut, uint2* __restrict__ rngFlagsOut,
                                                   float4* __restrict__ mediumOut, SelasRay* __restrict__ shadowRays,
                                                   float4* __restrict__ shadowPayload, uint32_t* __restrict__ counters,
                                                   float* __restrict__ film, const float* __restrict__ ptexRgb, const uint32_t* __restrict__ nPtr,
                                                   uint32_t* __restrict__ chunkCounter, uint32_t firstChunk) {
    // The queue length lives in device memory (nPtr; nMax is the launch's upper bound), so the host never waits for it: the grid is
    // persistent and walks the queue with a CTA-uniform stride.
    const uint32_t n = nPtr ? min(nMax, *nPtr) : nMax;
    // Two launch shapes.  ONESHOT: the grid covers the queue (one chunk of 128 slots per CTA, the hardware hands out CTAs in queue
    // order) -- possible when the host has an estimate of the queue length from an earlier pass (renderPathTracer); what an estimate
    // that came out too small leaves over is covered by a second, persistent launch starting at chunk `firstChunk` (a stride loop in the
    // one-shot kernel itself costs 0.26 ms per step: profiles/r2_ab_shade_loop.jsonl).  Otherwise a persistent grid whose CTAs pull chunks
    // from a counter, in queue order, so the resident CTAs always work on one contiguous window of the material-sorted queue.  Measured on
    // configs[2] (ms of shade + material sort per 4-spp step): static grid-wide stride 4.7 (CTAs drift apart across materials), per-warp
    // chunks of 32 slots 4.6, CTA chunks of 128 / 256 / 512 / 1024 slots 4.4 / 4.5 / 4.8 / 6.1, one-shot 4.0.
    __shared__ uint32_t sChunkBase;
  for (uint32_t turn = 0;; ++turn) {
    uint32_t base;
    if (ONESHOT) {
        base = (firstChunk + blockIdx.x) * blockDim.x;
    } else {
        if (threadIdx.x == 0) sChunkBase = atomicAdd(chunkCounter, blockDim.x);
        __syncthreads();
        base = firstChunk * blockDim.x + sChunkBase;
        __syncthreads();
    }
    if (base >= n) break;
    if (ONESHOT && turn > 0) break;
    {
    const uint32_t slotIndex = base + threadIdx.x;
    // material-sorted visiting order (binScatterKernel): lanes of a warp share a material, hence code path
    const uint32_t i = (order && slotIndex < n) ? __ldg(order + slotIndex) : slotIndex;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t ltMask = (1u << lane) - 1u;
    bool alive = false, wantShadow = false, wantShadow2 = false;
    SelasRay outRay;
    float4 outThrPix = make_float4(0.f, 0.f, 0.f, 0.f);
    uint2 outRngFlags = make_uint2(0u, 0u);
    float4 outMedium = make_float4(0.f, 0.f, 0.f, 0.f);
    SelasRay shRay, shRay2;
    float4 shPayload = make_float4(0.f, 0.f, 0.f, 0.f), shPayload2 = make_float4(0.f, 0.f, 0.f, 0.f);

    if (slotIndex < n) {
        const float4* rp = reinterpret_cast<const float4*>(rays + i);
        float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
        const uint4* hp = reinterpret_cast<const uint4*>(hits + i);
        uint4 h0 = __ldg(hp), h1 = __ldg(hp + 1);
        float4 tp = __ldg(thrPix + i);
        uint2 rf = __ldg(rngFlags + i);
        V3 o = v3(r0.x, r0.y, r0.z), d = v3(r1.x, r1.y, r1.z);
        V3 thr = v3(tp.x, tp.y, tp.z);
        const uint32_t pixel = __float_as_uint(tp.w);
        PathRng rng;
        rng.state = rf.x;
        uint32_t bounce = rf.y & 0xffffu;
        bool deltaOnly = (rf.y & FLAG_DELTA_ONLY) != 0;
        Medium medium{SELAS_MEDIUM_VACUUM, v3(0.0f)};
        if (INTEGRATOR == 0 && (rf.y & FLAG_IN_MEDIUM)) {
            float4 m = __ldg(mediumIn + i);
            medium.phase = SELAS_MEDIUM_ISOTROPIC;
            medium.extinction = v3(m.x, m.y, m.z);
        }
        float rayDistance = r1.w;  // the distance sampled for this segment (FLT_MAX in vacuum), PathTracer.cpp:160
        const uint32_t geomID = h1.x;
        const bool isHit = geomID != SELAS_INVALID_ID;
        float* filmRadiance = film + (prm.layerCount == 2 ? 3 * (size_t)prm.pixels : 0);

        Surface s;
        if (isHit) {
            surfaceFromHit(sc, o, d, __uint_as_float(h0.x), __uint_as_float(h0.y), __uint_as_float(h0.z), h0.w, geomID, h1.y,
                           ptexRgb ? ptexRgb + 3 * (size_t)i : nullptr, s, h1.w);
            rayDistance = length(s.position - o);  // :165-167
        }
        thr = thr * transmission(medium, rayDistance);  // :168-169

        if (isHit) {
            DisneyFrame frame;
            makeDisneyFrame(s, s.shader == SELAS_SHADER_DISNEY_THIN, frame);
            V3 view = -d;
            // ---- next-event estimation, :176-194 ----
            LightSample ls;
            ls.radiance = v3(0.0f);
            ls.index = 0; ls.direction = v3(0.f); ls.distance = 0.f; ls.pdfW = 0.f;
            uint32_t lightCount = 0;
            const SelasLight* lights = nullptr;
            if (s.lightSetIndex < sc.lightSetCount) {
                SelasLightSetRange range = sc.lightSets[s.lightSetIndex];
                lightCount = range.count;
                lights = sc.lights + range.start;
            }
            if (lightCount > 0) nextEventEstimation(rng, lights, lightCount, s.position, shadingNormal(s), ls);
            if (dot(ls.radiance, v3(1.0f)) > 0) {
                float fp, rpdf;
                V3 refl = evaluateBsdf(s, frame, view, ls.direction, fp, rpdf);
                // PathTracer.cpp:184 tests the reflectance, DeferredPathTracer.cpp:78 the finished sample (weight 1 in both)
                V3 smp = INTEGRATOR == 1 ? 1.0f * refl * ls.radiance * (1.0f / ls.pdfW) : refl * ls.radiance * (1.0f / ls.pdfW);
                if (dot(INTEGRATOR == 1 ? smp : refl, v3(1.0f)) > 0) {
                    // OcclusionRay, :71-92: origin offset 0.1*error along +-N, tnear = error, tfar = light distance
                    // (the deferred tracer's occlusion batch starts at tnear = 0, DeferredPathTracer.cpp:262)
                    V3 so = offsetRayOrigin(s, ls.direction, 0.1f);
                    V3 c = smp * thr;
                    shRay.ox = so.x; shRay.oy = so.y; shRay.oz = so.z; shRay.tnear = INTEGRATOR == 1 ? 0.0f : s.error;
                    shRay.dx = ls.direction.x; shRay.dy = ls.direction.y; shRay.dz = ls.direction.z; shRay.tfar = ls.distance;
                    shPayload = make_float4(c.x, c.y, c.z, __uint_as_float(pixel));
                    wantShadow = !(prm.skipNonFinite && !(isfinite(c.x) && isfinite(c.y) && isfinite(c.z)));
                }
            }
            if (INTEGRATOR == 1) {
                // ---- sky sample, DeferredPathTracer.cpp:91-111: SampleBackground (AreaLighting.cpp:392-400) + balance heuristic ----
                V3 skyDir, skyRadiance;
                float skyPdf;
                if (sc.ibl.present) {  // DirectIblLightSample, AreaLighting.cpp:293-314
                    float r0 = rng.next(), r1 = rng.next();
                    uint32_t x, y;
                    float dirPhi, dirTheta;
                    iblSample(sc.ibl, r0, r1, dirTheta, dirPhi, x, y, skyPdf);
                    float st = sinf(dirTheta), ct = cosf(dirTheta), sp = sinf(dirPhi), cp = cosf(dirPhi);
                    skyDir = v3(st * cp, ct, st * sp);  // Math::SphericalToCartesian, Projection.cpp:132-146
                    const float* t = sc.ibl.light + 3 * ((size_t)y * sc.ibl.width + x);
                    skyRadiance = sc.ibl.exposureScale * v3(__ldg(t), __ldg(t + 1), __ldg(t + 2));
                } else {  // BackgroundLightSample, AreaLighting.cpp:334-350; QUIRK: pdfW = 1, not 1/(4 pi)
                    skyDir = uniformSphere(rng);
                    skyRadiance = v3(sc.background[0], sc.background[1], sc.background[2]);
                    skyPdf = 1.0f;
                }
                if (dot(skyRadiance, v3(1.0f)) > 0) {
                    float fp = 0.f, rpdf = 0.f;
                    V3 refl = evaluateBsdf(s, frame, view, skyDir, fp, rpdf);
                    float misWeight = (1 * skyPdf) / (1 * skyPdf + 1 * fp);  // BalanceHeuristic, ImportanceSampling.cpp:19-22
                    V3 smp = misWeight * refl * skyRadiance * (1.0f / skyPdf);
                    if (dot(smp, v3(1.0f)) > 0) {
                        V3 so = offsetRayOrigin(s, skyDir, 0.1f);
                        V3 c = smp * thr;
                        shRay2.ox = so.x; shRay2.oy = so.y; shRay2.oz = so.z; shRay2.tnear = 0.0f;
                        shRay2.dx = skyDir.x; shRay2.dy = skyDir.y; shRay2.dz = skyDir.z; shRay2.tfar = 1e36f;
                        shPayload2 = make_float4(c.x, c.y, c.z, __uint_as_float(pixel));
                        wantShadow2 = !(prm.skipNonFinite && !(isfinite(c.x) && isfinite(c.y) && isfinite(c.z)));
                    }
                }
            }
            // ---- BSDF sampling, :197-226 ----
            BsdfResult bs;
            if (INTEGRATOR == 1) {  // DeferredPathTracer.cpp:113-146
                if (sampleBsdf(rng, s, frame, view, bs)) {
                    float skyPdfW = sc.ibl.present ? iblPdf(sc.ibl, bs.wi) : SELAS_INV_4PI;  // BackgroundLightingPdf, AreaLighting.cpp:403-411
                    float misWeight = (1 * bs.forwardPdfW) / (1 * bs.forwardPdfW + 1 * skyPdfW);
                    thr = misWeight * thr * bs.reflectance;
                    if (dot(thr, thr) != 0.0f) {
                        alive = true;
                        if (bounce >= 7u) {  // hit.trackedBounces >= MaxTrackedBounces_ (IntegratorContexts.h:43); tracked = min(7, bounces)
                            float p = maxf_(maxf_(thr.x, thr.y), thr.z);
                            if (rng.next() >= p) alive = false;
                            thr = thr * (1.0f / p);
                        }
                        deltaOnly = deltaOnly && (bs.flags & SELAS_EVENT_DIRAC);
                        o = offsetRayOrigin(s, bs.wi, 1.0f);
                        d = bs.wi;
                    }
                }
            } else if (sampleBsdf(rng, s, frame, view, bs)) {
                deltaOnly = deltaOnly && (bs.flags & SELAS_EVENT_DIRAC);
                if (bs.flags == SELAS_EVENT_TRANSMISSION) {  // exact equality, :206
                    if (medium.phase == SELAS_MEDIUM_VACUUM) {
                        medium.phase = bs.mediumPhase;
                        medium.extinction = bs.mediumExtinction;
                    } else {
                        medium.phase = SELAS_MEDIUM_VACUUM;
                        medium.extinction = v3(0.0f);
                    }
                }
                float weight = 1.0f;  // QUIRK :214-216: MIS weight forced to 1 (misMode 0 = reference)
                if (prm.misMode == 1) {  // the balance heuristic the reference computes and discards (ImportanceSampling.cpp:19-22)
                    float lightPdfW = lightingPdf(lights, lightCount, ls.index, s.position, bs.wi);
                    weight = bs.forwardPdfW / (bs.forwardPdfW + lightPdfW);
                }
                thr = weight * thr * bs.reflectance;
                V3 no = offsetRayOrigin(s, bs.wi, 1.0f);
                o = no;
                d = bs.wi;
                if (bounce == 0 && prm.layerCount == 2) {  // Ld[0], :223-225
                    atomicAdd(film + 3 * (size_t)pixel, bs.reflectance.x);
                    atomicAdd(film + 3 * (size_t)pixel + 1, bs.reflectance.y);
                    atomicAdd(film + 3 * (size_t)pixel + 2, bs.reflectance.z);
                }
                alive = true;
            }
        } else if (medium.phase != SELAS_MEDIUM_VACUUM) {  // scatter inside the medium, :228-233
            o = o + rayDistance * d;
            d = uniformSphere(rng);
            alive = true;
        } else {  // escaped: background, :234-243
            // EvaluateBackground / EvaluateBackgroundMiss, AreaLighting.cpp:415-437: the IBL when the scene has one (the
            // "miss" image for paths that only saw delta events), else the constant background
            V3 bg = v3(sc.background[0], sc.background[1], sc.background[2]);
            if (sc.ibl.present) bg = deltaOnly ? iblLookupMiss(sc.ibl, d) : iblLookup(sc.ibl, d);
            V3 c = bg * thr;
            if (prm.skipNonFinite && !(isfinite(c.x) && isfinite(c.y) && isfinite(c.z))) c = v3(0.0f);
            atomicAdd(filmRadiance + 3 * (size_t)pixel, c.x);
            atomicAdd(filmRadiance + 3 * (size_t)pixel + 1, c.y);
            atomicAdd(filmRadiance + 3 * (size_t)pixel + 2, c.z);
        }

        if (alive) {
            ++bounce;
            if (INTEGRATOR == 0 && bounce > prm.rrStartBounce) {  // Russian roulette, :248-254; QUIRK: p is not clamped to 1
                float p = maxf_(maxf_(thr.x, thr.y), thr.z);
                if (rng.next() > p) alive = false;
                thr = thr * (1.0f / p);
            }
            if (bounce >= prm.maxPathLength) alive = false;  // loop condition, :157
        }
        if (alive) {
            float nextDistance = sampleDistance(rng, medium);  // next iteration's :160
            outRay.ox = o.x; outRay.oy = o.y; outRay.oz = o.z; outRay.tnear = 0.0f;
            outRay.dx = d.x; outRay.dy = d.y; outRay.dz = d.z; outRay.tfar = nextDistance;
            outThrPix = make_float4(thr.x, thr.y, thr.z, __uint_as_float(pixel));
            uint32_t fl = (bounce & 0xffffu) | (deltaOnly ? FLAG_DELTA_ONLY : 0u) | (medium.phase != SELAS_MEDIUM_VACUUM ? FLAG_IN_MEDIUM : 0u);
            outRngFlags = make_uint2(rng.state, fl);
            outMedium = make_float4(medium.extinction.x, medium.extinction.y, medium.extinction.z, 0.f);
        }
    }

    // ---- warp-aggregated queue appends ----
    uint32_t aliveMask = __ballot_sync(0xffffffffu, alive);
    uint32_t shadowMask = __ballot_sync(0xffffffffu, wantShadow);
    uint32_t baseA = 0, baseS = 0;
    if (lane == 0) {
        if (aliveMask) baseA = atomicAdd(counters + 0, (uint32_t)__popc(aliveMask));
        if (shadowMask) baseS = atomicAdd(counters + 1, (uint32_t)__popc(shadowMask));
    }
    baseA = __shfl_sync(0xffffffffu, baseA, 0);
    baseS = __shfl_sync(0xffffffffu, baseS, 0);
    if (alive) {
        uint32_t slot = baseA + __popc(aliveMask & ltMask);
        float4* ro = reinterpret_cast<float4*>(raysOut + slot);
        ro[0] = make_float4(outRay.ox, outRay.oy, outRay.oz, outRay.tnear);
        ro[1] = make_float4(outRay.dx, outRay.dy, outRay.dz, outRay.tfar);
        thrPixOut[slot] = outThrPix;
        rngFlagsOut[slot] = outRngFlags;
        if (outRngFlags.y & FLAG_IN_MEDIUM) mediumOut[slot] = outMedium;
    }
    if (wantShadow) {
        uint32_t slot = baseS + __popc(shadowMask & ltMask);
        float4* so = reinterpret_cast<float4*>(shadowRays + slot);
        so[0] = make_float4(shRay.ox, shRay.oy, shRay.oz, shRay.tnear);
        so[1] = make_float4(shRay.dx, shRay.dy, shRay.dz, shRay.tfar);
        shadowPayload[slot] = shPayload;
    }
    if (INTEGRATOR == 1) {  // the sky shadow ray
        uint32_t mask2 = __ballot_sync(0xffffffffu, wantShadow2);
        uint32_t base2 = 0;
        if (lane == 0 && mask2) base2 = atomicAdd(counters + 1, (uint32_t)__popc(mask2));
        base2 = __shfl_sync(0xffffffffu, base2, 0);
        if (wantShadow2) {
            uint32_t slot = base2 + __popc(mask2 & ltMask);
            float4* so = reinterpret_cast<float4*>(shadowRays + slot);
            so[0] = make_float4(shRay2.ox, shRay2.oy, shRay2.oz, shRay2.tnear);
            so[1] = make_float4(shRay2.dx, shRay2.dy, shRay2.dz, shRay2.tfar);
            shadowPayload[slot] = shPayload2;
        }
    }
   }
  }
}

// Ray totals of one pass, summed on the device so that nothing has to be read back per bounce: closest-hit rays = sum of the queue lengths,
// occlusion rays = sum of the shadow-queue lengths (counter layout: see renderPathTracer).
__global__ void passTotalsKernel(uint32_t* __restrict__ counters, uint32_t iterations, unsigned long long* __restrict__ totals, uint32_t passSize) {
    unsigned long long closest = 0, shadow = 0;
    for (uint32_t it = threadIdx.x; it < iterations; it += 32) {
        closest += counters[kCounterStride * it];
        shadow += counters[kCounterStride * (it + 1) + 1];
    }
    for (int s = 16; s > 0; s >>= 1) {
        closest += __shfl_xor_sync(0xffffffffu, closest, s);
        shadow += __shfl_xor_sync(0xffffffffu, shadow, s);
    }
    if (threadIdx.x == 0) {
        totals[0] += closest;
        totals[1] += shadow;
        counters[5] = iterations;  // spare words of the first set: what the host needs to read the queue lengths as grid estimates
        counters[6] = passSize;
    }
}

// Sync with the render stream and hand out everything measured since the last collect: per-kernel event spans, ray totals summed on the
// device, path samples.  (selas_cuda_render_collect_stats; also the tail of a render call that was given a stats pointer.)
int collectRenderStats(SelasScene* scene, SelasRenderStats* out) {
    SelasContext* ctx = scene->ctx;
    SelasRenderStats st{};
    WavefrontBuffers* w = (WavefrontBuffers*)scene->wf;
    if (w) {
        cudaStream_t stream = w->lastStream ? w->lastStream : ctx->stream;
        if (w->spanOpen) cudaEventRecord(w->ev1, stream);
        SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(w->h_totals, w->totals, sizeof(unsigned long long) * 8, cudaMemcpyDeviceToHost, stream));
        SELAS_CUDA_CHECK(ctx, cudaMemsetAsync(w->totals, 0, sizeof(unsigned long long) * 8, stream));
        SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(stream));
        if (w->spanOpen) cudaEventElapsedTime(&st.renderMs, w->ev0, w->ev1);
        w->spanOpen = false;
        w->timers.collect(st);
        st.closestRays = w->h_totals[0];
        st.occlusionRays = w->h_totals[1];
        st.closestNodes = w->h_totals[2]; st.closestTris = w->h_totals[3]; st.shadowNodes = w->h_totals[4]; st.shadowTris = w->h_totals[5];
        st.pathSamples = w->pendingSamples;
        st.kernelLaunches = w->pendingLaunches;
        w->pendingSamples = 0;
        w->pendingLaunches = 0;
    }
    if (out) *out = st;
    return SELAS_OK;
}

// Enqueues the whole render on `stream` and returns without waiting for the device: queue lengths, ray totals and the launch shapes of
// every bounce stay on the device (persistent / grid-stride kernels that read their queue length from memory).  stats != null: the
// call then collects (and so waits); stats == null: nothing is waited for -- selas_cuda_render_collect_stats picks the numbers up later.
// Paths longer than 32 bounces (the reference's MaxBounceCount_ is 2048) look at the queue length every 16 bounces to stop early.
int renderPathTracer(SelasScene* scene, const SelasCamera& cam, const SelasRenderSettings& rs, float* d_film, SelasRenderStats* stats,
                     cudaStream_t stream) {
    SelasContext* ctx = scene->ctx;
    if (rs.width == 0 || rs.height == 0 || rs.layerCount < 1 || rs.layerCount > 2 || rs.maxPathLength > 65535u ||
        (uint64_t)rs.width * rs.height > 0x7fffffffull) {
        setError(ctx, "render settings out of range (width/height > 0, layerCount 1..2, maxPathLength <= 65535)");
        return SELAS_ERR_ARGUMENT;
    }
    if (rs.pixelSampler != 0) {
        const uint32_t m = rs.pixelSampler & 0xffffu, n = rs.pixelSampler >> 16;
        if (m == 0 || n == 0 || (uint64_t)rs.firstSample + rs.samplesPerPixel > (uint64_t)m * n) {
            setError(ctx, "pixelSampler: correlated multi-jitter needs an m x n grid (m, n > 0) holding firstSample + samplesPerPixel samples");
            return SELAS_ERR_ARGUMENT;
        }
    }
    if (rs.integrator > SELAS_INTEGRATOR_DEFERRED || (rs.integrator == SELAS_INTEGRATOR_DEFERRED && rs.layerCount != 1)) {
        setError(ctx, "integrator %u unknown, or the deferred integrator asked for %u layers (it has one, OutputLayers_)", rs.integrator, rs.layerCount);
        return SELAS_ERR_ARGUMENT;
    }
    if (rs.misMode > 1) {
        setError(ctx, "misMode %u unknown (0 = reference weighting, 1 = balance heuristic)", rs.misMode);
        return SELAS_ERR_ARGUMENT;
    }
    const uint32_t pixels = rs.width * rs.height;
    const uint32_t stride = rs.pixelStride > 1 ? rs.pixelStride : 1u;
    const uint32_t activePixels = (pixels + stride - 1) / stride;
    const uint64_t totalPaths = (uint64_t)activePixels * rs.samplesPerPixel;
    // paths in flight per pass: whole sample passes, about 8M
    uint32_t sppPerPass = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(rs.samplesPerPixel, (8u << 20) / std::max(1u, activePixels)));
    const uint32_t capacity = (uint32_t)std::min<uint64_t>(totalPaths, (uint64_t)activePixels * sppPerPass);
    if (capacity == 0 || rs.maxPathLength == 0) {
        if (stats) {
            *stats = SelasRenderStats{};
            stats->pathSamples = totalPaths;
        }
        return SELAS_OK;
    }
    const bool deferred = rs.integrator == SELAS_INTEGRATOR_DEFERRED;
    // counter slots: a long path length only needs as many as are enqueued between two early-out checks
    const uint32_t kCheckEvery = 16;
    const bool longPaths = rs.maxPathLength > 32;
    int rc = ensureWavefront(scene, capacity, longPaths ? kCheckEvery + 1 : rs.maxPathLength, deferred ? 2u : 1u, stream);
    if (rc) return rc;
    WavefrontBuffers* w = (WavefrontBuffers*)scene->wf;
    if (w->lastStream && w->lastStream != stream) cudaStreamSynchronize(w->lastStream);  // one queue set: calls on another stream wait for the last
    w->lastStream = stream;
    ShadeParams prm{rs.maxPathLength, rs.rrStartBounce, rs.layerCount, pixels, rs.misMode, (rs.flags & SELAS_RENDER_SKIP_NONFINITE) ? 1u : 0u};
    const bool counting = (rs.flags & SELAS_RENDER_COUNT_TRAVERSAL) != 0;
    unsigned long long* d_totals = counting ? w->totals + 2 : nullptr;
    KernelTimers& timers = w->timers;
    // tuning knob: first bounce whose hit queue is sorted by material before shading (SELAS_SORT_FROM_BOUNCE, default 0)
    static const uint32_t sortFromBounce = getenv("SELAS_SORT_FROM_BOUNCE") ? (uint32_t)atoi(getenv("SELAS_SORT_FROM_BOUNCE")) : 0u;
    // persistent shade grid: what is resident, twice over (the queue is walked with a grid-wide stride, so every CTA sees every material)
    static int shadeResident[2] = {0, 0};
    if (!shadeResident[deferred ? 1 : 0]) {
        int r = 0;
        cudaError_t e = deferred ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r, shadeKernel<1, false>, 128, 0)
                                 : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r, shadeKernel<0, false>, 128, 0);
        shadeResident[deferred ? 1 : 0] = (e == cudaSuccess && r > 0) ? r : 4;
    }
    // SELAS_RAY_SORT=b: from bounce b on, the ray queue is traced in direction-octant / origin-quadrant order (the reference sorts its ray
    // batches by direction category, PathTracingBatcher.cpp:162).  Default: off -- see DESIGN.md for the measurement.
    static const uint32_t raySortFrom = getenv("SELAS_RAY_SORT") ? (uint32_t)atoi(getenv("SELAS_RAY_SORT")) : 0xFFFFFFFFu;
    const float raySortCx = scene->boundsCenter[0], raySortCz = scene->boundsCenter[2];  // origin quadrants about the scene's centre
    // Warp-cooperative triangle stage (trace_coop.cuh) per bounce: camera rays are coherent and stay with traceKernel; everything after the
    // first bounce, and the shadow rays from the first hit on, go through traceCoopKernel.  SELAS_COOP_CLOSEST_FROM / SELAS_COOP_SHADOW_FROM
    // = first bounce that uses it (a large number switches it off).
    static const uint32_t coopClosestFrom = getenv("SELAS_COOP_CLOSEST_FROM") ? (uint32_t)atoi(getenv("SELAS_COOP_CLOSEST_FROM")) : 1u;
    static const uint32_t coopShadowFrom = getenv("SELAS_COOP_SHADOW_FROM") ? (uint32_t)atoi(getenv("SELAS_COOP_SHADOW_FROM")) : 0u;
    static const uint32_t kSortMinQueue = getenv("SELAS_SORT_MIN_QUEUE") ? (uint32_t)atoi(getenv("SELAS_SORT_MIN_QUEUE")) : 65536u;
    static const uint32_t shadowSortBounces = getenv("SELAS_SHADOW_SORT") ? (uint32_t)atoi(getenv("SELAS_SHADOW_SORT")) : 0u;
    static const bool shadeOneShot = !(getenv("SELAS_SHADE_ONESHOT") && atoi(getenv("SELAS_SHADE_ONESHOT")) == 0);
    static const int shadeWaves = getenv("SELAS_SHADE_WAVES") ? std::max(1, atoi(getenv("SELAS_SHADE_WAVES"))) : 2;

    if (!w->spanOpen) {
        cudaEventRecord(w->ev0, stream);
        w->spanOpen = true;
    }
    uint64_t launches = 0;
    for (uint64_t first = 0; first < totalPaths; first += capacity) {
        const uint32_t count = (uint32_t)std::min<uint64_t>(capacity, totalPaths - first);
        const uint32_t slotSets = w->counterSlots / kCounterStride;
        if (w->passCountsPending && cudaEventQuery(w->evPassCounts) == cudaSuccess) {  // an earlier pass's queue lengths have arrived
            w->passCountsPending = false;
            const uint32_t nb = std::min<uint32_t>(kEstimateBounces, w->h_passCounts[kCounterStride * 0 + 5]);
            w->estimateCount = w->h_passCounts[kCounterStride * 0 + 6];
            w->estimate.assign(nb, 0u);
            for (uint32_t b = 0; b < nb; ++b) w->estimate[b] = std::max<uint32_t>(1u, w->h_passCounts[kCounterStride * b]);
        }
        cudaGetLastError();  // cudaErrorNotReady of the query is not an error
        SELAS_CUDA_CHECK(ctx, cudaMemsetAsync(w->counters, 0, sizeof(uint32_t) * w->counterSlots, stream));
        SELAS_CUDA_CHECK(ctx, cudaMemsetAsync(w->bins, 0, sizeof(uint32_t) * 2 * kBinCount * (size_t)slotSets, stream));
        if (raySortFrom < rs.maxPathLength || shadowSortBounces > 0) SELAS_CUDA_CHECK(ctx, cudaMemsetAsync(w->rayBins, 0, sizeof(uint32_t) * 2 * kBinCount * (size_t)slotSets, stream));
        size_t t0 = timers.begin(stream);
        generateKernel<<<divUp(count, 256), 256, 0, stream>>>(cam, rs.width, rs.height, rs.seed, rs.firstSample, first, count, activePixels,
                                                              stride, rs.pixelSampler, w->rays[0], w->thrPix[0], w->rngFlags[0], w->counters);
        timers.end(0, t0, stream);
        ++launches;
        const int sortGrid = (int)std::min<uint32_t>(divUp(count, 2048), (uint32_t)ctx->smCount * 8);
        const int shadeGrid = (int)std::min<uint32_t>(divUp(count, 128), (uint32_t)(ctx->smCount * shadeResident[deferred ? 1 : 0] * shadeWaves));
        const unsigned ptexGrid = (unsigned)std::min<uint32_t>(divUp(count, 1024), (uint32_t)ctx->smCount * 8);
        int cur = 0;
        uint32_t slot = 0;  // counter set of the current bounce
        for (uint32_t it = 0; it < rs.maxPathLength; ++it, ++slot) {
            if (longPaths && slot == kCheckEvery) {
                // the sets in use are exhausted: fold them into the totals, look at the queue length, and start over at set 0
                passTotalsKernel<<<1, 32, 0, stream>>>(w->counters, slot, w->totals, count);
                SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(w->h_counts, w->counters + kCounterStride * (size_t)slot, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
                SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(stream));
                const uint32_t alive = w->h_counts[0];
                if (alive == 0) { slot = 0; break; }
                SELAS_CUDA_CHECK(ctx, cudaMemsetAsync(w->counters, 0, sizeof(uint32_t) * w->counterSlots, stream));
                SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(w->counters, w->h_counts, sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
                SELAS_CUDA_CHECK(ctx, cudaMemsetAsync(w->bins, 0, sizeof(uint32_t) * 2 * kBinCount * (size_t)slotSets, stream));
                slot = 0;
                ++launches;
            }
            uint32_t* ctr = w->counters + kCounterStride * (size_t)slot;  // [0] = this bounce's queue length
            uint32_t* next = ctr + kCounterStride;                         // [0] = next queue length, [1] = shadow rays of this bounce
            // closest hit over the dense path queue (optionally through a direction-sorted order)
            t0 = timers.begin(stream);
            const uint32_t* rayOrder = nullptr;
            if (it >= raySortFrom) {
                uint32_t* rb = w->rayBins + 2 * kBinCount * (size_t)slot;
                rayBinCountKernel<<<sortGrid, 256, 0, stream>>>(count, ctr, w->rays[cur], raySortCx, raySortCz, rb);
                binScanKernel<<<1, 32, 0, stream>>>(rb, rb + kBinCount);
                rayBinScatterKernel<<<sortGrid, 256, 0, stream>>>(count, ctr, w->rays[cur], raySortCx, raySortCz, rb + kBinCount, w->rayOrder);
                rayOrder = w->rayOrder;
                launches += 3;
            }
            // queue length of this bounce as an earlier pass of the same shape saw it (0 = not known yet)
            const uint32_t est = (!longPaths && w->estimateCount == count && it < w->estimate.size()) ? w->estimate[it] : 0u;
            rc = launchIntersectIndirect(scene, ctr, count, w->rays[cur], w->hits, ctr + 2, d_totals, stream, rayOrder, it >= coopClosestFrom ? 1 : 0, est);
            if (rc) return rc;
            timers.end(1, t0, stream);
            t0 = timers.begin(stream);
            const uint32_t* order = nullptr;
            // the material sort pays while a queue fills the machine; the last bounces' few thousand paths are shaded in queue order
            // (three launches less per bounce), and the binning grids follow the estimate instead of the pass size
            const bool sortThis = it >= sortFromBounce && !(est != 0u && est < kSortMinQueue);
            if (sortThis) {
                // (not below one CTA per SM: a stale estimate must not serialise a long queue)
                const int grid = est ? (int)std::min<uint32_t>((uint32_t)sortGrid, std::max<uint32_t>((uint32_t)ctx->smCount, divUp(est + est / 16u + 2048u, 2048))) : sortGrid;
                uint32_t* bins = w->bins + 2 * kBinCount * (size_t)slot;
                binCountKernel<<<grid, 256, 0, stream>>>(scene->gpu, count, ctr, w->hits, bins);
                binScanKernel<<<1, 32, 0, stream>>>(bins, bins + kBinCount);
                binScatterKernel<<<grid, 256, 0, stream>>>(scene->gpu, count, ctr, w->hits, bins + kBinCount, w->order);
                order = w->order;
                launches += 3;
            }
            if (scene->usesPtex) {  // Ptex stage: base-colour lookups for the whole hit queue
                if (scene->ptexMixedRes) ptexStageKernel<true><<<ptexGrid, 128, 0, stream>>>(scene->gpu, count, w->hits, w->ptexRgb, ctr);
                else ptexStageKernel<false><<<ptexGrid, 128, 0, stream>>>(scene->gpu, count, w->hits, w->ptexRgb, ctr);
                ++launches;
            }
            // one-shot grid when an earlier pass of this shape told how long this bounce's queue gets (3 % + 8 CTAs of slack; what a low
            // estimate leaves over is covered by a persistent launch behind it, which normally finds nothing), else the persistent shape
            const bool oneShot = est != 0u && shadeOneShot;
            // SELAS_SHADE_ESTIMATE_PERCENT scales the estimate (tests under-size it to exercise the launch behind the one-shot grid)
            const char* pct = getenv("SELAS_SHADE_ESTIMATE_PERCENT");
            const uint64_t estScaled = pct ? std::max<uint64_t>(1u, (uint64_t)est * (uint64_t)std::max(1, atoi(pct)) / 100u) : (uint64_t)est + est / 32u;
            const int grid = oneShot ? (int)std::min<uint64_t>(divUp(count, 128), divUp(estScaled, 128) + 8u) : shadeGrid;
            const bool tail = oneShot && (uint32_t)grid < divUp(count, 128);
#define SELAS_SHADE_LAUNCH(INTEGRATOR, ONESHOT, GRID, FIRST)                                                                                     \
    shadeKernel<INTEGRATOR, ONESHOT><<<GRID, 128, 0, stream>>>(scene->gpu, prm, order, count, w->rays[cur], w->hits, w->thrPix[cur], w->rngFlags[cur], \
                                                              w->medium[cur], w->rays[cur ^ 1], w->thrPix[cur ^ 1], w->rngFlags[cur ^ 1],       \
                                                              w->medium[cur ^ 1], w->shadowRays, w->shadowPayload, next, d_film,                 \
                                                              scene->usesPtex ? w->ptexRgb : nullptr, ctr, ctr + 4, FIRST)
            if (deferred) {
                if (oneShot) SELAS_SHADE_LAUNCH(1, true, grid, 0u);
                if (!oneShot || tail) SELAS_SHADE_LAUNCH(1, false, oneShot ? ctx->smCount : grid, oneShot ? (uint32_t)grid : 0u);
            } else {
                if (oneShot) SELAS_SHADE_LAUNCH(0, true, grid, 0u);
                if (!oneShot || tail) SELAS_SHADE_LAUNCH(0, false, oneShot ? ctx->smCount : grid, oneShot ? (uint32_t)grid : 0u);
            }
#undef SELAS_SHADE_LAUNCH
            if (tail) ++launches;
            timers.end(2, t0, stream);
            SELAS_CUDA_CHECK(ctx, cudaGetLastError());
            {
                float* filmRadiance = d_film + (rs.layerCount == 2 ? 3 * (size_t)pixels : 0);
                t0 = timers.begin(stream);
                const uint32_t* shadowOrder = nullptr;
                if (it < shadowSortBounces && w->shadowPerPath == 1 && raySortFrom >= rs.maxPathLength) {
                    // SELAS_SHADOW_SORT=b: the first b bounces' shadow rays are traced in direction-octant / origin-quadrant order (rays towards
                    // one light side by side); measured in DESIGN 4
                    uint32_t* rb = w->rayBins + 2 * kBinCount * (size_t)slot;
                    rayBinCountKernel<<<sortGrid, 256, 0, stream>>>(count, next + 1, w->shadowRays, raySortCx, raySortCz, rb);
                    binScanKernel<<<1, 32, 0, stream>>>(rb, rb + kBinCount);
                    rayBinScatterKernel<<<sortGrid, 256, 0, stream>>>(count, next + 1, w->shadowRays, raySortCx, raySortCz, rb + kBinCount, w->rayOrder);
                    shadowOrder = w->rayOrder;
                    launches += 3;
                }
                rc = launchShadowAccumulate(scene, next + 1, count * w->shadowPerPath, w->shadowRays, w->shadowPayload, filmRadiance, ctr + 3,
                                            d_totals ? d_totals + 2 : nullptr, stream, it >= coopShadowFrom ? 1 : 0, est ? est * w->shadowPerPath : 0u, shadowOrder);
                if (rc) return rc;
                timers.end(3, t0, stream);
            }
            launches += 3;
            cur ^= 1;
        }
        if (slot) {
            passTotalsKernel<<<1, 32, 0, stream>>>(w->counters, slot, w->totals, count);
            ++launches;
            if (!longPaths && !w->passCountsPending) {  // queue lengths of this pass for the grid estimates of later ones (not waited for)
                SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(w->h_passCounts, w->counters, sizeof(uint32_t) * kCounterStride * std::min<uint32_t>(slot, kEstimateBounces),
                                                      cudaMemcpyDeviceToHost, stream));
                SELAS_CUDA_CHECK(ctx, cudaEventRecord(w->evPassCounts, stream));
                w->passCountsPending = true;
            }
        }
    }
    w->pendingSamples += totalPaths;
    w->pendingLaunches += launches;
    SELAS_CUDA_CHECK(ctx, cudaGetLastError());
    if (stats) return collectRenderStats(scene, stats);
    return SELAS_OK;
}

}  // namespace selas
