// trace.cu -- BVH8 build driver and ray-cast launchers (translation unit compiled with FMA contraction on).
#include <cstdio>
#include <cstring>

#include "bvh_build.cuh"
#include "internal.h"
#include "trace.cuh"

namespace selas {

namespace {

template <typename T>
int devAlloc(SelasScene* scene, T** out, size_t count, bool keep) {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, count * sizeof(T) + 16);
    if (e != cudaSuccess) {
        setError(scene->ctx, "cudaMalloc(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(e));
        return SELAS_ERR_MEMORY;
    }
    if (keep) scene->allocations.push_back(p);
    *out = (T*)p;
    return SELAS_OK;
}

struct TempFree {
    std::vector<void*> ptrs;
    ~TempFree() {
        for (void* p : ptrs) cudaFree(p);
    }
};

inline uint32_t divUp(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

}  // namespace

int buildBvh(SelasScene* scene, const std::vector<MeshGpu>& meshes, const std::vector<InstanceGpu>& inst, const InstanceGpu* d_instances) {
    SelasContext* ctx = scene->ctx;
    cudaStream_t st = ctx->stream;
    cudaEvent_t ev0, ev1;
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    cudaEventRecord(ev0, st);

    // ---- count triangles per draw ----
    struct Draw { uint32_t mesh, inst, base, tris; };
    std::vector<Draw> draws;
    uint64_t total = 0;
    for (uint32_t ii = 0; ii < inst.size(); ++ii)
        for (uint32_t m = 0; m < meshes.size(); ++m) {
            if (inst[ii].mesh != 0xFFFFFFFFu && inst[ii].mesh != m) continue;
            uint32_t t = meshes[m].faceCount * (meshes[m].indicesPerFace == 4 ? 2u : 1u);
            if (t == 0) continue;
            draws.push_back({m, ii, (uint32_t)total, t});
            total += t;
        }
    if (total >= 0x7fffffffull) {
        setError(ctx, "scene has %llu triangles; the builder is limited to 2^31-1", (unsigned long long)total);
        return SELAS_ERR_UNSUPPORTED;
    }
    const uint32_t n = (uint32_t)total;
    scene->stats = SelasBvhStats{};
    scene->stats.triangleCount = n;
    scene->gpu.nodes = nullptr;
    scene->gpu.tris = nullptr;
    scene->gpu.nodeCount = 0;
    scene->gpu.triCount = 0;
    if (n == 0) return SELAS_OK;

    TempFree tmp;
    auto tempAlloc = [&](void** p, size_t bytes) -> int {
        cudaError_t e = cudaMalloc(p, bytes + 16);
        if (e != cudaSuccess) {
            setError(ctx, "cudaMalloc(%zu bytes) failed during BVH build: %s", bytes, cudaGetErrorString(e));
            return SELAS_ERR_MEMORY;
        }
        tmp.ptrs.push_back(*p);
        return SELAS_OK;
    };

    BuildTri* d_tris;
    uint32_t* d_bounds;
    uint64_t *d_keys, *d_keysAlt;
    uint32_t *d_vals, *d_valsAlt;
    int rc;
    if ((rc = tempAlloc((void**)&d_tris, sizeof(BuildTri) * (size_t)n))) return rc;
    if ((rc = tempAlloc((void**)&d_bounds, 6 * sizeof(uint32_t)))) return rc;
    if ((rc = tempAlloc((void**)&d_keys, sizeof(uint64_t) * (size_t)n))) return rc;
    if ((rc = tempAlloc((void**)&d_keysAlt, sizeof(uint64_t) * (size_t)n))) return rc;
    if ((rc = tempAlloc((void**)&d_vals, sizeof(uint32_t) * (size_t)n))) return rc;
    if ((rc = tempAlloc((void**)&d_valsAlt, sizeof(uint32_t) * (size_t)n))) return rc;

    uint32_t boundsInit[6] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u};
    SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(d_bounds, boundsInit, sizeof(boundsInit), cudaMemcpyHostToDevice, st));
    for (const Draw& d : draws) {
        MeshDev md{meshes[d.mesh].positions, meshes[d.mesh].indices, meshes[d.mesh].faceCount, meshes[d.mesh].indicesPerFace};
        gatherTrianglesKernel<<<divUp(md.faceCount, 256), 256, 0, st>>>(md, d_instances[d.inst].m, inst[d.inst].identity != 0, d.mesh, d.inst,
                                                                         d_tris, d.base, d_bounds);
    }
    mortonKernel<<<divUp(n, 256), 256, 0, st>>>(d_tris, n, d_bounds, d_keys, d_vals);
    SELAS_CUDA_CHECK(ctx, cudaGetLastError());

    // library radix sort (load-time step, not on the measured path)
    cub::DoubleBuffer<uint64_t> kb(d_keys, d_keysAlt);
    cub::DoubleBuffer<uint32_t> vb(d_vals, d_valsAlt);
    size_t sortBytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sortBytes, kb, vb, (int)n, 0, 63, st);
    void* d_sortTmp;
    if ((rc = tempAlloc(&d_sortTmp, sortBytes))) return rc;
    SELAS_CUDA_CHECK(ctx, cub::DeviceRadixSort::SortPairs(d_sortTmp, sortBytes, kb, vb, (int)n, 0, 63, st));
    const uint64_t* keys = kb.Current();
    const uint32_t* sorted = vb.Current();

    int2* d_children;
    uint32_t* d_parent;
    uint2* d_range;
    float4 *d_boxLo, *d_boxHi;
    uint32_t* d_arrivals;
    const size_t nodes2 = 2 * (size_t)n - 1;
    if ((rc = tempAlloc((void**)&d_children, sizeof(int2) * (size_t)n))) return rc;
    if ((rc = tempAlloc((void**)&d_range, sizeof(uint2) * (size_t)n))) return rc;
    if ((rc = tempAlloc((void**)&d_parent, sizeof(uint32_t) * nodes2))) return rc;
    if ((rc = tempAlloc((void**)&d_boxLo, sizeof(float4) * nodes2))) return rc;
    if ((rc = tempAlloc((void**)&d_boxHi, sizeof(float4) * nodes2))) return rc;
    if ((rc = tempAlloc((void**)&d_arrivals, sizeof(uint32_t) * (size_t)n))) return rc;
    SELAS_CUDA_CHECK(ctx, cudaMemsetAsync(d_arrivals, 0, sizeof(uint32_t) * (size_t)n, st));
    if (n > 1) radixTreeKernel<<<divUp(n - 1, 256), 256, 0, st>>>(keys, (int)n, d_children, d_parent, d_range);
    fitBoxesKernel<<<divUp(n, 256), 256, 0, st>>>(d_tris, sorted, (int)n, d_children, d_parent, d_boxLo, d_boxHi, d_arrivals);
    SELAS_CUDA_CHECK(ctx, cudaGetLastError());

    // ---- collapse ----
    // Wide nodes other than the root cover > 3 triangles.  Leaf wide nodes (no inner child) are disjoint: L <= n/4;
    // branching nodes B < L; chain nodes (one inner child) own >= 7 direct triangles: 4L + 7C <= n.  So the node
    // count is < 2L + C <= n/2.  The final triangle array has exactly n entries.
    const size_t maxNodes = (size_t)n / 2 + 2;
    float4 *d_nodes, *d_outTris;
    if ((rc = devAlloc(scene, &d_nodes, maxNodes * 5, true))) return rc;
    if ((rc = devAlloc(scene, &d_outTris, (size_t)n * 3, true))) return rc;
    CollapseTask *d_tasksA, *d_tasksB;
    uint32_t* d_ctr;  // [0]=nodeCounter [1]=triCounter [2..]=per-level next counts
    if ((rc = tempAlloc((void**)&d_tasksA, sizeof(CollapseTask) * maxNodes))) return rc;
    if ((rc = tempAlloc((void**)&d_tasksB, sizeof(CollapseTask) * maxNodes))) return rc;
    if ((rc = tempAlloc((void**)&d_ctr, sizeof(uint32_t) * 260))) return rc;
    std::vector<uint32_t> ctrInit(260, 0u);
    ctrInit[0] = 1;  // root occupies node 0
    SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(d_ctr, ctrInit.data(), sizeof(uint32_t) * 260, cudaMemcpyHostToDevice, st));
    CollapseTask root;
    root.bvh2 = (n == 1) ? 0u : 0u;  // internal node 0, or the only leaf (id n-1+0 = 0)
    root.node = 0;
    SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(d_tasksA, &root, sizeof(root), cudaMemcpyHostToDevice, st));
    uint32_t taskCount = 1, level = 0;
    while (taskCount > 0) {
        if (level >= 256) {
            setError(ctx, "BVH collapse exceeded 256 levels");
            return SELAS_ERR_UNSUPPORTED;
        }
        collapseLevelKernel<<<divUp(taskCount, 128), 128, 0, st>>>(d_tasksA, taskCount, (int)n, d_children, d_range, d_boxLo, d_boxHi, d_tris,
                                                                    sorted, d_nodes, d_outTris, d_ctr, d_ctr + 1, d_tasksB, d_ctr + 2 + level);
        SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(&taskCount, d_ctr + 2 + level, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
        std::swap(d_tasksA, d_tasksB);
        ++level;
    }
    uint32_t finalCtr[2];
    SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(finalCtr, d_ctr, sizeof(finalCtr), cudaMemcpyDeviceToHost, st));
    cudaEventRecord(ev1, st);
    SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    if (finalCtr[1] != n) {
        setError(ctx, "BVH collapse emitted %u of %u triangles", finalCtr[1], n);
        return SELAS_ERR_CUDA;
    }
    scene->gpu.nodes = d_nodes;
    scene->gpu.tris = d_outTris;
    scene->gpu.nodeCount = finalCtr[0];
    scene->gpu.triCount = n;
    scene->stats.nodeCount = finalCtr[0];
    scene->stats.maxDepth = level;
    scene->stats.buildMs = ms;
    scene->stats.deviceBytes = (uint64_t)finalCtr[0] * 80ull + (uint64_t)n * 48ull;
    return SELAS_OK;
}

static Bvh8 bvhOf(const SelasScene* s) {
    Bvh8 b;
    b.nodes = s->gpu.nodes;
    b.tris = s->gpu.tris;
    b.nodeCount = s->gpu.nodeCount;
    b.triCount = s->gpu.triCount;
    return b;
}

static int traceGrid(const SelasContext* ctx) { return ctx->smCount * SELAS_TRACE_CTAS_PER_SM; }

int launchIntersect(SelasScene* scene, uint32_t n, const SelasRay* d_rays, SelasHit* d_hits, float* d_ng4, uint32_t* d_counts,
                    uint32_t* d_rayCounter, cudaStream_t stream) {
    if (n == 0) return SELAS_OK;
    SelasContext* ctx = scene->ctx;
    SELAS_CUDA_CHECK(ctx, cudaMemsetAsync(d_rayCounter, 0, sizeof(uint32_t), stream));
    int grid = traceGrid(ctx);
    if (d_counts)
        traceKernel<0, true><<<grid, SELAS_TRACE_THREADS, 0, stream>>>(bvhOf(scene), d_rays, n, nullptr, d_hits, nullptr, (float4*)d_ng4, d_counts,
                                                                        nullptr, nullptr, d_rayCounter, nullptr);
    else
        traceKernel<0, false><<<grid, SELAS_TRACE_THREADS, 0, stream>>>(bvhOf(scene), d_rays, n, nullptr, d_hits, nullptr, (float4*)d_ng4, nullptr,
                                                                         nullptr, nullptr, d_rayCounter, nullptr);
    SELAS_CUDA_CHECK(ctx, cudaGetLastError());
    return SELAS_OK;
}

int launchOccluded(SelasScene* scene, uint32_t n, const SelasRay* d_rays, uint8_t* d_occ, uint32_t* d_counts, uint32_t* d_rayCounter,
                   cudaStream_t stream) {
    if (n == 0) return SELAS_OK;
    SelasContext* ctx = scene->ctx;
    SELAS_CUDA_CHECK(ctx, cudaMemsetAsync(d_rayCounter, 0, sizeof(uint32_t), stream));
    int grid = traceGrid(ctx);
    if (d_counts)
        traceKernel<1, true><<<grid, SELAS_TRACE_THREADS, 0, stream>>>(bvhOf(scene), d_rays, n, nullptr, nullptr, d_occ, nullptr, d_counts, nullptr,
                                                                        nullptr, d_rayCounter, nullptr);
    else
        traceKernel<1, false><<<grid, SELAS_TRACE_THREADS, 0, stream>>>(bvhOf(scene), d_rays, n, nullptr, nullptr, d_occ, nullptr, nullptr, nullptr,
                                                                         nullptr, d_rayCounter, nullptr);
    SELAS_CUDA_CHECK(ctx, cudaGetLastError());
    return SELAS_OK;
}

// Wavefront variants: the queue length lives in device memory (or n is exact when d_n is null); d_rayCounter must
// already be zero.  totals != null selects the counting build of the same kernel (SELAS_RENDER_COUNT_TRAVERSAL).
int launchIntersectIndirect(SelasScene* scene, const uint32_t* d_n, uint32_t maxN, const SelasRay* d_rays, SelasHit* d_hits,
                            uint32_t* d_rayCounter, unsigned long long* d_totals, cudaStream_t stream) {
    SelasContext* ctx = scene->ctx;
    if (d_totals)
        traceKernel<0, true><<<traceGrid(ctx), SELAS_TRACE_THREADS, 0, stream>>>(bvhOf(scene), d_rays, maxN, d_n, d_hits, nullptr, nullptr, nullptr,
                                                                                  nullptr, nullptr, d_rayCounter, d_totals);
    else
        traceKernel<0, false><<<traceGrid(ctx), SELAS_TRACE_THREADS, 0, stream>>>(bvhOf(scene), d_rays, maxN, d_n, d_hits, nullptr, nullptr, nullptr,
                                                                                   nullptr, nullptr, d_rayCounter, nullptr);
    SELAS_CUDA_CHECK(ctx, cudaGetLastError());
    return SELAS_OK;
}

int launchShadowAccumulate(SelasScene* scene, const uint32_t* d_n, uint32_t maxN, const SelasRay* d_rays, const float4* d_payload,
                           float* d_film, uint32_t* d_rayCounter, unsigned long long* d_totals, cudaStream_t stream) {
    SelasContext* ctx = scene->ctx;
    if (d_totals)
        traceKernel<2, true><<<traceGrid(ctx), SELAS_TRACE_THREADS, 0, stream>>>(bvhOf(scene), d_rays, maxN, d_n, nullptr, nullptr, nullptr, nullptr,
                                                                                  d_payload, d_film, d_rayCounter, d_totals);
    else
        traceKernel<2, false><<<traceGrid(ctx), SELAS_TRACE_THREADS, 0, stream>>>(bvhOf(scene), d_rays, maxN, d_n, nullptr, nullptr, nullptr, nullptr,
                                                                                   d_payload, d_film, d_rayCounter, nullptr);
    SELAS_CUDA_CHECK(ctx, cudaGetLastError());
    return SELAS_OK;
}

}  // namespace selas
