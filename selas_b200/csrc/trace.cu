// trace.cu -- BVH8 build driver and ray-cast launchers (translation unit compiled with FMA contraction on).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "bvh_build.cuh"
#include "bvh_sah.cuh"
#include "internal.h"
#include "trace.cuh"
#include "trace_coop.cuh"

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
// build primitives per face: a quad is two triangles, a curve segment SELAS_CURVE_PARTS cones (curve.cuh)
inline uint32_t primitivesPerFace(uint32_t indicesPerFace) { return indicesPerFace == 4 ? 2u : (indicesPerFace == 1 ? (uint32_t)SELAS_CURVE_PARTS : 1u); }

}  // namespace

// One BVH8 over n BuildTri records (Morton sort -> Karras radix tree -> refit -> per-level collapse).  out.nodes / out.tris are
// plain cudaMalloc allocations owned by the caller; node and triangle indices inside them are relative to their own start.
struct TreeOut {
    float4* nodes = nullptr;
    float4* tris = nullptr;
    uint32_t nodeCount = 0, triCount = 0, levels = 0;
    std::vector<uint32_t> levelEnds;  // node count after each level: nodes [0, levelEnds[L]) are levels 0..L (breadth-first numbering)
};

// ---- binned-SAH tree over the Morton clusters (bvh_sah.cuh) -------------------------------------------------------------------
// Runs generations of sahSplitKernel until no segment is left.  items[0] holds the segments of generation 0; the buffers alternate.
static int runSahGenerations(SelasContext* ctx, cudaStream_t st, SahItem* items[2], SahSeg* segs[2], uint32_t segCount, uint32_t* d_genCounts /* 256, zeroed */,
                             const SahTree& tree, uint32_t* generations) {
    uint32_t gen = 0;
    while (segCount > 0) {
        if (gen >= 255) {
            setError(ctx, "SAH build exceeded 255 generations");
            return SELAS_ERR_UNSUPPORTED;
        }
        sahSplitKernel<<<divUp(segCount, kSahWarpsPerCta), kSahWarpsPerCta * 32, 0, st>>>(items[gen & 1], items[(gen + 1) & 1], segs[gen & 1], segCount,
                                                                                          segs[(gen + 1) & 1], d_genCounts + gen + 1, tree);
        SELAS_CUDA_CHECK(ctx, cudaGetLastError());
        SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(&segCount, d_genCounts + gen + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
        ++gen;
    }
    if (generations) *generations = gen;
    return SELAS_OK;
}

struct Bvh2Arrays {  // the binary tree the collapse consumes
    const int2* children;
    const uint2* range;
    const float4* boxLo;
    const float4* boxHi;
    const uint32_t* sorted;
    uint32_t root;
};

// radix-tree arrays in, SAH tree out (device arrays registered with `tmp`).  n >= 1.
static int buildSahTree(SelasContext* ctx, cudaStream_t st, TempFree& tmp, const BuildTri* d_tris, uint32_t n, const uint32_t* d_sorted,
                        const uint2* d_range, const uint32_t* d_parent, const float4* d_boxLo, const float4* d_boxHi, uint32_t clusterMax,
                        Bvh2Arrays& out, SelasBvhStats* stats) {
    auto tempAlloc = [&](void** p, size_t bytes) -> int {
        cudaError_t e = cudaMalloc(p, bytes + 16);
        if (e != cudaSuccess) {
            setError(ctx, "cudaMalloc(%zu bytes) failed during the SAH build: %s", bytes, cudaGetErrorString(e));
            return SELAS_ERR_MEMORY;
        }
        tmp.ptrs.push_back(*p);
        return SELAS_OK;
    };
    int rc;
    uint32_t* d_ctr;  // [0] cluster count, [1] seed count, [2] top root, [3] final root, [8..264) top generations, [264..520) bottom generations
    if ((rc = tempAlloc((void**)&d_ctr, sizeof(uint32_t) * 520))) return rc;
    SELAS_CUDA_CHECK(ctx, cudaMemsetAsync(d_ctr, 0, sizeof(uint32_t) * 520, st));
    // ---- clusters: count, then fill ----
    const uint32_t total = 2u * n - 1u;
    sahClusterRootsKernel<<<divUp(total, 256), 256, 0, st>>>((int)n, clusterMax, d_range, d_parent, d_boxLo, d_boxHi, nullptr, nullptr, d_ctr, 0u);
    uint32_t C = 0;
    SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(&C, d_ctr, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    if (C == 0) {
        setError(ctx, "SAH build found no clusters over %u primitives", n);
        return SELAS_ERR_CUDA;
    }
    SahItem* topItems[2];
    uint2* d_clusters;
    if ((rc = tempAlloc((void**)&topItems[0], sizeof(SahItem) * (size_t)C))) return rc;
    if ((rc = tempAlloc((void**)&topItems[1], sizeof(SahItem) * (size_t)C))) return rc;
    if ((rc = tempAlloc((void**)&d_clusters, sizeof(uint2) * (size_t)C))) return rc;
    SELAS_CUDA_CHECK(ctx, cudaMemsetAsync(d_ctr, 0, sizeof(uint32_t), st));
    sahClusterRootsKernel<<<divUp(total, 256), 256, 0, st>>>((int)n, clusterMax, d_range, d_parent, d_boxLo, d_boxHi, topItems[0], d_clusters, d_ctr, C);
    SELAS_CUDA_CHECK(ctx, cudaGetLastError());

    // ---- top tree over the clusters ----
    SahTree top{};
    uint32_t* d_topOrder;
    if ((rc = tempAlloc((void**)&top.children, sizeof(int2) * (size_t)C))) return rc;
    if ((rc = tempAlloc((void**)&top.range, sizeof(uint2) * (size_t)C))) return rc;
    if ((rc = tempAlloc((void**)&top.boxLo, sizeof(float4) * 2 * (size_t)C))) return rc;
    if ((rc = tempAlloc((void**)&top.boxHi, sizeof(float4) * 2 * (size_t)C))) return rc;
    if ((rc = tempAlloc((void**)&d_topOrder, sizeof(uint32_t) * (size_t)C))) return rc;
    top.order = d_topOrder;
    top.leafBase = C - 1;
    top.rootId = d_ctr + 2;
    SahSeg* segs[2];
    const size_t segCap = std::max<size_t>((size_t)n / 2 + 2, (size_t)C + 2);
    if ((rc = tempAlloc((void**)&segs[0], sizeof(SahSeg) * segCap))) return rc;
    if ((rc = tempAlloc((void**)&segs[1], sizeof(SahSeg) * segCap))) return rc;
    uint32_t topGenerations = 0, bottomGenerations = 0;
    if (C >= 2) {
        const SahSeg rootSeg{0u, C, kSahNone, 0u};
        SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(segs[0], &rootSeg, sizeof(rootSeg), cudaMemcpyHostToDevice, st));
        if ((rc = runSahGenerations(ctx, st, topItems, segs, 1u, d_ctr + 8, top, &topGenerations))) return rc;
    } else {
        const uint32_t zero = 0;
        SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(d_topOrder, &zero, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    }
    // ---- place the clusters in top-tree leaf order ----
    uint32_t *d_counts, *d_excl, *d_incl, *d_sums;
    if ((rc = tempAlloc((void**)&d_counts, sizeof(uint32_t) * ((size_t)C + 1)))) return rc;
    if ((rc = tempAlloc((void**)&d_excl, sizeof(uint32_t) * ((size_t)C + 1)))) return rc;
    if ((rc = tempAlloc((void**)&d_incl, sizeof(uint32_t) * ((size_t)C + 1)))) return rc;
    if ((rc = tempAlloc((void**)&d_sums, sizeof(uint32_t) * (size_t)scanScratchWords(C + 1)))) return rc;
    SELAS_CUDA_CHECK(ctx, cudaMemsetAsync(d_counts, 0, sizeof(uint32_t) * ((size_t)C + 1), st));
    sahOrderedCountsKernel<<<divUp(C, 256), 256, 0, st>>>(C, d_topOrder, d_clusters, d_counts);
    if ((rc = launchScanU32(ctx, d_counts, C + 1, d_sums, d_incl, d_excl, st))) return rc;  // excl[C] = n: the graft reads excl[k + 1]
    // ---- the final tree ----
    SahItem* items[2];
    if ((rc = tempAlloc((void**)&items[0], sizeof(SahItem) * (size_t)n))) return rc;
    if ((rc = tempAlloc((void**)&items[1], sizeof(SahItem) * (size_t)n))) return rc;
    SahTree fin{};
    uint32_t* d_order;
    if ((rc = tempAlloc((void**)&fin.children, sizeof(int2) * (size_t)n))) return rc;
    if ((rc = tempAlloc((void**)&fin.range, sizeof(uint2) * (size_t)n))) return rc;
    if ((rc = tempAlloc((void**)&fin.boxLo, sizeof(float4) * 2 * (size_t)n))) return rc;
    if ((rc = tempAlloc((void**)&fin.boxHi, sizeof(float4) * 2 * (size_t)n))) return rc;
    if ((rc = tempAlloc((void**)&d_order, sizeof(uint32_t) * (size_t)n))) return rc;
    fin.order = d_order;
    fin.leafBase = n - 1;
    fin.rootId = d_ctr + 3;
    sahSeedItemsKernel<<<divUp(n, 256), 256, 0, st>>>(n, C, d_topOrder, d_clusters, d_excl, d_sorted, d_tris, items[0]);
    sahGraftTopKernel<<<divUp(std::max(C, 2u) - 1, 256), 256, 0, st>>>(n, C, top.children, top.range, top.boxLo, top.boxHi, d_excl, d_counts, items[0], fin,
                                                                       segs[0], d_ctr + 1, d_ctr + 2);
    SELAS_CUDA_CHECK(ctx, cudaGetLastError());
    uint32_t seedCount = 0;
    SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(&seedCount, d_ctr + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    if ((rc = runSahGenerations(ctx, st, items, segs, seedCount, d_ctr + 264, fin, &bottomGenerations))) return rc;
    uint32_t root = 0;
    SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(&root, d_ctr + 3, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    out.children = fin.children;
    out.range = fin.range;
    out.boxLo = fin.boxLo;
    out.boxHi = fin.boxHi;
    out.sorted = d_order;
    out.root = root;
    if (stats) {
        stats->sahClusters += C;
        stats->sahGenerations += topGenerations + bottomGenerations;
    }
    return SELAS_OK;
}

static int buildTree(SelasScene* scene, const BuildTri* d_tris, const uint32_t* d_bounds, uint32_t n, TreeOut& out, bool topLevel = false) {
    SelasContext* ctx = scene->ctx;
    cudaStream_t st = ctx->stream;
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
    uint64_t *d_keys, *d_keysAlt;
    uint32_t *d_vals, *d_valsAlt;
    int rc;
    if ((rc = tempAlloc((void**)&d_keys, sizeof(uint64_t) * (size_t)n))) return rc;
    if ((rc = tempAlloc((void**)&d_keysAlt, sizeof(uint64_t) * (size_t)n))) return rc;
    if ((rc = tempAlloc((void**)&d_vals, sizeof(uint32_t) * (size_t)n))) return rc;
    if ((rc = tempAlloc((void**)&d_valsAlt, sizeof(uint32_t) * (size_t)n))) return rc;
    static const int mortonCube = getenv("SELAS_MORTON_CUBE") ? atoi(getenv("SELAS_MORTON_CUBE")) : 1;  // A/B switch; cubic cells by default
    mortonKernel<<<divUp(n, 256), 256, 0, st>>>(d_tris, n, d_bounds, d_keys, d_vals, mortonCube);
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

    // The binary tree to collapse: the radix tree itself (SELAS_BVH_BUILD=lbvh), or -- the default -- a binned-SAH tree whose top is built
    // over the radix tree's clusters of <= SELAS_SAH_CLUSTER primitives and whose clusters are rebuilt by binned SAH (bvh_sah.cuh).
    static const bool useSah = !(getenv("SELAS_BVH_BUILD") && !strcmp(getenv("SELAS_BVH_BUILD"), "lbvh"));
    static const uint32_t clusterMax = getenv("SELAS_SAH_CLUSTER") ? (uint32_t)std::max(2, atoi(getenv("SELAS_SAH_CLUSTER"))) : 1024u;
    Bvh2Arrays b2{d_children, d_range, d_boxLo, d_boxHi, sorted, 0u};
    if (useSah && !topLevel) {
        if ((rc = buildSahTree(ctx, st, tmp, d_tris, n, sorted, d_range, d_parent, d_boxLo, d_boxHi, clusterMax, b2, &scene->stats))) return rc;
    }

    // ---- collapse ----
    // Wide nodes other than the root cover > 3 triangles.  Leaf wide nodes (no inner child) are disjoint: L <= n/4;
    // branching nodes B < L; chain nodes (one inner child) own >= 7 direct triangles: 4L + 7C <= n.  So the node
    // count is < 2L + C <= n/2.  The final triangle array has exactly n entries.
    const size_t maxNodes = (size_t)n / 2 + 2;
    float4 *d_nodes = nullptr, *d_outTris = nullptr;
    if (cudaMalloc((void**)&d_nodes, maxNodes * 5 * sizeof(float4) + 16) != cudaSuccess ||
        cudaMalloc((void**)&d_outTris, (size_t)n * 3 * sizeof(float4) + 16) != cudaSuccess) {
        cudaFree(d_nodes);
        setError(ctx, "cudaMalloc failed for the BVH8 arrays (%u primitives)", n);
        return SELAS_ERR_MEMORY;
    }
    out.nodes = d_nodes;
    out.tris = d_outTris;
    // cost-optimal slot assignment (bvh_build.cuh: collapseCostKernel) for wide nodes over at most SELAS_COLLAPSE_DP_MAX primitives, the
    // largest-box-first expansion above them (and everywhere with SELAS_COLLAPSE=greedy): see the measurement note at collapseCostKernel
    static const bool greedyCollapse = getenv("SELAS_COLLAPSE") && !strcmp(getenv("SELAS_COLLAPSE"), "greedy");
    static const uint32_t collapseDpMax = getenv("SELAS_COLLAPSE_DP_MAX") ? (uint32_t)atoi(getenv("SELAS_COLLAPSE_DP_MAX")) : 32768u;
    static const float cPrim = getenv("SELAS_COLLAPSE_CPRIM") ? (float)atof(getenv("SELAS_COLLAPSE_CPRIM")) : 0.3f;
    uint32_t* d_decision = nullptr;
    if (!greedyCollapse && n > 1) {
        uint32_t *d_parent2, *d_arrivals2;
        float* d_cost7;
        if ((rc = tempAlloc((void**)&d_parent2, sizeof(uint32_t) * (2 * (size_t)n - 1)))) return rc;
        if ((rc = tempAlloc((void**)&d_arrivals2, sizeof(uint32_t) * (size_t)n))) return rc;
        if ((rc = tempAlloc((void**)&d_cost7, sizeof(float) * 7 * (size_t)n))) return rc;
        if ((rc = tempAlloc((void**)&d_decision, sizeof(uint32_t) * (size_t)n))) return rc;
        SELAS_CUDA_CHECK(ctx, cudaMemsetAsync(d_arrivals2, 0, sizeof(uint32_t) * (size_t)n, st));
        bvh2ParentsKernel<<<divUp(n - 1, 256), 256, 0, st>>>((int)n, b2.children, b2.root, d_parent2);
        collapseCostKernel<<<divUp(n, 256), 256, 0, st>>>((int)n, b2.children, b2.range, d_parent2, b2.boxLo, b2.boxHi, 1.0f, cPrim, d_cost7, d_decision,
                                                           d_arrivals2);
        SELAS_CUDA_CHECK(ctx, cudaGetLastError());
    }
    CollapseTask *d_tasksA, *d_tasksB;
    uint32_t* d_ctr;  // [0]=nodeCounter [1]=triCounter [2..]=per-level next counts
    if ((rc = tempAlloc((void**)&d_tasksA, sizeof(CollapseTask) * maxNodes))) return rc;
    if ((rc = tempAlloc((void**)&d_tasksB, sizeof(CollapseTask) * maxNodes))) return rc;
    if ((rc = tempAlloc((void**)&d_ctr, sizeof(uint32_t) * 260))) return rc;
    std::vector<uint32_t> ctrInit(260, 0u);
    ctrInit[0] = 1;  // root occupies node 0
    SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(d_ctr, ctrInit.data(), sizeof(uint32_t) * 260, cudaMemcpyHostToDevice, st));
    CollapseTask root;
    root.bvh2 = b2.root;  // radix tree: internal node 0, or the only leaf (id n-1+0 = 0); SAH tree: the id its root segment reported
    root.node = 0;
    SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(d_tasksA, &root, sizeof(root), cudaMemcpyHostToDevice, st));
    uint32_t taskCount = 1, level = 0, nodesSoFar = 0;
    while (taskCount > 0) {
        nodesSoFar += taskCount;
        out.levelEnds.push_back(nodesSoFar);
        if (level >= 256) {
            setError(ctx, "BVH collapse exceeded 256 levels");
            return SELAS_ERR_UNSUPPORTED;
        }
        collapseLevelKernel<<<divUp(taskCount, 128), 128, 0, st>>>(d_tasksA, taskCount, (int)n, b2.children, b2.range, b2.boxLo, b2.boxHi, d_tris,
                                                                    b2.sorted, d_nodes, d_outTris, d_ctr, d_ctr + 1, d_tasksB, d_ctr + 2 + level, d_decision, collapseDpMax);
        SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(&taskCount, d_ctr + 2 + level, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
        std::swap(d_tasksA, d_tasksB);
        ++level;
    }
    uint32_t finalCtr[2];
    SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(finalCtr, d_ctr, sizeof(finalCtr), cudaMemcpyDeviceToHost, st));
    SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    if (finalCtr[1] != n) {
        setError(ctx, "BVH collapse emitted %u of %u triangles", finalCtr[1], n);
        return SELAS_ERR_CUDA;
    }
    out.nodeCount = finalCtr[0];
    out.triCount = n;
    out.levels = level;

    // ---- depth-first relayout (bvh_build.cuh), SELAS_BVH_LAYOUT=dfs.  Measured neutral (configs[2] 15.65 against 15.69 ms per step, configs[1]
    // within noise for camera, shadow and incoherent rays: profiles/r2_ab_layout.jsonl) -- the level order already keeps each level in the
    // tree's spatial order and the kernels are bound by issue slots, not by where the nodes lie -- so the collapse's order is kept by default.
    static const bool keepLevelOrder = !(getenv("SELAS_BVH_LAYOUT") && !strcmp(getenv("SELAS_BVH_LAYOUT"), "dfs"));
    if (!keepLevelOrder && out.nodeCount > 1) {
        const uint32_t N = out.nodeCount;
        uint32_t *d_below, *d_newIndex, *d_descStart, *d_leafCounts, *d_triStart;
        if ((rc = tempAlloc((void**)&d_below, sizeof(uint32_t) * (size_t)N))) return rc;
        if ((rc = tempAlloc((void**)&d_newIndex, sizeof(uint32_t) * (size_t)N))) return rc;
        if ((rc = tempAlloc((void**)&d_descStart, sizeof(uint32_t) * (size_t)N))) return rc;
        if ((rc = tempAlloc((void**)&d_leafCounts, sizeof(uint32_t) * (size_t)N))) return rc;
        if ((rc = tempAlloc((void**)&d_triStart, sizeof(uint32_t) * (size_t)N))) return rc;
        float4 *d_nodes2 = nullptr, *d_tris2 = nullptr;
        if (cudaMalloc((void**)&d_nodes2, (size_t)N * 5 * sizeof(float4) + 16) != cudaSuccess ||
            cudaMalloc((void**)&d_tris2, (size_t)n * 3 * sizeof(float4) + 16) != cudaSuccess) {
            cudaFree(d_nodes2);
            cudaGetLastError();  // out of memory for the second copy: keep the level order
        } else {
            for (int L = (int)out.levelEnds.size() - 1; L >= 0; --L) {
                const uint32_t b = L ? out.levelEnds[L - 1] : 0u, e = out.levelEnds[L];
                if (e > b) relayoutSubtreeKernel<<<divUp(e - b, 256), 256, 0, st>>>(d_nodes, b, e, d_below);
            }
            for (size_t L = 0; L < out.levelEnds.size(); ++L) {
                const uint32_t b = L ? out.levelEnds[L - 1] : 0u, e = out.levelEnds[L];
                if (e > b) relayoutPlaceKernel<<<divUp(e - b, 256), 256, 0, st>>>(d_nodes, b, e, d_below, d_newIndex, d_descStart);
            }
            relayoutLeafCountKernel<<<divUp(N, 256), 256, 0, st>>>(d_nodes, N, d_newIndex, d_leafCounts);
            size_t scanBytes = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, scanBytes, d_leafCounts, d_triStart, (int)N, st);
            void* d_scanTemp;
            if ((rc = tempAlloc(&d_scanTemp, scanBytes))) { cudaFree(d_nodes2); cudaFree(d_tris2); return rc; }
            cub::DeviceScan::ExclusiveSum(d_scanTemp, scanBytes, d_leafCounts, d_triStart, (int)N, st);
            relayoutScatterKernel<<<divUp(N, 128), 128, 0, st>>>(d_nodes, d_outTris, N, d_newIndex, d_descStart, d_triStart, d_nodes2, d_tris2);
            SELAS_CUDA_CHECK(ctx, cudaGetLastError());
            SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
            cudaFree(d_nodes);
            cudaFree(d_outTris);
            out.nodes = d_nodes2;
            out.tris = d_tris2;
            out.levelEnds.clear();  // the first nodes are no longer whole levels: nothing to stage in shared memory
        }
    }
    return SELAS_OK;
}

// Whole levels of the tree (top-level tree for the two-level form) that fit the shared-memory budget of the trace kernels:
// SELAS_TRACE_TOP_NODES nodes of 80 bytes per CTA (default 0 = no staging: see the measurement note in trace.cuh).
static uint32_t chooseTopNodeCount(const std::vector<uint32_t>& levelEnds) {
    static const uint32_t budget = getenv("SELAS_TRACE_TOP_NODES") ? (uint32_t)std::max(0, atoi(getenv("SELAS_TRACE_TOP_NODES"))) : 0u;
    uint32_t m = 0;
    for (uint32_t e : levelEnds)
        if (e <= budget && e * 80u <= 200u * 1024u) m = e;
    return m;
}

// (the two-level variant of the kernel runs only when the scene was built that way: the flattened form keeps its register budget)
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
    uint64_t total = 0, curveSegments = 0;
    for (uint32_t ii = 0; ii < inst.size(); ++ii)
        for (uint32_t m = 0; m < meshes.size(); ++m) {
            if (inst[ii].mesh != 0xFFFFFFFFu && inst[ii].mesh != m) continue;
            uint32_t t = meshes[m].faceCount * primitivesPerFace(meshes[m].indicesPerFace);
            if (t == 0) continue;
            draws.push_back({m, ii, (uint32_t)total, t});
            total += t;
            if (meshes[m].indicesPerFace == 1) curveSegments += meshes[m].faceCount;  // a curve geometry: one pool entry per segment
        }
    if (total >= 0x3fffffffull || inst.size() >= 0x3fffffffull) {
        setError(ctx, "scene has %llu primitives; the builder is limited to 2^30-1", (unsigned long long)total);
        return SELAS_ERR_UNSUPPORTED;
    }
    const uint32_t n = (uint32_t)total;
    scene->stats = SelasBvhStats{};
    scene->stats.triangleCount = n;
    scene->gpu.nodes = nullptr;
    scene->gpu.tris = nullptr;
    scene->gpu.nodeCount = 0;
    scene->gpu.triCount = 0;
    scene->gpu.traceInstances = nullptr;
    scene->gpu.curvePool = nullptr;
    scene->topNodeCount = 0;
    if (n == 0) return SELAS_OK;

    TempFree tmp;
    BuildTri* d_tris = nullptr;
    uint32_t* d_bounds = nullptr;
    if (cudaMalloc((void**)&d_tris, sizeof(BuildTri) * (size_t)n + 16) != cudaSuccess || cudaMalloc((void**)&d_bounds, 6 * sizeof(uint32_t)) != cudaSuccess) {
        cudaFree(d_tris);
        setError(ctx, "cudaMalloc failed during BVH build (%u triangles)", n);
        return SELAS_ERR_MEMORY;
    }
    tmp.ptrs.push_back(d_tris);
    tmp.ptrs.push_back(d_bounds);
    uint32_t boundsInit[6] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u};
    SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(d_bounds, boundsInit, sizeof(boundsInit), cudaMemcpyHostToDevice, st));
    float4* d_pool = nullptr;
    if (curveSegments) {
        int rcPool = devAlloc(scene, &d_pool, (size_t)curveSegments * 4, true);
        if (rcPool) return rcPool;
    }
    uint32_t poolBase = 0;
    for (const Draw& d : draws) {
        MeshDev md{meshes[d.mesh].positions, meshes[d.mesh].indices, meshes[d.mesh].faceCount, meshes[d.mesh].indicesPerFace};
        if (md.indicesPerFace == 1) {
            gatherCurvesKernel<<<divUp(md.faceCount, 256), 256, 0, st>>>(md, d_instances[d.inst].m, inst[d.inst].identity != 0, d.mesh, d.inst, d_tris,
                                                                          d.base, d_pool, poolBase, d_bounds);
            poolBase += md.faceCount;
        } else
            gatherTrianglesKernel<<<divUp(md.faceCount, 256), 256, 0, st>>>(md, d_instances[d.inst].m, inst[d.inst].identity != 0, d.mesh, d.inst,
                                                                             d_tris, d.base, d_bounds);
    }
    scene->gpu.curvePool = d_pool;
    TreeOut tree;
    int rc = buildTree(scene, d_tris, d_bounds, n, tree);
    if (tree.nodes) scene->allocations.push_back(tree.nodes);
    if (tree.tris) scene->allocations.push_back(tree.tris);
    if (rc) return rc;
    cudaEventRecord(ev1, st);
    SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    {
        uint32_t hb[6];
        SELAS_CUDA_CHECK(ctx, cudaMemcpy(hb, d_bounds, sizeof(hb), cudaMemcpyDeviceToHost));
        for (int k = 0; k < 3; ++k) scene->boundsCenter[k] = 0.5f * (floatUnflip(hb[k]) + floatUnflip(hb[3 + k]));
    }
    scene->gpu.nodes = tree.nodes;
    scene->gpu.tris = tree.tris;
    scene->gpu.nodeCount = tree.nodeCount;
    scene->gpu.triCount = n;
    scene->stats.nodeCount = tree.nodeCount;
    scene->stats.maxDepth = tree.levels;
    scene->topNodeCount = d_pool ? 0u : chooseTopNodeCount(tree.levelEnds);
    if (tree.levels > SELAS_STACK_SIZE) {  // one stack entry per level at most: a deeper tree would drop pushes silently
        setError(ctx, "BVH8 has %u levels; the traversal stack holds %d", tree.levels, SELAS_STACK_SIZE);
        return SELAS_ERR_UNSUPPORTED;
    }
    scene->stats.buildMs = ms;
    scene->stats.deviceBytes = (uint64_t)tree.nodeCount * 80ull + (uint64_t)n * 48ull;
    return SELAS_OK;
}

// ---- two-level build (SELAS_SCENE_TWO_LEVEL_BVH) -------------------------------------------------------------------------------
// One object-space BVH8 per mesh that some instance uses, and a top-level BVH8 over the instances' world-space boxes, all in one node
// array and one triangle array: [top level][mesh 0][mesh 1]...  Geometry is stored once however many instances refer to it -- what the
// reference gets from rtcNewGeometry(RTC_GEOMETRY_TYPE_INSTANCE) / its user-geometry subscene instances (SubsceneResource.cpp:63-79,
// SceneResource.cpp:218-249).  The top level reuses the triangle builder: instance i is fed as a degenerate "triangle" (lo, hi, lo)
// whose bounds are the instance's box and whose primID is i.
static uint32_t hostFloatFlip(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

int buildBvhTwoLevel(SelasScene* scene, const std::vector<MeshGpu>& meshes, const std::vector<InstanceGpu>& inst) {
    SelasContext* ctx = scene->ctx;
    cudaStream_t st = ctx->stream;
    cudaEvent_t ev0, ev1;
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    cudaEventRecord(ev0, st);
    scene->stats = SelasBvhStats{};
    scene->gpu.nodes = nullptr;
    scene->gpu.tris = nullptr;
    scene->gpu.nodeCount = 0;
    scene->gpu.triCount = 0;
    scene->gpu.traceInstances = nullptr;
    scene->gpu.curvePool = nullptr;
    scene->topNodeCount = 0;

    struct Blas {
        TreeOut tree;
        float lo[3], hi[3];
        uint32_t nodeBase = 0, triBase = 0;
        bool used = false;
    };
    std::vector<Blas> blas(meshes.size());
    TreeOut top;
    struct TreeGuard {  // the per-tree arrays are temporaries: everything is copied into the two shared arrays
        std::vector<Blas>& b;
        TreeOut& top;
        ~TreeGuard() {
            for (Blas& x : b) {
                cudaFree(x.tree.nodes);
                cudaFree(x.tree.tris);
            }
            cudaFree(top.nodes);
            cudaFree(top.tris);
        }
    } guard{blas, top};
    for (const InstanceGpu& in : inst) {
        if (in.mesh >= meshes.size()) {
            setError(ctx, "two-level BVH: every instance must name its mesh");
            return SELAS_ERR_ARGUMENT;
        }
        blas[in.mesh].used = true;
    }
    uint64_t uniqueTris = 0, instancedTris = 0;
    uint64_t curveSegments2 = 0;
    for (uint32_t m = 0; m < meshes.size(); ++m)
        if (blas[m].used && meshes[m].indicesPerFace == 1) curveSegments2 += meshes[m].faceCount;
    float4* d_pool2 = nullptr;
    uint32_t poolBase2 = 0;
    if (curveSegments2) {
        int rcPool = devAlloc(scene, &d_pool2, (size_t)curveSegments2 * 4, true);
        if (rcPool) return rcPool;
    }
    for (uint32_t m = 0; m < meshes.size(); ++m) {
        const uint32_t t = meshes[m].faceCount * primitivesPerFace(meshes[m].indicesPerFace);
        if (!blas[m].used || t == 0) {
            blas[m].used = false;
            continue;
        }
        TempFree tmp;
        BuildTri* d_tris = nullptr;
        uint32_t* d_bounds = nullptr;
        if (cudaMalloc((void**)&d_tris, sizeof(BuildTri) * (size_t)t + 16) != cudaSuccess || cudaMalloc((void**)&d_bounds, 6 * sizeof(uint32_t)) != cudaSuccess) {
            cudaFree(d_tris);
            setError(ctx, "cudaMalloc failed during the BVH build of mesh %u (%u triangles)", m, t);
            return SELAS_ERR_MEMORY;
        }
        tmp.ptrs.push_back(d_tris);
        tmp.ptrs.push_back(d_bounds);
        uint32_t boundsInit[6] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u};
        SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(d_bounds, boundsInit, sizeof(boundsInit), cudaMemcpyHostToDevice, st));
        MeshDev md{meshes[m].positions, meshes[m].indices, meshes[m].faceCount, meshes[m].indicesPerFace};
        if (md.indicesPerFace == 1) {  // curve geometry: object-space control vertices into the pool
            gatherCurvesKernel<<<divUp(md.faceCount, 256), 256, 0, st>>>(md, nullptr, true, m, 0u, d_tris, 0u, d_pool2, poolBase2, d_bounds);
            poolBase2 += md.faceCount;
        } else
            gatherTrianglesKernel<<<divUp(md.faceCount, 256), 256, 0, st>>>(md, nullptr, true, m, 0u, d_tris, 0u, d_bounds);
        int rc = buildTree(scene, d_tris, d_bounds, t, blas[m].tree);
        if (rc) return rc;
        uint32_t hb[6];
        SELAS_CUDA_CHECK(ctx, cudaMemcpy(hb, d_bounds, sizeof(hb), cudaMemcpyDeviceToHost));
        for (int k = 0; k < 3; ++k) {
            blas[m].lo[k] = floatUnflip(hb[k]);
            blas[m].hi[k] = floatUnflip(hb[3 + k]);
        }
        uniqueTris += t;
    }
    // top level: one degenerate triangle per instance whose mesh has geometry
    std::vector<BuildTri> fake;
    float tlo[3] = {SELAS_FLOAT_MAX, SELAS_FLOAT_MAX, SELAS_FLOAT_MAX}, thi[3] = {-SELAS_FLOAT_MAX, -SELAS_FLOAT_MAX, -SELAS_FLOAT_MAX};
    for (uint32_t i = 0; i < inst.size(); ++i) {
        const Blas& b = blas[inst[i].mesh];
        if (!b.used) continue;
        float lo[3] = {SELAS_FLOAT_MAX, SELAS_FLOAT_MAX, SELAS_FLOAT_MAX}, hi[3] = {-SELAS_FLOAT_MAX, -SELAS_FLOAT_MAX, -SELAS_FLOAT_MAX};
        for (int c = 0; c < 8; ++c) {
            V3 p = v3((c & 1) ? b.hi[0] : b.lo[0], (c & 2) ? b.hi[1] : b.lo[1], (c & 4) ? b.hi[2] : b.lo[2]);
            V3 q = mulPoint44(p, inst[i].m);
            const float qq[3] = {q.x, q.y, q.z};
            for (int k = 0; k < 3; ++k) {
                lo[k] = std::min(lo[k], qq[k]);
                hi[k] = std::max(hi[k], qq[k]);
            }
        }
        for (int k = 0; k < 3; ++k) {  // the object-space walk rounds differently from this box: keep a relative margin
            const float pad = 1e-5f * (std::max(std::fabs(lo[k]), std::fabs(hi[k])) + (hi[k] - lo[k])) + 1e-30f;
            lo[k] -= pad;
            hi[k] += pad;
            tlo[k] = std::min(tlo[k], lo[k]);
            thi[k] = std::max(thi[k], hi[k]);
        }
        BuildTri t;
        t.a = make_float4(lo[0], lo[1], lo[2], 0.f);
        t.b = make_float4(hi[0], hi[1], hi[2], 0.f);
        t.c = t.a;
        memcpy(&t.a.w, &i, 4);             // primID = flat instance index
        memcpy(&t.b.w, &inst[i].mesh, 4);  // geomID = mesh
        memcpy(&t.c.w, &i, 4);
        fake.push_back(t);
        instancedTris += meshes[inst[i].mesh].faceCount * primitivesPerFace(meshes[inst[i].mesh].indicesPerFace);
    }
    scene->stats.triangleCount = (uint32_t)std::min<uint64_t>(uniqueTris, 0xFFFFFFFFull);
    if (fake.empty()) return SELAS_OK;
    {
        TempFree tmp;
        BuildTri* d_fake = nullptr;
        uint32_t* d_bounds = nullptr;
        if (cudaMalloc((void**)&d_fake, sizeof(BuildTri) * fake.size() + 16) != cudaSuccess || cudaMalloc((void**)&d_bounds, 6 * sizeof(uint32_t)) != cudaSuccess) {
            cudaFree(d_fake);
            setError(ctx, "cudaMalloc failed during the top-level BVH build (%zu instances)", fake.size());
            return SELAS_ERR_MEMORY;
        }
        tmp.ptrs.push_back(d_fake);
        tmp.ptrs.push_back(d_bounds);
        uint32_t hb[6] = {hostFloatFlip(tlo[0]), hostFloatFlip(tlo[1]), hostFloatFlip(tlo[2]), hostFloatFlip(thi[0]), hostFloatFlip(thi[1]), hostFloatFlip(thi[2])};
        SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(d_bounds, hb, sizeof(hb), cudaMemcpyHostToDevice, st));
        SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(d_fake, fake.data(), sizeof(BuildTri) * fake.size(), cudaMemcpyHostToDevice, st));
        int rc = buildTree(scene, d_fake, d_bounds, (uint32_t)fake.size(), top, true);  // instance boxes: the radix tree is kept
        if (rc) return rc;
    }
    // one node array, one triangle array: [top level][mesh 0][mesh 1]...
    uint64_t totalNodes = top.nodeCount, totalTris = top.triCount;
    for (Blas& b : blas)
        if (b.used) {
            b.nodeBase = (uint32_t)totalNodes;
            b.triBase = (uint32_t)totalTris;
            totalNodes += b.tree.nodeCount;
            totalTris += b.tree.triCount;
        }
    if (totalNodes >= 0x7fffffffull || totalTris >= 0x7fffffffull) {
        setError(ctx, "two-level BVH: %llu nodes / %llu triangles exceed the 2^31-1 index range", (unsigned long long)totalNodes, (unsigned long long)totalTris);
        return SELAS_ERR_UNSUPPORTED;
    }
    float4 *d_nodes, *d_tris;
    int rc;
    if ((rc = devAlloc(scene, &d_nodes, (size_t)totalNodes * 5, true))) return rc;
    if ((rc = devAlloc(scene, &d_tris, (size_t)totalTris * 3, true))) return rc;
    SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(d_nodes, top.nodes, sizeof(float4) * 5 * (size_t)top.nodeCount, cudaMemcpyDeviceToDevice, st));
    SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(d_tris, top.tris, sizeof(float4) * 3 * (size_t)top.triCount, cudaMemcpyDeviceToDevice, st));
    for (const Blas& b : blas)
        if (b.used) {
            SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(d_nodes + (size_t)b.nodeBase * 5, b.tree.nodes, sizeof(float4) * 5 * (size_t)b.tree.nodeCount,
                                                  cudaMemcpyDeviceToDevice, st));
            SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(d_tris + (size_t)b.triBase * 3, b.tree.tris, sizeof(float4) * 3 * (size_t)b.tree.triCount,
                                                  cudaMemcpyDeviceToDevice, st));
        }
    // per instance: world -> object (inverse of the affine localToWorld, in double) and where its mesh's tree starts
    std::vector<InstanceTrace> table(inst.size());
    for (uint32_t i = 0; i < inst.size(); ++i) {
        InstanceTrace& t = table[i];
        memset(&t, 0, sizeof(t));
        const float* m = inst[i].m;
        const double a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], k = m[10];
        const double det = a * (e * k - f * h) - b * (d * k - f * g) + c * (d * h - e * g);
        if (blas[inst[i].mesh].used && !(std::fabs(det) > 0.0)) {
            setError(ctx, "two-level BVH: instance %u has a singular localToWorld", i);
            return SELAS_ERR_ARGUMENT;
        }
        const double id = det != 0.0 ? 1.0 / det : 0.0;
        const double w[9] = {(e * k - f * h) * id, (c * h - b * k) * id, (b * f - c * e) * id,
                             (f * g - d * k) * id, (a * k - c * g) * id, (c * d - a * f) * id,
                             (d * h - e * g) * id, (b * g - a * h) * id, (a * e - b * d) * id};
        for (int r = 0; r < 9; ++r) t.w2l[r] = (float)w[r];
        const double tx = m[12], ty = m[13], tz = m[14];
        t.w2l[9] = (float)-(tx * w[0] + ty * w[3] + tz * w[6]);
        t.w2l[10] = (float)-(tx * w[1] + ty * w[4] + tz * w[7]);
        t.w2l[11] = (float)-(tx * w[2] + ty * w[5] + tz * w[8]);
        t.nodeBase = blas[inst[i].mesh].nodeBase;
        t.triBase = blas[inst[i].mesh].triBase;
        t.mesh = inst[i].mesh;
    }
    InstanceTrace* d_table;
    if ((rc = devAlloc(scene, &d_table, table.size(), true))) return rc;
    SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(d_table, table.data(), sizeof(InstanceTrace) * table.size(), cudaMemcpyHostToDevice, st));
    cudaEventRecord(ev1, st);
    SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    scene->gpu.nodes = d_nodes;
    scene->gpu.tris = d_tris;
    scene->gpu.nodeCount = (uint32_t)totalNodes;
    scene->gpu.triCount = (uint32_t)totalTris;
    scene->gpu.traceInstances = d_table;
    for (int k = 0; k < 3; ++k) scene->boundsCenter[k] = 0.5f * (tlo[k] + thi[k]);
    scene->gpu.curvePool = d_pool2;
    scene->topNodeCount = d_pool2 ? 0u : chooseTopNodeCount(top.levelEnds);
    scene->stats.nodeCount = (uint32_t)totalNodes;
    uint32_t deepestMesh = 0;
    for (const Blas& b : blas)
        if (b.used) deepestMesh = std::max(deepestMesh, b.tree.levels);
    // the two-level walk shares one stack: top-level entries, the two entries parked on entering an instance, the mesh tree's entries
    scene->stats.maxDepth = top.levels + 2 + deepestMesh;
    if (scene->stats.maxDepth > SELAS_STACK_SIZE) {
        setError(ctx, "two-level BVH8 needs %u stack entries (top %u + 2 + mesh %u); the traversal stack holds %d", scene->stats.maxDepth, top.levels,
                 deepestMesh, SELAS_STACK_SIZE);
        return SELAS_ERR_UNSUPPORTED;
    }
    scene->stats.buildMs = ms;
    scene->stats.reserved = (uint32_t)std::min<uint64_t>(instancedTris, 0xFFFFFFFFull);  // triangles a flattened build would have stored
    scene->stats.deviceBytes = totalNodes * 80ull + totalTris * 48ull + table.size() * sizeof(InstanceTrace);
    return SELAS_OK;
}

static Bvh8 bvhOf(const SelasScene* s, bool publicIds = false) {
    Bvh8 b;
    b.instanceIds = publicIds ? s->gpu.instanceIds : nullptr;
    b.curves = s->gpu.curvePool;
    b.nodes = s->gpu.nodes;
    b.tris = s->gpu.tris;
    b.nodeCount = s->gpu.nodeCount;
    b.triCount = s->gpu.triCount;
    b.instances = s->gpu.traceInstances;
    b.topCount = s->topNodeCount;
    return b;
}

// Launch shape of a trace kernel: the staged top levels are dynamic shared memory; the grid is what is resident with that much of it.
struct TraceLaunch {
    int grid;
    size_t smem;
};
template <typename K>
static TraceLaunch traceLaunch(const SelasScene* scene, K kernel, uint32_t expected = 0) {
    TraceLaunch l;
    l.smem = (size_t)scene->topNodeCount * 80u;
    if (l.smem > 48u * 1024u) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l.smem);
    int perSm = SELAS_TRACE_CTAS_PER_SM;
    if (l.smem) {
        int resident = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kernel, SELAS_TRACE_THREADS, l.smem) == cudaSuccess && resident > 0)
            perSm = std::min(perSm, resident);
    }
    l.grid = scene->ctx->smCount * perSm;
    // a short queue (the late bounces' few thousand rays) does not need the whole machine's worth of CTAs started: one CTA per 64 expected rays
    // (never below two CTAs per SM: the estimate comes from an earlier pass and may be far too low after a camera change)
    if (expected) l.grid = std::max(2 * scene->ctx->smCount, std::min(l.grid, (int)((expected + expected / 8u + 63u) / 64u)));
    return l;
}
// SELAS_TRACE_COOP=1: flattened trees without curves go through the kernel with the warp-cooperative triangle stage (trace_coop.cuh)
// SELAS_TRACE_COOP: 1 / 0 force it on / off for every launch; unset: the public intersect / occluded calls use traceKernel (camera rays
// of triangle scenes are coherent) except on scenes with curve geometry, where the cooperative kernel wins for every kind of ray (a curve
// part costs three times a triangle: 200 000 blades, camera rays 459 -> 709 M/s, shadow rays 1.64 -> 2.71 G/s, profiles/r2_curves.txt).
static int traceCoopEnv() {
    static const int v = getenv("SELAS_TRACE_COOP") ? (atoi(getenv("SELAS_TRACE_COOP")) != 0 ? 1 : 0) : -1;
    return v;
}
static bool traceCoopEnabled(const SelasScene* scene) { return traceCoopEnv() >= 0 ? traceCoopEnv() == 1 : scene->gpu.curvePool != nullptr; }
// the wavefront asks per launch (coop: -1 = the default above, 0 / 1 = no / yes): primary rays are coherent and run better
// through traceKernel, everything after the first bounce and the shadow rays through traceCoopKernel (measurements in trace_coop.cuh)
static bool traceCoopFor(int coop, const SelasScene* scene) {
    if (traceCoopEnv() >= 0) return traceCoopEnv() == 1;
    return coop < 0 ? traceCoopEnabled(scene) : (coop != 0 || scene->gpu.curvePool != nullptr);
}
// kernel variant by scene: two-level form only when the scene was built that way (the flattened form keeps its register budget), staged
// top levels only when asked for
#define SELAS_TRACE_GO2(MODE, COUNT, INST, TOP, CURVES, ...)                                                       \
    do {                                                                                                           \
        TraceLaunch tl = traceLaunch(scene, traceKernel<MODE, COUNT, INST, TOP, CURVES>, expected__);              \
        traceKernel<MODE, COUNT, INST, TOP, CURVES><<<tl.grid, SELAS_TRACE_THREADS, tl.smem, stream>>>(__VA_ARGS__); \
    } while (0)
#define SELAS_TRACE_COOP_GO(MODE, COUNT, CURVES, INST, ...)                                                          \
    do {                                                                                                           \
        TraceLaunch tl = traceLaunch(scene, traceCoopKernel<MODE, COUNT, CURVES, INST>, expected__);               \
        traceCoopKernel<MODE, COUNT, CURVES, INST><<<tl.grid, SELAS_TRACE_THREADS, 0, stream>>>(__VA_ARGS__);      \
    } while (0)
#define SELAS_TRACE_GO(MODE, COUNT, ...)                                                                            \
    do {                                                                                                           \
        const bool inst__ = scene->gpu.traceInstances != nullptr, top__ = scene->topNodeCount != 0;                \
        const bool curves__ = scene->gpu.curvePool != nullptr;                                                     \
        if (coop__ && !top__ && scene->gpu.triCount < (1u << 27)) { /* queue entries hold lane | triangle << 5 */  \
            if (curves__ && inst__) SELAS_TRACE_COOP_GO(MODE, COUNT, true, true, __VA_ARGS__);                     \
            else if (curves__) SELAS_TRACE_COOP_GO(MODE, COUNT, true, false, __VA_ARGS__);                         \
            else if (inst__) SELAS_TRACE_COOP_GO(MODE, COUNT, false, true, __VA_ARGS__);                           \
            else SELAS_TRACE_COOP_GO(MODE, COUNT, false, false, __VA_ARGS__);                                      \
        } else if (curves__) { /* scenes with curve geometry: the variant that carries the curve leaf test */      \
            if (inst__) SELAS_TRACE_GO2(MODE, COUNT, true, false, true, __VA_ARGS__);                              \
            else SELAS_TRACE_GO2(MODE, COUNT, false, false, true, __VA_ARGS__);                                    \
        } else if (inst__ && top__) SELAS_TRACE_GO2(MODE, COUNT, true, true, false, __VA_ARGS__);                  \
        else if (inst__) SELAS_TRACE_GO2(MODE, COUNT, true, false, false, __VA_ARGS__);                            \
        else if (top__) SELAS_TRACE_GO2(MODE, COUNT, false, true, false, __VA_ARGS__);                             \
        else SELAS_TRACE_GO2(MODE, COUNT, false, false, false, __VA_ARGS__);                                       \
    } while (0)

int launchIntersect(SelasScene* scene, uint32_t n, const SelasRay* d_rays, SelasHit* d_hits, float* d_ng4, uint32_t* d_counts,
                    uint32_t* d_rayCounter, cudaStream_t stream) {
    if (n == 0) return SELAS_OK;
    SelasContext* ctx = scene->ctx;
    SELAS_CUDA_CHECK(ctx, cudaMemsetAsync(d_rayCounter, 0, sizeof(uint32_t), stream));
    const bool coop__ = traceCoopEnabled(scene);
    const uint32_t expected__ = 0;
    if (d_counts)
        SELAS_TRACE_GO(0, true, bvhOf(scene, true), d_rays, n, nullptr, d_hits, nullptr, (float4*)d_ng4, d_counts, nullptr, nullptr, d_rayCounter, nullptr);
    else
        SELAS_TRACE_GO(0, false, bvhOf(scene, true), d_rays, n, nullptr, d_hits, nullptr, (float4*)d_ng4, nullptr, nullptr, nullptr, d_rayCounter, nullptr);
    SELAS_CUDA_CHECK(ctx, cudaGetLastError());
    return SELAS_OK;
}

int launchOccluded(SelasScene* scene, uint32_t n, const SelasRay* d_rays, uint8_t* d_occ, uint32_t* d_counts, uint32_t* d_rayCounter,
                   cudaStream_t stream) {
    if (n == 0) return SELAS_OK;
    SelasContext* ctx = scene->ctx;
    SELAS_CUDA_CHECK(ctx, cudaMemsetAsync(d_rayCounter, 0, sizeof(uint32_t), stream));
    const bool coop__ = traceCoopEnabled(scene);
    const uint32_t expected__ = 0;
    if (d_counts)
        SELAS_TRACE_GO(1, true, bvhOf(scene), d_rays, n, nullptr, nullptr, d_occ, nullptr, d_counts, nullptr, nullptr, d_rayCounter, nullptr);
    else
        SELAS_TRACE_GO(1, false, bvhOf(scene), d_rays, n, nullptr, nullptr, d_occ, nullptr, nullptr, nullptr, nullptr, d_rayCounter, nullptr);
    SELAS_CUDA_CHECK(ctx, cudaGetLastError());
    return SELAS_OK;
}

// Wavefront variants: the queue length lives in device memory (or n is exact when d_n is null); d_rayCounter must
// already be zero.  totals != null selects the counting build of the same kernel (SELAS_RENDER_COUNT_TRAVERSAL).
int launchIntersectIndirect(SelasScene* scene, const uint32_t* d_n, uint32_t maxN, const SelasRay* d_rays, SelasHit* d_hits,
                            uint32_t* d_rayCounter, unsigned long long* d_totals, cudaStream_t stream, const uint32_t* d_order, int coop, uint32_t expected) {
    SelasContext* ctx = scene->ctx;
    const bool coop__ = traceCoopFor(coop, scene);
    const uint32_t expected__ = expected;
    if (d_totals)
        SELAS_TRACE_GO(0, true, bvhOf(scene), d_rays, maxN, d_n, d_hits, nullptr, nullptr, nullptr, nullptr, nullptr, d_rayCounter, d_totals, d_order);
    else
        SELAS_TRACE_GO(0, false, bvhOf(scene), d_rays, maxN, d_n, d_hits, nullptr, nullptr, nullptr, nullptr, nullptr, d_rayCounter, nullptr, d_order);
    SELAS_CUDA_CHECK(ctx, cudaGetLastError());
    return SELAS_OK;
}

int launchShadowAccumulate(SelasScene* scene, const uint32_t* d_n, uint32_t maxN, const SelasRay* d_rays, const float4* d_payload,
                           float* d_film, uint32_t* d_rayCounter, unsigned long long* d_totals, cudaStream_t stream, int coop, uint32_t expected, const uint32_t* d_order) {
    SelasContext* ctx = scene->ctx;
    const bool coop__ = traceCoopFor(coop, scene);
    const uint32_t expected__ = expected;
    if (d_totals)
        SELAS_TRACE_GO(2, true, bvhOf(scene), d_rays, maxN, d_n, nullptr, nullptr, nullptr, nullptr, d_payload, d_film, d_rayCounter, d_totals, d_order);
    else
        SELAS_TRACE_GO(2, false, bvhOf(scene), d_rays, maxN, d_n, nullptr, nullptr, nullptr, nullptr, d_payload, d_film, d_rayCounter, nullptr, d_order);
    SELAS_CUDA_CHECK(ctx, cudaGetLastError());
    return SELAS_OK;
}

}  // namespace selas
