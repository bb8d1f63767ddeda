// bvh_sah.cuh -- binned-SAH binary tree construction on the device (the tree Embree commits at ModelResource.cpp:461-468 is a
// SAH tree; the LBVH of bvh_build.cuh costs 13 % / 19 % more node tests per camera / incoherent ray on configs[2],
// profiles/r1_tree_quality.json).
//
// Shape of the build (all kernels, no host geometry work):
//   1. the Morton order + Karras radix tree + refit of bvh_build.cuh give contiguous clusters: the radix-tree nodes with at most
//      SELAS_SAH_CLUSTER primitives whose parent has more (sahClusterRootsKernel);
//   2. a top tree is built by binned SAH over the cluster boxes, weighted by their primitive counts (sahSplitKernel, one warp per
//      segment, one launch per generation);
//   3. the clusters are laid out in the top tree's leaf order (scan of the counts) and every cluster is built by the same
//      sahSplitKernel over its primitives -- thousands of independent segments per launch from the first generation on;
//   4. top-tree nodes are grafted onto the primitive positions (sahGraftTopKernel) so that ONE binary tree results, in the
//      {children, range, box} form the BVH2 -> BVH8 collapse of bvh_build.cuh consumes.
// Node ids: an internal node is named after the gap it splits, id = (position of its right child's first primitive) - 1, which is
// unique over the whole tree and needs no allocation; leaf j (final position) is (n - 1) + j, as in the radix tree.
//
// sahSplitKernel: a warp owns a segment [first, first + count) of 32-byte items {lo, payload | hi, weight}.  Pass 1 reduces the
// centroid and geometry bounds; pass 2 bins the items on all three axes (16 bins, shared-memory atomics); 45 lanes evaluate the
// 45 split planes (area(L) * weight(L) + area(R) * weight(R)); pass 3 partitions the items into the other buffer of a ping-pong pair
// (left from the front, right from the back), so generations never copy back.  Children of one item are finished on the spot
// (leaf box, final order); segments of two items need no binning.
#pragma once
#include "bvh_build.cuh"

namespace selas {

#ifndef SELAS_SAH_BINS
#define SELAS_SAH_BINS 16
#endif
constexpr uint32_t kSahNone = 0xFFFFFFFFu;
constexpr int kSahWarpsPerCta = 4;

struct SahItem {
    float4 lo;  // box min, w = payload bits (primitive index / cluster index)
    float4 hi;  // box max, w = weight (primitive count)
};
struct SahSeg {
    uint32_t first, count;
    uint32_t parent;  // node id whose child this segment is, kSahNone for a root
    uint32_t side;    // 0 = left, 1 = right
};
struct SahTree {  // outputs, indexed by node id
    int2* children;
    uint2* range;
    float4* boxLo;
    float4* boxHi;
    uint32_t* order;    // payload per final position
    uint32_t leafBase;  // id of the leaf at position 0
    uint32_t* rootId;   // written by the segment whose parent is kSahNone
};

SELAS_D void sahLinkChild(const SahTree& t, uint32_t parent, uint32_t side, uint32_t id) {
    if (parent == kSahNone) *t.rootId = id;
    else if (side == 0) t.children[parent].x = (int)id;
    else t.children[parent].y = (int)id;
}
SELAS_D void sahFinishLeaf(const SahTree& t, uint32_t parent, uint32_t side, uint32_t pos, const SahItem& it) {
    const uint32_t id = t.leafBase + pos;
    t.order[pos] = __float_as_uint(it.lo.w);
    t.boxLo[id] = make_float4(it.lo.x, it.lo.y, it.lo.z, 0.f);
    t.boxHi[id] = make_float4(it.hi.x, it.hi.y, it.hi.z, 0.f);
    sahLinkChild(t, parent, side, id);
}
// Bin of a centroid coordinate.  Written with explicit round-to-nearest intrinsics: the binning pass and the partition pass must put an
// item in the same bin, so the compiler may not contract the two sites differently.
SELAS_D int sahBin(float lo, float hi, float cLo, float scale) {
    const float c = __fmul_rn(0.5f, __fadd_rn(lo, hi));
    return min(SELAS_SAH_BINS - 1, (int)__fmul_rn(__fsub_rn(c, cLo), scale));
}
SELAS_D float sahHalfArea(const float* lo, const float* hi) {
    const float x = fmaxf(hi[0] - lo[0], 0.f), y = fmaxf(hi[1] - lo[1], 0.f), z = fmaxf(hi[2] - lo[2], 0.f);
    return x * y + y * z + z * x;
}

__global__ void __launch_bounds__(kSahWarpsPerCta * 32) sahSplitKernel(const SahItem* __restrict__ src, SahItem* __restrict__ dst,
                                                                       const SahSeg* __restrict__ segs, uint32_t segCount,
                                                                       SahSeg* __restrict__ nextSegs, uint32_t* __restrict__ nextCount,
                                                                       SahTree tree) {
    constexpr int B = SELAS_SAH_BINS;
    __shared__ uint32_t sLo[kSahWarpsPerCta][3][B][3], sHi[kSahWarpsPerCta][3][B][3];
    __shared__ uint32_t sCnt[kSahWarpsPerCta][3][B];
    __shared__ uint32_t sW[kSahWarpsPerCta][3][B];  // weights are primitive counts
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t ltMask = (1u << lane) - 1u;
    const uint32_t si = blockIdx.x * kSahWarpsPerCta + warp;
    if (si >= segCount) return;
    const SahSeg seg = segs[si];
    const uint32_t first = seg.first, count = seg.count;

    if (count == 2) {  // one node over two leaves: no search
        if (lane == 0) {
            const SahItem a = src[first], b = src[first + 1];
            const uint32_t id = first;
            tree.boxLo[id] = make_float4(fminf(a.lo.x, b.lo.x), fminf(a.lo.y, b.lo.y), fminf(a.lo.z, b.lo.z), 0.f);
            tree.boxHi[id] = make_float4(fmaxf(a.hi.x, b.hi.x), fmaxf(a.hi.y, b.hi.y), fmaxf(a.hi.z, b.hi.z), 0.f);
            tree.range[id] = make_uint2(first, first + 1);
            sahLinkChild(tree, seg.parent, seg.side, id);
            sahFinishLeaf(tree, id, 0, first, a);
            sahFinishLeaf(tree, id, 1, first + 1, b);
        }
        return;
    }

    // ---- pass 1: centroid bounds and geometry bounds ----
    float cLo[3] = {SELAS_FLOAT_MAX, SELAS_FLOAT_MAX, SELAS_FLOAT_MAX}, cHi[3] = {-SELAS_FLOAT_MAX, -SELAS_FLOAT_MAX, -SELAS_FLOAT_MAX};
    float gLo[3] = {SELAS_FLOAT_MAX, SELAS_FLOAT_MAX, SELAS_FLOAT_MAX}, gHi[3] = {-SELAS_FLOAT_MAX, -SELAS_FLOAT_MAX, -SELAS_FLOAT_MAX};
    for (uint32_t i = lane; i < count; i += 32) {
        const SahItem it = src[first + i];
        const float l[3] = {it.lo.x, it.lo.y, it.lo.z}, h[3] = {it.hi.x, it.hi.y, it.hi.z};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float c = __fmul_rn(0.5f, __fadd_rn(l[a], h[a]));
            cLo[a] = fminf(cLo[a], c); cHi[a] = fmaxf(cHi[a], c);
            gLo[a] = fminf(gLo[a], l[a]); gHi[a] = fmaxf(gHi[a], h[a]);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
        for (int s = 16; s > 0; s >>= 1) {
            cLo[a] = fminf(cLo[a], __shfl_xor_sync(0xffffffffu, cLo[a], s));
            cHi[a] = fmaxf(cHi[a], __shfl_xor_sync(0xffffffffu, cHi[a], s));
            gLo[a] = fminf(gLo[a], __shfl_xor_sync(0xffffffffu, gLo[a], s));
            gHi[a] = fmaxf(gHi[a], __shfl_xor_sync(0xffffffffu, gHi[a], s));
        }
    float scale[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float e = cHi[a] - cLo[a];
        scale[a] = e > 0.f ? (float)B * 0.999f / e : 0.f;  // a degenerate axis puts everything in bin 0: it offers no plane
    }

    // ---- pass 2: binning on the three axes ----
    for (uint32_t i = lane; i < 3 * B; i += 32) {
        const uint32_t a = i / B, b = i % B;
        sCnt[warp][a][b] = 0u;
        sW[warp][a][b] = 0u;
#pragma unroll
        for (int k = 0; k < 3; ++k) { sLo[warp][a][b][k] = 0xFFFFFFFFu; sHi[warp][a][b][k] = 0u; }
    }
    __syncwarp();
    // Items arrive in Morton order, so the 32 lanes of a step mostly share a bin: 32-way same-address shared atomics made the first
    // version of this pass 20 x slower than the rest of the build.  Lanes with equal bins are grouped (match.any), reduced with the
    // integer warp reductions (redux.sync on order-preserving float keys; weights are integer counts) and only the group leader touches
    // shared memory.
    for (uint32_t base = 0; base < count; base += 32) {
        const uint32_t i = base + lane;
        const bool valid = i < count;
        const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
        if (!valid) continue;
        const SahItem it = src[first + i];
        const float l[3] = {it.lo.x, it.lo.y, it.lo.z}, h[3] = {it.hi.x, it.hi.y, it.hi.z};
        const uint32_t fl[3] = {floatFlip(l[0]), floatFlip(l[1]), floatFlip(l[2])}, fh[3] = {floatFlip(h[0]), floatFlip(h[1]), floatFlip(h[2])};
        const uint32_t w = (uint32_t)it.hi.w;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int b = sahBin(l[a], h[a], cLo[a], scale[a]);
            const uint32_t peers = __match_any_sync(vmask, b);
            const uint32_t cnt = (uint32_t)__popc(peers);
            const uint32_t wsum = __reduce_add_sync(peers, w);
            const uint32_t m0 = __reduce_min_sync(peers, fl[0]), m1 = __reduce_min_sync(peers, fl[1]), m2 = __reduce_min_sync(peers, fl[2]);
            const uint32_t x0 = __reduce_max_sync(peers, fh[0]), x1 = __reduce_max_sync(peers, fh[1]), x2 = __reduce_max_sync(peers, fh[2]);
            if (lane == (uint32_t)(__ffs(peers) - 1)) {
                atomicAdd(&sCnt[warp][a][b], cnt);
                atomicAdd(&sW[warp][a][b], wsum);
                atomicMin(&sLo[warp][a][b][0], m0); atomicMin(&sLo[warp][a][b][1], m1); atomicMin(&sLo[warp][a][b][2], m2);
                atomicMax(&sHi[warp][a][b][0], x0); atomicMax(&sHi[warp][a][b][1], x1); atomicMax(&sHi[warp][a][b][2], x2);
            }
        }
    }
    __syncwarp();

    // ---- the 3 x (B - 1) split planes, one per lane (two rounds when B - 1 > 10) ----
    float bestCost = SELAS_FLOAT_MAX;
    uint32_t bestPlane = kSahNone;  // axis * B + first right bin
    for (uint32_t cand = lane; cand < 3 * (B - 1); cand += 32) {
        const uint32_t a = cand / (B - 1), p = cand % (B - 1) + 1;
        float lo[3] = {SELAS_FLOAT_MAX, SELAS_FLOAT_MAX, SELAS_FLOAT_MAX}, hi[3] = {-SELAS_FLOAT_MAX, -SELAS_FLOAT_MAX, -SELAS_FLOAT_MAX};
        uint32_t wL = 0, wR = 0;
        uint32_t nL = 0, nR = 0;
        for (uint32_t b = 0; b < p; ++b)
            if (sCnt[warp][a][b]) {
                nL += sCnt[warp][a][b]; wL += sW[warp][a][b];
#pragma unroll
                for (int k = 0; k < 3; ++k) { lo[k] = fminf(lo[k], floatUnflip(sLo[warp][a][b][k])); hi[k] = fmaxf(hi[k], floatUnflip(sHi[warp][a][b][k])); }
            }
        const float aL = sahHalfArea(lo, hi);
#pragma unroll
        for (int k = 0; k < 3; ++k) { lo[k] = SELAS_FLOAT_MAX; hi[k] = -SELAS_FLOAT_MAX; }
        for (uint32_t b = p; b < (uint32_t)B; ++b)
            if (sCnt[warp][a][b]) {
                nR += sCnt[warp][a][b]; wR += sW[warp][a][b];
#pragma unroll
                for (int k = 0; k < 3; ++k) { lo[k] = fminf(lo[k], floatUnflip(sLo[warp][a][b][k])); hi[k] = fmaxf(hi[k], floatUnflip(sHi[warp][a][b][k])); }
            }
        const float aR = sahHalfArea(lo, hi);
        if (nL && nR) {
            const float cost = aL * (float)wL + aR * (float)wR;
            if (cost < bestCost) { bestCost = cost; bestPlane = a * B + p; }
        }
    }
    for (int s = 16; s > 0; s >>= 1) {
        const float oc = __shfl_xor_sync(0xffffffffu, bestCost, s);
        const uint32_t op = __shfl_xor_sync(0xffffffffu, bestPlane, s);
        if (oc < bestCost || (oc == bestCost && op < bestPlane)) { bestCost = oc; bestPlane = op; }
    }

    // ---- pass 3: partition into the other buffer (left from the front, right from the back) ----
    uint32_t nl = 0, nr = 0;
    if (bestPlane == kSahNone) {  // all centroids coincide: split the current order in the middle
        nl = count / 2;
        for (uint32_t i = lane; i < count; i += 32) dst[first + i] = src[first + i];
    } else {
        const int axis = (int)(bestPlane / B), plane = (int)(bestPlane % B);
        for (uint32_t base = 0; base < count; base += 32) {
            const uint32_t i = base + lane;
            const bool valid = i < count;
            SahItem it;
            bool left = false;
            if (valid) {
                it = src[first + i];
                const float l = axis == 0 ? it.lo.x : (axis == 1 ? it.lo.y : it.lo.z), h = axis == 0 ? it.hi.x : (axis == 1 ? it.hi.y : it.hi.z);
                const float cl = axis == 0 ? cLo[0] : (axis == 1 ? cLo[1] : cLo[2]);
                const float sc = axis == 0 ? scale[0] : (axis == 1 ? scale[1] : scale[2]);
                left = sahBin(l, h, cl, sc) < plane;
            }
            const uint32_t lm = __ballot_sync(0xffffffffu, valid && left), rm = __ballot_sync(0xffffffffu, valid && !left);
            if (valid) {
                if (left) dst[first + nl + __popc(lm & ltMask)] = it;
                else dst[first + count - 1 - (nr + __popc(rm & ltMask))] = it;
            }
            nl += __popc(lm);
            nr += __popc(rm);
        }
    }
    if (nl == 0 || nl == count) nl = count / 2;  // cannot happen (the plane had items on both sides); kept as a guard
    __syncwarp();
    __threadfence_block();

    // ---- the node, and what becomes of its two children ----
    const uint32_t id = first + nl - 1;
    if (lane == 0) {
        tree.boxLo[id] = make_float4(gLo[0], gLo[1], gLo[2], 0.f);
        tree.boxHi[id] = make_float4(gHi[0], gHi[1], gHi[2], 0.f);
        tree.range[id] = make_uint2(first, first + count - 1);
        sahLinkChild(tree, seg.parent, seg.side, id);
    }
    if (lane < 2) {
        const uint32_t f = lane == 0 ? first : first + nl, c = lane == 0 ? nl : count - nl;
        if (c == 1) sahFinishLeaf(tree, id, lane, f, dst[f]);
        else {
            const uint32_t slot = atomicAdd(nextCount, 1u);
            nextSegs[slot] = SahSeg{f, c, id, lane};
        }
    }
}

// ---- cluster roots of the radix tree ------------------------------------------------------------------------------------------
// Radix-tree node ids: internal i in [0, n - 1), leaf j -> (n - 1) + j.  A cluster root covers <= maxCount primitives and its parent
// covers more (the root of the whole tree is a cluster when n <= maxCount).
__global__ void sahClusterRootsKernel(int n, uint32_t maxCount, const uint2* __restrict__ range, const uint32_t* __restrict__ parent,
                                      const float4* __restrict__ boxLo, const float4* __restrict__ boxHi, SahItem* __restrict__ items,
                                      uint2* __restrict__ clusters /* first (Morton position), count */, uint32_t* __restrict__ clusterCount,
                                      uint32_t capacity) {
    const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t total = 2u * (uint32_t)n - 1u;
    if (id >= total) return;
    const uint32_t leafBase = (uint32_t)n - 1u;
    uint32_t first, cnt;
    if (id >= leafBase) { first = id - leafBase; cnt = 1u; }
    else { const uint2 r = range[id]; first = r.x; cnt = r.y - r.x + 1u; }
    if (cnt > maxCount) return;
    const uint32_t p = n > 1 ? parent[id] : kSahNone;
    if (p != kSahNone) {
        const uint2 pr = range[p];
        if (pr.y - pr.x + 1u <= maxCount) return;
    }
    const uint32_t c = atomicAdd(clusterCount, 1u);
    if (c >= capacity) return;
    clusters[c] = make_uint2(first, cnt);
    const float4 lo = boxLo[id], hi = boxHi[id];
    SahItem it;
    it.lo = make_float4(lo.x, lo.y, lo.z, __uint_as_float(c));
    it.hi = make_float4(hi.x, hi.y, hi.z, (float)cnt);
    items[c] = it;
}

// counts of the clusters in top-tree leaf order (input of the scan that places them)
__global__ void sahOrderedCountsKernel(uint32_t C, const uint32_t* __restrict__ topOrder, const uint2* __restrict__ clusters,
                                       uint32_t* __restrict__ counts) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < C) counts[p] = clusters[topOrder[p]].y;
}

// Primitive items in their cluster's final slot: position -> cluster by binary search over the exclusive scan.
__global__ void sahSeedItemsKernel(uint32_t n, uint32_t C, const uint32_t* __restrict__ topOrder, const uint2* __restrict__ clusters,
                                   const uint32_t* __restrict__ excl, const uint32_t* __restrict__ sorted, const BuildTri* __restrict__ tris,
                                   SahItem* __restrict__ items) {
    const uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= n) return;
    uint32_t lo = 0, hi = C - 1;  // last p with excl[p] <= pos
    while (lo < hi) {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (excl[mid] <= pos) lo = mid;
        else hi = mid - 1;
    }
    const uint2 cl = clusters[topOrder[lo]];
    const uint32_t prim = sorted[cl.x + (pos - excl[lo])];
    const BuildTri t = tris[prim];
    float plo[3], phi[3];
    primBounds(t, plo, phi);
    SahItem it;
    it.lo = make_float4(plo[0], plo[1], plo[2], __uint_as_float(prim));
    it.hi = make_float4(phi[0], phi[1], phi[2], 1.0f);
    items[pos] = it;
}

// Top-tree node k (ids in cluster positions: internal = gap id, leaf = (C - 1) + position) -> node of the final tree, id = excl[k + 1] - 1.
// Cluster leaves of one primitive are finished here; larger ones become the root segments of the per-cluster builds.
__global__ void sahGraftTopKernel(uint32_t n, uint32_t C, const int2* __restrict__ topChildren, const uint2* __restrict__ topRange,
                                  const float4* __restrict__ topLo, const float4* __restrict__ topHi, const uint32_t* __restrict__ excl,
                                  const uint32_t* __restrict__ counts, const SahItem* __restrict__ items, SahTree tree,
                                  SahSeg* __restrict__ seeds, uint32_t* __restrict__ seedCount, const uint32_t* __restrict__ topRoot) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;  // top internal node id (gap between cluster positions k and k + 1)
    if (C == 1) {
        if (k == 0) {
            if (counts[0] == 1) {
                sahFinishLeaf(tree, kSahNone, 0, 0, items[0]);
            } else {
                const uint32_t s = atomicAdd(seedCount, 1u);
                seeds[s] = SahSeg{0u, counts[0], kSahNone, 0u};
            }
        }
        return;
    }
    if (k >= C - 1) return;
    const uint32_t g = excl[k + 1] - 1u;
    const int2 ch = topChildren[k];
    const uint2 r = topRange[k];
    tree.boxLo[g] = topLo[k];
    tree.boxHi[g] = topHi[k];
    tree.range[g] = make_uint2(excl[r.x], excl[r.y] + counts[r.y] - 1u);
    if (k == *topRoot) *tree.rootId = g;
    const uint32_t refs[2] = {(uint32_t)ch.x, (uint32_t)ch.y};
    for (uint32_t side = 0; side < 2; ++side) {
        const uint32_t ref = refs[side];
        if (ref < C - 1) {
            sahLinkChild(tree, g, side, excl[ref + 1] - 1u);
        } else {
            const uint32_t p = ref - (C - 1);
            if (counts[p] == 1) sahFinishLeaf(tree, g, side, excl[p], items[excl[p]]);
            else {
                const uint32_t s = atomicAdd(seedCount, 1u);
                seeds[s] = SahSeg{excl[p], counts[p], g, side};
            }
        }
    }
}

}  // namespace selas
