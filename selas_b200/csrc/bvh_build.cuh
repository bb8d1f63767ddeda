// bvh_build.cuh -- device-side construction of the compressed BVH8.
//
// Replaces the Embree scene build the reference triggers through rtcSetSharedGeometryBuffer /
// rtcCommitGeometry / rtcCommitScene (ModelResource.cpp:164-200,:461-468; SubsceneResource.cpp:141-156;
// SceneResource.cpp:218-249).  Load-time work, timed separately from the render.
//
// Pipeline (all on the GPU):
//   1. gatherTriangles   faces -> world-space triangles (quads -> 2, Embree split v0v1v3 / v2v3v1) + bounds
//   2. mortonKeys        63-bit Morton code of the triangle-box centre; cub::DeviceRadixSort (library sort)
//   3. radixTree         Karras 2012 binary radix tree over the sorted keys
//   4. fitBoxes          bottom-up AABB refit with per-node arrival counters
//   5. collapseLevel     top-down, one launch per BVH8 level: greedily open the largest-area members until 8
//                        children, assign octant-ordered slots, quantise child boxes to 8 bits, emit the
//                        80-byte node and its <=24 leaf triangles as {v0,e1,e2} records
#pragma once
#include <cub/cub.cuh>

#include "bvh8.cuh"

namespace selas {

struct MeshDev {
    const float* positions;
    const uint32_t* indices;
    uint32_t faceCount, indicesPerFace;
};

struct BuildTri {  // 48 bytes: raw world-space vertices + ids
    float4 a, b, c;  // a.w = primID, b.w = geomID, c.w = instID | curve << 30 | second << 31
};
constexpr uint32_t kPrimCurveFlag = 1u << 30;  // the record is a curve segment: a = box min, b = box max, c.x = index into the curve pool

// bounds of a build record: the three vertices of a triangle, or the stored box of a curve segment
SELAS_D void primBounds(const BuildTri& t, float* lo, float* hi) {
    if (__float_as_uint(t.c.w) & kPrimCurveFlag) {
        lo[0] = t.a.x; lo[1] = t.a.y; lo[2] = t.a.z;
        hi[0] = t.b.x; hi[1] = t.b.y; hi[2] = t.b.z;
    } else {
        lo[0] = fminf(t.a.x, fminf(t.b.x, t.c.x)); hi[0] = fmaxf(t.a.x, fmaxf(t.b.x, t.c.x));
        lo[1] = fminf(t.a.y, fminf(t.b.y, t.c.y)); hi[1] = fmaxf(t.a.y, fmaxf(t.b.y, t.c.y));
        lo[2] = fminf(t.a.z, fminf(t.b.z, t.c.z)); hi[2] = fmaxf(t.a.z, fmaxf(t.b.z, t.c.z));
    }
}

SELAS_D uint32_t floatFlip(float f) {  // order-preserving float -> uint
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
SELAS_HD float floatUnflip(uint32_t u) {
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

__global__ void gatherTrianglesKernel(MeshDev mesh, const float* __restrict__ xform, bool identity, uint32_t geomID,
                                      uint32_t instID, BuildTri* __restrict__ out, uint32_t outBase,
                                      uint32_t* __restrict__ bounds /* 6 flipped floats: min xyz, max xyz */) {
    uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {SELAS_FLOAT_MAX, SELAS_FLOAT_MAX, SELAS_FLOAT_MAX}, hi[3] = {-SELAS_FLOAT_MAX, -SELAS_FLOAT_MAX, -SELAS_FLOAT_MAX};
    if (f < mesh.faceCount) {
        const uint32_t* ix = mesh.indices + (size_t)f * mesh.indicesPerFace;
        V3 p[4];
        float m[16];
        if (!identity)
            for (int k = 0; k < 16; ++k) m[k] = xform[k];
        for (uint32_t k = 0; k < mesh.indicesPerFace; ++k) {
            V3 q = v3(mesh.positions + 3 * (size_t)ix[k]);
            p[k] = identity ? q : mulPoint44(q, m);
        }
        auto emit = [&](uint32_t slot, V3 a, V3 b, V3 c, uint32_t second) {
            BuildTri t;
            t.a = make_float4(a.x, a.y, a.z, __uint_as_float(f));
            t.b = make_float4(b.x, b.y, b.z, __uint_as_float(geomID));
            t.c = make_float4(c.x, c.y, c.z, __uint_as_float(instID | (second << 31)));
            out[slot] = t;
            lo[0] = fminf(lo[0], fminf(a.x, fminf(b.x, c.x))); hi[0] = fmaxf(hi[0], fmaxf(a.x, fmaxf(b.x, c.x)));
            lo[1] = fminf(lo[1], fminf(a.y, fminf(b.y, c.y))); hi[1] = fmaxf(hi[1], fmaxf(a.y, fmaxf(b.y, c.y)));
            lo[2] = fminf(lo[2], fminf(a.z, fminf(b.z, c.z))); hi[2] = fmaxf(hi[2], fmaxf(a.z, fmaxf(b.z, c.z)));
        };
        if (mesh.indicesPerFace == 3) {
            emit(outBase + f, p[0], p[1], p[2], 0);
        } else {
            emit(outBase + 2 * f, p[0], p[1], p[3], 0);
            emit(outBase + 2 * f + 1, p[2], p[3], p[1], 1);
        }
    }
    // warp reduce then one atomic per warp per component
    for (int k = 0; k < 3; ++k) {
        float a = lo[k], b = hi[k];
        for (int s = 16; s > 0; s >>= 1) {
            a = fminf(a, __shfl_xor_sync(0xffffffffu, a, s));
            b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, s));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(bounds + k, floatFlip(a));
            atomicMax(bounds + 3 + k, floatFlip(b));
        }
    }
}

// Curve segments of one curve geometry (MeshDev.indicesPerFace == 1: positions = float4 control vertices, indices = first vertex of each
// segment): the four control vertices go through the instance transform into the curve pool (radius scaled by the cube root of the
// transform's volume change).  Every segment becomes SELAS_CURVE_PARTS build records, one per cone of curve.cuh, each with the box of
// its two end spheres: blades of grass are long, thin and oblique, and with one box per segment (the first version) a ray through a
// lawn met 60 segment boxes and ran the whole 8-cone test in each (48 M primary rays/s; per-part boxes: profiles/r2_curves.txt).
__global__ void gatherCurvesKernel(MeshDev mesh, const float* __restrict__ xform, bool identity, uint32_t geomID, uint32_t instID,
                                   BuildTri* __restrict__ out, uint32_t outBase, float4* __restrict__ pool, uint32_t poolBase,
                                   uint32_t* __restrict__ bounds) {
    uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {SELAS_FLOAT_MAX, SELAS_FLOAT_MAX, SELAS_FLOAT_MAX}, hi[3] = {-SELAS_FLOAT_MAX, -SELAS_FLOAT_MAX, -SELAS_FLOAT_MAX};
    if (f < mesh.faceCount) {
        const float4* cv = reinterpret_cast<const float4*>(mesh.positions) + mesh.indices[f];
        float m[16];
        float rscale = 1.0f;
        if (!identity) {
            for (int k = 0; k < 16; ++k) m[k] = xform[k];
            const float det = m[0] * (m[5] * m[10] - m[6] * m[9]) - m[1] * (m[4] * m[10] - m[6] * m[8]) + m[2] * (m[4] * m[9] - m[5] * m[8]);
            rscale = cbrtf(fabsf(det));
        }
        float4 c[4];
        for (int k = 0; k < 4; ++k) {
            const float4 p = cv[k];
            V3 q = v3(p.x, p.y, p.z);
            if (!identity) q = mulPoint44(q, m);
            c[k] = make_float4(q.x, q.y, q.z, p.w * rscale);
            pool[4 * (size_t)(poolBase + f) + k] = c[k];
        }
        float4 a = curvePoint(c[0], c[1], c[2], c[3], 0.0f);
        for (int part = 0; part < SELAS_CURVE_PARTS; ++part) {
            const float4 b = curvePoint(c[0], c[1], c[2], c[3], (float)(part + 1) * (1.0f / SELAS_CURVE_PARTS));
            const float pa[3] = {a.x, a.y, a.z}, pb[3] = {b.x, b.y, b.z};
            const float ra = fabsf(a.w), rb = fabsf(b.w);
            float plo[3], phi[3];
            for (int k = 0; k < 3; ++k) {  // the two end spheres, plus a relative margin for the fp32 evaluation of the curve points
                plo[k] = fminf(pa[k] - ra, pb[k] - rb);
                phi[k] = fmaxf(pa[k] + ra, pb[k] + rb);
                const float pad = 1e-4f * fmaxf(ra, rb) + 2e-6f * fmaxf(fabsf(plo[k]), fabsf(phi[k]));
                plo[k] -= pad;
                phi[k] += pad;
                lo[k] = fminf(lo[k], plo[k]);
                hi[k] = fmaxf(hi[k], phi[k]);
            }
            BuildTri t;
            t.a = make_float4(plo[0], plo[1], plo[2], __uint_as_float(f));
            t.b = make_float4(phi[0], phi[1], phi[2], __uint_as_float(geomID));
            t.c = make_float4(__uint_as_float(poolBase + f), __uint_as_float((uint32_t)part), 0.f, __uint_as_float(instID | kPrimCurveFlag));
            out[outBase + (size_t)f * SELAS_CURVE_PARTS + part] = t;
            a = b;
        }
    }
    for (int k = 0; k < 3; ++k) {
        float a = lo[k], b2 = hi[k];
        for (int s = 16; s > 0; s >>= 1) {
            a = fminf(a, __shfl_xor_sync(0xffffffffu, a, s));
            b2 = fmaxf(b2, __shfl_xor_sync(0xffffffffu, b2, s));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(bounds + k, floatFlip(a));
            atomicMax(bounds + 3 + k, floatFlip(b2));
        }
    }
}

SELAS_D uint64_t expandBits21(uint32_t v) {
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

// cube != 0: all three axes are quantised on the grid of the longest one (cubic cells), so the axes with a short extent do not
// contribute their leading bits -- on a flat scene (terrain) per-axis normalisation splits by height every third level and the
// resulting strata overlap over the whole ground plane.
__global__ void mortonKernel(const BuildTri* __restrict__ tris, uint32_t n, const uint32_t* __restrict__ bounds,
                             uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, int cube) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float lo[3], ext[3];
    for (int k = 0; k < 3; ++k) {
        lo[k] = floatUnflip(bounds[k]);
        ext[k] = floatUnflip(bounds[3 + k]) - lo[k];
    }
    if (cube) {
        const float longest = fmaxf(ext[0], fmaxf(ext[1], ext[2]));
        ext[0] = ext[1] = ext[2] = longest;
    }
    BuildTri t = tris[i];
    float plo[3], phi[3];
    primBounds(t, plo, phi);
    float c[3] = {0.5f * (plo[0] + phi[0]), 0.5f * (plo[1] + phi[1]), 0.5f * (plo[2] + phi[2])};
    uint32_t q[3];
    for (int k = 0; k < 3; ++k) {
        float nrm = ext[k] > 0.f ? (c[k] - lo[k]) / ext[k] : 0.f;
        q[k] = (uint32_t)fminf(fmaxf(nrm * 2097152.0f, 0.0f), 2097151.0f);
    }
    keys[i] = (expandBits21(q[0]) << 2) | (expandBits21(q[1]) << 1) | expandBits21(q[2]);
    vals[i] = i;
}

// ---- Karras radix tree -------------------------------------------------------------------------------------
// Node ids: internal i in [0, n-1); leaf j -> (n-1) + j.
SELAS_D int deltaKeys(const uint64_t* keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz((uint32_t)i ^ (uint32_t)j);
    return __clzll((long long)(a ^ b));
}

__global__ void radixTreeKernel(const uint64_t* __restrict__ keys, int n, int2* __restrict__ children, uint32_t* __restrict__ parent,
                                uint2* __restrict__ range) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = (deltaKeys(keys, n, i, i + 1) - deltaKeys(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = deltaKeys(keys, n, i, i - d);
    int lmax = 2;
    while (deltaKeys(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (deltaKeys(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = deltaKeys(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (deltaKeys(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int first = min(i, j), last = max(i, j);
    int left = (first == gamma) ? (n - 1) + gamma : gamma;
    int right = (last == gamma + 1) ? (n - 1) + gamma + 1 : gamma + 1;
    children[i] = make_int2(left, right);
    parent[left] = (uint32_t)i;
    parent[right] = (uint32_t)i;
    range[i] = make_uint2((uint32_t)first, (uint32_t)last);
    if (i == 0) parent[0] = 0xFFFFFFFFu;
}

__global__ void fitBoxesKernel(const BuildTri* __restrict__ tris, const uint32_t* __restrict__ sorted, int n,
                               const int2* __restrict__ children, const uint32_t* __restrict__ parent, float4* __restrict__ boxLo,
                               float4* __restrict__ boxHi, uint32_t* __restrict__ arrivals) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    BuildTri t = tris[sorted[j]];
    float plo[3], phi[3];
    primBounds(t, plo, phi);
    float4 lo = make_float4(plo[0], plo[1], plo[2], 0.f);
    float4 hi = make_float4(phi[0], phi[1], phi[2], 0.f);
    uint32_t id = (uint32_t)(n - 1) + (uint32_t)j;
    boxLo[id] = lo;
    boxHi[id] = hi;
    if (n == 1) return;
    uint32_t p = parent[id];
    while (p != 0xFFFFFFFFu) {
        __threadfence();
        if (atomicAdd(arrivals + p, 1u) == 0u) return;  // first child to arrive leaves; the second one merges
        int2 ch = children[p];
        float4 l0 = __ldcg(boxLo + ch.x), l1 = __ldcg(boxLo + ch.y), h0 = __ldcg(boxHi + ch.x), h1 = __ldcg(boxHi + ch.y);
        lo = make_float4(fminf(l0.x, l1.x), fminf(l0.y, l1.y), fminf(l0.z, l1.z), 0.f);
        hi = make_float4(fmaxf(h0.x, h1.x), fmaxf(h0.y, h1.y), fmaxf(h0.z, h1.z), 0.f);
        boxLo[p] = lo;
        boxHi[p] = hi;
        p = parent[p];
    }
}

// ---- BVH2 -> BVH8 collapse ---------------------------------------------------------------------------------
struct CollapseTask {
    uint32_t bvh2;  // BVH2 node id this wide node covers
    uint32_t node;  // index of the wide node to write
};

SELAS_D float halfArea(float4 lo, float4 hi) {
    float x = hi.x - lo.x, y = hi.y - lo.y, z = hi.z - lo.z;
    return x * y + y * z + z * x;
}

// ---- cost-optimal collapse (after Ylitie, Karras, Laine 2017, section 3.1) ------------------------------------------------------------
// C(x, j) = the smallest SAH cost of representing BVH2 subtree x with at most j child slots of its wide parent:
//   C(x, 1) = min( leaf slot: A(x) * count * cPrim  (count <= 3),   wide node: A(x) * cNode + D(x, 8) )
//   C(x, j) = min( C(x, j - 1), D(x, j) ),   D(x, j) = min over k of C(left, k) + C(right, j - k)
// computed bottom-up (the second child to arrive does the parent, as in the refit), with the choices packed into one word per internal
// node: bit 0 = leaf slot, bits 1-3 = k of D(x, 8), bits 4 + 3 (j - 2) .. = k of D(x, j) for j = 2..7 or 0 when C(x, j - 1) is kept.
// The greedy collapse (open the largest box until eight slots are taken) dissolved small subtrees into many wide nodes of two to five
// slots; with the costs in hand the bottom of the tree is made of well-filled nodes: configs[1] 144 k -> 96 k nodes, 12.18 -> 11.85
// nodes per camera ray; configs[2] 12.29 -> 11.93 per closest-hit ray, 13.14 -> 12.68 per shadow ray, step 16.24 -> 15.90 ms
// (profiles/r2_ab_collapse*.jsonl).  Used for wide nodes over at most `decisionMaxCount` primitives (default 32768): applied all the way to
// the root it costs camera rays 8-20 % MORE node visits (13.1-14.8 per ray on configs[1] for any cPrim from 0.2 to 4) -- the cost model
// knows nothing about front-to-back order, and the uneven children it picks near the root (one huge subtree beside seven small ones)
// defeat the octant-ordered slots; any-hit rays do not care (4.27-4.35 G/s either way).
__global__ void bvh2ParentsKernel(int n, const int2* __restrict__ children, uint32_t root, uint32_t* __restrict__ parent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int2 ch = children[i];
    parent[ch.x] = (uint32_t)i;
    parent[ch.y] = (uint32_t)i;
    if ((uint32_t)i == root) parent[i] = 0xFFFFFFFFu;
}

__global__ void collapseCostKernel(int n, const int2* __restrict__ children, const uint2* __restrict__ range, const uint32_t* __restrict__ parent,
                                   const float4* __restrict__ boxLo, const float4* __restrict__ boxHi, float cNode, float cPrim,
                                   float* __restrict__ cost7, uint32_t* __restrict__ decision, uint32_t* __restrict__ arrivals) {
    const int j0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (j0 >= n || n == 1) return;
    const uint32_t leafBase = (uint32_t)(n - 1);
    uint32_t p = parent[leafBase + (uint32_t)j0];
    while (p != 0xFFFFFFFFu) {
        __threadfence();
        if (atomicAdd(arrivals + p, 1u) == 0u) return;
        const int2 ch = children[p];
        float cl[8], cr[8];
        for (int side = 0; side < 2; ++side) {
            const uint32_t id = (uint32_t)(side ? ch.y : ch.x);
            float* c = side ? cr : cl;
            if (id >= leafBase) {
                const float a = halfArea(boxLo[id], boxHi[id]) * cPrim;
                for (int k = 1; k <= 7; ++k) c[k] = a;
            } else {
                for (int k = 1; k <= 7; ++k) c[k] = __ldcg(cost7 + (size_t)id * 7 + (k - 1));
            }
        }
        float D[9];
        uint32_t K[9];
        for (int j = 2; j <= 8; ++j) {
            float best = SELAS_FLOAT_MAX;
            uint32_t bk = 1;
            for (int k = (j - 7 > 1 ? j - 7 : 1); k <= (j - 1 < 7 ? j - 1 : 7); ++k) {
                const float v = cl[k] + cr[j - k];
                if (v < best) { best = v; bk = (uint32_t)k; }
            }
            D[j] = best;
            K[j] = bk;
        }
        const float area = halfArea(boxLo[p], boxHi[p]);
        const uint2 r = range[p];
        const uint32_t count = r.y - r.x + 1u;
        const float cInt = area * cNode + D[8];
        const float cLeaf = count <= 3u ? area * (float)count * cPrim : SELAS_FLOAT_MAX;
        uint32_t word = (cLeaf <= cInt ? 1u : 0u) | (K[8] << 1);
        float c = fminf(cLeaf, cInt);
        cost7[(size_t)p * 7] = c;
        for (int j = 2; j <= 7; ++j) {
            if (D[j] < c) {
                c = D[j];
                word |= K[j] << (4 + 3 * (j - 2));
            }
            cost7[(size_t)p * 7 + (j - 1)] = c;
        }
        decision[p] = word;
        p = parent[p];
    }
}

__global__ void collapseLevelKernel(const CollapseTask* __restrict__ tasks, uint32_t taskCount, int n,
                                    const int2* __restrict__ children, const uint2* __restrict__ range,
                                    const float4* __restrict__ boxLo, const float4* __restrict__ boxHi,
                                    const BuildTri* __restrict__ tris, const uint32_t* __restrict__ sorted,
                                    float4* __restrict__ outNodes, float4* __restrict__ outTris, uint32_t* __restrict__ nodeCounter,
                                    uint32_t* __restrict__ triCounter, CollapseTask* __restrict__ nextTasks,
                                    uint32_t* __restrict__ nextCount, const uint32_t* __restrict__ decision, uint32_t decisionMaxCount) {
    uint32_t ti = blockIdx.x * blockDim.x + threadIdx.x;
    if (ti >= taskCount) return;
    CollapseTask task = tasks[ti];
    const uint32_t leafBase = (uint32_t)(n - 1);

    auto countOf = [&](uint32_t id) -> uint32_t {
        if (id >= leafBase) return 1u;
        uint2 r = range[id];
        return r.y - r.x + 1u;
    };
    auto firstOf = [&](uint32_t id) -> uint32_t { return id >= leafBase ? id - leafBase : range[id].x; };

    uint32_t member[8];
    uint32_t cnt[8];
    float area[8];
    int m = 0;
    if (task.bvh2 >= leafBase) {  // single-triangle scene
        member[0] = task.bvh2; cnt[0] = 1; area[0] = 0.f; m = 1;
    } else if (decision && countOf(task.bvh2) <= decisionMaxCount) {
        // the slots the cost recurrence chose: (subtree, slots) pairs are refined until every pair holds one slot
        uint32_t stackId[8];
        uint32_t stackJ[8];
        int top = 0;
        {
            const int2 ch = children[task.bvh2];
            const uint32_t k8 = (decision[task.bvh2] >> 1) & 7u;
            stackId[top] = (uint32_t)ch.y; stackJ[top++] = 8u - k8;
            stackId[top] = (uint32_t)ch.x; stackJ[top++] = k8;
        }
        while (top > 0) {
            const uint32_t x = stackId[--top];
            uint32_t j = stackJ[top];
            if (x >= leafBase || j <= 1u) {
                member[m] = x; cnt[m] = countOf(x); area[m] = 0.f;
                ++m;
                continue;
            }
            const uint32_t word = decision[x];
            uint32_t k = 0;
            while (j >= 2u && (k = (word >> (4 + 3 * (j - 2u))) & 7u) == 0u) --j;  // C(x, j) == C(x, j - 1): fewer slots do as well
            if (j <= 1u) {
                member[m] = x; cnt[m] = countOf(x); area[m] = 0.f;
                ++m;
                continue;
            }
            const int2 ch = children[x];
            stackId[top] = (uint32_t)ch.y; stackJ[top++] = j - k;
            stackId[top] = (uint32_t)ch.x; stackJ[top++] = k;
        }
    } else {
        int2 ch = children[task.bvh2];
        member[0] = (uint32_t)ch.x; member[1] = (uint32_t)ch.y;
        for (int k = 0; k < 2; ++k) { cnt[k] = countOf(member[k]); area[k] = halfArea(boxLo[member[k]], boxHi[member[k]]); }
        m = 2;
        // phase A opens members that are too big to be leaves (> 3 triangles); phase B keeps splitting leaves
        // while slots remain -- extra children cost nothing in an 8-wide node and tighten the boxes.
        for (int phase = 0; phase < 2; ++phase) {
            while (m < 8) {
                int best = -1;
                float bestArea = -1.f;
                for (int k = 0; k < m; ++k) {
                    bool openable = phase == 0 ? cnt[k] > 3u : cnt[k] > 1u;
                    if (openable && area[k] > bestArea) { bestArea = area[k]; best = k; }
                }
                if (best < 0) break;
                int2 c2 = children[member[best]];
                member[best] = (uint32_t)c2.x;
                member[m] = (uint32_t)c2.y;
                cnt[best] = countOf(member[best]); area[best] = halfArea(boxLo[member[best]], boxHi[member[best]]);
                cnt[m] = countOf(member[m]); area[m] = halfArea(boxLo[member[m]], boxHi[member[m]]);
                ++m;
            }
        }
    }

    float4 nlo = boxLo[task.bvh2], nhi = boxHi[task.bvh2];
    float cx = 0.5f * (nlo.x + nhi.x), cy = 0.5f * (nlo.y + nhi.y), cz = 0.5f * (nlo.z + nhi.z);

    // octant-ordered slot assignment: slot s should hold the child met first by rays whose direction signs
    // are (s&4 ? - : +, s&2 ? - : +, s&1 ? - : +); greedy minimum of dot(childCentre - nodeCentre, dir_s).
    float4 clo[8], chi[8];
    float cost[8][8];
    for (int k = 0; k < m; ++k) {
        clo[k] = boxLo[member[k]];
        chi[k] = boxHi[member[k]];
        float dx = 0.5f * (clo[k].x + chi[k].x) - cx, dy = 0.5f * (clo[k].y + chi[k].y) - cy, dz = 0.5f * (clo[k].z + chi[k].z) - cz;
        // (the box's entry corner instead of its centre only shifts a child's costs by a constant: the same assignment)
        for (int s = 0; s < 8; ++s) cost[k][s] = ((s & 4) ? -dx : dx) + ((s & 2) ? -dy : dy) + ((s & 1) ? -dz : dz);
    }
    int slotOf[8];
    int childInSlot[8];
    for (int s = 0; s < 8; ++s) childInSlot[s] = -1;
    for (int k = 0; k < 8; ++k) slotOf[k] = -1;
    for (int it = 0; it < m; ++it) {
        float bestC = SELAS_FLOAT_MAX;
        int bk = -1, bs = -1;
        for (int k = 0; k < m; ++k) {
            if (slotOf[k] >= 0) continue;
            for (int s = 0; s < 8; ++s) {
                if (childInSlot[s] >= 0) continue;
                if (cost[k][s] < bestC) { bestC = cost[k][s]; bk = k; bs = s; }
            }
        }
        slotOf[bk] = bs;
        childInSlot[bs] = bk;
    }
#ifndef SELAS_SLOT_REFINE
#define SELAS_SLOT_REFINE 1
#endif
#if SELAS_SLOT_REFINE
    // the greedy pick is not the cheapest assignment: improve it by moves to free slots and pairwise swaps until none lowers the total
    // (configs[1]: 11.86 -> 11.60 nodes per camera ray, +2 % rays/s; configs[2] unchanged: profiles/r2_ab_slot_refine.jsonl)
    for (int round = 0; round < 6; ++round) {
        bool changed = false;
        for (int k = 0; k < m; ++k) {
            for (int s = 0; s < 8; ++s) {
                const int k2 = childInSlot[s];
                if (k2 == k) continue;
                const int s1 = slotOf[k];
                if (k2 < 0) {
                    if (cost[k][s] < cost[k][s1]) { childInSlot[s1] = -1; childInSlot[s] = k; slotOf[k] = s; changed = true; }
                } else if (cost[k][s] + cost[k2][s1] < cost[k][s1] + cost[k2][s]) {
                    childInSlot[s] = k; childInSlot[s1] = k2; slotOf[k] = s; slotOf[k2] = s1; changed = true;
                }
            }
        }
        if (!changed) break;
    }
#endif

    uint32_t innerCount = 0, triTotal = 0;
    for (int k = 0; k < m; ++k) {
        if (cnt[k] > 3u) ++innerCount;
        else triTotal += cnt[k];
    }
    uint32_t childBase = innerCount ? atomicAdd(nodeCounter, innerCount) : 0u;
    uint32_t triBase = triTotal ? atomicAdd(triCounter, triTotal) : 0u;
    uint32_t taskBase = innerCount ? atomicAdd(nextCount, innerCount) : 0u;

    // exponents: smallest e with 255 * 2^e >= extent
    float ext[3] = {nhi.x - nlo.x, nhi.y - nlo.y, nhi.z - nlo.z};
    uint32_t eb[3];
    float scale[3];
    for (int a = 0; a < 3; ++a) {
        int e = -126;
        if (ext[a] > 0.f) {
            e = (int)ceilf(log2f(ext[a] / 255.0f));
            while (ldexpf(255.0f, e) < ext[a]) ++e;  // guard against log2f rounding
        }
        e = max(-126, min(127, e));
        eb[a] = (uint32_t)(e + 127);
        scale[a] = __uint_as_float(eb[a] << 23);
    }

    uint32_t imask = 0;
    uint32_t meta[8], qlo[3][8], qhi[3][8];
    uint32_t innerRank = 0, triOffset = 0;
    for (int s = 0; s < 8; ++s) {
        int k = childInSlot[s];
        meta[s] = 0;
        for (int a = 0; a < 3; ++a) { qlo[a][s] = 255; qhi[a][s] = 0; }  // empty: inverted box, never hit
        if (k < 0) continue;
        float lo3[3] = {clo[k].x, clo[k].y, clo[k].z}, hi3[3] = {chi[k].x, chi[k].y, chi[k].z};
        float p3[3] = {nlo.x, nlo.y, nlo.z};
        for (int a = 0; a < 3; ++a) {
            float xl = (lo3[a] - p3[a]) / scale[a], xh = (hi3[a] - p3[a]) / scale[a];
            float fl = floorf(xl), ch = ceilf(xh);
            if (xl - fl < 4e-3f) fl -= 1.0f;  // keep a hair of slack against rounding in the traversal's slab test (bvh8.cuh: 2^-9 step)
            if (ch - xh < 4e-3f) ch += 1.0f;
            qlo[a][s] = (uint32_t)fminf(fmaxf(fl, 0.f), 255.f);
            qhi[a][s] = (uint32_t)fminf(fmaxf(ch, 0.f), 255.f);
        }
        if (cnt[k] > 3u) {
            imask |= 1u << s;
            meta[s] = (1u << 5) | (24u + (uint32_t)s);
            CollapseTask nt;
            nt.bvh2 = member[k];
            nt.node = childBase + innerRank;
            nextTasks[taskBase + innerRank] = nt;
            ++innerRank;
        } else {
            uint32_t unary = cnt[k] == 1u ? 1u : (cnt[k] == 2u ? 3u : 7u);
            meta[s] = (unary << 5) | triOffset;
            uint32_t first = firstOf(member[k]);
            for (uint32_t q = 0; q < cnt[k]; ++q) {
                BuildTri t = tris[sorted[first + q]];
                float4* o = outTris + (size_t)(triBase + triOffset + q) * 3;
                if (__float_as_uint(t.c.w) & kPrimCurveFlag) {  // part of a curve segment: the leaf record names the pool entry and the part
                    o[0] = make_float4(t.c.x, t.c.y, 0.f, t.a.w);
                    o[1] = make_float4(0.f, 0.f, 0.f, t.b.w);
                    o[2] = make_float4(0.f, 0.f, 0.f, t.c.w);
                    continue;
                }
                o[0] = t.a;
                o[1] = make_float4(t.b.x - t.a.x, t.b.y - t.a.y, t.b.z - t.a.z, t.b.w);
                o[2] = make_float4(t.c.x - t.a.x, t.c.y - t.a.y, t.c.z - t.a.z, t.c.w);
            }
            triOffset += cnt[k];
        }
    }
    auto pack4 = [](const uint32_t* b) { return b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24); };
    float4* o = outNodes + (size_t)task.node * 5;
    o[0] = make_float4(nlo.x, nlo.y, nlo.z, __uint_as_float(eb[0] | (eb[1] << 8) | (eb[2] << 16) | (imask << 24)));
    o[1] = make_float4(__uint_as_float(childBase), __uint_as_float(triBase), __uint_as_float(pack4(meta)), __uint_as_float(pack4(meta + 4)));
    o[2] = make_float4(__uint_as_float(pack4(qlo[0])), __uint_as_float(pack4(qlo[0] + 4)), __uint_as_float(pack4(qhi[0])), __uint_as_float(pack4(qhi[0] + 4)));
    o[3] = make_float4(__uint_as_float(pack4(qlo[1])), __uint_as_float(pack4(qlo[1] + 4)), __uint_as_float(pack4(qhi[1])), __uint_as_float(pack4(qhi[1] + 4)));
    o[4] = make_float4(__uint_as_float(pack4(qlo[2])), __uint_as_float(pack4(qlo[2] + 4)), __uint_as_float(pack4(qhi[2])), __uint_as_float(pack4(qhi[2] + 4)));
}

// ---- depth-first relayout of the finished BVH8 ------------------------------------------------------------------------------------------
// The collapse numbers wide nodes level by level, so a ray's root-to-leaf walk touches one distant region of the node array per level.
// Here every subtree becomes one contiguous range: a node's block of inner children, then the ranges of those children in order
// (descStart(child i) = descStart(n) + k + sum of the inner-node counts below the earlier children).  The leaf triangles follow the same
// order.  Only positions change -- childBase / triBase words are rewritten -- so traversal results are identical.
SELAS_D uint32_t wideNodeInnerCount(const float4* nodes, uint32_t i) { return (uint32_t)__popc(__float_as_uint(nodes[(size_t)i * 5].w) >> 24); }
SELAS_D uint32_t wideNodeLeafTriangles(const float4* nodes, uint32_t i) {
    const float4 n1 = nodes[(size_t)i * 5 + 1];
    uint32_t total = 0;
    for (int h = 0; h < 2; ++h) {
        const uint32_t meta4 = __float_as_uint(h == 0 ? n1.z : n1.w);
        for (int j = 0; j < 4; ++j) {
            const uint32_t meta = (meta4 >> (8 * j)) & 0xffu;
            if (meta != 0u && (meta & 0x1fu) < 24u) total += (uint32_t)__popc(meta >> 5);  // leaf slot: unary count | triangle offset
        }
    }
    return total;
}
__global__ void relayoutSubtreeKernel(const float4* __restrict__ nodes, uint32_t begin, uint32_t end, uint32_t* __restrict__ below) {
    const uint32_t i = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= end) return;
    const uint32_t k = wideNodeInnerCount(nodes, i), base = __float_as_uint(nodes[(size_t)i * 5 + 1].x);
    uint32_t sum = 0;
    for (uint32_t c = 0; c < k; ++c) sum += 1u + below[base + c];
    below[i] = sum;
}
__global__ void relayoutPlaceKernel(const float4* __restrict__ nodes, uint32_t begin, uint32_t end, const uint32_t* __restrict__ below,
                                    uint32_t* __restrict__ newIndex, uint32_t* __restrict__ descStart) {
    const uint32_t i = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= end) return;
    if (i == 0u) { newIndex[0] = 0u; descStart[0] = 1u; }
    const uint32_t k = wideNodeInnerCount(nodes, i), base = __float_as_uint(nodes[(size_t)i * 5 + 1].x);
    const uint32_t start = i == 0u ? 1u : descStart[i];
    uint32_t next = start + k;
    for (uint32_t c = 0; c < k; ++c) {
        newIndex[base + c] = start + c;
        descStart[base + c] = next;
        next += below[base + c];
    }
}
__global__ void relayoutLeafCountKernel(const float4* __restrict__ nodes, uint32_t n, const uint32_t* __restrict__ newIndex, uint32_t* __restrict__ countsNewOrder) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    countsNewOrder[newIndex[i]] = wideNodeLeafTriangles(nodes, i);
}
__global__ void relayoutScatterKernel(const float4* __restrict__ nodes, const float4* __restrict__ tris, uint32_t n, const uint32_t* __restrict__ newIndex,
                                      const uint32_t* __restrict__ descStart, const uint32_t* __restrict__ triStartNewOrder, float4* __restrict__ outNodes,
                                      float4* __restrict__ outTris) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t j = newIndex[i];
    float4 n0 = nodes[(size_t)i * 5], n1 = nodes[(size_t)i * 5 + 1];
    const uint32_t oldTriBase = __float_as_uint(n1.y), newTriBase = triStartNewOrder[j];
    const uint32_t count = wideNodeLeafTriangles(nodes, i);
    for (uint32_t t = 0; t < count; ++t)
        for (int q = 0; q < 3; ++q) outTris[(size_t)(newTriBase + t) * 3 + q] = tris[(size_t)(oldTriBase + t) * 3 + q];
    n1.x = __uint_as_float(wideNodeInnerCount(nodes, i) ? descStart[i] : 0u);
    n1.y = __uint_as_float(count ? newTriBase : 0u);
    outNodes[(size_t)j * 5] = n0;
    outNodes[(size_t)j * 5 + 1] = n1;
    for (int q = 2; q < 5; ++q) outNodes[(size_t)j * 5 + q] = nodes[(size_t)i * 5 + q];
}

}  // namespace selas
