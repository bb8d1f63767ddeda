// trace.cuh -- ray-cast kernels over the BVH8: persistent warps with per-lane refill.
//
// Launch shape: grid = SMs x resident CTAs, 128 threads per CTA; every warp keeps pulling rays from a
// global counter.  A lane that finishes its ray waits until fewer than REFILL_THRESHOLD lanes of its warp
// are still traversing, then the idle lanes are refilled together (one atomicAdd per refill, ballot-ranked
// slots) -- the warp-ballot compaction the north star asks for.  Ray and hit records are 32 bytes and are
// moved as two 16-byte vectors.
#pragma once
#include "bvh8.cuh"
#include "internal.h"

namespace selas {

#ifndef SELAS_TRACE_THREADS
#define SELAS_TRACE_THREADS 128
#endif
#ifndef SELAS_REFILL_THRESHOLD
#define SELAS_REFILL_THRESHOLD 20
#endif
#ifndef SELAS_TRACE_CTAS_PER_SM
#define SELAS_TRACE_CTAS_PER_SM 8
#endif
// Triangle postponing (after Ylitie, Karras, Laine 2017, section 4.3).  ncu (profiles/r1_trace_*): tested right after the node that
// produced them, leaf triangles took 28 % of the kernel's issue slots at 3 of 32 lanes.  Instead a lane parks the triangle group of
// its last node in a register pair and the warp votes every iteration: the parked groups are tested together once at least
// 1/SELAS_TRI_POSTPONE_DIV of the traversing lanes have one (or a lane has nothing else left, or needs the slot for a new group).
// 0 disables -- and is the default: measured on configs[2] (scripts/ab_trace.sh, DIV = 0/2/3/5/8) the closest-hit time stayed within
// noise (9.47-9.68 ms per step) while instructions went UP 6 % (2.32 -> 2.45 G warp instructions, lanes/instr 18.8 -> 19.6): leaf
// triangles are met at the end of a ray's traversal, when the lane has nothing else left, so most parked groups are tested alone anyway.
#ifndef SELAS_TRI_POSTPONE_DIV
#define SELAS_TRI_POSTPONE_DIV 0
#endif

// MODE 0: closest hit -> hits (+Ng); 1: any hit -> occluded byte; 2: any hit -> film[pixel] += payload.rgb when
// unoccluded (the NEE shadow ray of PathTracer.cpp:186-191 fused with FramebufferWriter_Write).
// nPtr != null: the ray count is read from device memory (wavefront queues), n is then only an upper bound.
// INST: two-level traversal (SceneInstanceIntersectFunction / InstanceOccludedFunction, SceneResource.cpp:87-187): a top-level leaf is an
// instance; the lane keeps what is left of the top-level node on its stack, maps the ray into the instance's object space
// (MatrixMultiplyPoint / MatrixMultiplyVector with worldToLocal -- the direction is not renormalised, so t means the same in both
// spaces, as in the reference), walks the mesh's tree with the same stack, and restores the world ray when the stack is back at the mark.
SELAS_D void setupRayDirection(V3 d, V3& idir, bool& negx, bool& negy, bool& negz, uint32_t& octInv4) {
    const float tiny = 1e-30f;
    idir = v3(1.0f / (fabsf(d.x) > tiny ? d.x : copysignf(tiny, d.x)), 1.0f / (fabsf(d.y) > tiny ? d.y : copysignf(tiny, d.y)),
              1.0f / (fabsf(d.z) > tiny ? d.z : copysignf(tiny, d.z)));
    negx = d.x < 0.0f; negy = d.y < 0.0f; negz = d.z < 0.0f;
    uint32_t oct = (negx ? 4u : 0u) | (negy ? 2u : 0u) | (negz ? 1u : 0u);
    octInv4 = (7u - oct) * 0x01010101u;
}

// Leaf triangles deferred by one node (round 2).  ncu on the bounce-0 launch (profiles/r2_trace_bounce0_sass_summary.txt): the triangle
// tests were 30 % of the kernel's warp instructions at 3.3 lanes and HALF of its stall samples, most of them on the first use of the
// triangle's LDGs -- a lane loaded the node, waited, tested it, then loaded the leaf triangles and waited again, twice the memory latency
// per step with almost nothing in flight the second time.  With SELAS_TRI_DEFER the leaf hits of a node are parked in two registers and
// tested in the NEXT iteration, after the loads of the next node have been issued: both latencies overlap, and because the parked
// triangles are still tested before that node's slab test, the test sees the updated tmax (at most the choice of the next node is one
// hit out of date).  Two-level scenes keep the direct order (a top-level "triangle" is an instance entry that changes the ray).
// MEASURED (profiles/r2_ab_tri_defer*.jsonl): slower both ways -- holding the next node's five vectors in registers across the triangle
// test spills the ray constants (closest hit 8.25 -> 8.72 ms per step), and the register-free form below (prefetch.global.L1 of the
// triangles when they are found and of the next node before they are tested) costs the same (8.71): the extra turn of the loop for the
// parked group and the prefetch instructions outweigh the latency, which the other seven warps of the scheduler were already hiding
// (issue slots are 70 % busy).  Off; the stall samples on the triangle loads were not lost time.
#ifndef SELAS_TRI_DEFER
#define SELAS_TRI_DEFER 0
#endif

SELAS_D void prefetchL1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// The triangles (or curve parts) of one leaf group against the lane's ray; returns true when an any-hit ray is done.
template <bool ANY, bool COUNT, bool CURVES>
SELAS_D bool testLeafGroup(const Bvh8& bvh, uint2 triGroup, V3 o, V3 d, float tnear, float& tmax, TraceResult& res, uint32_t& nTris) {
    while (triGroup.y != 0u) {
        const uint32_t bit = 31u - __clz(triGroup.y);
        triGroup.y &= ~(1u << bit);
        const uint32_t ti = triGroup.x + bit;
        const float4* tp = bvh.tris + (size_t)ti * 3;
        const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
        if (COUNT) ++nTris;
        float t, u, v;
        if (CURVES && (__float_as_uint(c.w) & kLeafCurveFlag)) {
            const float4* cp = bvh.curves + 4 * (size_t)__float_as_uint(a.x);
            CurveHit ch;
            if (intersectCurve(__ldg(cp), __ldg(cp + 1), __ldg(cp + 2), __ldg(cp + 3), __float_as_uint(a.y), o, d, tnear, tmax, ch)) {
                tmax = ch.t;
                res.t = ch.t; res.u = ch.u; res.v = 0.f; res.tri = ti; res.element = ch.element; res.ng = ch.ng;
                if (ANY) return true;
            }
        } else if (intersectTri(a, b, c, o, d, tnear, tmax, t, u, v)) {
            tmax = t;
            res.t = t; res.u = u; res.v = v; res.tri = ti;
            if (ANY) return true;
        }
    }
    return false;
}

template <int MODE, bool COUNT, bool INST = false, bool TOP = false, bool CURVES = false>
__global__ void __launch_bounds__(SELAS_TRACE_THREADS, (INST || COUNT || CURVES) ? 6 : SELAS_TRACE_CTAS_PER_SM) traceKernel(Bvh8 bvh, const SelasRay* __restrict__ rays, uint32_t n,
                                                                   const uint32_t* __restrict__ nPtr,
                                                                   SelasHit* __restrict__ hits, uint8_t* __restrict__ occluded,
                                                                   float4* __restrict__ ng, uint32_t* __restrict__ counts,
                                                                   const float4* __restrict__ payload, float* __restrict__ film,
                                                                   uint32_t* __restrict__ nextRay,
                                                                   unsigned long long* __restrict__ totals,
                                                                   const uint32_t* __restrict__ order = nullptr) {
    constexpr bool ANY = MODE != 0;
    static_assert(!(INST && SELAS_TRI_POSTPONE_DIV), "parked triangle groups are not carried across instance boundaries");
    // TOP: hot top levels in shared memory (north star (1)): every ray starts at the root and the first levels are visited by all of
    // them; from shared memory 32 lanes on 32 different nodes are served bank-wise instead of one L1 tag lookup per line.  Measured on
    // B200 (profiles/r2_ab_smem_top.jsonl) it does not pay: 9 / 73 / 585 / 1200 staged nodes at 128 / 256 / 512 threads per CTA are all
    // 0.2 .. 2 % SLOWER than no staging on configs[2] and configs[1] (the top levels already hit in L1, the shared-memory carve-out
    // shrinks that L1, and the select costs registers), so the variant is compiled but off by default (SELAS_TRACE_TOP_NODES=n enables).
    extern __shared__ float4 sTopNodes[];
    const uint32_t topCount = TOP ? bvh.topCount : 0u;
    if constexpr (TOP) {
        for (uint32_t i = threadIdx.x; i < bvh.topCount * 5u; i += blockDim.x) sTopNodes[i] = __ldg(bvh.nodes + i);
        __syncthreads();
    }
    if (nPtr) n = min(n, *nPtr);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t ltMask = (1u << lane) - 1u;

    // per-lane traversal state
    uint2 stack[SELAS_STACK_SIZE];
    int sp = 0;
    uint2 group = make_uint2(0u, 0u);
    uint2 pendTri = make_uint2(0u, 0u);  // parked leaf triangles (SELAS_TRI_DEFER / SELAS_TRI_POSTPONE_DIV)
    (void)pendTri;
    V3 o = v3(0.f), d = v3(0.f), idir = v3(0.f);
    float tnear = 0.f, tfar = 0.f, tmax = 0.f;
    bool negx = false, negy = false, negz = false;
    uint32_t octInv4 = 0;
    TraceResult res;
    res.t = 0.f; res.u = 0.f; res.v = 0.f; res.tri = 0xFFFFFFFFu; res.inst = 0u; res.element = 0u; res.ng = v3(0.f);
    // two-level state (INST only): world ray while inside an instance, stack mark, the instance's node / triangle bases
    V3 ow = v3(0.f), dw = v3(0.f);
    int spMark = 0;
    uint32_t nodeBase = 0u, triBase = 0u, curInst = 0u;
    bool inBlas = false;
    uint32_t rayIdx = 0xFFFFFFFFu;
    uint32_t nNodes = 0, nTris = 0;
    unsigned long long sumNodes = 0, sumTris = 0;
    bool active = false;
    bool exhausted = false;

    for (;;) {
        // ---- refill idle lanes ----
        uint32_t idle = __ballot_sync(0xffffffffu, !active);
        if (idle && !exhausted) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(nextRay, (uint32_t)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base >= n) exhausted = true;
            if (!active) {
                uint32_t idx = base + __popc(idle & ltMask);
                if (idx < n) {
                    // order != null: the queue is visited in a direction-sorted order (rays of one octant and region side by side in a
                    // warp); results still land at the ray's own index
                    if (order) idx = __ldg(order + idx);
                    const float4* rp = reinterpret_cast<const float4*>(rays + idx);
                    float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
                    o = v3(r0.x, r0.y, r0.z);
                    d = v3(r1.x, r1.y, r1.z);
                    tnear = r0.w;
                    tfar = r1.w;
                    tmax = tfar;
                    setupRayDirection(d, idir, negx, negy, negz, octInv4);
                    res.t = tfar; res.u = 0.f; res.v = 0.f; res.tri = 0xFFFFFFFFu;
                    sp = 0;
                    if (INST) { inBlas = false; nodeBase = 0u; triBase = 0u; }
                    // A ray with a non-finite origin / direction would pass every slab test (NaN compares false both ways)
                    // and walk the whole tree; Embree reports such rays as misses, and so does this kernel, immediately.
                    const bool valid = isfinite(o.x) && isfinite(o.y) && isfinite(o.z) && isfinite(d.x) && isfinite(d.y) && isfinite(d.z) &&
                                       tnear <= tfar;
                    group = make_uint2(0u, (bvh.nodeCount && valid) ? 0x80000000u : 0u);
                    rayIdx = idx;
                    nNodes = 0; nTris = 0;
                    pendTri = make_uint2(0u, 0u);
                    active = true;
                }
            }
        }
        if (!__any_sync(0xffffffffu, active)) break;

        // ---- traverse until this lane finishes or the warp runs thin ----
        if constexpr (SELAS_TRI_DEFER && !INST && !TOP && !SELAS_TRI_POSTPONE_DIV) {
            while (active) {
                bool finished = false;
                const bool hasNode = (group.y & 0xff000000u) != 0u;
                const float4* np = bvh.nodes;
                if (hasNode) {
                    const uint32_t hitsImask = group.y;
                    const uint32_t childBit = nextChildBit<MODE == 2>(hitsImask);
                    group.y &= ~(1u << childBit);
                    if ((group.y & 0xff000000u) && sp < SELAS_STACK_SIZE) stack[sp++] = group;
                    const uint32_t slot = (childBit - 24u) ^ (octInv4 & 0xffu);
                    const uint32_t rel = __popc(hitsImask & ~(0xffffffffu << slot));
                    np = bvh.nodes + (size_t)(group.x + rel) * 5;
                    if (pendTri.y != 0u) {  // the node's lines start their way up while the parked triangles are tested
                        prefetchL1(np);
                        prefetchL1(np + 4);
                    }
                    if (COUNT) ++nNodes;
                }
                // the previous node's leaf hits (prefetched when they were found)
                if (pendTri.y != 0u) {
                    finished = testLeafGroup<ANY, COUNT, CURVES>(bvh, pendTri, o, d, tnear, tmax, res, nTris);
                    pendTri = make_uint2(0u, 0u);
                }
                if (hasNode && !finished) {
                    const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
                    const uint32_t hitMask = intersectNode(n0, n1, n2, n3, n4, o, idir, negx, negy, negz, octInv4, tnear, tmax);
                    group.x = __float_as_uint(n1.x);
                    group.y = (hitMask & 0xff000000u) | extractByte(__float_as_uint(n0.w), 3);
                    pendTri = make_uint2(__float_as_uint(n1.y), hitMask & 0x00ffffffu);
                    if (pendTri.y != 0u) {
                        const float4* tlo = bvh.tris + (size_t)(pendTri.x + (uint32_t)__ffs((int)pendTri.y) - 1u) * 3;
                        const float4* thi = bvh.tris + (size_t)(pendTri.x + 31u - (uint32_t)__clz(pendTri.y)) * 3;
                        prefetchL1(tlo);
                        prefetchL1(tlo + 2);
                        prefetchL1(thi);
                        prefetchL1(thi + 2);
                    }
                }
                if (!finished && (group.y & 0xff000000u) == 0u) {
                    if (sp > 0) group = stack[--sp];
                    else {
                        group = make_uint2(0u, 0u);
                        finished = pendTri.y == 0u;  // else one more turn for the parked triangles
                    }
                }
                if (finished) {
                    if (MODE == 1) occluded[rayIdx] = res.tri != 0xFFFFFFFFu ? 1 : 0;
                    else if (MODE == 2) {
                        if (res.tri == 0xFFFFFFFFu) {
                            float4 pl = __ldg(payload + rayIdx);
                            float* px = film + 3 * (size_t)__float_as_uint(pl.w);
                            atomicAdd(px, pl.x);
                            atomicAdd(px + 1, pl.y);
                            atomicAdd(px + 2, pl.z);
                        }
                    } else writeHit<INST>(bvh, res, tfar, hits + rayIdx, ng ? ng + rayIdx : nullptr);
                    if (COUNT) {
                        if (counts) { counts[2 * (size_t)rayIdx] = nNodes; counts[2 * (size_t)rayIdx + 1] = nTris; }
                        sumNodes += nNodes;
                        sumTris += nTris;
                    }
                    active = false;
                }
                if (!exhausted && __popc(__activemask()) < SELAS_REFILL_THRESHOLD) break;
            }
        } else
        while (active) {
            bool finished = false;
            uint2 triGroup;

            if (group.y & 0xff000000u) {
                uint32_t hitsImask = group.y;
                uint32_t childBit = nextChildBit<MODE == 2>(hitsImask);
                group.y &= ~(1u << childBit);
                if ((group.y & 0xff000000u) && sp < SELAS_STACK_SIZE) stack[sp++] = group;
                uint32_t slot = (childBit - 24u) ^ (octInv4 & 0xffu);
                uint32_t rel = __popc(hitsImask & ~(0xffffffffu << slot));
                const uint32_t nodeIdx = (INST ? nodeBase : 0u) + group.x + rel;
                float4 n0, n1, n2, n3, n4;
                if (TOP && nodeIdx < topCount) {
                    const float4* sp = sTopNodes + nodeIdx * 5u;
                    n0 = sp[0]; n1 = sp[1]; n2 = sp[2]; n3 = sp[3]; n4 = sp[4];
                } else {
                    const float4* np = bvh.nodes + (size_t)nodeIdx * 5;
                    n0 = __ldg(np); n1 = __ldg(np + 1); n2 = __ldg(np + 2); n3 = __ldg(np + 3); n4 = __ldg(np + 4);
                }
                if (COUNT) ++nNodes;
                uint32_t hitMask = intersectNode(n0, n1, n2, n3, n4, o, idir, negx, negy, negz, octInv4, tnear, tmax);
                group.x = __float_as_uint(n1.x);
                triGroup.x = __float_as_uint(n1.y);
                group.y = (hitMask & 0xff000000u) | extractByte(__float_as_uint(n0.w), 3);
                triGroup.y = hitMask & 0x00ffffffu;
            } else {
                triGroup = group;
                group = make_uint2(0u, 0u);
            }
#if SELAS_TRI_POSTPONE_DIV
            {
                if (triGroup.y != 0u) {  // new leaf triangles: park them; if the slot is taken, the older group is tested now
                    uint2 older = pendTri;
                    pendTri = triGroup;
                    triGroup = older;
                }
                const uint32_t amask = __activemask();
                const uint32_t parked = __ballot_sync(amask, pendTri.y != 0u);
                const bool nothingElse = (group.y & 0xff000000u) == 0u && sp == 0;
                if (triGroup.y == 0u && pendTri.y != 0u && (__popc(parked) * SELAS_TRI_POSTPONE_DIV >= __popc(amask) || nothingElse)) {
                    triGroup = pendTri;
                    pendTri = make_uint2(0u, 0u);
                }
            }
#endif
            while (triGroup.y != 0u) {
                uint32_t bit = 31u - __clz(triGroup.y);
                triGroup.y &= ~(1u << bit);
                uint32_t ti = triGroup.x + bit;
                if (INST && !inBlas) {
                    // top-level leaf: enter the instance.  What is left of this leaf group and of the current node goes on the stack
                    // (below the mark), then the lane continues in object space from the root of the mesh's tree.
                    const uint32_t ii = __float_as_uint(__ldg(bvh.tris + (size_t)ti * 3).w);
                    if (triGroup.y != 0u && sp < SELAS_STACK_SIZE) stack[sp++] = triGroup;
                    if ((group.y & 0xff000000u) && sp < SELAS_STACK_SIZE) stack[sp++] = group;
                    spMark = sp;
                    ow = o; dw = d;
                    const float4* ip = reinterpret_cast<const float4*>(bvh.instances + ii);
                    float4 m0 = __ldg(ip), m1 = __ldg(ip + 1), m2 = __ldg(ip + 2);
                    uint4 m3 = __ldg(reinterpret_cast<const uint4*>(ip + 3));
                    // w2l rows: r0 = (m0.x, m0.y, m0.z), r1 = (m0.w, m1.x, m1.y), r2 = (m1.z, m1.w, m2.x), t = (m2.y, m2.z, m2.w)
                    o = v3(ow.x * m0.x + ow.y * m0.w + ow.z * m1.z + m2.y, ow.x * m0.y + ow.y * m1.x + ow.z * m1.w + m2.z,
                           ow.x * m0.z + ow.y * m1.y + ow.z * m2.x + m2.w);
                    d = v3(dw.x * m0.x + dw.y * m0.w + dw.z * m1.z, dw.x * m0.y + dw.y * m1.x + dw.z * m1.w, dw.x * m0.z + dw.y * m1.y + dw.z * m2.x);
                    setupRayDirection(d, idir, negx, negy, negz, octInv4);
                    nodeBase = m3.x; triBase = m3.y; curInst = ii;
                    inBlas = true;
                    group = make_uint2(0u, 0x80000000u);
                    triGroup = make_uint2(0u, 0u);
                    break;
                }
                if (INST) ti += triBase;
                const float4* tp = bvh.tris + (size_t)ti * 3;
                float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
                if (COUNT) ++nTris;
                float t, u, v;
                if (CURVES && (__float_as_uint(c.w) & kLeafCurveFlag)) {
                    // part of a round B-spline curve segment (ModelResource.cpp:300-333): the leaf record names the segment's control vertices + the part
                    const float4* cp = bvh.curves + 4 * (size_t)__float_as_uint(a.x);
                    CurveHit ch;
                    if (intersectCurve(__ldg(cp), __ldg(cp + 1), __ldg(cp + 2), __ldg(cp + 3), __float_as_uint(a.y), o, d, tnear, tmax, ch)) {
                        tmax = ch.t;
                        res.t = ch.t; res.u = ch.u; res.v = 0.f; res.tri = ti; res.element = ch.element; res.ng = ch.ng;
                        if (INST) res.inst = curInst;
                        if (ANY) { finished = true; break; }
                    }
                } else if (intersectTri(a, b, c, o, d, tnear, tmax, t, u, v)) {
                    tmax = t;
                    res.t = t; res.u = u; res.v = v; res.tri = ti;
                    if (INST) res.inst = curInst;
                    if (ANY) { finished = true; break; }
                }
            }
            if (INST && inBlas && !finished && (group.y & 0xff000000u) == 0u && sp == spMark) {
                // the instance's tree is exhausted: back to world space and to what was left at the top level
                o = ow; d = dw;
                setupRayDirection(d, idir, negx, negy, negz, octInv4);
                inBlas = false;
                nodeBase = 0u; triBase = 0u;
            }
            if (!finished && (group.y & 0xff000000u) == 0u) {
#if SELAS_TRI_POSTPONE_DIV
                if (sp == 0) {  // parked triangles are tested next iteration (nothingElse); the spent group must not be read as one
                    group = make_uint2(0u, 0u);
                    finished = pendTri.y == 0u;
                }
#else
                if (sp == 0) finished = true;
#endif
                else group = stack[--sp];
            }
            if (finished) {
                if (MODE == 1) occluded[rayIdx] = res.tri != 0xFFFFFFFFu ? 1 : 0;
                else if (MODE == 2) {
                    if (res.tri == 0xFFFFFFFFu) {
                        float4 pl = __ldg(payload + rayIdx);
                        float* px = film + 3 * (size_t)__float_as_uint(pl.w);
                        atomicAdd(px, pl.x);
                        atomicAdd(px + 1, pl.y);
                        atomicAdd(px + 2, pl.z);
                    }
                } else writeHit<INST>(bvh, res, tfar, hits + rayIdx, ng ? ng + rayIdx : nullptr);
                if (COUNT) {
                    if (counts) { counts[2 * (size_t)rayIdx] = nNodes; counts[2 * (size_t)rayIdx + 1] = nTris; }
                    sumNodes += nNodes;
                    sumTris += nTris;
                }
                active = false;
            }
            if (!exhausted && __popc(__activemask()) < SELAS_REFILL_THRESHOLD) break;
        }
    }
    if (COUNT && totals) {
        for (int s = 16; s > 0; s >>= 1) {
            sumNodes += __shfl_xor_sync(0xffffffffu, sumNodes, s);
            sumTris += __shfl_xor_sync(0xffffffffu, sumTris, s);
        }
        if (lane == 0) {
            atomicAdd(totals, sumNodes);
            atomicAdd(totals + 1, sumTris);
        }
    }
}

}  // namespace selas
