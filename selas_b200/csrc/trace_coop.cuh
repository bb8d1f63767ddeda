// trace_coop.cuh -- the BVH8 ray-cast kernel with a warp-cooperative triangle stage (round 2).  The integrators use it from the first
// bounce on and for all shadow rays, scenes with curve geometry for every launch; SELAS_TRACE_COOP=1 / 0 forces it on / off (trace.cu).
//
// Why: ncu on the closest-hit launch of configs[2] (profiles/r2_trace_bounce0_sass_summary.txt) shows the leaf triangle tests as 30 % of
// the kernel's warp instructions at 3.3 of 32 lanes: a ray meets its 3.2 leaf triangles at a handful of its 12 node visits, so in any
// one iteration only a few lanes of the warp have triangles and each runs Moller-Trumbore alone.  Here a lane does not test the leaf
// triangles its node test finds; it appends (lane, triangle) pairs to a per-warp queue in shared memory and goes on traversing.  When
// the queue holds enough pairs (or rays are waiting on theirs) the WHOLE warp -- idle lanes included -- tests one pair per lane: the
// source lane's ray comes over with eight shuffles, the hit goes back through a 64-bit atomicMin on {t, triangle} in shared memory.
// Closest-hit results do not depend on the order of the tests (minimum over all candidates; equal t: the lower triangle index, so
// reruns are bit-identical although the queue order is not), and tmax is only an optimisation for the slab test, so letting it lag a
// few node visits is safe.
//
// Queue layout, measured: one 32-bit (lane, triangle) pair per slot, written by the lane that found it (a short loop at 3-4 lanes, 13 %
// of the kernel's instructions).  Queueing {triangle base, hit mask, lane} once per lane and node instead and expanding to pairs inside
// the flush (scan of the counts + __reduce_or of the entry heads per window of 32 pairs) removes that loop but makes every flush round
// dearer: 16.00-16.05 ms per step against 15.88 (profiles/r2_ab_trace_coop_group_queue.jsonl).  Pairs stay.
//
// A per-pixel shadow cache (the triangle that blocked the pixel's last first-bounce shadow ray, tried before any traversal) was built and
// measured: 12.06 -> 12.04 nodes per shadow ray and 2 % slower (profiles/r2_ab_shadow_cache.jsonl) -- too few of configs[2]'s shadow rays
// are blocked at all, and by the same triangle even fewer.  Removed.
//
// Same interface, launch shape and ray pulling (ballot-ranked refill) as traceKernel.  CURVES: leaves may be parts of round B-spline
// curves (curve.cuh); the element code and normal of a curve hit travel back beside (u, v).  INST: two-level trees (trace.cuh): a lane
// enters and leaves instances as in traceKernel; the pairs it queues name the instance, and the lane that tests a pair maps the source
// lane's WORLD ray into that instance's space itself (t means the same in both spaces), so a pair stays valid after its ray has left
// the instance again.
#pragma once
#include "trace.cuh"

namespace selas {

// Tuning (configs[2], ms per step, profiles/r2_ab_trace_coop*.jsonl): flush at 12 / 16 / 24 / 32 queued pairs 16.29 / 16.21 / 16.34 / 16.59;
// flush when 1 / 2 / 4 / 6 / 8 rays wait on their pairs 16.98 / 16.80 / 16.21 / 16.21 / 16.74; refill below 16 / 20 / 24 / 28 / 32 busy lanes
// 16.72 / 16.34 / 16.23 / 16.36 / 17.48; queue capacity 64 / 96: the same.
#ifndef SELAS_COOP_QCAP
#define SELAS_COOP_QCAP 96
#endif
#ifndef SELAS_COOP_FLUSH_MIN
#define SELAS_COOP_FLUSH_MIN 16
#endif
#ifndef SELAS_COOP_REFILL_THRESHOLD
#define SELAS_COOP_REFILL_THRESHOLD 24
#endif
#ifndef SELAS_COOP_WAIT_MAX
#define SELAS_COOP_WAIT_MAX 4
#endif
SELAS_D uint32_t orderedFloatBits(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
SELAS_D float fromOrderedFloatBits(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

template <int MODE, bool COUNT, bool CURVES = false, bool INST = false>
__global__ void __launch_bounds__(SELAS_TRACE_THREADS, (COUNT || CURVES || INST) ? 6 : SELAS_TRACE_CTAS_PER_SM)
    traceCoopKernel(Bvh8 bvh, const SelasRay* __restrict__ rays, uint32_t n, const uint32_t* __restrict__ nPtr, SelasHit* __restrict__ hits,
                    uint8_t* __restrict__ occluded, float4* __restrict__ ng, uint32_t* __restrict__ counts, const float4* __restrict__ payload,
                    float* __restrict__ film, uint32_t* __restrict__ nextRay, unsigned long long* __restrict__ totals,
                    const uint32_t* __restrict__ order = nullptr) {
    constexpr bool ANY = MODE != 0;
    constexpr int WARPS = SELAS_TRACE_THREADS / 32;
    constexpr uint32_t FULL = 0xffffffffu;
    __shared__ uint32_t sQueue[WARPS][SELAS_COOP_QCAP];          // lane | triangle << 5
    __shared__ uint32_t sTail[WARPS];                            // queue length
    __shared__ float2 sUv[WARPS][32];                            // per source lane: (u, v) of its best hit
    __shared__ unsigned long long sBest[WARPS][32];              // per source lane: ordered(t) << 32 | triangle
    __shared__ float4 sCurve[CURVES ? WARPS : 1][32];            // per source lane: curve normal and element code of its best hit
    __shared__ uint32_t sQueueInst[INST ? WARPS : 1][INST ? SELAS_COOP_QCAP : 1];  // the pair's instance
    __shared__ uint32_t sInst[INST ? WARPS : 1][32];             // per source lane: instance of its best hit
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t ltMask = (1u << lane) - 1u;
    uint32_t* const q = sQueue[warp];
    float2* const quv = sUv[warp];
    unsigned long long* const best = sBest[warp];
    float4* const qcurve = sCurve[CURVES ? warp : 0];
    uint32_t* const qinst = sQueueInst[INST ? warp : 0];
    uint32_t* const bestInst = sInst[INST ? warp : 0];
    uint32_t* const tail = &sTail[warp];
    if (lane == 0) *tail = 0u;
    __syncwarp();
    if (nPtr) n = min(n, *nPtr);

    uint2 stack[SELAS_STACK_SIZE];
    int sp = 0;
    uint2 group = make_uint2(0u, 0u);
    V3 o = v3(0.f), d = v3(0.f), idir = v3(0.f);
    float tnear = 0.f, tfar = 0.f, tmax = 0.f;
    bool negx = false, negy = false, negz = false;
    uint32_t octInv4 = 0;
    TraceResult res;
    res.t = 0.f; res.u = 0.f; res.v = 0.f; res.tri = 0xFFFFFFFFu; res.inst = 0u; res.element = 0u; res.ng = v3(0.f);
    uint32_t rayIdx = 0xFFFFFFFFu;
    uint32_t nNodes = 0, nTris = 0;
    unsigned long long sumNodes = 0, sumTris = 0;
    bool active = false, exhausted = false;
    bool outstanding = false;  // this lane has pairs in the queue
    bool done = false;         // any-hit: a queued triangle was hit
    // two-level state (INST): the world ray (always), stack mark, the instance's node / triangle bases; a popped remainder of top-level leaf hits
    V3 ow = v3(0.f), dw = v3(0.f);
    int spMark = 0;
    uint32_t nodeBase = 0u, triBase = 0u, curInst = 0u;
    bool inBlas = false, pendingLeaf = false;
    uint32_t qCount = 0;       // warp-uniform

    // the whole warp tests the queued pairs, one per lane and round
    auto flush = [&]() {
        __syncwarp();
        best[lane] = ((unsigned long long)orderedFloatBits(tmax) << 32) | 0xFFFFFFFFull;
        __syncwarp();
        for (uint32_t base = 0; base < qCount; base += 32u) {
            const uint32_t e = base + lane;
            const bool valid = e < qCount;
            const uint32_t entry = valid ? q[e] : 0u;
            const uint32_t src = entry & 31u;
            V3 so = v3(__shfl_sync(FULL, INST ? ow.x : o.x, src), __shfl_sync(FULL, INST ? ow.y : o.y, src), __shfl_sync(FULL, INST ? ow.z : o.z, src));
            V3 sd = v3(__shfl_sync(FULL, INST ? dw.x : d.x, src), __shfl_sync(FULL, INST ? dw.y : d.y, src), __shfl_sync(FULL, INST ? dw.z : d.z, src));
            const float stn = __shfl_sync(FULL, tnear, src), stm = __shfl_sync(FULL, tmax, src);
            uint32_t pairInst = 0u;
            if (INST && valid) {  // the world ray into the pair's instance (SceneInstanceIntersectFunction, SceneResource.cpp:87-142)
                pairInst = qinst[e];
                const float4* ip = reinterpret_cast<const float4*>(bvh.instances + pairInst);
                const float4 m0 = __ldg(ip), m1 = __ldg(ip + 1), m2 = __ldg(ip + 2);
                const V3 wo = so, wd = sd;
                so = v3(wo.x * m0.x + wo.y * m0.w + wo.z * m1.z + m2.y, wo.x * m0.y + wo.y * m1.x + wo.z * m1.w + m2.z,
                        wo.x * m0.z + wo.y * m1.y + wo.z * m2.x + m2.w);
                sd = v3(wd.x * m0.x + wd.y * m0.w + wd.z * m1.z, wd.x * m0.y + wd.y * m1.x + wd.z * m1.w, wd.x * m0.z + wd.y * m1.y + wd.z * m2.x);
            }
            bool hit = false;
            unsigned long long key = 0ull;
            float t, u, v;
            float4 curveOut = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) {
                const float4* tp = bvh.tris + (size_t)(entry >> 5) * 3;
                const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
                if (CURVES && (__float_as_uint(c.w) & kLeafCurveFlag)) {
                    const float4* cp = bvh.curves + 4 * (size_t)__float_as_uint(a.x);
                    CurveHit ch;
                    hit = intersectCurve(__ldg(cp), __ldg(cp + 1), __ldg(cp + 2), __ldg(cp + 3), __float_as_uint(a.y), so, sd, stn, stm, ch);
                    if (hit) {
                        t = ch.t; u = ch.u; v = 0.f;
                        curveOut = make_float4(ch.ng.x, ch.ng.y, ch.ng.z, __uint_as_float(ch.element));
                    }
                } else {
                    hit = intersectTri(a, b, c, so, sd, stn, stm, t, u, v);
                }
                if (hit) {
                    key = ((unsigned long long)orderedFloatBits(t) << 32) | (unsigned long long)(entry >> 5);
                    atomicMin(&best[src], key);
                }
            }
            __syncwarp();
            if (hit && best[src] == key) {  // the best so far (a later round's better hit overwrites it)
                quv[src] = make_float2(u, v);
                if (CURVES) qcurve[src] = curveOut;
                if (INST) bestInst[src] = pairInst;
            }
            __syncwarp();
        }
        if (outstanding) {
            const unsigned long long b = best[lane];
            const uint32_t tri = (uint32_t)b;
            if (tri != 0xFFFFFFFFu) {
                const float t = fromOrderedFloatBits((uint32_t)(b >> 32));
                const float2 uv = quv[lane];
                tmax = t;
                res.t = t; res.u = uv.x; res.v = uv.y; res.tri = tri;
                if (INST) res.inst = bestInst[lane];
                if (CURVES) {
                    const float4 cv = qcurve[lane];
                    res.ng = v3(cv.x, cv.y, cv.z);
                    res.element = __float_as_uint(cv.w);
                }
                if (ANY) done = true;
            }
            outstanding = false;
        }
        qCount = 0;
        if (lane == 0) *tail = 0u;
        __syncwarp();
    };

    for (;;) {
        // ---- refill idle lanes once the warp runs thin ----
        const uint32_t idle = __ballot_sync(FULL, !active);
        if (!exhausted && 32 - __popc(idle) < SELAS_COOP_REFILL_THRESHOLD) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(nextRay, (uint32_t)__popc(idle));
            base = __shfl_sync(FULL, base, 0);
            if (base >= n) exhausted = true;
            if (!active) {
                uint32_t idx = base + __popc(idle & ltMask);
                if (idx < n) {
                    if (order) idx = __ldg(order + idx);
                    const float4* rp = reinterpret_cast<const float4*>(rays + idx);
                    const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
                    o = v3(r0.x, r0.y, r0.z);
                    d = v3(r1.x, r1.y, r1.z);
                    tnear = r0.w;
                    tfar = r1.w;
                    tmax = tfar;
                    setupRayDirection(d, idir, negx, negy, negz, octInv4);
                    res.t = tfar; res.u = 0.f; res.v = 0.f; res.tri = 0xFFFFFFFFu;
                    sp = 0;
                    if (INST) { ow = o; dw = d; inBlas = false; pendingLeaf = false; nodeBase = 0u; triBase = 0u; }
                    const bool valid = isfinite(o.x) && isfinite(o.y) && isfinite(o.z) && isfinite(d.x) && isfinite(d.y) && isfinite(d.z) &&
                                       tnear <= tfar;
                    group = make_uint2(0u, (bvh.nodeCount && valid) ? 0x80000000u : 0u);
                    rayIdx = idx;
                    nNodes = 0; nTris = 0;
                    outstanding = false;
                    done = false;
                    active = true;
                }
            }
        }
        if (!__any_sync(FULL, active)) break;

        // ---- one node per traversing lane ----
        uint32_t leafMask = 0u, leafBase = 0u;
        if (active && (group.y & 0xff000000u)) {
            const uint32_t hitsImask = group.y;
            const uint32_t childBit = nextChildBit<MODE == 2>(hitsImask);
            group.y &= ~(1u << childBit);
            if ((group.y & 0xff000000u) && sp < SELAS_STACK_SIZE) stack[sp++] = group;
            const uint32_t slot = (childBit - 24u) ^ (octInv4 & 0xffu);
            const uint32_t rel = __popc(hitsImask & ~(0xffffffffu << slot));
            const float4* np = bvh.nodes + (size_t)((INST ? nodeBase : 0u) + group.x + rel) * 5;
            const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
            if (COUNT) ++nNodes;
            const uint32_t hitMask = intersectNode(n0, n1, n2, n3, n4, o, idir, negx, negy, negz, octInv4, tnear, tmax);
            group.x = __float_as_uint(n1.x);
            group.y = (hitMask & 0xff000000u) | extractByte(__float_as_uint(n0.w), 3);
            leafBase = __float_as_uint(n1.y);
            leafMask = hitMask & 0x00ffffffu;
        } else if (INST && active && pendingLeaf) {  // what an earlier top-level node left of its instance hits
            leafBase = group.x;
            leafMask = group.y & 0x00ffffffu;
            group = make_uint2(0u, 0u);
            pendingLeaf = false;
        }
        if (INST && leafMask != 0u && !inBlas) {
            // top-level leaf hits are instances: enter the first one now (what is left of the hits and of the node waits on the stack, below
            // the mark), as traceKernel does
            const uint32_t bit = 31u - __clz(leafMask);
            leafMask &= ~(1u << bit);
            const uint32_t ii = __float_as_uint(__ldg(bvh.tris + (size_t)(leafBase + bit) * 3).w);
            if (leafMask != 0u && sp < SELAS_STACK_SIZE) stack[sp++] = make_uint2(leafBase, leafMask);
            if ((group.y & 0xff000000u) && sp < SELAS_STACK_SIZE) stack[sp++] = group;
            spMark = sp;
            const float4* ip = reinterpret_cast<const float4*>(bvh.instances + ii);
            const float4 m0 = __ldg(ip), m1 = __ldg(ip + 1), m2 = __ldg(ip + 2);
            const uint4 m3 = __ldg(reinterpret_cast<const uint4*>(ip + 3));
            o = v3(ow.x * m0.x + ow.y * m0.w + ow.z * m1.z + m2.y, ow.x * m0.y + ow.y * m1.x + ow.z * m1.w + m2.z,
                   ow.x * m0.z + ow.y * m1.y + ow.z * m2.x + m2.w);
            d = v3(dw.x * m0.x + dw.y * m0.w + dw.z * m1.z, dw.x * m0.y + dw.y * m1.x + dw.z * m1.w, dw.x * m0.z + dw.y * m1.y + dw.z * m2.x);
            setupRayDirection(d, idir, negx, negy, negz, octInv4);
            nodeBase = m3.x; triBase = m3.y; curInst = ii;
            inBlas = true;
            group = make_uint2(0u, 0x80000000u);
            leafMask = 0u;
        }

        // ---- leaf hits go to the warp's queue (slots by a shared-memory atomic: the order within the queue does not matter) ----
        const uint32_t cnt = (uint32_t)__popc(leafMask);
        const uint32_t total = __reduce_add_sync(FULL, cnt);
        if (total) {
            if (qCount + total > SELAS_COOP_QCAP && qCount) flush();
            if (total > SELAS_COOP_QCAP) {
                // more leaf hits in one step than the queue holds (up to 24 per lane): every lane tests its own, as traceKernel does
                if (cnt) {
                    uint32_t dummy = 0;
                    const uint32_t before = res.tri;
                    if (testLeafGroup<ANY, false, CURVES>(bvh, make_uint2(leafBase + (INST ? triBase : 0u), leafMask), o, d, tnear, tmax, res, dummy)) done = true;
                    if (INST && res.tri != before) res.inst = curInst;
                }
            } else {
                if (cnt) {
                    uint32_t w = atomicAdd(tail, cnt);
                    const uint32_t entryBase = lane | ((leafBase + (INST ? triBase : 0u)) << 5);
                    uint32_t m = leafMask;
                    do {
                        const uint32_t bit = 31u - __clz(m);
                        m ^= 1u << bit;
                        if (INST) qinst[w] = curInst;
                        q[w++] = entryBase + (bit << 5);
                    } while (m);
                    outstanding = true;
                }
                qCount += total;
            }
            if (COUNT) nTris += cnt;
        }

        // ---- next group; flush when the queue is full enough, rays are waiting on it, or nobody has a node left ----
        if (active && (group.y & 0xff000000u) == 0u) {
            if (INST && inBlas && sp == spMark) {  // the instance's tree is exhausted: back to world space and to what was left at the top level
                o = ow; d = dw;
                setupRayDirection(d, idir, negx, negy, negz, octInv4);
                inBlas = false;
                nodeBase = 0u; triBase = 0u;
            }
            if (sp > 0) {
                group = stack[--sp];
                if (INST) pendingLeaf = (group.y & 0xff000000u) == 0u;  // only remainders of top-level leaf hits are pushed without inner bits
            }
        }
        const bool hasWork = active && !done && ((group.y & 0xff000000u) != 0u || (INST && pendingLeaf));
        const uint32_t waiting = __ballot_sync(FULL, active && !hasWork && outstanding);
        const bool anyWork = __any_sync(FULL, hasWork);
        if (qCount && (qCount >= SELAS_COOP_FLUSH_MIN || __popc(waiting) >= SELAS_COOP_WAIT_MAX || !anyWork)) flush();

        // ---- rays with nothing left ----
        if (active && !outstanding && (done || ((group.y & 0xff000000u) == 0u && !(INST && pendingLeaf)))) {
            if (MODE == 1) occluded[rayIdx] = res.tri != 0xFFFFFFFFu ? 1 : 0;
            else if (MODE == 2) {
                if (res.tri == 0xFFFFFFFFu) {
                    const float4 pl = __ldg(payload + rayIdx);
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
    }
    if (COUNT && totals) {
        for (int s = 16; s > 0; s >>= 1) {
            sumNodes += __shfl_xor_sync(FULL, sumNodes, s);
            sumTris += __shfl_xor_sync(FULL, sumTris, s);
        }
        if (lane == 0) {
            atomicAdd(totals, sumNodes);
            atomicAdd(totals + 1, sumTris);
        }
    }
}

}  // namespace selas
