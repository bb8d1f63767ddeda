// bvh8.cuh -- compressed 8-wide BVH: node / triangle records and the closest-hit / any-hit traversal.
//
// Stands in for rtcIntersect1 / rtcOccluded1 (call sites PathTracer.cpp:88,:116; packet form
// DeferredPathTracer.cpp:191,:267).  Embree semantics kept: closest hit with t in [tnear,tfar], no back-face
// culling, (u,v) = weights of v1,v2, quads as (v0,v1,v3)+(v2,v3,v1) with u,v flipped on the second half,
// Ng = cross(v1-v0, v2-v0) unnormalised (share/man/man3/rtcIntersect1.3embree3:24,:53,
// RTC_GEOMETRY_TYPE_QUAD.3embree3:38-41).
//
// Node: 80 bytes = 5 x 16-byte vectors (compressed wide BVH after Ylitie, Karras, Laine 2017):
//   n0 = { origin.x, origin.y, origin.z, (ex | ey<<8 | ez<<16 | imask<<24) }
//   n1 = { childBase, triBase, meta[0..3], meta[4..7] }
//   n2 = { qlo.x[0..3], qlo.x[4..7], qhi.x[0..3], qhi.x[4..7] }   n3 = same for y   n4 = same for z
//   child box = origin + q * 2^(e-127) ; meta: 0 = empty, inner = 001|(24+slot), leaf = unary count|triOffset
// Triangle: 48 bytes = 3 x float4 = { v0.xyz, primID }, { e1.xyz, geomID }, { e2.xyz, instID | second<<31 }.
// A node's leaf triangles are contiguous from triBase (<= 24 per node, <= 3 per child).
//
// Algorithmic bytes per ray: 32 in + 32 out + 80 per visited node + 48 per tested triangle.
#pragma once
#include <cuda_fp16.h>

#include "../../include/selas_cuda.h"
#include "curve.cuh"
#include "vec.cuh"

namespace selas {

struct InstanceTrace;  // internal.h

struct Bvh8 {
    const float4* nodes;  // 5 per node
    const float4* tris;   // 3 per triangle
    uint32_t nodeCount, triCount;
    // two-level form: the top-level tree's leaves are instance references (triangle records whose a.w is the flat instance index) and
    // every mesh has its own object-space tree elsewhere in the same arrays (internal.h: InstanceTrace); null for the flattened form
    const InstanceTrace* instances;
    // the first topCount nodes (whole levels of the tree, breadth first: the collapse numbers nodes level by level) are staged in
    // shared memory by every CTA of the trace kernels; 0 = none
    uint32_t topCount;
    // public entry points of a SELAS_SCENE_INSTANCE_IDS scene: per flat instance the {instID0, instID1, geomID} a hit reports (the
    // reference's two instancing levels); null inside the wavefront, whose queues keep {flat instance, mesh index}
    const uint4* instanceIds;
    // curve pool: four float4 control vertices {x, y, z, radius} per curve segment of the tree (world space for the flattened form,
    // object space under the two-level form); a leaf record with bit 30 of its third id word set names its entry in a.x
    const float4* curves;
};
constexpr uint32_t kLeafCurveFlag = 1u << 30;

#ifndef SELAS_STACK_SIZE
#define SELAS_STACK_SIZE 48
#endif

SELAS_D uint32_t extractByte(uint32_t x, uint32_t i) { return (x >> (i * 8)) & 0xff; }
SELAS_D uint32_t signExtendS8x4(uint32_t x) {
    uint32_t r;
    asm("prmt.b32 %0, %1, 0x0, 0x0000BA98;" : "=r"(r) : "r"(x));
    return r;
}
SELAS_D float fmaxMax(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }
SELAS_D float fminMin(float a, float b, float c) { return fminf(fminf(a, b), c); }

struct TraceResult {
    float t, u, v;
    uint32_t tri;   // index into tris, 0xFFFFFFFF = none
    uint32_t inst;  // two-level form: flat instance index of the hit (the triangle record's own instID is 0 there)
    // curve hits (kernels built with CURVES): element code of curve.cuh and the unnormalised surface normal; element == 0: a triangle
    uint32_t element;
    V3 ng;
};

// Quantised plane byte -> float without the conversion (XU) pipe.  The first version used (float)byte = I2F.U8, 48 per
// node, and ncu showed the XU pipe as the busiest unit of the kernel (65 % of peak, profiles/r1a).  Here two bytes at a
// time are dropped into the mantissas of a half2 holding {1024, 1024} (one PRMT): 0x6400 | b == 1024 + b exactly; the
// half -> float step is an HADD2.F32 on the FMA pipe, and the -1024 is folded into the plane offset (o - 1024 s), whose
// rounding error is 2^-14 of a quantisation step, far inside the builder's 1e-3-step slack.
#ifndef SELAS_NODE_BYTE_PRMT
#define SELAS_NODE_BYTE_PRMT 0
#endif
#if SELAS_NODE_BYTE_PRMT
// Round-2 experiment, OFF: one PRMT per byte and no conversion at all.  The byte is dropped into mantissa bits 8..15 of a float holding 2^23
// ({0x00, b, 0x00, 0x4B}): the value is 2^23 + 256 b exactly, i.e. the plane index biased by 2^15 in units of 1/256 of a step.  The
// slab test folds the 1/256 into the step and the bias into the plane offset (o - 2^23 s / 256); rounding that offset costs 2^-9 of a
// quantisation step, inside the builder's 4e-3-step slack (bvh_build.cuh).  Per node: 48 PRMT instead of 24 PRMT + 48 HADD2.F32,
// 16 instructions fewer in the kernel (640 -> 624 SASS) -- and 1.8 % SLOWER on configs[2] (closest hit 8.41 against 8.26 ms per step,
// profiles/r2_ab_node_prmt.jsonl): PRMT issues on the ALU pipe, which LOP3 / SEL / FMNMX / FSETP / SHF already fill, while HADD2.F32
// rides the FMA pipe next to the FFMAs.  The half2 path below stays the product.
constexpr float kNodeByteBias = 8388608.0f, kNodeByteScale = 1.0f / 256.0f;
SELAS_D void bytesToBiasedFloats(uint32_t w, float& f0, float& f1, float& f2, float& f3) {
    f0 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7404));
    f1 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7414));
    f2 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7424));
    f3 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7434));
}
#else
constexpr float kNodeByteBias = 1024.0f, kNodeByteScale = 1.0f;
SELAS_D void bytesToBiasedFloats(uint32_t w, float& f0, float& f1, float& f2, float& f3) {
    uint32_t lo = __byte_perm(w, 0x64646464u, 0x4140);  // {b0, 0x64, b1, 0x64}
    uint32_t hi = __byte_perm(w, 0x64646464u, 0x4342);  // {b2, 0x64, b3, 0x64}
    __half2 hl = *reinterpret_cast<__half2*>(&lo), hh = *reinterpret_cast<__half2*>(&hi);
    f0 = __low2float(hl);
    f1 = __high2float(hl);
    f2 = __low2float(hh);
    f3 = __high2float(hh);
}

#endif

// Ray-vs-8-children slab test on the quantised boxes; returns the hit mask in traversal-bit order.
SELAS_D uint32_t intersectNode(const float4 n0, const float4 n1, const float4 n2, const float4 n3, const float4 n4, V3 o,
                               V3 idir, bool negx, bool negy, bool negz, uint32_t octInv4, float tmin, float tmax) {
    uint32_t eimask = __float_as_uint(n0.w);
    float sx = __uint_as_float(extractByte(eimask, 0) << 23) * (idir.x * kNodeByteScale);
    float sy = __uint_as_float(extractByte(eimask, 1) << 23) * (idir.y * kNodeByteScale);
    float sz = __uint_as_float(extractByte(eimask, 2) << 23) * (idir.z * kNodeByteScale);
    // plane t = (origin + q*scale - o) * idir = q*s + ox, with q delivered as bias + q (in units of kNodeByteScale steps)
    float ox = fmaf(-kNodeByteBias, sx, (n0.x - o.x) * idir.x), oy = fmaf(-kNodeByteBias, sy, (n0.y - o.y) * idir.y),
          oz = fmaf(-kNodeByteBias, sz, (n0.z - o.z) * idir.z);
    uint32_t hitMask = 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        uint32_t meta4 = __float_as_uint(i == 0 ? n1.z : n1.w);
        uint32_t isInner4 = (meta4 & (meta4 << 1)) & 0x10101010;
        uint32_t innerMask4 = signExtendS8x4(isInner4 << 3);
        uint32_t bitIndex4 = (meta4 ^ (octInv4 & innerMask4)) & 0x1f1f1f1f;
        uint32_t childBits4 = (meta4 >> 5) & 0x07070707;
        uint32_t qlox = __float_as_uint(i == 0 ? n2.x : n2.y), qhix = __float_as_uint(i == 0 ? n2.z : n2.w);
        uint32_t qloy = __float_as_uint(i == 0 ? n3.x : n3.y), qhiy = __float_as_uint(i == 0 ? n3.z : n3.w);
        uint32_t qloz = __float_as_uint(i == 0 ? n4.x : n4.y), qhiz = __float_as_uint(i == 0 ? n4.z : n4.w);
        uint32_t xmin = negx ? qhix : qlox, xmax = negx ? qlox : qhix;
        uint32_t ymin = negy ? qhiy : qloy, ymax = negy ? qloy : qhiy;
        uint32_t zmin = negz ? qhiz : qloz, zmax = negz ? qloz : qhiz;
        float x0[4], x1[4], y0[4], y1[4], z0[4], z1[4];
        bytesToBiasedFloats(xmin, x0[0], x0[1], x0[2], x0[3]);
        bytesToBiasedFloats(xmax, x1[0], x1[1], x1[2], x1[3]);
        bytesToBiasedFloats(ymin, y0[0], y0[1], y0[2], y0[3]);
        bytesToBiasedFloats(ymax, y1[0], y1[1], y1[2], y1[3]);
        bytesToBiasedFloats(zmin, z0[0], z0[1], z0[2], z0[3]);
        bytesToBiasedFloats(zmax, z1[0], z1[1], z1[2], z1[3]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float tx0 = fmaf(x0[j], sx, ox), tx1 = fmaf(x1[j], sx, ox);
            float ty0 = fmaf(y0[j], sy, oy), ty1 = fmaf(y1[j], sy, oy);
            float tz0 = fmaf(z0[j], sz, oz), tz1 = fmaf(z1[j], sz, oz);
            float t0 = fmaxMax(tx0, ty0, fmaxf(tz0, tmin));
            float t1 = fminMin(tx1, ty1, fminf(tz1, tmax));
            if (t0 <= t1) {
                uint32_t childBits = extractByte(childBits4, j);
                uint32_t bitIndex = extractByte(bitIndex4, j);
                hitMask |= childBits << bitIndex;
            }
        }
    }
    return hitMask;
}

// Which inner child of a group comes next.  The slots are octant-ordered, so the highest pending bit is the child a ray of this octant
// meets first: closest-hit rays take it (front to back: tmax shrinks early), and so do the generic any-hit calls (selas_cuda_occluded).
// The integrators' shadow rays (MODE 2: next-event estimation, VCM connections) take the LOWEST bit instead -- back to front, i.e. from
// the light's end.  Such a ray leaves a surface on the side the light is on (the BSDF is zero otherwise and no ray is cast), so whatever
// blocks it is rarely at its origin and usually further along; walking from the far end finds the blocker sooner and ends the ray:
// configs[2] 12.73 -> 12.06 nodes per shadow ray, shadow kernel 3.83 -> 3.58 ms per step, VCM 430 -> 459 M sub-paths/s
// (profiles/r2_ab_anyhit_order*.jsonl).  For shadow rays cast from ALL hit points regardless of facing (the configs[1] side benchmark,
// half of whose rays start into the body they left) the near end is the better one (4.44 against 4.01 G rays/s), hence the split by mode.
// SELAS_ANYHIT_BACK_TO_FRONT=0 restores front to back everywhere.
#ifndef SELAS_ANYHIT_BACK_TO_FRONT
#define SELAS_ANYHIT_BACK_TO_FRONT 1
#endif
template <bool BACK_TO_FRONT>
SELAS_D uint32_t nextChildBit(uint32_t hitsImask) {
    if (BACK_TO_FRONT && SELAS_ANYHIT_BACK_TO_FRONT) return (uint32_t)__ffs((int)(hitsImask & 0xff000000u)) - 1u;
    return 31u - __clz(hitsImask);
}

// Moller-Trumbore on the {v0,e1,e2} record.  Accepts t in [tnear, tmax].
SELAS_D bool intersectTri(const float4 a, const float4 b, const float4 c, V3 o, V3 d, float tnear, float tmax, float& t,
                          float& u, float& v) {
    V3 v0 = v3(a.x, a.y, a.z), e1 = v3(b.x, b.y, b.z), e2 = v3(c.x, c.y, c.z);
    V3 p = cross(d, e2);
    float det = dot(e1, p);
    if (det == 0.0f) return false;
    float inv = 1.0f / det;
    V3 s = o - v0;
    float uu = dot(s, p) * inv;
    if (uu < 0.0f || uu > 1.0f) return false;
    V3 q = cross(s, e1);
    float vv = dot(d, q) * inv;
    if (vv < 0.0f || uu + vv > 1.0f) return false;
    float tt = dot(e2, q) * inv;
    if (!(tt >= tnear && tt <= tmax)) return false;
    t = tt;
    u = uu;
    v = vv;
    return true;
}

// One ray through the BVH8.  ANY: stop at the first hit.  COUNT: tally visited nodes / tested triangles.
template <bool ANY, bool COUNT>
SELAS_D void traceRay(const Bvh8& bvh, V3 o, V3 d, float tnear, float tfar, TraceResult& res, uint32_t* counters) {
    res.t = tfar;
    res.u = res.v = 0.0f;
    res.tri = 0xFFFFFFFFu;
    if (bvh.nodeCount == 0) return;
    uint2 stack[SELAS_STACK_SIZE];
    int sp = 0;
    // reciprocal direction; a zero component becomes a huge finite value so 0*inf never appears
    const float tiny = 1e-30f;
    V3 idir = v3(1.0f / (fabsf(d.x) > tiny ? d.x : copysignf(tiny, d.x)), 1.0f / (fabsf(d.y) > tiny ? d.y : copysignf(tiny, d.y)),
                 1.0f / (fabsf(d.z) > tiny ? d.z : copysignf(tiny, d.z)));
    bool negx = d.x < 0.0f, negy = d.y < 0.0f, negz = d.z < 0.0f;
    uint32_t oct = (negx ? 4u : 0u) | (negy ? 2u : 0u) | (negz ? 1u : 0u);
    uint32_t octInv4 = (7u - oct) * 0x01010101u;
    float tmax = tfar;
    uint32_t nNodes = 0, nTris = 0;

    uint2 group = make_uint2(0u, 0x80000000u);  // root: child base 0, one pending inner child in bit 31
    for (;;) {
        uint2 triGroup;
        if (group.y & 0xff000000u) {
            uint32_t hitsImask = group.y;
            uint32_t childBit = nextChildBit<false>(hitsImask);
            group.y &= ~(1u << childBit);
            if (group.y & 0xff000000u) {
                if (sp < SELAS_STACK_SIZE) stack[sp++] = group;
            }
            uint32_t slot = (childBit - 24u) ^ (octInv4 & 0xffu);
            uint32_t rel = __popc(hitsImask & ~(0xffffffffu << slot));
            uint32_t nodeIdx = group.x + rel;
            const float4* np = bvh.nodes + (size_t)nodeIdx * 5;
            float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
            if (COUNT) ++nNodes;
            uint32_t hitMask = intersectNode(n0, n1, n2, n3, n4, o, idir, negx, negy, negz, octInv4, tnear, tmax);
            uint32_t imask = extractByte(__float_as_uint(n0.w), 3);
            group.x = __float_as_uint(n1.x);
            triGroup.x = __float_as_uint(n1.y);
            group.y = (hitMask & 0xff000000u) | imask;
            triGroup.y = hitMask & 0x00ffffffu;
        } else {
            triGroup = group;
            group = make_uint2(0u, 0u);
        }
        while (triGroup.y != 0u) {
            uint32_t bit = 31u - __clz(triGroup.y);
            triGroup.y &= ~(1u << bit);
            uint32_t ti = triGroup.x + bit;
            const float4* tp = bvh.tris + (size_t)ti * 3;
            float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
            if (COUNT) ++nTris;
            float t, u, v;
            if (intersectTri(a, b, c, o, d, tnear, tmax, t, u, v)) {
                tmax = t;
                res.t = t;
                res.u = u;
                res.v = v;
                res.tri = ti;
                if (ANY) {
                    if (COUNT) { counters[0] = nNodes; counters[1] = nTris; }
                    return;
                }
            }
        }
        if ((group.y & 0xff000000u) == 0u) {
            if (sp == 0) break;
            group = stack[--sp];
        }
    }
    if (COUNT) { counters[0] = nNodes; counters[1] = nTris; }
}

// Fill the 32-byte hit record (and optional Ng) from a trace result.  The ids travel as integers: bit-casting a
// small id to float and masking it lets the compiler emit abs.ftz.f32, which flushes the (denormal) pattern to 0.
template <bool INST = false>
SELAS_D void writeHit(const Bvh8& bvh, const TraceResult& r, float tfar, SelasHit* out, float4* ng) {
    uint4 h0, h1;
    if (r.tri == 0xFFFFFFFFu) {
        h0 = make_uint4(__float_as_uint(tfar), 0u, 0u, SELAS_INVALID_ID);
        h1 = make_uint4(SELAS_INVALID_ID, SELAS_INVALID_ID, SELAS_INVALID_ID, 0u);
        if (ng) *ng = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        const uint4* tp = reinterpret_cast<const uint4*>(bvh.tris + (size_t)r.tri * 3);
        uint4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
        bool second = (c.w >> 31) != 0;
        const bool curve = (c.w & kLeafCurveFlag) != 0;
        float u = second ? 1.0f - r.u : r.u, v = second ? 1.0f - r.v : r.v;
        h0 = make_uint4(__float_as_uint(r.t), __float_as_uint(u), __float_as_uint(v), a.w);
        const uint32_t flat = INST ? r.inst : (c.w & 0x3fffffffu);
        // round curves report v = 0 (RTC_GEOMETRY_TYPE_CURVE.3embree3); the last word carries the element that was hit, so that the
        // shading stage rebuilds the normal exactly (curve.cuh); it is 0 for triangles
        h1 = make_uint4(b.w, flat, 0u, curve ? r.element : 0u);
        if (curve) h0.z = 0u;
        if (bvh.instanceIds) {  // SceneInstanceIntersectFunction: instID[0] = subscene instance, instID[1] = the subscene hit's instID[0]
            const uint4 ids = __ldg(bvh.instanceIds + flat);
            h1 = make_uint4(ids.z, ids.x, ids.y, h1.w);
        }
        if (ng) {
            V3 n = curve ? r.ng
                         : cross(v3(__uint_as_float(b.x), __uint_as_float(b.y), __uint_as_float(b.z)),
                                 v3(__uint_as_float(c.x), __uint_as_float(c.y), __uint_as_float(c.z)));
            *ng = make_float4(n.x, n.y, n.z, 0.f);
        }
    }
    uint4* o4 = reinterpret_cast<uint4*>(out);
    o4[0] = h0;
    o4[1] = h1;
}

}  // namespace selas
