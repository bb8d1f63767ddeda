// curve.cuh -- ray vs round cubic B-spline curve segment (RTC_GEOMETRY_TYPE_ROUND_BSPLINE_CURVE, the reference's hair / grass geometry:
// ModelResource.cpp:300-333).  Embree's semantics (share/man/man3/RTC_GEOMETRY_TYPE_CURVE.3embree3): four consecutive control vertices
// {x, y, z, radius} starting at the segment's index; the surface is the sweep of a circle of the interpolated radius along the uniform
// cubic B-spline; a hit reports u = parameter along the segment in [0, 1], v = 0 and the unnormalised surface normal as Ng.
//
// The segment is cut into SELAS_CURVE_PARTS parts.  Each part is a truncated cone between two curve points (radii interpolated with
// the same basis) and the joints are closed with spheres, so the pieces form one watertight tube; every piece is a quadratic in t.
// A part is one leaf primitive of the BVH8 with its own box (bvh_build.cuh: gatherCurvesKernel), the four control vertices of the segment
// are shared in a pool.  The ray origin is first moved next to the part (t0 = projection of its start point on the ray): with origins
// hundreds of units away and radii of 1e-3 the quadratic's constant term would otherwise cancel catastrophically in fp32.
// The element that was hit (2 * part + 1 = cone, 2 * joint = sphere) travels with the hit so that the shading stage can rebuild the
// normal exactly instead of storing it.  Ray-cast parity of curves is unpinned like the rest of the ray casting (Embree is absent);
// the tests check against a double-precision intersector over a 16 times finer subdivision.
#pragma once
#include "vec.cuh"

namespace selas {

#ifndef SELAS_CURVE_PARTS
#define SELAS_CURVE_PARTS 8
#endif

struct CurveHit {
    float t, u;
    uint32_t element;  // 1 + (2 * part + 1 for a cone, 2 * joint for a sphere); 0 = none
    V3 ng;             // unnormalised surface normal, in the space of the control points
};

// uniform cubic B-spline basis at s in [0, 1]
SELAS_D void bsplineBasis(float s, float& b0, float& b1, float& b2, float& b3) {
    const float s2 = s * s, s3 = s2 * s, om = 1.0f - s;
    b0 = om * om * om * (1.0f / 6.0f);
    b1 = (3.0f * s3 - 6.0f * s2 + 4.0f) * (1.0f / 6.0f);
    b2 = (-3.0f * s3 + 3.0f * s2 + 3.0f * s + 1.0f) * (1.0f / 6.0f);
    b3 = s3 * (1.0f / 6.0f);
}
SELAS_D float4 curvePoint(const float4 c0, const float4 c1, const float4 c2, const float4 c3, float s) {
    float b0, b1, b2, b3;
    bsplineBasis(s, b0, b1, b2, b3);
    return make_float4(b0 * c0.x + b1 * c1.x + b2 * c2.x + b3 * c3.x, b0 * c0.y + b1 * c1.y + b2 * c2.y + b3 * c3.y,
                       b0 * c0.z + b1 * c1.z + b2 * c2.z + b3 * c3.z, b0 * c0.w + b1 * c1.w + b2 * c2.w + b3 * c3.w);
}

// Normal of element `element` (CurveHit encoding) at point p: cone: q - rho k axis (gradient of |q|^2 - rho^2), sphere: p - centre.
SELAS_D V3 curveElementNormal(const float4 c0, const float4 c1, const float4 c2, const float4 c3, uint32_t element, V3 p) {
    const uint32_t e = element - 1u;
    const uint32_t i = e >> 1;
    const float4 a = curvePoint(c0, c1, c2, c3, (float)i * (1.0f / SELAS_CURVE_PARTS));
    const V3 pa = v3(a.x, a.y, a.z);
    if (!(e & 1u)) return p - pa;
    const float4 b = curvePoint(c0, c1, c2, c3, (float)(i + 1) * (1.0f / SELAS_CURVE_PARTS));
    V3 axis = v3(b.x, b.y, b.z) - pa;
    const float len = length(axis);
    if (!(len > 0.0f)) return p - pa;
    axis = axis * (1.0f / len);
    const float k = (b.w - a.w) / len;
    const V3 m = p - pa;
    const float s = dot(m, axis);
    const V3 q = m - s * axis;
    const float rho = a.w + k * s;
    return q - (rho * k) * axis;
}

// One part of a segment (the BVH's leaf primitive): the cone from curve point `part` to `part + 1`, the joint sphere at its start and,
// for the last part, the one at its end.
static __device__ __noinline__ bool intersectCurve(const float4 c0, const float4 c1, const float4 c2, const float4 c3, uint32_t part, V3 o, V3 d,
                                                   float tnear, float tmax, CurveHit& hit) {
    const float dd = dot(d, d);
    if (!(dd > 0.0f)) return false;
    const float4 a = curvePoint(c0, c1, c2, c3, (float)part * (1.0f / SELAS_CURVE_PARTS));
    const float4 b = curvePoint(c0, c1, c2, c3, (float)(part + 1u) * (1.0f / SELAS_CURVE_PARTS));
    const V3 pa = v3(a.x, a.y, a.z), pb = v3(b.x, b.y, b.z);
    // move the origin next to the part
    const float t0 = dot(pa - o, d) / dd;
    const V3 oo = o + t0 * d;
    const float lo = tnear - t0, hi = tmax - t0;
    float best = hi;
    bool found = false;
    const V3 m = oo - pa;
#pragma unroll 1
    for (int e = 0; e < 2; ++e) {  // joint spheres |c + t d|^2 = r^2: the part's start, and the segment's very end
        if (e == 1 && part != SELAS_CURVE_PARTS - 1) break;
        const V3 c = e == 0 ? m : oo - pb;
        const float r = e == 0 ? a.w : b.w;
        const float B = dot(c, d), C = dot(c, c) - r * r;
        const float disc = B * B - dd * C;
        if (disc >= 0.0f && r > 0.0f) {
            const float sq = sqrtf(disc);
            float t = (-B - sq) / dd;
            if (t < lo) t = (-B + sq) / dd;
            if (t >= lo && t <= best) {
                best = t;
                found = true;
                hit.element = 1u + 2u * (part + (uint32_t)e);
                hit.u = (float)(part + (uint32_t)e) * (1.0f / SELAS_CURVE_PARTS);
                hit.ng = c + t * d;
            }
        }
    }
    V3 axis = pb - pa;
    const float len = length(axis);
    if (len > 0.0f) {
        axis = axis * (1.0f / len);
        const float k = (b.w - a.w) / len;
        const float s0 = dot(m, axis), s1 = dot(d, axis);
        const float r0 = a.w + k * s0;
        // |m + t d|^2 - (s0 + t s1)^2 - (ra + k (s0 + t s1))^2 = 0
        const float A = dd - s1 * s1 - k * k * s1 * s1;
        const float B = dot(m, d) - s0 * s1 - k * s1 * r0;
        const float C = dot(m, m) - s0 * s0 - r0 * r0;
        const float disc = B * B - A * C;
        if (disc >= 0.0f && A != 0.0f) {
            const float sq = sqrtf(disc);
            const float ta = (-B - sq) / A, tb = (-B + sq) / A;
            const float t1 = fminf(ta, tb), t2 = fmaxf(ta, tb);
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const float t = r == 0 ? t1 : t2;
                const float s = s0 + t * s1;
                const float rho = a.w + k * s;
                if (t >= lo && t <= best && s >= 0.0f && s <= len && rho >= 0.0f) {
                    best = t;
                    found = true;
                    hit.element = 2u + 2u * part;
                    hit.u = ((float)part + s / len) * (1.0f / SELAS_CURVE_PARTS);
                    const V3 q = (m + t * d) - s * axis;
                    hit.ng = q - (rho * k) * axis;
                    break;
                }
            }
        }
    }
    if (found) hit.t = best + t0;
    return found;
}

}  // namespace selas
