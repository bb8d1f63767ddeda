// oracle/orc_curves.h -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Double-precision checker for round cubic B-spline curve geometry (RTC_GEOMETRY_TYPE_ROUND_BSPLINE_CURVE, the reference's hair / grass
// geometry, ModelResource.cpp:300-333).  The reference's intersector is Embree's, which is not available: parity of curves
// is UNPINNED like the rest of the ray casting.  This file restates the documented semantics
// (share/man/man3/RTC_GEOMETRY_TYPE_CURVE.3embree3): four consecutive control vertices {x, y, z, r} make a uniform cubic B-spline segment,
// the surface is the sweep of a circle of the interpolated radius along it, a hit reports u = parameter along the segment, v = 0 and the
// unnormalised surface normal.  The sweep is evaluated as 128 truncated cones with spherical joints per segment, in double (the CUDA
// path uses 8 parts in fp32): the two differ by the finer subdivision, which is what the parity tests bound.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace orc {

constexpr int kCurveParts = 128;

struct CurveSeg {
    double c[4][4];    // control vertices in world space (radius scaled by the instance's uniform scale)
    double lo[3], hi[3];
    uint32_t primID, geomID, instID;
};
struct CurveHitD {
    double t, u;
    double ng[3];
};

static inline void curvePointD(const CurveSeg& s, double x, double out[4]) {
    const double x2 = x * x, x3 = x2 * x, om = 1.0 - x;
    const double b0 = om * om * om / 6.0, b1 = (3.0 * x3 - 6.0 * x2 + 4.0) / 6.0, b2 = (-3.0 * x3 + 3.0 * x2 + 3.0 * x + 1.0) / 6.0, b3 = x3 / 6.0;
    for (int k = 0; k < 4; ++k) out[k] = b0 * s.c[0][k] + b1 * s.c[1][k] + b2 * s.c[2][k] + b3 * s.c[3][k];
}

static inline void curveBounds(CurveSeg& s) {
    for (int k = 0; k < 3; ++k) { s.lo[k] = 1e300; s.hi[k] = -1e300; }
    for (int i = 0; i <= kCurveParts; ++i) {
        double p[4];
        curvePointD(s, (double)i / kCurveParts, p);
        for (int k = 0; k < 3; ++k) {
            s.lo[k] = std::fmin(s.lo[k], p[k] - std::fabs(p[3]));
            s.hi[k] = std::fmax(s.hi[k], p[k] + std::fabs(p[3]));
        }
    }
}

static inline bool curveBoxHit(const CurveSeg& s, const double o[3], const double d[3], double tnear, double tfar) {
    double t0 = tnear, t1 = tfar;
    for (int k = 0; k < 3; ++k) {
        if (d[k] == 0.0) {
            if (o[k] < s.lo[k] || o[k] > s.hi[k]) return false;
            continue;
        }
        double a = (s.lo[k] - o[k]) / d[k], b = (s.hi[k] - o[k]) / d[k];
        if (a > b) std::swap(a, b);
        t0 = std::fmax(t0, a);
        t1 = std::fmin(t1, b);
    }
    return t0 <= t1;
}

static inline bool intersectCurveD(const CurveSeg& s, const double o[3], const double d[3], double tnear, double tmax, CurveHitD& hit) {
    const double dd = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
    if (!(dd > 0.0)) return false;
    double best = tmax;
    bool found = false;
    double a[4];
    curvePointD(s, 0.0, a);
    for (int i = 0; i <= kCurveParts; ++i) {
        const double m[3] = {o[0] - a[0], o[1] - a[1], o[2] - a[2]};
        const double md = m[0] * d[0] + m[1] * d[1] + m[2] * d[2], mm = m[0] * m[0] + m[1] * m[1] + m[2] * m[2];
        {  // joint sphere
            const double C = mm - a[3] * a[3], disc = md * md - dd * C;
            if (disc >= 0.0 && a[3] > 0.0) {
                const double sq = std::sqrt(disc);
                double t = (-md - sq) / dd;
                if (t < tnear) t = (-md + sq) / dd;
                if (t >= tnear && t <= best) {
                    best = t;
                    found = true;
                    hit.u = (double)i / kCurveParts;
                    for (int k = 0; k < 3; ++k) hit.ng[k] = m[k] + t * d[k];
                }
            }
        }
        if (i == kCurveParts) break;
        double b[4];
        curvePointD(s, (double)(i + 1) / kCurveParts, b);
        double ax[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
        const double len = std::sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
        if (len > 0.0) {
            for (int k = 0; k < 3; ++k) ax[k] /= len;
            const double kk = (b[3] - a[3]) / len;
            const double s0 = m[0] * ax[0] + m[1] * ax[1] + m[2] * ax[2], s1 = d[0] * ax[0] + d[1] * ax[1] + d[2] * ax[2];
            const double r0 = a[3] + kk * s0;
            const double A = dd - s1 * s1 - kk * kk * s1 * s1, B = md - s0 * s1 - kk * s1 * r0, C = mm - s0 * s0 - r0 * r0;
            const double disc = B * B - A * C;
            if (disc >= 0.0 && A != 0.0) {
                const double sq = std::sqrt(disc);
                const double ta = (-B - sq) / A, tb = (-B + sq) / A;
                const double ts[2] = {std::fmin(ta, tb), std::fmax(ta, tb)};
                for (int r = 0; r < 2; ++r) {
                    const double t = ts[r], sp = s0 + t * s1, rho = a[3] + kk * sp;
                    if (t >= tnear && t <= best && sp >= 0.0 && sp <= len && rho >= 0.0) {
                        best = t;
                        found = true;
                        hit.u = ((double)i + sp / len) / kCurveParts;
                        for (int k = 0; k < 3; ++k) hit.ng[k] = (m[k] + t * d[k]) - sp * ax[k] - rho * kk * ax[k];
                        break;
                    }
                }
            }
        }
        for (int k = 0; k < 4; ++k) a[k] = b[k];
    }
    if (found) hit.t = best;
    return found;
}

// closest (or any) hit over a list of segments: box rejection, then the exact test
static inline bool traceCurves(const std::vector<CurveSeg>& segs, const float of[3], const float df[3], double tnear, double tfar, bool anyHit,
                               uint32_t& seg, CurveHitD& out) {
    const double o[3] = {of[0], of[1], of[2]}, d[3] = {df[0], df[1], df[2]};
    bool found = false;
    double best = tfar;
    for (uint32_t i = 0; i < segs.size(); ++i) {
        if (!curveBoxHit(segs[i], o, d, tnear, best)) continue;
        CurveHitD h;
        if (intersectCurveD(segs[i], o, d, tnear, best, h)) {
            best = h.t;
            out = h;
            seg = i;
            found = true;
            if (anyHit) return true;
        }
    }
    return found;
}

}  // namespace orc
