// lights.cuh -- quad-light next-event estimation, light pdf, homogeneous medium, camera rays, path RNG.
//
// Replaces (reference paths relative to Source/Core):
//   NextEventEstimation / SampleRectangleLightSolidAngle   Shading/AreaLighting.cpp:113-153, :353-373
//   LightingPdf / QuadLightSolidAnglePdf                   Shading/AreaLighting.cpp:49-66, :376-390
//   InitializeRectangleLightSampler / SampleRectangleLight GeometryLib/RectangulerLightSampler.cpp:17-97
//   Intersection::RayQuad                                  MathLib/GeometryIntersection.cpp:183-236
//   SampleDistance / Transmission / UniformSphere          Shading/IsotropicScattering.cpp:20-72, MathLib/Sampler.cpp:44-54
//   ImageToWorldDirection / JitteredCameraRay              GeometryLib/Camera.cpp:56-74
#pragma once
#include "../../include/selas_cuda.h"
#include "internal.h"
#include "vec.cuh"

namespace selas {

// ---- path RNG (DESIGN.md "Random numbers") ------------------------------------------------------------------
// The reference draws from one std::mt19937 per worker thread (MathLib/Random.cpp:22-55); on the GPU every
// path owns a 32-bit PCG-RXS-M-XS stream keyed by (seed, pixel, sample), carried in the path state.
SELAS_HD uint32_t mix32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}
struct PathRng {
    uint32_t state;
    SELAS_HD void init(uint32_t seed, uint32_t pixel, uint32_t sample) {
        state = mix32(seed ^ mix32(pixel ^ mix32(sample + 0x9e3779b9u)));
    }
    SELAS_HD float next() {
        state = state * 747796405u + 2891336453u;
        uint32_t w = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
        w = (w >> 22u) ^ w;
        return (float)(w >> 8) * (1.0f / 16777216.0f);
    }
};
struct ArrayRng {  // scripted stream for the batch parity kernels
    const float* u;
    uint32_t pos;
    SELAS_HD float next() { return u[pos++]; }
};

// ---- spherical-rectangle sampling (Urena 2013) ------------------------------------------------------------
struct RectSampler {
    V3 o, x, y, z;
    float z0, z0sq, x0, y0, y0sq, x1, y1, y1sq, b0, b1, b0sq, k, S;
};
static SELAS_HDN void initRectSampler(V3 s, V3 eX, V3 eY, V3 o, RectSampler& r) {
    r.o = o;
    float exl = length(eX), eyl = length(eY);
    r.x = eX * (1.0f / exl);
    r.y = eY * (1.0f / eyl);
    r.z = cross(r.x, r.y);
    V3 d = s - o;
    r.z0 = dot(d, r.z);
    if (r.z0 > 0) {
        r.z = -r.z;
        r.z0 *= -1;
    }
    r.z0sq = r.z0 * r.z0;
    r.x0 = dot(d, r.x);
    r.y0 = dot(d, r.y);
    r.x1 = r.x0 + exl;
    r.y1 = r.y0 + eyl;
    r.y0sq = r.y0 * r.y0;
    r.y1sq = r.y1 * r.y1;
    V3 v00 = v3(r.x0, r.y0, r.z0), v01 = v3(r.x0, r.y1, r.z0), v10 = v3(r.x1, r.y0, r.z0), v11 = v3(r.x1, r.y1, r.z0);
    V3 n0 = normalize(cross(v00, v10));
    V3 n1 = normalize(cross(v10, v11));
    V3 n2 = normalize(cross(v11, v01));
    V3 n3 = normalize(cross(v01, v00));
    float g0 = acosf(-dot(n0, n1));
    float g1 = acosf(-dot(n1, n2));
    float g2 = acosf(-dot(n2, n3));
    float g3 = acosf(-dot(n3, n0));
    r.b0 = n0.z;
    r.b1 = n2.z;
    r.b0sq = r.b0 * r.b0;
    r.k = 2 * SELAS_PI - g2 - g3;
    r.S = g0 + g1 - r.k;
}
static SELAS_HDN V3 sampleRect(const RectSampler& r, float u, float v) {
    float au = u * r.S + r.k;
    float fu = (cosf(au) * r.b0 - r.b1) / sinf(au);
    float cu = 1 / sqrtf(fu * fu + r.b0sq) * (fu > 0 ? +1 : -1);
    cu = clampf_(cu, -1.0f, 1.0f);
    float xu = -(cu * r.z0) / sqrtf(1 - cu * cu);
    xu = clampf_(xu, r.x0, r.x1);
    float d = sqrtf(xu * xu + r.z0sq);
    float h0 = r.y0 / sqrtf(d * d + r.y0sq);
    float h1 = r.y1 / sqrtf(d * d + r.y1sq);
    float hv = h0 + v * (h1 - h0), hv2 = hv * hv;
    float yv = (hv2 < 1 - 1e-10f) ? (hv * d) / sqrtf(1 - hv2) : r.y1;
    return (r.o + xu * r.x + yv * r.y + r.z0 * r.z);
}

struct LightSample {  // LightDirectSample, AreaLighting.h:47-54
    uint32_t index;
    V3 direction, radiance;
    float distance, pdfW;
};

SELAS_HD V3 lightF3(const float* p) { return v3(p[0], p[1], p[2]); }

// NextEventEstimation for a light set of `count` quads starting at `lights`.  Draws p0 then (u,v).
// QUIRK AreaLighting.cpp:144: dot(normal, l) is not clamped; :368-370: uint(p0*count) may reach count at
// p0 == 1.0 -- clamped here.
template <typename Rng>
SELAS_HD void nextEventEstimation(Rng& rng, const SelasLight* lights, uint32_t count, V3 position, V3 normal, LightSample& o) {
    o.index = 0;
    o.direction = v3(0.0f);
    o.radiance = v3(0.0f);
    o.distance = 0.0f;
    o.pdfW = 0.0f;
    if (count == 0) return;
    float perLight = 1.0f / count;
    float p0 = rng.next();
    uint32_t idx = (uint32_t)(p0 * count);
    if (idx >= count) idx = count - 1;
    const SelasLight& L = lights[idx];
    V3 eX = lightF3(L.x), eZ = lightF3(L.z);
    V3 s = lightF3(L.position) - 0.5f * eX - 0.5f * eZ;
    RectSampler q;
    initRectSampler(s, eX, eZ, position, q);
    float pdf = 1.0f / q.S;
    float u = rng.next(), v = rng.next();
    V3 lp = sampleRect(q, u, v);
    V3 ul = lp - position;
    float distSq = dot(ul, ul);
    float dist = sqrtf(distSq);
    V3 l = (1.0f / dist) * ul;
    float areaMeasure = dot(lightF3(L.direction), -l) / distSq;
    if (areaMeasure > 0) {
        o.direction = l;
        o.distance = dist;
        o.radiance = lightF3(L.radiance) * areaMeasure * dot(normal, l);
        o.pdfW = pdf * areaMeasure;
    } else {
        o.pdfW = 1.0f;
    }
    o.pdfW *= perLight;
    o.index = idx;
}

static SELAS_HDN bool rayQuad(V3 o, V3 d, V3 v00, V3 v10, V3 v01, V3 v11) {
    const float eps = 0.00001f;
    V3 e01 = v10 - v00, e03 = v01 - v00;
    V3 p = cross(d, e03);
    float det = dot(e01, p);
    if (fabsf(det) < eps) return false;
    V3 t = o - v00;
    float alpha = dot(t, p) / det;
    if (alpha < 0 || alpha > 1) return false;
    V3 q = cross(t, e01);
    float beta = dot(d, q) / det;
    if (beta < 0 || beta > 1) return false;
    float tau = dot(e03, q) / det;
    if (tau < 0) return false;
    if ((alpha + beta) > 1) {
        V3 e23 = v01 - v11, e21 = v10 - v11;
        V3 pp = cross(d, e21);
        float detp = dot(e23, pp);
        if (fabsf(detp) < eps) return false;
        V3 tp = o - v11;
        float alphap = dot(tp, pp) / detp;
        if (alphap < 0) return false;
        V3 qp = cross(tp, e23);
        float betap = dot(d, qp) / detp;
        if (betap < 0.0f) return false;
    }
    return true;
}

SELAS_HD float lightingPdf(const SelasLight* lights, uint32_t count, uint32_t index, V3 position, V3 wi) {
    if (count == 0) return 0.0f;
    float perLight = 1.0f / count;
    const SelasLight& L = lights[index];
    V3 eX = lightF3(L.x), eZ = lightF3(L.z), c = lightF3(L.position);
    V3 v00 = c - 0.5f * eX - 0.5f * eZ;
    V3 v01 = c - 0.5f * eX + 0.5f * eZ;
    V3 v10 = c + 0.5f * eX - 0.5f * eZ;
    V3 v11 = c + 0.5f * eX + 0.5f * eZ;
    if (rayQuad(position, wi, v00, v10, v01, v11)) {
        RectSampler q;
        initRectSampler(v00, eX, eZ, position, q);
        return (1.0f / q.S) * perLight;
    }
    return 0.0f;
}

// ---- homogeneous medium ---------------------------------------------------------------------------------------
struct Medium {
    uint32_t phase;
    V3 extinction;
};
template <typename Rng>
SELAS_HD float sampleDistance(Rng& rng, const Medium& m) {
    if (m.phase == SELAS_MEDIUM_VACUUM) return SELAS_FLOAT_MAX;
    float ps = dot(m.extinction, v3(1.0f));
    float pr = m.extinction.x / ps, pg = m.extinction.y / ps;
    float c;
    float r0 = rng.next();
    if (r0 < pr)
        c = m.extinction.x;
    else if (r0 < pr + pg)
        c = m.extinction.y;
    else
        c = m.extinction.z;
    return -logf(rng.next()) / c;
}
SELAS_HD V3 transmission(const Medium& m, float distance) {
    if (m.phase == SELAS_MEDIUM_VACUUM) return v3(1.0f);
    return v3(expf(-m.extinction.x * distance), expf(-m.extinction.y * distance), expf(-m.extinction.z * distance));
}
template <typename Rng>
SELAS_HD V3 uniformSphere(Rng& rng) {
    float r0 = rng.next(), r1 = rng.next();
    float u = 2.0f * r1 - 1.0f;
    float norm = sqrtf(maxf_(0.0f, 1.0f - u * u));
    float theta = SELAS_TWO_PI * r0;
    return v3(norm * cosf(theta), u, norm * sinf(theta));
}

// ---- camera ---------------------------------------------------------------------------------------------------
SELAS_HD V3 imageToWorldDirection(const SelasCamera& cam, float x, float y) {
    float cx = ((x / (uint32_t)cam.width) * 2.0f - 1.0f);
    float cy = 1.0f - 2.0f * y / (uint32_t)cam.height;
    return normalize(lightF3(cam.cameraZ) - cx * lightF3(cam.cameraX) + cy * lightF3(cam.cameraY));
}

// ---- image based light: SceneLib/ImageBasedLightResource.cpp:81-262 ------------------------------------------------------
SELAS_D void iblSampleDistribution(const float* __restrict__ dist, int count, float random01, uint32_t& index, float& pdf) {  // :81-103
    index = 0xFFFFFFFFu;
    int low = 0, high = count - 1;
    while (low <= high) {
        int mid = (low + high) / 2;
        float d = __ldg(dist + mid);
        if (d >= random01) {
            index = (uint32_t)mid;
            pdf = (mid > 0) ? d - __ldg(dist + mid - 1) : d;
            high = mid - 1;
        } else {
            low = mid + 1;
        }
    }
}
SELAS_D void iblSample(const IblGpu& ibl, float r0, float r1, float& theta, float& phi, uint32_t& x, uint32_t& y, float& pdf) {  // Ibl(), :118-151
    float widthf = (float)ibl.width, heightf = (float)ibl.height;
    float mdf = 0.f, cdf = 0.f;
    iblSampleDistribution(ibl.marginal, (int)ibl.height, r0, y, mdf);
    if (y == 0xFFFFFFFFu) y = ibl.height - 1;
    iblSampleDistribution(ibl.conditional + (size_t)y * ibl.width, (int)ibl.width, r1, x, cdf);
    if (x == 0xFFFFFFFFu) x = ibl.width - 1;
    theta = (y + 0.5f) * SELAS_PI / heightf;
    phi = fmodf(((x + 0.5f) * SELAS_TWO_PI / widthf) + ibl.rotation, SELAS_TWO_PI) - SELAS_PI;
    float invJacobian = (widthf * heightf) / SELAS_TWO_PI;
    float sinTheta = sinf(theta);
    pdf = (sinTheta > 0) ? mdf * cdf * invJacobian / sinTheta : 0.0f;
}
// Texel addressing of SampleIbl / SampleIblMiss / SampleIBlPdf (:160-169).  QUIRK: the reference clamps to [0, width] / [0, height]
// (one past the image); those values cannot be reached for a unit direction, and the reads are kept in range here.
SELAS_D void iblTexel(V3 wi, float rotation, uint32_t width, uint32_t height, int& x, int& y, float& theta) {
    float widthf = (float)(int)width, heightf = (float)(int)height;
    theta = acosf(wi.y);  // NormalizedCartesianToSpherical, MathLib/Projection.cpp:109-113
    float phi = atan2f(wi.z, wi.x);
    phi = fmodf(phi + SELAS_PI + rotation, SELAS_TWO_PI);
    x = (int)(phi * widthf / SELAS_TWO_PI - 0.5f);
    y = (int)(theta * heightf / SELAS_PI - 0.5f);
    x = x < 0 ? 0 : (x > (int)width - 1 ? (int)width - 1 : x);
    y = y < 0 ? 0 : (y > (int)height - 1 ? (int)height - 1 : y);
}
SELAS_D float iblPdf(const IblGpu& ibl, V3 wi) {  // SampleIBlPdf, :221-262
    int x, y;
    float theta;
    iblTexel(wi, ibl.rotation, ibl.width, ibl.height, x, y, theta);
    float widthf = (float)(int)ibl.width, heightf = (float)(int)ibl.height;
    float mdf = (y > 0) ? __ldg(ibl.marginal + y) - __ldg(ibl.marginal + y - 1) : __ldg(ibl.marginal + y);
    const float* row = ibl.conditional + (size_t)y * ibl.width;
    float cdf = (x > 0) ? __ldg(row + x) - __ldg(row + x - 1) : __ldg(row + x);
    float invJacobian = (widthf * heightf) / SELAS_TWO_PI;
    float sinTheta = sinf(theta);
    return (sinTheta > 0) ? mdf * cdf * invJacobian / sinTheta : 0.0f;
}
SELAS_D V3 iblLookup(const IblGpu& ibl, V3 wi) {  // SampleIbl(ibl, wi, pdf) without the pdf, :154-195
    int x, y;
    float theta;
    iblTexel(wi, ibl.rotation, ibl.width, ibl.height, x, y, theta);
    const float* t = ibl.light + 3 * ((size_t)y * ibl.width + x);
    return ibl.exposureScale * v3(__ldg(t), __ldg(t + 1), __ldg(t + 2));
}
SELAS_D V3 iblLookupMiss(const IblGpu& ibl, V3 wi) {  // SampleIblMiss, :198-218 (no exposure scale there)
    int x, y;
    float theta;
    iblTexel(wi, ibl.rotation, ibl.missWidth, ibl.missHeight, x, y, theta);
    const float* t = ibl.miss + 3 * ((size_t)y * ibl.missWidth + x);
    return v3(__ldg(t), __ldg(t + 1), __ldg(t + 2));
}

// ---- correlated multi-jitter (MathLib/CorrelatedMultiJitter.cpp:12-62; Kensler 2013) ---------------------------------
SELAS_HD float cmjFloat(uint32_t i, uint32_t p) {
    i ^= p;
    i ^= i >> 17;
    i ^= i >> 10; i *= 0xb36534e5u;
    i ^= i >> 12;
    i ^= i >> 21; i *= 0x93fc4795u;
    i ^= 0xdf6e307fu;
    i ^= i >> 17; i *= 1u | p >> 18;
    return i * (1.0f / 4294967808.0f);
}
SELAS_HD uint32_t cmjPermute(uint32_t i, uint32_t l, uint32_t p) {
    uint32_t w = l - 1;
    w |= w >> 1; w |= w >> 2; w |= w >> 4; w |= w >> 8; w |= w >> 16;
    do {
        i ^= p; i *= 0xe170893du;
        i ^= p >> 16;
        i ^= (i & w) >> 4;
        i ^= p >> 8; i *= 0x0929eb3fu;
        i ^= p >> 23;
        i ^= (i & w) >> 1; i *= 1u | p >> 27;
        i *= 0x6935fa69u;
        i ^= (i & w) >> 11; i *= 0x74dcb303u;
        i ^= (i & w) >> 2; i *= 0x9e501cc3u;
        i ^= (i & w) >> 2; i *= 0xc860a3dfu;
        i &= w;
        i ^= i >> 5;
    } while (i >= l);
    return (i + p) % l;
}
SELAS_HD void correlatedMultiJitter(int32_t s, int32_t m, int32_t n, int32_t p, float& jx, float& jy) {
    int sx = (int)cmjPermute((uint32_t)(s % m), (uint32_t)m, (uint32_t)p * 0xa511e9b3u);
    int sy = (int)cmjPermute((uint32_t)(s / m), (uint32_t)n, (uint32_t)p * 0x63d83595u);
    float fx = cmjFloat((uint32_t)s, (uint32_t)p * 0xa399d265u);
    float fy = cmjFloat((uint32_t)s, (uint32_t)p * 0x711ad6a5u);
    jx = (s % m + (sy + fx) / n) / m;
    jy = (s / m + (sx + fy) / m) / n;
}

}  // namespace selas
