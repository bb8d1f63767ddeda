// oracle/orc_shading.h -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Scalar CPU restatement of the reference's shading arithmetic for the path-tracing hot path: Disney BSDF
// evaluate/sample, GGX helpers, Fresnel, the dirac-transparent shader, quad-light next-event estimation,
// the homogeneous medium, the camera model and correlated multi-jitter.  Every function names the
// reference lines it follows (paths relative to /root/reference/Source).  It is pinned against the
// reference's own objects (oracle/_ref, built by oracle/Makefile) in tests/test_oracle_vs_ref.py and
// against the committed vectors in tests/golden/.
//
// All arithmetic is fp32 in the reference's operation order, so that on the same CPU the two agree to the
// last bit wherever libm is the only source of difference.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#include "../include/selas_cuda.h"

namespace orc {

struct V3 {
    float x, y, z;
};
static inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
static inline V3 v3(float s) { return V3{s, s, s}; }
static inline V3 v3(const float* p) { return V3{p[0], p[1], p[2]}; }
static inline void store(float* p, V3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
static inline V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
static inline V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
static inline V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
static inline V3 operator*(float s, V3 a) { return V3{a.x * s, a.y * s, a.z * s}; }
static inline V3 operator/(V3 a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }  // FloatStructs: per-component divide
// Reference quirk (MathLib/FloatFuncs.h:147-151): "float / float3" is defined as rhs / dividend, i.e. the
// operands are swapped.  CalculateExtinction (Disney.cpp:342) therefore yields s*scatterDistance, not its
// reciprocal.  Reproduced as is.
static inline V3 operator/(float s, V3 a) { return V3{a.x / s, a.y / s, a.z / s}; }
// MathLib/FloatFuncs.h:213-216, :198-206, :258-261, :401-407
static inline float dot(V3 a, V3 b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }
static inline V3 cross(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static inline float length(V3 a) { return sqrtf(dot(a, a)); }
static inline V3 normalize(V3 a) {
    float inv = 1.f / length(a);
    return V3{a.x * inv, a.y * inv, a.z * inv};
}
static inline float absf(float x) { return fabsf(x); }
static inline float sq(float x) { return x * x; }
static inline float maxf(float a, float b) { return (a > b) ? a : b; }  // SystemLib/MinMax.h:16-19
static inline float minf(float a, float b) { return (a < b) ? a : b; }
static inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }  // MinMax.h:22-29
static inline float saturate(float x) { return x < 0.f ? 0.f : (x > 1.f ? 1.f : x); }                // FloatFuncs.h:389-399
static inline float lerp(float a, float b, float t) { return (1.0f - t) * a + t * b; }              // FloatFuncs.h:369-372
static inline V3 lerp(V3 a, V3 b, float t) { return (1.0f - t) * a + t * b; }
static inline V3 vsqrt(V3 a) { return V3{sqrtf(a.x), sqrtf(a.y), sqrtf(a.z)}; }
static inline float signf(float x) { return x < 0.f ? -1.f : (x > 0.f ? 1.f : 0.f); }  // Trigonometric.cpp:121-132

static const float kPi = 3.141592653589793f;  // Trigonometric.h:14-18
static const float kTwoPi = kPi * 2.0f;
static const float kInvPi = 1.0f / kPi;
static const float kInv4Pi = 1.0f / (4.0f * kPi);
static const float kFloatMax = 3.402823466e+38F;

// row-vector * 3x3 (rows r0,r1,r2): FloatFuncs.cpp:356-364
static inline V3 mul33(V3 v, const float* m) {
    return V3{v.x * m[0] + v.y * m[3] + v.z * m[6], v.x * m[1] + v.y * m[4] + v.z * m[7],
              v.x * m[2] + v.y * m[5] + v.z * m[8]};
}
// row-vector * transpose(3x3): what MatrixMultiply(wi, MatrixTranspose(worldToTangent)) evaluates to
static inline V3 mul33t(V3 v, const float* m) {
    return V3{v.x * m[0] + v.y * m[1] + v.z * m[2], v.x * m[3] + v.y * m[4] + v.z * m[5],
              v.x * m[6] + v.y * m[7] + v.z * m[8]};
}
// FloatFuncs.h:348-367
static inline V3 reflect(V3 n, V3 l) { return 2.0f * dot(n, l) * n - l; }
static inline bool transmit(V3 wm, V3 wi, float n, V3& wo) {
    float c = dot(wi, wm);
    if (c < 0.0f) {
        c = -c;
        wm = -wm;
    }
    float root = 1.0f - n * n * (1.0f - c * c);
    if (root <= 0) return false;
    wo = (n * c - sqrtf(root)) * wm - n * wi;
    return true;
}

// ---- tangent-space trig (y-up): MathLib/Trigonometric.h:51-108 ------------------------------------------
static inline float cosTheta(V3 w) { return w.y; }
static inline float cos2Theta(V3 w) { return w.y * w.y; }
static inline float absCosTheta(V3 w) { return absf(w.y); }
static inline float sin2Theta(V3 w) { return maxf(0.0f, 1.0f - cos2Theta(w)); }
static inline float sinTheta(V3 w) { return sqrtf(sin2Theta(w)); }
static inline float tanTheta(V3 w) { return sinTheta(w) / cosTheta(w); }
static inline float cosPhi(V3 w) {
    float s = sinTheta(w);
    return (s == 0) ? 1.0f : clampf(w.x / s, -1.0f, 1.0f);
}
static inline float sinPhi(V3 w) {
    float s = sinTheta(w);
    return (s == 0) ? 1.0f : clampf(w.z / s, -1.0f, 1.0f);
}

// ---- uniform-number stream -------------------------------------------------------------------------------
struct Stream {
    const float* u;
    uint32_t pos;
    float next() { return u[pos++]; }
};

// ---- Fresnel: Shading/Fresnel.cpp ---------------------------------------------------------------------
static inline float schlickWeight(float u) {  // :33-38
    float m = saturate(1.0f - u);
    float m2 = m * m;
    return m * m2 * m2;
}
static inline float schlickScalar(float r0, float radians) { return lerp(1.0f, schlickWeight(radians), r0); }  // :26-29
static inline V3 schlickV(V3 r0, float radians) {  // :19-23 (powf form)
    float e = powf(1.0f - radians, 5.0f);
    return r0 + (v3(1.0f) - r0) * e;
}
static inline float schlickR0FromRelativeIOR(float eta) { return sq(eta - 1.0f) / sq(eta + 1.0f); }  // :78-82
static inline float fresnelDielectric(float cosThetaI, float ni, float nt) {  // :47-76
    cosThetaI = clampf(cosThetaI, -1.0f, 1.0f);
    if (cosThetaI < 0.0f) {
        float tmp = ni;
        ni = nt;
        nt = tmp;
        cosThetaI = -cosThetaI;
    }
    float sinThetaI = sqrtf(maxf(0.0f, 1.0f - cosThetaI * cosThetaI));
    float sinThetaT = ni / nt * sinThetaI;
    if (sinThetaT >= 1) return 1;
    float cosThetaT = sqrtf(maxf(0.0f, 1.0f - sinThetaT * sinThetaT));
    float rPar = ((nt * cosThetaI) - (ni * cosThetaT)) / ((nt * cosThetaI) + (ni * cosThetaT));
    float rPerp = ((ni * cosThetaI) - (nt * cosThetaT)) / ((ni * cosThetaI) + (nt * cosThetaT));
    return (rPar * rPar + rPerp * rPerp) / 2;
}

// ---- GGX: Shading/Ggx.cpp -------------------------------------------------------------------------------
static inline float smithG1Aniso(V3 w, float ax, float ay) {  // :19-37 (wm unused there)
    float absTan = absf(tanTheta(w));
    if (std::isinf(absTan)) return 0.0f;
    float c = cosPhi(w), s = sinPhi(w);
    float a = sqrtf(c * c * ax * ax + s * s * ay * ay);
    float a2Tan2 = sq(a * absTan);
    float lambda = 0.5f * (-1.0f + sqrtf(1.0f + a2Tan2));
    return 1.0f / (1.0f + lambda);
}
static inline float smithG1Iso(V3 w, float a) {  // :40-46
    float a2 = a * a;
    float c = absCosTheta(w);
    return 2.0f / (1.0f + sqrtf(a2 + (1 - a2) * c * c));
}
static inline float ggxAnisoD(V3 wm, float ax, float ay) {  // :75-84
    float hx2 = sq(wm.x), hy2 = sq(wm.z), c2 = cos2Theta(wm);
    float ax2 = sq(ax), ay2 = sq(ay);
    return 1.0f / (kPi * ax * ay * sq(hx2 / ax2 + hy2 / ay2 + c2));
}
static inline V3 sampleGgxVndfAniso(V3 wo, float ax, float ay, float u1, float u2) {  // :105-126
    V3 v = normalize(v3(wo.x * ax, wo.y, wo.z * ay));
    V3 t1 = (v.y < 0.9999f) ? normalize(cross(v, v3(0, 1, 0))) : v3(1, 0, 0);
    V3 t2 = cross(t1, v);
    float a = 1.0f / (1.0f + v.y);
    float r = sqrtf(u1);
    float phi = (u2 < a) ? (u2 / a) * kPi : kPi + (u2 - a) / (1.0f - a) * kPi;
    float p1 = r * cosf(phi);
    float p2 = r * sinf(phi) * ((u2 < a) ? 1.0f : v.y);
    V3 n = p1 * t1 + p2 * t2 + sqrtf(maxf(0.0f, 1.0f - p1 * p1 - p2 * p2)) * v;
    return normalize(v3(ax * n.x, n.y, ay * n.z));
}
static inline void ggxVndfAnisoPdf(V3 wi, V3 wm, V3 wo, float ax, float ay, float& fwd, float& rev) {  // :141-155
    float D = ggxAnisoD(wm, ax, ay);
    float absNL = absCosTheta(wi), absHL = absf(dot(wm, wi));
    float g1v = smithG1Aniso(wo, ax, ay);
    fwd = g1v * absHL * D / absNL;
    float absNV = absCosTheta(wo), absHV = absf(dot(wm, wo));
    float g1l = smithG1Aniso(wi, ax, ay);
    rev = g1l * absHV * D / absNV;
}

// ---- Disney: Shading/Disney.cpp ---------------------------------------------------------------------------
struct Surf {  // working copy of SelasSurface
    const SelasSurface* s;
    V3 baseColor, transmittance;
};

static inline void lobePdfs(const SelasSurface& s, float& pSpec, float& pDiff, float& pCoat, float& pTrans) {  // :26-44
    float metallicBRDF = s.metallic;
    float specularBSDF = (1.0f - s.metallic) * s.specTrans;
    float dielectricBRDF = (1.0f - s.specTrans) * (1.0f - s.metallic);
    float wSpec = metallicBRDF + dielectricBRDF;
    float wTrans = specularBSDF;
    float wDiff = dielectricBRDF;
    float wCoat = 1.0f * saturate(s.clearcoat);
    float norm = 1.0f / (wSpec + wTrans + wDiff + wCoat);
    pSpec = wSpec * norm;
    pTrans = wTrans * norm;
    pDiff = wDiff * norm;
    pCoat = wCoat * norm;
}
static inline float thinTransmissionRoughness(float ior, float roughness) {  // :47-53
    return saturate((0.65f * ior - 0.35f) * roughness);
}
static inline void anisoParams(float roughness, float anisotropic, float& ax, float& ay) {  // :56-61
    float aspect = sqrtf(1.0f - 0.9f * anisotropic);
    ax = maxf(0.001f, sq(roughness) / aspect);
    ay = maxf(0.001f, sq(roughness) * aspect);
}
static inline V3 tintOf(V3 baseColor) {  // :64-70
    float lum = dot(v3(0.3f, 0.6f, 1.0f), baseColor);
    return (lum > 0.0f) ? baseColor * (1.0f / lum) : v3(1.0f);
}
static inline float gtr1(float absDotHL, float a) {  // :74-82 (log2, as the reference has it)
    if (a >= 1) return kInvPi;
    float a2 = a * a;
    return (a2 - 1.0f) / (kPi * log2f(a2) * (1.0f + (a2 - 1.0f) * absDotHL * absDotHL));
}
static inline float evalClearcoat(float clearcoat, float alpha, V3 wo, V3 wm, V3 wi, float& fPdf, float& rPdf) {  // :85-106
    if (clearcoat <= 0.0f) return 0.0f;
    float absNH = absCosTheta(wm);
    float dotHL = dot(wm, wi);
    float d = gtr1(absNH, lerp(0.1f, 0.001f, alpha));
    float f = schlickScalar(0.04f, dotHL);
    float gl = smithG1Iso(wi, 0.25f);
    float gv = smithG1Iso(wo, 0.25f);
    fPdf = d / (4.0f * absf(dot(wo, wm)));
    rPdf = d / (4.0f * absf(dot(wi, wm)));
    return 0.25f * clearcoat * d * f * gl * gv;
}
static inline V3 evalSheen(const SelasSurface& s, V3 wo, V3 wm, V3 wi) {  // :109-119
    if (s.sheen <= 0.0f) return v3(0.0f);
    float dotHL = absf(dot(wm, wi));
    V3 tint = tintOf(v3(s.baseColor));
    return s.sheen * lerp(v3(1.0f), tint, s.sheenTint) * schlickWeight(dotHL);
}
static inline V3 disneyFresnel(const SelasSurface& s, V3 wo, V3 wm, V3 wi) {  // :122-137
    float dotHV = dot(wm, wo);
    V3 tint = tintOf(v3(s.baseColor));
    V3 R0 = schlickR0FromRelativeIOR(s.relativeIOR) * lerp(v3(1.0f), tint, s.specularTint);
    R0 = lerp(R0, v3(s.baseColor), s.metallic);
    float dielectric = fresnelDielectric(dotHV, 1.0f, s.ior);
    V3 metallic = schlickV(R0, dot(wi, wm));
    return lerp(v3(dielectric), metallic, s.metallic);
}
static inline V3 evalDisneyBRDF(const SelasSurface& s, V3 wo, V3 wm, V3 wi, float& fPdf, float& rPdf) {  // :140-166
    fPdf = 0.0f;
    rPdf = 0.0f;
    float dotNL = cosTheta(wi), dotNV = cosTheta(wo);
    if (dotNL <= 0.0f || dotNV <= 0.0f) return v3(0.0f);
    float ax, ay;
    anisoParams(s.roughness, s.anisotropic, ax, ay);
    float d = ggxAnisoD(wm, ax, ay);
    float gl = smithG1Aniso(wi, ax, ay);
    float gv = smithG1Aniso(wo, ax, ay);
    V3 f = disneyFresnel(s, wo, wm, wi);
    ggxVndfAnisoPdf(wi, wm, wo, ax, ay, fPdf, rPdf);
    fPdf *= (1.0f / (4 * absf(dot(wo, wm))));
    rPdf *= (1.0f / (4 * absf(dot(wi, wm))));
    return d * gl * gv * f / (4.0f * dotNL * dotNV);
}
static inline V3 evalSpecTransmission(const SelasSurface& s, V3 wo, V3 wm, V3 wi, float ax, float ay, bool thin) {  // :214-245
    float relIor = s.relativeIOR;
    float n2 = relIor * relIor;
    float absNL = absCosTheta(wi), absNV = absCosTheta(wo);
    float dotHL = dot(wm, wi), dotHV = dot(wm, wo);
    float absHL = absf(dotHL), absHV = absf(dotHV);
    float d = ggxAnisoD(wm, ax, ay);
    float gl = smithG1Aniso(wi, ax, ay);
    float gv = smithG1Aniso(wo, ax, ay);
    float f = fresnelDielectric(dotHV, 1.0f, s.ior);
    V3 color = thin ? vsqrt(v3(s.baseColor)) : v3(s.baseColor);
    float c = (absHL * absHV) / (absNL * absNV);
    float t = (n2 / sq(dotHL + relIor * dotHV));
    return color * c * t * (1.0f - f) * gl * gv * d;
}
static inline float evalRetroDiffuse(const SelasSurface& s, V3 wo, V3 wi) {  // :248-260
    float dotNL = absCosTheta(wi), dotNV = absCosTheta(wo);
    float roughness = s.roughness * s.roughness;
    float rr = 0.5f + 2.0f * dotNL * dotNL * roughness;
    float fl = schlickWeight(dotNL), fv = schlickWeight(dotNV);
    return rr * (fl + fv + fl * fv * (rr - 1.0f));
}
static inline float evalDisneyDiffuse(const SelasSurface& s, V3 wo, V3 wm, V3 wi, bool thin) {  // :263-290
    float dotNL = absCosTheta(wi), dotNV = absCosTheta(wo);
    float fl = schlickWeight(dotNL), fv = schlickWeight(dotNV);
    float hk = 0.0f;
    if (thin && s.flatness > 0.0f) {
        float roughness = s.roughness * s.roughness;
        float dotHL = dot(wm, wi);
        float fss90 = dotHL * dotHL * roughness;
        float fss = lerp(1.0f, fss90, fl) * lerp(1.0f, fss90, fv);
        hk = 1.25f * (fss * (1.0f / (dotNL + dotNV) - 0.5f) + 0.5f);
    }
    float lambert = 1.0f;
    float retro = evalRetroDiffuse(s, wo, wi);
    float subsurfaceApprox = lerp(lambert, hk, thin ? s.flatness : 0.0f);
    return kInvPi * (retro + subsurfaceApprox * (1.0f - 0.5f * fl) * (1.0f - 0.5f * fv));
}
static inline V3 extinctionOf(V3 apparentColor, float scatterDistance) {  // :337-343
    V3 a = apparentColor;
    V3 s = v3(1.9f) - a + 3.5f * (a - v3(0.8f)) * (a - v3(0.8f));
    return 1.0f / (s * scatterDistance);
}

static inline void failSample(SelasBsdfSample& o) {
    o.forwardPdfW = 0.0f;
    o.reversePdfW = 0.0f;
    store(o.reflectance, v3(0.0f));
    store(o.wi, v3(0.0f));
}

static inline bool sampleDisneyBRDF(Stream& rng, const SelasSurface& s, V3 v, SelasBsdfSample& o) {  // :169-211
    V3 wo = normalize(mul33(v, s.worldToTangent));
    float ax, ay;
    anisoParams(s.roughness, s.anisotropic, ax, ay);
    float r0 = rng.next(), r1 = rng.next();
    V3 wm = sampleGgxVndfAniso(wo, ax, ay, r0, r1);
    V3 wi = normalize(reflect(wm, wo));
    if (cosTheta(wi) <= 0.0f) {
        failSample(o);
        return false;
    }
    V3 F = disneyFresnel(s, wo, wm, wi);
    float G1v = smithG1Aniso(wo, ax, ay);
    o.flags = SELAS_EVENT_SCATTER;
    store(o.reflectance, G1v * F);
    store(o.wi, normalize(mul33t(wi, s.worldToTangent)));
    ggxVndfAnisoPdf(wi, wm, wo, ax, ay, o.forwardPdfW, o.reversePdfW);
    o.forwardPdfW *= (1.0f / (4 * absf(dot(wo, wm))));
    o.reversePdfW *= (1.0f / (4 * absf(dot(wi, wm))));
    return true;
}
static inline bool sampleDisneyClearcoat(Stream& rng, const SelasSurface& s, V3 v, SelasBsdfSample& o) {  // :293-334
    V3 wo = normalize(mul33(v, s.worldToTangent));
    float a = 0.25f, a2 = a * a;
    float r0 = rng.next(), r1 = rng.next();
    float cosT = sqrtf(maxf(0.f, (1.0f - powf(a2, 1.0f - r0)) / (1.0f - a2)));
    float sinT = sqrtf(maxf(0.f, 1.0f - cosT * cosT));
    float phi = kTwoPi * r1;
    V3 wm = v3(sinT * cosf(phi), cosT, sinT * sinf(phi));
    if (dot(wm, wo) < 0.0f) wm = -wm;
    V3 wi = reflect(wm, wo);
    if (dot(wi, wo) < 0.0f) return false;
    float dotNH = cosTheta(wm), dotLH = dot(wm, wi);
    float d = gtr1(absf(dotNH), lerp(0.1f, 0.001f, s.clearcoatGloss));
    float f = schlickScalar(0.04f, dotLH);
    float g = smithG1Iso(wi, 0.25f) * smithG1Iso(wo, 0.25f);
    float fPdf = d / (4.0f * dot(wo, wm));
    store(o.reflectance, v3(0.25f * s.clearcoat * g * f * d) / fPdf);
    store(o.wi, normalize(mul33t(wi, s.worldToTangent)));
    o.forwardPdfW = fPdf;
    o.reversePdfW = d / (4.0f * dot(wi, wm));
    // flags are left as the caller initialised them (reference never writes them here; Scattering.h:36)
    return true;
}
static inline bool sampleDisneySpecTransmission(Stream& rng, const SelasSurface& s, V3 v, bool thin, SelasBsdfSample& o) {  // :346-452
    V3 wo = mul33(v, s.worldToTangent);
    if (cosTheta(wo) == 0.0) {
        failSample(o);
        return false;
    }
    float rscaled = thin ? thinTransmissionRoughness(s.ior, s.roughness) : s.roughness;
    float tax, tay;
    anisoParams(rscaled, s.anisotropic, tax, tay);
    float r0 = rng.next(), r1 = rng.next();
    V3 wm = sampleGgxVndfAniso(wo, tax, tay, r0, r1);
    float dotVH = dot(wo, wm);
    if (wm.y < 0.0f) dotVH = -dotVH;
    float ni = wo.y > 0.0f ? 1.0f : s.ior;
    float nt = wo.y > 0.0f ? s.ior : 1.0f;
    float relativeIOR = ni / nt;
    float F = fresnelDielectric(dotVH, 1.0f, s.ior);
    float G1v = smithG1Aniso(wo, tax, tay);
    float pdf;
    V3 wi;
    if (rng.next() <= F) {
        wi = normalize(reflect(wm, wo));
        o.flags = SELAS_EVENT_SCATTER;
        store(o.reflectance, G1v * v3(s.baseColor));
        float jacobian = (4 * absf(dot(wo, wm)));
        pdf = F / jacobian;
    } else {
        if (thin) {
            wi = reflect(wm, wo);
            wi.y = -wi.y;
            store(o.reflectance, G1v * vsqrt(v3(s.baseColor)));
            o.flags = SELAS_EVENT_SCATTER;
        } else {
            if (transmit(wm, wo, relativeIOR, wi)) {
                o.flags = SELAS_EVENT_TRANSMISSION;
                o.mediumPhase = dotVH > 0.0f ? SELAS_MEDIUM_ISOTROPIC : SELAS_MEDIUM_VACUUM;
                store(o.mediumExtinction, extinctionOf(v3(s.transmittanceColor), s.scatterDistance));
            } else {
                o.flags = SELAS_EVENT_SCATTER;
                wi = reflect(wm, wo);
            }
            store(o.reflectance, G1v * v3(s.baseColor));
        }
        wi = normalize(wi);
        float dotLH = absf(dot(wi, wm));
        float jacobian = dotLH / (sq(dotLH + s.relativeIOR * dotVH));
        pdf = (1.0f - F) / jacobian;
    }
    if (cosTheta(wi) == 0.0f) {
        failSample(o);
        return false;
    }
    if (s.roughness < 0.01f) o.flags |= SELAS_EVENT_DIRAC;
    ggxVndfAnisoPdf(wi, wm, wo, tax, tay, o.forwardPdfW, o.reversePdfW);
    o.forwardPdfW *= pdf;
    o.reversePdfW *= pdf;
    store(o.wi, normalize(mul33t(wi, s.worldToTangent)));
    return true;
}
static inline V3 cosineHemisphere(float r0, float r1) {  // :455-461
    float r = sqrtf(r0);
    float theta = kTwoPi * r1;
    return v3(r * cosf(theta), sqrtf(maxf(0.0f, 1 - r0)), r * sinf(theta));
}
static inline bool sampleDisneyDiffuse(Stream& rng, const SelasSurface& s, V3 v, bool thin, SelasBsdfSample& o) {  // :464-522
    V3 wo = mul33(v, s.worldToTangent);
    float sign = signf(cosTheta(wo));
    float r0 = rng.next(), r1 = rng.next();
    V3 wi = sign * cosineHemisphere(r0, r1);
    V3 wm = normalize(wi + wo);
    float dotNL = cosTheta(wi);
    if (dotNL == 0.0f) {
        failSample(o);
        return false;
    }
    float dotNV = cosTheta(wo);
    float pdf;
    uint32_t eventType = SELAS_EVENT_SCATTER;
    V3 color = v3(s.baseColor);
    float p = rng.next();
    if (p <= s.diffTrans) {
        wi = -wi;
        pdf = s.diffTrans;
        if (thin)
            color = vsqrt(color);
        else {
            eventType = SELAS_EVENT_TRANSMISSION;
            o.mediumPhase = SELAS_MEDIUM_ISOTROPIC;
            store(o.mediumExtinction, extinctionOf(v3(s.transmittanceColor), s.scatterDistance));
        }
    } else {
        pdf = (1.0f - s.diffTrans);
    }
    V3 sheen = evalSheen(s, wo, wm, wi);
    float diffuse = evalDisneyDiffuse(s, wo, wm, wi, thin);
    store(o.reflectance, sheen + color * (diffuse / pdf));
    store(o.wi, normalize(mul33t(wi, s.worldToTangent)));
    o.forwardPdfW = absf(dotNL) * pdf;
    o.reversePdfW = absf(dotNV) * pdf;
    o.flags = eventType;
    return true;
}

static inline V3 evaluateDisney(const SelasSurface& s, V3 v, V3 l, bool thin, float& fwd, float& rev) {  // :525-614
    V3 wo = normalize(mul33(v, s.worldToTangent));
    V3 wi = normalize(mul33(l, s.worldToTangent));
    V3 wm = normalize(wo + wi);
    float dotNV = cosTheta(wo), dotNL = cosTheta(wi);
    V3 refl = v3(0.0f);
    fwd = 0.0f;
    rev = 0.0f;
    float pBRDF, pDiffuse, pClearcoat, pSpecTrans;
    lobePdfs(s, pBRDF, pDiffuse, pClearcoat, pSpecTrans);
    float metallic = s.metallic, specTrans = s.specTrans;
    float ax, ay;
    anisoParams(s.roughness, s.anisotropic, ax, ay);
    float diffuseWeight = (1.0f - metallic) * (1.0f - specTrans);
    float transWeight = (1.0f - metallic) * specTrans;
    bool upper = dotNL > 0.0f && dotNV > 0.0f;
    if (upper && s.clearcoat > 0.0f) {
        float f, r;
        float cc = evalClearcoat(s.clearcoat, s.clearcoatGloss, wo, wm, wi, f, r);
        refl = refl + v3(cc);
        fwd += pClearcoat * f;
        rev += pClearcoat * r;
    }
    if (diffuseWeight > 0.0f) {
        float f = absCosTheta(wi), r = absCosTheta(wo);
        float diffuse = evalDisneyDiffuse(s, wo, wm, wi, thin);
        V3 sheen = evalSheen(s, wo, wm, wi);
        refl = refl + diffuseWeight * (diffuse * v3(s.baseColor) + sheen);
        fwd += pDiffuse * f;
        rev += pDiffuse * r;
    }
    if (transWeight > 0.0f) {
        float rscaled = thin ? thinTransmissionRoughness(s.ior, s.roughness) : s.roughness;
        float tax, tay;
        anisoParams(rscaled, s.anisotropic, tax, tay);
        V3 tr = evalSpecTransmission(s, wo, wm, wi, tax, tay, thin);
        refl = refl + transWeight * tr;
        float f, r;
        ggxVndfAnisoPdf(wi, wm, wo, tax, tay, f, r);
        float dotLH = dot(wm, wi), dotVH = dot(wm, wo);
        fwd += pSpecTrans * f / (sq(dotLH + s.relativeIOR * dotVH));
        rev += pSpecTrans * r / (sq(dotVH + s.relativeIOR * dotLH));
    }
    if (upper) {
        float f, r;
        V3 spec = evalDisneyBRDF(s, wo, wm, wi, f, r);
        refl = refl + spec;
        fwd += pBRDF * f / (4 * absf(dot(wo, wm)));
        rev += pBRDF * r / (4 * absf(dot(wi, wm)));
    }
    return refl * absf(dotNL);
}

static inline bool sampleDisney(Stream& rng, const SelasSurface& s, V3 v, bool thin, SelasBsdfSample& o) {  // :617-659
    float pSpec, pDiff, pCoat, pTrans;
    lobePdfs(s, pSpec, pDiff, pCoat, pTrans);
    bool ok = false;
    float pLobe = 0.0f;
    float p = rng.next();
    if (p <= pSpec) {
        ok = sampleDisneyBRDF(rng, s, v, o);
        pLobe = pSpec;
    } else if (p > pSpec && p <= (pSpec + pCoat)) {
        ok = sampleDisneyClearcoat(rng, s, v, o);
        pLobe = pCoat;
    } else if (p > pSpec + pCoat && p <= (pSpec + pCoat + pDiff)) {
        ok = sampleDisneyDiffuse(rng, s, v, thin, o);
        pLobe = pDiff;
    } else if (pTrans >= 0.0f) {
        ok = sampleDisneySpecTransmission(rng, s, v, thin, o);
        pLobe = pTrans;
    } else {
        store(o.reflectance, v3(1000000.0f, 0.0f, 0.0f));
        o.forwardPdfW = 0.000000001f;
        o.reversePdfW = 0.000000001f;
    }
    if (pLobe > 0.0f) {
        store(o.reflectance, v3(o.reflectance) * (1.0f / pLobe));
        o.forwardPdfW *= pLobe;
        o.reversePdfW *= pLobe;
    }
    return ok;
}

// Shading/SurfaceParameters.cpp:391-394 -- despite its name this is the shading normal (column 1)
static inline V3 geometricNormal(const SelasSurface& s) {
    return v3(s.worldToTangent[1], s.worldToTangent[4], s.worldToTangent[7]);
}
// Shading/DiracTransparent.cpp:20-68
static inline bool sampleDiracTransparent(Stream& rng, const SelasSurface& s, V3 v, SelasBsdfSample& o) {
    V3 normal = geometricNormal(s);
    float dotVH = dot(v, normal);
    float ni = dotVH > 0.0f ? 1.0f : s.ior;
    float nt = dotVH > 0.0f ? s.ior : 1.0f;
    float n = ni / nt;
    float F = fresnelDielectric(dotVH, ni, nt);
    float pdf;
    V3 wi;
    if (rng.next() < F) {
        wi = normalize(reflect(normal, v));
        o.flags = SELAS_EVENT_SCATTER;
        pdf = F;
    } else {
        if (transmit(normal, v, n, wi)) {
            o.flags = SELAS_EVENT_TRANSMISSION;
        } else {
            wi = reflect(normal, v);
            o.flags = SELAS_EVENT_SCATTER;
        }
        wi = normalize(wi);
        pdf = (1.0f - F);
    }
    o.flags |= SELAS_EVENT_DIRAC;
    store(o.reflectance, v3(s.baseColor) / pdf);
    store(o.wi, wi);
    o.forwardPdfW = pdf;
    o.reversePdfW = pdf;
    return true;
}

// Shading/SurfaceScattering.cpp:41-61
static inline V3 evaluateBsdf(const SelasSurface& s, V3 v, V3 l, float& fwd, float& rev) {
    if (s.shader == SELAS_SHADER_DISNEY_THIN) return evaluateDisney(s, v, l, true, fwd, rev);
    if (s.shader == SELAS_SHADER_DISNEY_SOLID) return evaluateDisney(s, v, l, false, fwd, rev);
    return v3(0.0f);  // eDiracTransparent: pdfs untouched by the reference; callers pre-zero them
}
// Shading/SurfaceScattering.cpp:18-38.  The BsdfSample defaults (Scattering.h:34-43) are applied first;
// flags has no default there, this restatement defines it as eScatterEvent.
static inline bool sampleBsdf(Stream& rng, const SelasSurface& s, V3 v, SelasBsdfSample& o) {
    o.flags = SELAS_EVENT_SCATTER;
    o.mediumPhase = SELAS_MEDIUM_VACUUM;
    store(o.mediumExtinction, v3(0.0f));
    store(o.reflectance, v3(0.0f));
    store(o.wi, v3(0.0f));
    o.forwardPdfW = 0.0f;
    o.reversePdfW = 0.0f;
    uint32_t start = rng.pos;
    bool ok = false;
    if (s.shader == SELAS_SHADER_DISNEY_THIN)
        ok = sampleDisney(rng, s, v, true, o);
    else if (s.shader == SELAS_SHADER_DISNEY_SOLID)
        ok = sampleDisney(rng, s, v, false, o);
    else if (s.shader == SELAS_SHADER_DIRAC_TRANSPARENT)
        ok = sampleDiracTransparent(rng, s, v, o);
    o.valid = ok ? 1u : 0u;
    o.randomsUsed = rng.pos - start;
    return ok;
}

// Shading/SurfaceParameters.cpp:403-408
static inline V3 offsetRayOrigin(const SelasSurface& s, V3 direction, float biasScale) {
    V3 n = geometricNormal(s);
    float dirOffset = dot(direction, n) < 0.0f ? -1.0f : 1.0f;
    V3 offset = dirOffset * s.error * biasScale * n;
    return v3(s.position) + offset;
}

// ---- quad lights ------------------------------------------------------------------------------------------
struct RectSampler {  // GeometryLib/RectangulerLightSampler.h:12-21
    V3 o, x, y, z;
    float z0, z0sq, x0, y0, y0sq, x1, y1, y1sq, b0, b1, b0sq, k, S;
};
static inline void initRectSampler(V3 s, V3 eX, V3 eY, V3 o, RectSampler& r) {  // RectangulerLightSampler.cpp:17-73
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
    r.k = 2 * kPi - g2 - g3;
    r.S = g0 + g1 - r.k;
}
static inline V3 sampleRect(const RectSampler& r, float u, float v) {  // RectangulerLightSampler.cpp:75-96
    float au = u * r.S + r.k;
    float fu = (cosf(au) * r.b0 - r.b1) / sinf(au);
    float cu = 1 / sqrtf(fu * fu + r.b0sq) * (fu > 0 ? +1 : -1);
    cu = clampf(cu, -1.0f, 1.0f);
    float xu = -(cu * r.z0) / sqrtf(1 - cu * cu);
    xu = clampf(xu, r.x0, r.x1);
    float d = sqrtf(xu * xu + r.z0sq);
    float h0 = r.y0 / sqrtf(d * d + r.y0sq);
    float h1 = r.y1 / sqrtf(d * d + r.y1sq);
    float hv = h0 + v * (h1 - h0), hv2 = hv * hv;
    float yv = (hv2 < 1 - 1e-10f) ? (hv * d) / sqrtf(1 - hv2) : r.y1;
    return (r.o + xu * r.x + yv * r.y + r.z0 * r.z);
}
// Shading/AreaLighting.cpp:113-153
static inline void sampleRectLightSolidAngle(Stream& rng, V3 position, V3 normal, const SelasLight& L, SelasLightSample& o) {
    V3 eX = v3(L.x), eZ = v3(L.z);
    V3 s = v3(L.position) - 0.5f * eX - 0.5f * eZ;
    RectSampler q;
    initRectSampler(s, eX, eZ, position, q);
    float pdf = 1.0f / q.S;
    float u = rng.next(), v = rng.next();
    V3 lp = sampleRect(q, u, v);
    V3 ul = lp - position;
    float distSq = dot(ul, ul);
    float dist = sqrtf(distSq);
    V3 l = (1.0f / dist) * ul;
    float areaMeasure = dot(v3(L.direction), -l) / distSq;
    if (areaMeasure > 0) {
        store(o.direction, l);
        o.distance = dist;
        store(o.radiance, v3(L.radiance) * areaMeasure * dot(normal, l));
        o.pdfW = pdf * areaMeasure;
    } else {
        store(o.direction, v3(0.0f));
        o.distance = 0.0f;
        store(o.radiance, v3(0.0f));
        o.pdfW = 1.0f;
    }
}
// Shading/AreaLighting.cpp:353-373.  uint(p0*count) can reach count when p0 == 1.0 (appendix B #20); the
// restatement clamps, as the CUDA path does.  Fields other than radiance are untouched on the early-outs in
// the reference; here they are zeroed so comparisons are defined.
static inline void nextEventEstimation(Stream& rng, const SelasLight* lights, uint32_t count, V3 position, V3 normal,
                                       SelasLightSample& o) {
    memset(&o, 0, sizeof(o));
    if (count == 0) return;
    float perLight = 1.0f / count;
    float p0 = rng.next();
    uint32_t idx = (uint32_t)(p0 * count);
    if (idx >= count) idx = count - 1;
    sampleRectLightSolidAngle(rng, position, normal, lights[idx], o);
    o.pdfW *= perLight;
    o.index = idx;
}
// MathLib/GeometryIntersection.cpp:183-236
static inline bool rayQuad(V3 o, V3 d, V3 v00, V3 v10, V3 v01, V3 v11) {
    const float eps = 0.00001f;
    V3 e01 = v10 - v00, e03 = v01 - v00;
    V3 p = cross(d, e03);
    float det = dot(e01, p);
    if (absf(det) < eps) return false;
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
        if (absf(detp) < eps) return false;
        V3 tp = o - v11;
        float alphap = dot(tp, pp) / detp;
        if (alphap < 0) return false;
        V3 qp = cross(tp, e23);
        float betap = dot(d, qp) / detp;
        if (betap < 0.0f) return false;
    }
    return true;
}
// Shading/AreaLighting.cpp:49-66 and :376-390
static inline float lightingPdf(const SelasLight* lights, uint32_t count, uint32_t index, V3 position, V3 wi) {
    if (count == 0) return 0.0f;
    float perLight = 1.0f / count;
    const SelasLight& L = lights[index];
    V3 eX = v3(L.x), eZ = v3(L.z), c = v3(L.position);
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

// ---- homogeneous medium: Shading/IsotropicScattering.cpp:20-72, VolumetricScattering.cpp:14-60 ------------
struct Medium {
    uint32_t phase;
    V3 extinction;
};
static inline float sampleDistance(Stream& rng, const Medium& m) {
    if (m.phase == SELAS_MEDIUM_VACUUM) return kFloatMax;
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
static inline V3 transmission(const Medium& m, float distance) {
    if (m.phase == SELAS_MEDIUM_VACUUM) return v3(1.0f);
    return v3(expf(-m.extinction.x * distance), expf(-m.extinction.y * distance), expf(-m.extinction.z * distance));
}
static inline V3 uniformSphere(Stream& rng) {  // MathLib/Sampler.cpp:44-54
    float r0 = rng.next(), r1 = rng.next();
    float u = 2.0f * r1 - 1.0f;
    float norm = sqrtf(maxf(0.0f, 1.0f - u * u));
    float theta = kTwoPi * r0;
    return v3(norm * cosf(theta), u, norm * sinf(theta));
}

// ---- image based light: SceneLib/ImageBasedLightResource.cpp:81-262 ------------------------------------------------
struct Ibl {  // host copy of SelasIblDesc
    uint32_t width = 0, height = 0, missWidth = 0, missHeight = 0;
    const float *marginal = nullptr, *conditional = nullptr, *light = nullptr, *miss = nullptr;
    float rotation = 0.f, exposureScale = 1.f;
};
static inline void sampleDistribution(const float* dist, int64_t count, float random01, uint32_t& index, float& pdf) {  // :81-103
    index = 0xFFFFFFFFu;
    int64_t low = 0, high = count - 1;
    while (low <= high) {
        int64_t mid = (low + high) / 2;
        if (dist[mid] >= random01) {
            index = (uint32_t)mid;
            pdf = (mid > 0) ? dist[mid] - dist[mid - 1] : dist[mid];
            high = mid - 1;
        } else {
            low = mid + 1;
        }
    }
}
static inline void iblSample(const Ibl& ibl, float r0, float r1, float& theta, float& phi, uint32_t& x, uint32_t& y, float& pdf) {  // :118-151
    float widthf = (float)ibl.width, heightf = (float)ibl.height;
    float mdf = 0.f, cdf = 0.f;
    sampleDistribution(ibl.marginal, ibl.height, r0, y, mdf);
    if (y == 0xFFFFFFFFu) y = ibl.height - 1;  // r0 above the last CDF entry cannot happen (it is forced to 1); guard anyway
    sampleDistribution(ibl.conditional + (size_t)y * ibl.width, ibl.width, r1, x, cdf);
    if (x == 0xFFFFFFFFu) x = ibl.width - 1;
    theta = (y + 0.5f) * kPi / heightf;
    phi = fmodf(((x + 0.5f) * kTwoPi / widthf) + ibl.rotation, kTwoPi) - kPi;
    float invJacobian = (widthf * heightf) / kTwoPi;
    float sinTheta = sinf(theta);
    pdf = (sinTheta > 0) ? mdf * cdf * invJacobian / sinTheta : 0.0f;
}
// texel addressing shared by SampleIbl / SampleIblMiss / SampleIBlPdf (:160-169).  QUIRK (appendix B #23): Clamp(.., 0, width)
// lets x reach width (the first texel of the next row) and y reach height; the linear index is clamped to the last texel here,
// as the CUDA path does, where the reference would read past the image.
static inline void iblTexel(V3 wi, float rotation, uint32_t width, uint32_t height, int32_t& x, int32_t& y, float& theta) {
    float widthf = (float)(int32_t)width, heightf = (float)(int32_t)height;
    theta = acosf(wi.y);                 // NormalizedCartesianToSpherical, MathLib/Projection.cpp:109-113
    float phi = atan2f(wi.z, wi.x);
    phi = fmodf(phi + kPi + rotation, kTwoPi);
    x = (int32_t)(phi * widthf / kTwoPi - 0.5f);
    y = (int32_t)(theta * heightf / kPi - 0.5f);
    x = x < 0 ? 0 : (x > (int32_t)width ? (int32_t)width : x);
    y = y < 0 ? 0 : (y > (int32_t)height ? (int32_t)height : y);
}
static inline size_t iblIndex(int32_t x, int32_t y, uint32_t width, uint32_t height) {
    size_t i = (size_t)y * width + (size_t)x, last = (size_t)width * height - 1;
    return i > last ? last : i;
}
static inline float iblPdf(const Ibl& ibl, V3 wi) {  // SampleIBlPdf, :221-262
    int32_t x, y;
    float theta;
    iblTexel(wi, ibl.rotation, ibl.width, ibl.height, x, y, theta);
    float widthf = (float)(int32_t)ibl.width, heightf = (float)(int32_t)ibl.height;
    if (y > (int32_t)ibl.height - 1) y = (int32_t)ibl.height - 1;  // unreachable for theta in [0, pi]; keeps the reads in range
    if (x > (int32_t)ibl.width - 1) x = (int32_t)ibl.width - 1;
    float mdf = (y > 0) ? ibl.marginal[y] - ibl.marginal[y - 1] : ibl.marginal[y];
    size_t i = iblIndex(x, y, ibl.width, ibl.height);
    float cdf = (x > 0) ? ibl.conditional[i] - ibl.conditional[i - 1] : ibl.conditional[i];
    float invJacobian = (widthf * heightf) / kTwoPi;
    float sinTheta = sinf(theta);
    return (sinTheta > 0) ? mdf * cdf * invJacobian / sinTheta : 0.0f;
}
static inline V3 iblLookup(const Ibl& ibl, V3 wi, float& pdf) {  // SampleIbl(ibl, wi, pdf), :154-195
    pdf = iblPdf(ibl, wi);
    int32_t x, y;
    float theta;
    iblTexel(wi, ibl.rotation, ibl.width, ibl.height, x, y, theta);
    return ibl.exposureScale * v3(ibl.light + 3 * iblIndex(x, y, ibl.width, ibl.height));
}
static inline V3 iblLookupMiss(const Ibl& ibl, V3 wi, float& pdf) {  // SampleIblMiss, :198-218 (no exposure scale there)
    pdf = iblPdf(ibl, wi);
    int32_t x, y;
    float theta;
    iblTexel(wi, ibl.rotation, ibl.missWidth, ibl.missHeight, x, y, theta);
    return v3(ibl.miss + 3 * iblIndex(x, y, ibl.missWidth, ibl.missHeight));
}

// ---- camera: GeometryLib/Camera.cpp:56-60, :89-115 -------------------------------------------------------
static inline void initCamera(const SelasCameraSettings& st, uint32_t width, uint32_t height, SelasCamera& cam) {
    float widthf = (float)width, heightf = (float)height;
    float aspect = widthf / heightf;
    float hfov = st.fovDegrees * (kPi / 180.f);
    float hLen = tanf(hfov * 0.5f);
    float vfov = 2 * atanf(hLen / aspect);
    float vLen = tanf(vfov * 0.5f);
    V3 fwd = normalize(v3(st.lookAt) - v3(st.position));
    V3 right = normalize(cross(v3(st.up), fwd));
    store(cam.cameraX, hLen * right);
    store(cam.cameraY, vLen * cross(fwd, right));
    store(cam.cameraZ, fwd);
    store(cam.forward, normalize(v3(st.lookAt) - v3(st.position)));
    cam.viewportWidth = widthf;
    cam.viewportHeight = heightf;
    store(cam.position, v3(st.position));
    cam.znear = st.znear;
    cam.zfar = st.zfar;
    cam.virtualImagePlaneDistance = widthf / (2.0f * tanf(hfov));
    cam.width = width;
    cam.height = height;
    cam.aspect = aspect;
}
static inline V3 imageToWorldDirection(const SelasCamera& cam, float x, float y) {
    float cx = ((x / cam.width) * 2.0f - 1.0f);
    float cy = 1.0f - 2.0f * y / cam.height;
    return normalize(v3(cam.cameraZ) - cx * v3(cam.cameraX) + cy * v3(cam.cameraY));
}

// ---- correlated multi-jitter: MathLib/CorrelatedMultiJitter.cpp:12-62 --------------------------------------
static inline float cmjFloat(uint32_t i, uint32_t p) {
    i ^= p;
    i ^= i >> 17;
    i ^= i >> 10;
    i *= 0xb36534e5;
    i ^= i >> 12;
    i ^= i >> 21;
    i *= 0x93fc4795;
    i ^= 0xdf6e307f;
    i ^= i >> 17;
    i *= 1 | p >> 18;
    return i * (1.0f / 4294967808.0f);
}
static inline uint32_t cmjPermute(uint32_t i, uint32_t l, uint32_t p) {
    uint32_t w = l - 1;
    w |= w >> 1;
    w |= w >> 2;
    w |= w >> 4;
    w |= w >> 8;
    w |= w >> 16;
    do {
        i ^= p;
        i *= 0xe170893d;
        i ^= p >> 16;
        i ^= (i & w) >> 4;
        i ^= p >> 8;
        i *= 0x0929eb3f;
        i ^= p >> 23;
        i ^= (i & w) >> 1;
        i *= 1 | p >> 27;
        i *= 0x6935fa69;
        i ^= (i & w) >> 11;
        i *= 0x74dcb303;
        i ^= (i & w) >> 2;
        i *= 0x9e501cc3;
        i ^= (i & w) >> 2;
        i *= 0xc860a3df;
        i &= w;
        i ^= i >> 5;
    } while (i >= l);
    return (i + p) % l;
}
static inline void cmj(int32_t s, int32_t m, int32_t n, int32_t p, float out[2]) {
    int sx = cmjPermute(s % m, m, p * 0xa511e9b3);
    int sy = cmjPermute(s / m, n, p * 0x63d83595);
    float jx = cmjFloat(s, p * 0xa399d265);
    float jy = cmjFloat(s, p * 0x711ad6a5);
    out[0] = (s % m + (sy + jx) / n) / m;
    out[1] = (s / m + (sx + jy) / m) / n;
}


// ---- VCM building blocks ------------------------------------------------------------------------------------------
// MathLib/Projection.cpp:132-146
static inline V3 sphericalToCartesian(float theta, float phi) {
    float sintheta = sinf(theta), costheta = cosf(theta), sinphi = sinf(phi), cosphi = cosf(phi);
    return v3(sintheta * cosphi, costheta, sintheta * sinphi);
}
// GeometryLib/CoordinateSystem.h:14-21
static inline void makeOrthogonalCoordinateSystem(V3 v1, V3& v2, V3& v3out) {
    if (absf(v1.x) > absf(v1.y))
        v2 = v3(-v1.z, 0, v1.x) * (1.0f / sqrtf(v1.x * v1.x + v1.z * v1.z));
    else
        v2 = v3(0, v1.z, -v1.y) * (1.0f / sqrtf(v1.y * v1.y + v1.z * v1.z));
    v3out = cross(v1, v2);
}
// GeometryLib/Disc.cpp:13-56 (Shirley-Chiu concentric map)
static inline void sampleConcentricDisc(float r0, float r1, float& dx, float& dy) {
    float phi, r;
    float a = 2.0f * r0 - 1.0f, b = 2.0f * r1 - 1.0f;
    if (a > -b) {
        if (a > b) { r = a; phi = (kPi / 4.0f) * (b / a); }
        else { r = b; phi = (kPi / 4.0f) * (2.0f - (a / b)); }
    } else {
        if (a < b) { r = -a; phi = (kPi / 4.0f) * (4.0f + (b / a)); }
        else {
            r = -b;
            if (absf(b) > 0.0f) phi = (kPi / 4.0f) * (6.0f - (a / b));
            else phi = 0;
        }
    }
    dx = r * cosf(phi);
    dy = r * sinf(phi);
}
struct LightEmission {  // LightEmissionSample, Shading/AreaLighting.h:33-45
    V3 position, direction, radiance;
    float emissionPdfW, directionPdfA, cosThetaLight;
};
// EmitIblLightSample, Shading/AreaLighting.cpp:246-290 (four uniforms r0..r3)
static inline void emitIblLightSample(Stream& rng, const Ibl& ibl, const float sphere[4], LightEmission& o) {
    float r0 = rng.next(), r1 = rng.next(), r2 = rng.next(), r3 = rng.next();
    uint32_t x, y;
    float dirPhi, dirTheta;
    iblSample(ibl, r0, r1, dirTheta, dirPhi, x, y, o.directionPdfA);
    V3 toIbl = sphericalToCartesian(dirTheta, dirPhi);
    V3 radiance = ibl.exposureScale * v3(ibl.light + 3 * ((size_t)y * ibl.width + x));  // SampleIbl(ibl, x, y), ImageBasedLightResource.cpp:264-267
    V3 dX, dZ;
    makeOrthogonalCoordinateSystem(toIbl, dX, dZ);
    float R = sphere[3];
    V3 center = v3(sphere);
    float sx, sy;
    sampleConcentricDisc(r2, r3, sx, sy);
    float discPdf = kInvPi;
    o.position = center + R * (toIbl + sx * dX + sy * dZ);
    float pdfPosition = discPdf * (1.0f / (R * R));
    float pdfDirection = o.directionPdfA;
    o.direction = -toIbl;
    o.radiance = radiance;
    o.emissionPdfW = pdfPosition * pdfDirection;
    o.cosThetaLight = 1.0f;
}
// DirectIblLightSample, Shading/AreaLighting.cpp:293-314 (two uniforms)
static inline void directIblLightSample(Stream& rng, const Ibl& ibl, SelasLightSample& o) {
    float r0 = rng.next(), r1 = rng.next();
    uint32_t x, y;
    float dirPhi, dirTheta;
    iblSample(ibl, r0, r1, dirTheta, dirPhi, x, y, o.pdfW);
    V3 toIbl = sphericalToCartesian(dirTheta, dirPhi);
    o.index = 0;
    o.distance = 1e36f;
    store(o.direction, toIbl);
    store(o.radiance, ibl.exposureScale * v3(ibl.light + 3 * ((size_t)y * ibl.width + x)));
}
// IblCalculateRadiance, Shading/AreaLighting.cpp:317-331
static inline V3 iblCalculateRadiance(const Ibl& ibl, const float sphere[4], V3 direction, float& directPdfA, float& emissionPdfW) {
    float iblPdfA;
    V3 radiance = iblLookup(ibl, direction, iblPdfA);
    float R = sphere[3];
    float pdfPosition = kInvPi * (1.0f / (R * R));
    directPdfA = iblPdfA;
    emissionPdfW = pdfPosition * iblPdfA;
    return radiance;
}
// ContinuationProbability, Shading/SurfaceParameters.cpp:420-424
static inline float continuationProbability(const SelasSurface& s) {
    return saturate(maxf(maxf(s.baseColor[0], s.baseColor[1]), s.baseColor[2]));
}
// OffsetRayOrigin with the signed bias distance, Shading/SurfaceParameters.cpp:411-417
static inline V3 offsetRayOriginBias(const SelasSurface& s, V3 direction, float biasScale, float& signedBiasDistance) {
    V3 n = geometricNormal(s);
    float dirOffset = dot(direction, n) < 0.0f ? -1.0f : 1.0f;
    signedBiasDistance = dirOffset * s.error * biasScale;
    V3 offset = signedBiasDistance * n;
    return v3(s.position) + offset;
}
// VCMCommon::SearchRadius / CalculateIterationConstants, Applications/Selas/Source/VCMCommon.cpp:91-112
struct VcmConstants {
    uint32_t vmCount, vcCount;
    float vmSearchRadius, vmSearchRadiusSqr, vmNormalization, vmWeight, vcWeight;
};
static inline VcmConstants vcmIterationConstants(uint32_t vmCount, uint32_t vcCount, float baseRadius, float radiusAlpha, float iterationIndex) {
    float r = baseRadius / powf(iterationIndex, 0.5f * (1.0f - radiusAlpha));
    VcmConstants c;
    c.vmCount = vmCount;
    c.vcCount = vcCount;
    c.vmSearchRadius = r;
    c.vmSearchRadiusSqr = r * r;
    c.vmNormalization = 1.0f / (kPi * c.vmSearchRadiusSqr * vmCount);
    c.vmWeight = kPi * c.vmSearchRadiusSqr * vmCount / vcCount;
    c.vcWeight = vcCount / (kPi * c.vmSearchRadiusSqr * vmCount);
    return c;
}
// WorldToImage: the reference's VCM calls it (VCM.cpp:199) but the function no longer exists in GeometryLib/Camera.cpp; it is
// restated as the inverse of ImageToWorldDirection (Camera.cpp:56-60): dir ~ Z - cx X + cy Y with X, Y, Z mutually orthogonal
// (InitializeRayCastCamera, Camera.cpp:89-115), x = (cx + 1) w / 2, y = (1 - cy) h / 2, truncated toward -inf.
static inline void worldToImage(const SelasCamera& cam, V3 p, int32_t& ix, int32_t& iy) {
    V3 to = p - v3(cam.position);
    V3 X = v3(cam.cameraX), Y = v3(cam.cameraY), Z = v3(cam.cameraZ);
    float z = dot(to, Z) / dot(Z, Z);
    float cx = -(dot(to, X) / dot(X, X)) / z;
    float cy = (dot(to, Y) / dot(Y, Y)) / z;
    float x = (cx + 1.0f) * 0.5f * (float)(uint32_t)cam.width;
    float y = (1.0f - cy) * 0.5f * (float)(uint32_t)cam.height;
    ix = (int32_t)floorf(x);
    iy = (int32_t)floorf(y);
}


// ---- non-Ptex base colour textures ------------------------------------------------------------------------------------
struct Texture {  // TextureResourceData level 0 (TextureLib/TextureResource.h:15-39)
    const float* texels = nullptr;
    uint32_t width = 0, height = 0, format = 0;
};
// TextureFiltering::Sample, WrapMode::Repeat (TextureFiltering.h:35-52): `s % w` with int32 s and uint32 w is an unsigned modulo
static inline size_t textureTexel(const Texture& t, int32_t s, int32_t tt) {
    uint32_t ss = (uint32_t)s % t.width, ts = (uint32_t)tt % t.height;
    return (size_t)ts * t.width + ss;
}
static inline float srgbToLinearPrecise(float x) {  // MathLib/ColorSpace.cpp:16-23
    if (x <= 0.04045f) return x / 12.92f;
    return powf((x + 0.055f) / (1.0f + 0.055f), 2.4f);
}
// SampleTextureFloat3 (Shading/SurfaceParameters.cpp:97-118) over TextureFiltering::Triangle at level 0 (TextureFiltering.h:77-104)
static inline V3 sampleTextureFloat3(const Texture& tex, float u, float v, bool sRGB) {
    if (tex.format == SELAS_TEXTURE_FLOAT2) return v3(0.0f);
    float s = u * tex.width - 0.5f, t = v * tex.height - 0.5f;
    int32_t s0 = (int32_t)floorf(s), t0 = (int32_t)floorf(t);
    float ds = s - s0, dt = t - t0;
    const float w00 = (1 - ds) * (1 - dt), w01 = (1 - ds) * dt, w10 = ds * (1 - dt), w11 = ds * dt;
    const uint32_t comps = tex.format + 1u;
    const size_t i00 = textureTexel(tex, s0, t0) * comps, i01 = textureTexel(tex, s0, t0 + 1) * comps,
                 i10 = textureTexel(tex, s0 + 1, t0) * comps, i11 = textureTexel(tex, s0 + 1, t0 + 1) * comps;
    float r[3] = {0, 0, 0};
    const uint32_t used = comps == 1 ? 1u : 3u;
    for (uint32_t c = 0; c < used; ++c) {
        float x = w00 * tex.texels[i00 + c] + w01 * tex.texels[i01 + c] + w10 * tex.texels[i10 + c] + w11 * tex.texels[i11 + c];
        r[c] = sRGB ? srgbToLinearPrecise(x) : x;
    }
    return comps == 1 ? v3(r[0]) : v3(r[0], r[1], r[2]);
}

}  // namespace orc
