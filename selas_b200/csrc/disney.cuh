// disney.cuh -- device-side Disney BSDF (evaluate / sample / pdf), the dirac-transparent shader and the
// SampleBsdfFunction / EvaluateBsdf dispatch, for the wavefront shade kernel and the batch parity kernels.
//
// Replaces (reference paths relative to Source/Core/Shading):
//   EvaluateBsdf / SampleBsdfFunction   SurfaceScattering.cpp:18-61
//   EvaluateDisney / SampleDisney       Disney.cpp:525-659 and the lobes at :26-522
//   GGX helpers                         Ggx.cpp:19-155
//   Fresnel                             Fresnel.cpp:19-82
//   SampleDiracTransparent              DiracTransparent.cpp:20-68
// Design: one DisneyFrame per shading point caches everything that depends only on (surface, wo) -- lobe
// probabilities, anisotropic alphas, tint -- so evaluate (for NEE) and sample (for the bounce) share it
// inside the fused shade kernel.  Values are computed in the reference's operation order; the shading
// translation unit is compiled with -fmad=false -ftz=true so +,*,/ and sqrt round exactly like the CPU
// reference (FTZ/DAZ is on there too, main.cpp:84-85).  Reference quirks are kept and marked QUIRK.
#pragma once
#include "../../include/selas_cuda.h"
#include "vec.cuh"

namespace selas {

struct Surface {  // register-resident copy of SurfaceParameters (SurfaceParameters.h:18-49)
    V3 position;
    float w2t[9];
    float error;
    V3 baseColor, transmittance;
    float sheen, sheenTint, clearcoat, clearcoatGloss, metallic, specTrans, diffTrans, flatness, anisotropic,
        relativeIOR, specularTint, roughness, scatterDistance, ior;
    uint32_t shader, materialFlags, lightSetIndex;
};

SELAS_HD void loadSurface(const SelasSurface& s, Surface& o) {
    o.position = v3(s.position);
    for (int i = 0; i < 9; ++i) o.w2t[i] = s.worldToTangent[i];
    o.error = s.error;
    o.baseColor = v3(s.baseColor);
    o.transmittance = v3(s.transmittanceColor);
    o.sheen = s.sheen; o.sheenTint = s.sheenTint; o.clearcoat = s.clearcoat; o.clearcoatGloss = s.clearcoatGloss;
    o.metallic = s.metallic; o.specTrans = s.specTrans; o.diffTrans = s.diffTrans; o.flatness = s.flatness;
    o.anisotropic = s.anisotropic; o.relativeIOR = s.relativeIOR; o.specularTint = s.specularTint;
    o.roughness = s.roughness; o.scatterDistance = s.scatterDistance; o.ior = s.ior;
    o.shader = s.shader; o.materialFlags = s.materialFlags; o.lightSetIndex = s.lightSetIndex;
}

struct BsdfResult {  // BsdfSample (Scattering.h:34-43)
    uint32_t flags;
    uint32_t mediumPhase;
    V3 mediumExtinction;
    V3 reflectance;
    V3 wi;
    float forwardPdfW, reversePdfW;
};

// ---- Fresnel (Fresnel.cpp) ---------------------------------------------------------------------------------
SELAS_HD float schlickWeight(float u) {  // :33-38
    float m = saturate_(1.0f - u);
    float m2 = m * m;
    return m * m2 * m2;
}
// QUIRK Fresnel.cpp:26-29: lerp(1, weight, r0) -- returns 1 at r0 = 0
SELAS_HD float schlickScalar(float r0, float c) { return lerpf(1.0f, schlickWeight(c), r0); }
SELAS_HD V3 schlickColor(V3 r0, float c) {  // :19-23, powf form
    // Math::Powf(1 - c, 5) of the C library is the correctly rounded fifth power; the product in double rounded once gives the same float
    // (53-bit factors: the error before the final rounding is below 2^-50 relative) for 3 DMUL instead of a powf (2.7 % of the shade kernel)
    const double b = (double)(1.0f - c);
    const double b2 = b * b;
    float e = (float)(b2 * b2 * b);
    return r0 + (v3(1.0f) - r0) * e;
}
SELAS_HD float schlickR0(float eta) { return sq(eta - 1.0f) / sq(eta + 1.0f); }  // :78-82
static SELAS_HDN float fresnelDielectric(float cosI, float ni, float nt) {               // :47-76
    cosI = clampf_(cosI, -1.0f, 1.0f);
    if (cosI < 0.0f) {
        float t = ni;
        ni = nt;
        nt = t;
        cosI = -cosI;
    }
    float sinI = sqrtf(maxf_(0.0f, 1.0f - cosI * cosI));
    float sinT = ni / nt * sinI;
    if (sinT >= 1) return 1;
    float cosT = sqrtf(maxf_(0.0f, 1.0f - sinT * sinT));
    float rPar = ((nt * cosI) - (ni * cosT)) / ((nt * cosI) + (ni * cosT));
    float rPerp = ((ni * cosI) - (nt * cosT)) / ((ni * cosI) + (nt * cosT));
    return (rPar * rPar + rPerp * rPerp) / 2;
}

// ---- GGX (Ggx.cpp) --------------------------------------------------------------------------------------------
static SELAS_HDN float smithG1Aniso(V3 w, float ax, float ay) {  // :19-37; QUIRK: the half vector argument is unused
    float absTan = fabsf(tanTheta(w));
    if (isinf(absTan)) return 0.0f;
    float c = cosPhi(w), s = sinPhi(w);
    float a = sqrtf(c * c * ax * ax + s * s * ay * ay);
    float a2Tan2 = sq(a * absTan);
    float lambda = 0.5f * (-1.0f + sqrtf(1.0f + a2Tan2));
    return 1.0f / (1.0f + lambda);
}
SELAS_HD float smithG1Iso(V3 w, float a) {  // :40-46
    float a2 = a * a;
    float c = absCosTheta(w);
    return 2.0f / (1.0f + sqrtf(a2 + (1 - a2) * c * c));
}
static SELAS_HDN float ggxD(V3 wm, float ax, float ay) {  // :75-84
    float hx2 = sq(wm.x), hy2 = sq(wm.z), c2 = cos2Theta(wm);
    return 1.0f / (SELAS_PI * ax * ay * sq(hx2 / sq(ax) + hy2 / sq(ay) + c2));
}
// sin / cos evaluated in double and rounded once (see sampleVndf); -DSELAS_FLOAT_SINCOS builds the plain sinf / cosf form for A/B
#ifdef SELAS_FLOAT_SINCOS
SELAS_HD float sinRounded(float x) { return sinf(x); }
SELAS_HD float cosRounded(float x) { return cosf(x); }
#else
SELAS_HD float sinRounded(float x) { return (float)sin((double)x); }
SELAS_HD float cosRounded(float x) { return (float)cos((double)x); }
#endif
// both at once (one argument reduction): every sampling site wants the pair
SELAS_HD void sinCosRounded(float x, float& s, float& c) {
#if defined(SELAS_FLOAT_SINCOS)
    s = sinf(x); c = cosf(x);
#else
    double ds, dc;
    sincos((double)x, &ds, &dc);
    s = (float)ds; c = (float)dc;
#endif
}
static SELAS_HDN V3 sampleVndf(V3 wo, float ax, float ay, float u1, float u2) {  // :105-126
    V3 v = normalize(v3(wo.x * ax, wo.y, wo.z * ay));
    V3 t1 = (v.y < 0.9999f) ? normalize(cross(v, v3(0.f, 1.f, 0.f))) : v3(1.f, 0.f, 0.f);
    V3 t2 = cross(t1, v);
    float a = 1.0f / (1.0f + v.y);
    float r = sqrtf(u1);
    float phi = (u2 < a) ? (u2 / a) * SELAS_PI : SELAS_PI + (u2 - a) / (1.0f - a) * SELAS_PI;
    // sin / cos of the disc angle in double, rounded once: glibc's sinf / cosf (what the reference links) are correctly rounded in
    // practice, CUDA's are within 2 ulp -- and for small u1 this sampler turns a 1-ulp difference of p1 / p2 into a jump of a whole
    // float32 quantum of the result (1.5e-5 in wi, up to 6e-4 in the pdf: tests/test_gpu_parity.py::test_bsdf_sample_parity)
    float sinPhi, cosPhi;
    sinCosRounded(phi, sinPhi, cosPhi);
    float p1 = r * cosPhi;
    float p2 = r * sinPhi * ((u2 < a) ? 1.0f : v.y);
    V3 n = p1 * t1 + p2 * t2 + sqrtf(maxf_(0.0f, 1.0f - p1 * p1 - p2 * p2)) * v;
    return normalize(v3(ax * n.x, n.y, ay * n.z));
}
static SELAS_HDN void vndfPdf(V3 wi, V3 wm, V3 wo, float ax, float ay, float& fwd, float& rev) {  // :141-155
    float D = ggxD(wm, ax, ay);
    fwd = smithG1Aniso(wo, ax, ay) * fabsf(dot(wm, wi)) * D / absCosTheta(wi);
    rev = smithG1Aniso(wi, ax, ay) * fabsf(dot(wm, wo)) * D / absCosTheta(wo);
}

// ---- per-shading-point cache ------------------------------------------------------------------------------------
struct DisneyFrame {
    float pSpec, pDiff, pCoat, pTrans;  // CalculateLobePdfs, Disney.cpp:26-44
    float ax, ay;                       // CalculateAnisotropicParams(roughness), :56-61
    float tax, tay;                     // ... of the (thin-scaled) transmission roughness, :359-362/:583-585
    V3 tint;                            // CalculateTint, :64-70
    bool thin;
};

SELAS_HD void anisoParams(float roughness, float anisotropic, float& ax, float& ay) {
    float aspect = sqrtf(1.0f - 0.9f * anisotropic);
    ax = maxf_(0.001f, sq(roughness) / aspect);
    ay = maxf_(0.001f, sq(roughness) * aspect);
}

SELAS_HD void makeDisneyFrame(const Surface& s, bool thin, DisneyFrame& f) {
    float metallicBRDF = s.metallic;
    float specularBSDF = (1.0f - s.metallic) * s.specTrans;
    float dielectricBRDF = (1.0f - s.specTrans) * (1.0f - s.metallic);
    float wSpec = metallicBRDF + dielectricBRDF;
    float wCoat = 1.0f * saturate_(s.clearcoat);
    float norm = 1.0f / (wSpec + specularBSDF + dielectricBRDF + wCoat);
    f.pSpec = wSpec * norm;
    f.pTrans = specularBSDF * norm;
    f.pDiff = dielectricBRDF * norm;
    f.pCoat = wCoat * norm;
    anisoParams(s.roughness, s.anisotropic, f.ax, f.ay);
    // QUIRK Disney.cpp:47-53: thin-surface roughness remap saturate((0.65*ior - 0.35)*roughness)
    float rscaled = thin ? saturate_((0.65f * s.ior - 0.35f) * s.roughness) : s.roughness;
    anisoParams(rscaled, s.anisotropic, f.tax, f.tay);
    // QUIRK Disney.cpp:68: "luminance" weights are (0.3, 0.6, 1.0)
    float lum = dot(v3(0.3f, 0.6f, 1.0f), s.baseColor);
    f.tint = (lum > 0.0f) ? s.baseColor * (1.0f / lum) : v3(1.0f);
    f.thin = thin;
}

// QUIRK Disney.cpp:74-82: GTR1 normalisation uses log2, not ln
SELAS_HD float gtr1(float absDotHL, float a) {
    if (a >= 1) return SELAS_INV_PI;
    float a2 = a * a;
    return (a2 - 1.0f) / (SELAS_PI * log2f(a2) * (1.0f + (a2 - 1.0f) * absDotHL * absDotHL));
}

SELAS_HD V3 sheenTerm(const Surface& s, const DisneyFrame& f, V3 wm, V3 wi) {  // :109-119
    if (s.sheen <= 0.0f) return v3(0.0f);
    float dotHL = fabsf(dot(wm, wi));
    return s.sheen * lerp3(v3(1.0f), f.tint, s.sheenTint) * schlickWeight(dotHL);
}

static SELAS_HDN V3 disneyFresnel(const Surface& s, const DisneyFrame& f, V3 wo, V3 wm, V3 wi) {  // :122-137
    float dotHV = dot(wm, wo);
    V3 R0 = schlickR0(s.relativeIOR) * lerp3(v3(1.0f), f.tint, s.specularTint);
    R0 = lerp3(R0, s.baseColor, s.metallic);
    float dielectric = fresnelDielectric(dotHV, 1.0f, s.ior);
    V3 metallic = schlickColor(R0, dot(wi, wm));
    return lerp3(v3(dielectric), metallic, s.metallic);
}

static SELAS_HDN float diffuseTerm(const Surface& s, V3 wo, V3 wm, V3 wi, bool thin) {  // :248-290
    float dotNL = absCosTheta(wi), dotNV = absCosTheta(wo);
    float fl = schlickWeight(dotNL), fv = schlickWeight(dotNV);
    float hk = 0.0f;
    if (thin && s.flatness > 0.0f) {
        float rough = s.roughness * s.roughness;
        float dotHL = dot(wm, wi);
        float fss90 = dotHL * dotHL * rough;
        float fss = lerpf(1.0f, fss90, fl) * lerpf(1.0f, fss90, fv);
        hk = 1.25f * (fss * (1.0f / (dotNL + dotNV) - 0.5f) + 0.5f);
    }
    // retro-reflection, :248-260
    float rough = s.roughness * s.roughness;
    float rr = 0.5f + 2.0f * dotNL * dotNL * rough;
    float retro = rr * (fl + fv + fl * fv * (rr - 1.0f));
    float subsurfaceApprox = lerpf(1.0f, hk, thin ? s.flatness : 0.0f);
    return SELAS_INV_PI * (retro + subsurfaceApprox * (1.0f - 0.5f * fl) * (1.0f - 0.5f * fv));
}

// QUIRK Disney.cpp:337-343 + FloatFuncs.h:147: "1.0f / (s*scatterDistance)" evaluates to (s*scatterDistance)/1
SELAS_HD V3 extinctionOf(V3 a, float scatterDistance) {
    V3 s = v3(1.9f) - a + 3.5f * (a - v3(0.8f)) * (a - v3(0.8f));
    return refScalarOverVec(1.0f, s * scatterDistance);
}

// EvaluateDisney, Disney.cpp:525-614.  v,l in world space; returns f*|cos(theta_l)|.
SELAS_HD V3 evaluateDisney(const Surface& s, const DisneyFrame& f, V3 v, V3 l, float& fwd, float& rev) {
    V3 wo = normalize(mul33(v, s.w2t));
    V3 wi = normalize(mul33(l, s.w2t));
    V3 wm = normalize(wo + wi);
    float dotNV = cosTheta(wo), dotNL = cosTheta(wi);
    V3 refl = v3(0.0f);
    fwd = 0.0f;
    rev = 0.0f;
    float diffuseWeight = (1.0f - s.metallic) * (1.0f - s.specTrans);
    float transWeight = (1.0f - s.metallic) * s.specTrans;
    bool upper = dotNL > 0.0f && dotNV > 0.0f;

    if (upper && s.clearcoat > 0.0f) {  // EvaluateDisneyClearcoat, :85-106
        float d = gtr1(absCosTheta(wm), lerpf(0.1f, 0.001f, s.clearcoatGloss));
        float fr = schlickScalar(0.04f, dot(wm, wi));
        float gl = smithG1Iso(wi, 0.25f), gv = smithG1Iso(wo, 0.25f);
        float fp = d / (4.0f * fabsf(dot(wo, wm)));
        float rp = d / (4.0f * fabsf(dot(wi, wm)));
        refl = refl + v3(0.25f * s.clearcoat * d * fr * gl * gv);
        fwd += f.pCoat * fp;
        rev += f.pCoat * rp;
    }
    if (diffuseWeight > 0.0f) {
        float diffuse = diffuseTerm(s, wo, wm, wi, f.thin);
        V3 sheen = sheenTerm(s, f, wm, wi);
        refl = refl + diffuseWeight * (diffuse * s.baseColor + sheen);
        fwd += f.pDiff * absCosTheta(wi);
        rev += f.pDiff * absCosTheta(wo);
    }
    if (transWeight > 0.0f) {  // EvaluateDisneySpecTransmission, :214-245
        float relIor = s.relativeIOR;
        float n2 = relIor * relIor;
        float absNL = absCosTheta(wi), absNV = absCosTheta(wo);
        float dotHL = dot(wm, wi), dotHV = dot(wm, wo);
        float d = ggxD(wm, f.tax, f.tay);
        float gl = smithG1Aniso(wi, f.tax, f.tay), gv = smithG1Aniso(wo, f.tax, f.tay);
        float fr = fresnelDielectric(dotHV, 1.0f, s.ior);
        V3 color = f.thin ? sqrt3(s.baseColor) : s.baseColor;
        // QUIRK :239-244: no 1/n^2 spreading factor
        float c = (fabsf(dotHL) * fabsf(dotHV)) / (absNL * absNV);
        float t = (n2 / sq(dotHL + relIor * dotHV));
        refl = refl + transWeight * (color * c * t * (1.0f - fr) * gl * gv * d);
        float fp, rp;
        vndfPdf(wi, wm, wo, f.tax, f.tay, fp, rp);
        fwd += f.pTrans * fp / (sq(dotHL + s.relativeIOR * dotHV));
        rev += f.pTrans * rp / (sq(dotHV + s.relativeIOR * dotHL));
    }
    if (upper) {  // EvaluateDisneyBRDF, :140-166 (dotNL, dotNV > 0 already established)
        float d = ggxD(wm, f.ax, f.ay);
        float gl = smithG1Aniso(wi, f.ax, f.ay), gv = smithG1Aniso(wo, f.ax, f.ay);
        V3 fr = disneyFresnel(s, f, wo, wm, wi);
        float fp, rp;
        vndfPdf(wi, wm, wo, f.ax, f.ay, fp, rp);
        fp *= (1.0f / (4 * fabsf(dot(wo, wm))));
        rp *= (1.0f / (4 * fabsf(dot(wi, wm))));
        refl = refl + d * gl * gv * fr / (4.0f * dotNL * dotNV);
        // QUIRK :607-608: divided by 4|wo.wm| a second time
        fwd += f.pSpec * fp / (4 * fabsf(dot(wo, wm)));
        rev += f.pSpec * rp / (4 * fabsf(dot(wi, wm)));
    }
    return refl * fabsf(dotNL);
}

SELAS_HD void failSample(BsdfResult& o) {
    o.forwardPdfW = 0.0f;
    o.reversePdfW = 0.0f;
    o.reflectance = v3(0.0f);
    o.wi = v3(0.0f);
}

// SampleDisney, Disney.cpp:617-659.  Rng must provide float next(); numbers are drawn in the reference's
// order: lobe pick, then 2 (specular / clearcoat) or 2+1 (diffuse, transmission).
template <typename Rng>
SELAS_HD bool sampleDisney(Rng& rng, const Surface& s, const DisneyFrame& f, V3 v, BsdfResult& o) {
    bool ok = false;
    float pLobe = 0.0f;
    float p = rng.next();
    if (p <= f.pSpec) {
        // ---- SampleDisneyBRDF, :169-211 ----
        pLobe = f.pSpec;
        V3 wo = normalize(mul33(v, s.w2t));
        float r0 = rng.next(), r1 = rng.next();
        V3 wm = sampleVndf(wo, f.ax, f.ay, r0, r1);
        V3 wi = normalize(reflect_(wm, wo));
        if (cosTheta(wi) <= 0.0f) {
            failSample(o);
        } else {
            V3 F = disneyFresnel(s, f, wo, wm, wi);
            float G1v = smithG1Aniso(wo, f.ax, f.ay);
            o.flags = SELAS_EVENT_SCATTER;
            o.reflectance = G1v * F;
            o.wi = normalize(mul33t(wi, s.w2t));
            vndfPdf(wi, wm, wo, f.ax, f.ay, o.forwardPdfW, o.reversePdfW);
            o.forwardPdfW *= (1.0f / (4 * fabsf(dot(wo, wm))));
            o.reversePdfW *= (1.0f / (4 * fabsf(dot(wi, wm))));
            ok = true;
        }
    } else if (p > f.pSpec && p <= (f.pSpec + f.pCoat)) {
        // ---- SampleDisneyClearcoat, :293-334 ----
        pLobe = f.pCoat;
        V3 wo = normalize(mul33(v, s.w2t));
        float a = 0.25f, a2 = a * a;
        float r0 = rng.next(), r1 = rng.next();
        float cosT = sqrtf(maxf_(0.f, (1.0f - powf(a2, 1.0f - r0)) / (1.0f - a2)));
        float sinT = sqrtf(maxf_(0.f, 1.0f - cosT * cosT));
        float phi = SELAS_TWO_PI * r1;
        float sinPhi, cosPhi;
        sinCosRounded(phi, sinPhi, cosPhi);
        V3 wm = v3(sinT * cosPhi, cosT, sinT * sinPhi);  // rounded once, like sampleVndf
        if (dot(wm, wo) < 0.0f) wm = -wm;
        V3 wi = reflect_(wm, wo);
        if (dot(wi, wo) < 0.0f) {
            ok = false;  // QUIRK :312-314: returns without touching the sample
        } else {
            float d = gtr1(fabsf(cosTheta(wm)), lerpf(0.1f, 0.001f, s.clearcoatGloss));
            float fr = schlickScalar(0.04f, dot(wm, wi));
            float g = smithG1Iso(wi, 0.25f) * smithG1Iso(wo, 0.25f);
            float fPdf = d / (4.0f * dot(wo, wm));  // QUIRK :326: signed dot
            o.reflectance = v3(0.25f * s.clearcoat * g * fr * d) / fPdf;
            o.wi = normalize(mul33t(wi, s.w2t));
            o.forwardPdfW = fPdf;
            o.reversePdfW = d / (4.0f * dot(wi, wm));
            // QUIRK: flags never written here (Scattering.h:36 has no default) -> caller's eScatterEvent stays
            ok = true;
        }
    } else if (p > f.pSpec + f.pCoat && p <= (f.pSpec + f.pCoat + f.pDiff)) {
        // ---- SampleDisneyDiffuse, :464-522 ----
        pLobe = f.pDiff;
        V3 wo = mul33(v, s.w2t);
        float sign = signf_(cosTheta(wo));
        float r0 = rng.next(), r1 = rng.next();
        float r = sqrtf(r0);
        float theta = SELAS_TWO_PI * r1;
        float sinTh, cosTh;
        sinCosRounded(theta, sinTh, cosTh);
        V3 wi = sign * v3(r * cosTh, sqrtf(maxf_(0.0f, 1 - r0)), r * sinTh);  // rounded once
        V3 wm = normalize(wi + wo);
        float dotNL = cosTheta(wi);
        if (dotNL == 0.0f) {
            failSample(o);
        } else {
            float dotNV = cosTheta(wo);
            float pdf;
            uint32_t eventType = SELAS_EVENT_SCATTER;
            V3 color = s.baseColor;
            float pp = rng.next();
            if (pp <= s.diffTrans) {
                wi = -wi;
                pdf = s.diffTrans;
                if (f.thin)
                    color = sqrt3(color);
                else {
                    eventType = SELAS_EVENT_TRANSMISSION;
                    o.mediumPhase = SELAS_MEDIUM_ISOTROPIC;
                    o.mediumExtinction = extinctionOf(s.transmittance, s.scatterDistance);
                }
            } else {
                pdf = (1.0f - s.diffTrans);
            }
            V3 sheen = sheenTerm(s, f, wm, wi);
            float diffuse = diffuseTerm(s, wo, wm, wi, f.thin);
            o.reflectance = sheen + color * (diffuse / pdf);
            o.wi = normalize(mul33t(wi, s.w2t));
            o.forwardPdfW = fabsf(dotNL) * pdf;
            o.reversePdfW = fabsf(dotNV) * pdf;
            o.flags = eventType;
            ok = true;
        }
    } else if (f.pTrans >= 0.0f) {
        // ---- SampleDisneySpecTransmission, :346-452 ----
        pLobe = f.pTrans;
        V3 wo = mul33(v, s.w2t);
        if (cosTheta(wo) == 0.0f) {
            failSample(o);
        } else {
            float r0 = rng.next(), r1 = rng.next();
            V3 wm = sampleVndf(wo, f.tax, f.tay, r0, r1);
            float dotVH = dot(wo, wm);
            if (wm.y < 0.0f) dotVH = -dotVH;
            float ni = wo.y > 0.0f ? 1.0f : s.ior;
            float nt = wo.y > 0.0f ? s.ior : 1.0f;
            float relativeIOR = ni / nt;
            float F = fresnelDielectric(dotVH, 1.0f, s.ior);
            float G1v = smithG1Aniso(wo, f.tax, f.tay);
            float pdf;
            V3 wi;
            if (rng.next() <= F) {
                wi = normalize(reflect_(wm, wo));
                o.flags = SELAS_EVENT_SCATTER;
                o.reflectance = G1v * s.baseColor;
                float jacobian = (4 * fabsf(dot(wo, wm)));
                pdf = F / jacobian;
            } else {
                if (f.thin) {
                    wi = reflect_(wm, wo);
                    wi.y = -wi.y;
                    o.reflectance = G1v * sqrt3(s.baseColor);
                    o.flags = SELAS_EVENT_SCATTER;
                } else {
                    if (transmit_(wm, wo, relativeIOR, wi)) {
                        o.flags = SELAS_EVENT_TRANSMISSION;
                        o.mediumPhase = dotVH > 0.0f ? SELAS_MEDIUM_ISOTROPIC : SELAS_MEDIUM_VACUUM;
                        o.mediumExtinction = extinctionOf(s.transmittance, s.scatterDistance);
                    } else {
                        o.flags = SELAS_EVENT_SCATTER;
                        wi = reflect_(wm, wo);
                    }
                    o.reflectance = G1v * s.baseColor;
                }
                wi = normalize(wi);
                float dotLH = fabsf(dot(wi, wm));
                // QUIRK :426-428: Jacobian uses surface.relativeIOR, not the local ni/nt
                float jacobian = dotLH / (sq(dotLH + s.relativeIOR * dotVH));
                pdf = (1.0f - F) / jacobian;
            }
            if (cosTheta(wi) == 0.0f) {
                failSample(o);
            } else {
                if (s.roughness < 0.01f) o.flags |= SELAS_EVENT_DIRAC;
                vndfPdf(wi, wm, wo, f.tax, f.tay, o.forwardPdfW, o.reversePdfW);
                o.forwardPdfW *= pdf;
                o.reversePdfW *= pdf;
                o.wi = normalize(mul33t(wi, s.w2t));
                ok = true;
            }
        }
    } else {
        o.reflectance = v3(1000000.0f, 0.0f, 0.0f);  // reference's "make sure we notice" sentinel, :645-650
        o.forwardPdfW = 0.000000001f;
        o.reversePdfW = 0.000000001f;
    }
    if (pLobe > 0.0f) {
        o.reflectance = o.reflectance * (1.0f / pLobe);
        o.forwardPdfW *= pLobe;
        o.reversePdfW *= pLobe;
    }
    return ok;
}

// QUIRK SurfaceParameters.cpp:391-394: "GeometricNormal" is the shading normal (column 1 of worldToTangent)
SELAS_HD V3 shadingNormal(const Surface& s) { return v3(s.w2t[1], s.w2t[4], s.w2t[7]); }

// SampleDiracTransparent, DiracTransparent.cpp:20-68
template <typename Rng>
SELAS_HD bool sampleDiracTransparent(Rng& rng, const Surface& s, V3 v, BsdfResult& o) {
    V3 normal = shadingNormal(s);
    float dotVH = dot(v, normal);
    float ni = dotVH > 0.0f ? 1.0f : s.ior;
    float nt = dotVH > 0.0f ? s.ior : 1.0f;
    float n = ni / nt;
    float F = fresnelDielectric(dotVH, ni, nt);
    float pdf;
    V3 wi;
    if (rng.next() < F) {
        wi = normalize(reflect_(normal, v));
        o.flags = SELAS_EVENT_SCATTER;
        pdf = F;
    } else {
        if (transmit_(normal, v, n, wi)) {
            o.flags = SELAS_EVENT_TRANSMISSION;
        } else {
            wi = reflect_(normal, v);
            o.flags = SELAS_EVENT_SCATTER;
        }
        wi = normalize(wi);
        pdf = (1.0f - F);
    }
    o.flags |= SELAS_EVENT_DIRAC;
    o.reflectance = s.baseColor / pdf;
    o.wi = wi;
    o.forwardPdfW = pdf;
    o.reversePdfW = pdf;
    return true;
}

SELAS_HD void defaultBsdfResult(BsdfResult& o) {  // BsdfSample defaults, Scattering.h:34-43
    o.flags = SELAS_EVENT_SCATTER;                // no default in the reference; defined (appendix B #4)
    o.mediumPhase = SELAS_MEDIUM_VACUUM;
    o.mediumExtinction = v3(0.0f);
    o.reflectance = v3(0.0f);
    o.wi = v3(0.0f);
    o.forwardPdfW = 0.0f;
    o.reversePdfW = 0.0f;
}

// EvaluateBsdf, SurfaceScattering.cpp:41-61
SELAS_HD V3 evaluateBsdf(const Surface& s, const DisneyFrame& f, V3 v, V3 l, float& fwd, float& rev) {
    if (s.shader == SELAS_SHADER_DISNEY_THIN || s.shader == SELAS_SHADER_DISNEY_SOLID) return evaluateDisney(s, f, v, l, fwd, rev);
    return v3(0.0f);
}
// SampleBsdfFunction, SurfaceScattering.cpp:18-38
template <typename Rng>
SELAS_HD bool sampleBsdf(Rng& rng, const Surface& s, const DisneyFrame& f, V3 v, BsdfResult& o) {
    defaultBsdfResult(o);
    if (s.shader == SELAS_SHADER_DISNEY_THIN || s.shader == SELAS_SHADER_DISNEY_SOLID) return sampleDisney(rng, s, f, v, o);
    if (s.shader == SELAS_SHADER_DIRAC_TRANSPARENT) return sampleDiracTransparent(rng, s, v, o);
    return false;
}

// OffsetRayOrigin, SurfaceParameters.cpp:403-408
SELAS_HD V3 offsetRayOrigin(const Surface& s, V3 direction, float biasScale) {
    V3 n = shadingNormal(s);
    float dirOffset = dot(direction, n) < 0.0f ? -1.0f : 1.0f;
    V3 offset = dirOffset * s.error * biasScale * n;
    return s.position + offset;
}

}  // namespace selas
