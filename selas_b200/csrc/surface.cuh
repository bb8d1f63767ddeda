// surface.cuh -- hit -> shading point (frame + material), the device restatement of
// CalculateSurfaceParams (Source/Core/Shading/SurfaceParameters.cpp:144-247) and of the HitParameters
// fill-in done by RayPick (Source/Applications/Selas/Source/PathTracer.cpp:119-136).
//
// rtcInterpolate0 (SurfaceParameters.cpp:162-189) becomes direct gathers from the per-mesh attribute arrays:
// triangles   t = (1-u-v) t0 + u t1 + v t2                              RTC_GEOMETRY_TYPE_TRIANGLE.3embree3:47
// quads       t = (1-v)((1-u) t0 + u t1) + v ((1-u) t3 + u t2)          RTC_GEOMETRY_TYPE_QUAD.3embree3:52
#pragma once
#include "disney.cuh"
#include "curve.cuh"
#include "internal.h"
#include "ptex_eval.cuh"

namespace selas {

SELAS_D float interpAttr(const float* a, const uint32_t* ix, uint32_t ipf, int comps, int c, float u, float v) {
    if (ipf == 3) {
        float t0 = __ldg(a + (size_t)ix[0] * comps + c), t1 = __ldg(a + (size_t)ix[1] * comps + c), t2 = __ldg(a + (size_t)ix[2] * comps + c);
        return (1.0f - u - v) * t0 + u * t1 + v * t2;
    }
    float t0 = __ldg(a + (size_t)ix[0] * comps + c), t1 = __ldg(a + (size_t)ix[1] * comps + c), t2 = __ldg(a + (size_t)ix[2] * comps + c),
          t3 = __ldg(a + (size_t)ix[3] * comps + c);
    return (1.0f - v) * ((1.0f - u) * t0 + u * t1) + v * ((1.0f - u) * t3 + u * t2);
}

// ---- non-Ptex base colour textures ----------------------------------------------------------------------------------------
// TextureFiltering::Sample with WrapMode::Repeat (TextureFiltering.h:35-52).  QUIRK: `s % w` mixes int32 and uint32, so a negative
// coordinate is reduced as a 32-bit unsigned number: -1 wraps to w - 1 only when w is a power of two.
SELAS_D size_t textureTexel(const TextureGpu& tex, int32_t s, int32_t t) {
    uint32_t ss = (uint32_t)s % tex.width, tt = (uint32_t)t % tex.height;
    return (size_t)tt * tex.width + ss;
}
// Math::SrgbToLinearPrecise, MathLib/ColorSpace.cpp:16-23
SELAS_D float srgbToLinearPrecise(float x) {
    if (x <= 0.04045f) return x / 12.92f;
    return powf((x + 0.055f) / (1.0f + 0.055f), 2.4f);
}
// SampleTextureFloat3(texture, uvs, sRGB, default) (SurfaceParameters.cpp:97-118) over TextureFiltering::Triangle at level 0
// (TextureFiltering.h:77-104): four taps weighted (1-ds)(1-dt), (1-ds)dt, ds(1-dt), ds dt, summed in that order.
SELAS_D V3 sampleTextureFloat3(const TextureGpu& tex, float u, float v, bool sRGB) {
    if (tex.format == SELAS_TEXTURE_FLOAT2) return v3(0.0f);  // Assert_(false) branch, :116-117
    float s = u * tex.width - 0.5f, t = v * tex.height - 0.5f;
    int32_t s0 = (int32_t)floorf(s), t0 = (int32_t)floorf(t);
    float ds = s - s0, dt = t - t0;
    const float w00 = (1 - ds) * (1 - dt), w01 = (1 - ds) * dt, w10 = ds * (1 - dt), w11 = ds * dt;
    const uint32_t comps = tex.format + 1u;
    const size_t i00 = textureTexel(tex, s0, t0) * comps, i01 = textureTexel(tex, s0, t0 + 1) * comps,
                 i10 = textureTexel(tex, s0 + 1, t0) * comps, i11 = textureTexel(tex, s0 + 1, t0 + 1) * comps;
    float r[3];
    const uint32_t used = comps == 1 ? 1u : 3u;
    for (uint32_t c = 0; c < used; ++c) {
        float x = w00 * __ldg(tex.texels + i00 + c) + w01 * __ldg(tex.texels + i01 + c) + w10 * __ldg(tex.texels + i10 + c) +
                  w11 * __ldg(tex.texels + i11 + c);
        r[c] = sRGB ? srgbToLinearPrecise(x) : x;
    }
    return comps == 1 ? v3(r[0]) : v3(r[0], r[1], r[2]);
}

// ptexRgb: this hit's Ptex lookup, made by the Ptex stage kernel (ptexStageKernel) before shading, or null.
// hit->normal of a hit on a curve geometry = Embree's Ng of a round curve: the surface normal (RTC_GEOMETRY_TYPE_CURVE.3embree3); curve
// geometries carry no attribute buffers (ModelResource.cpp:314-319).  Rebuilt from the element that was hit, with the control vertices
// taken to world space (for rotations and uniform scales the same as sending the object-space normal through localToWorld).  Out of
// line: scenes without curves never come here, and the shade kernels stay as lean as they were (inlined it cost them 6 registers and
// 6 % more instructions on configs[2], which has no curves at all).
static __device__ __noinline__ V3 curveSurfaceNormal(const MeshGpu& mesh, const InstanceGpu* inst, uint32_t primID, uint32_t element, V3 p) {
    const bool identity = inst->identity != 0;
    const float4* cv = reinterpret_cast<const float4*>(mesh.positions) + __ldg(mesh.indices + primID);
    float4 c[4];
    float rscale = 1.0f;
    const float* m = inst->m;
    if (!identity) rscale = cbrtf(fabsf(m[0] * (m[5] * m[10] - m[6] * m[9]) - m[1] * (m[4] * m[10] - m[6] * m[8]) + m[2] * (m[4] * m[9] - m[5] * m[8])));
    for (int k = 0; k < 4; ++k) {
        const float4 q4 = __ldg(cv + k);
        V3 q = v3(q4.x, q4.y, q4.z);
        if (!identity) q = mulPoint44(q, m);
        c[k] = make_float4(q.x, q.y, q.z, q4.w * rscale);
    }
    return curveElementNormal(c[0], c[1], c[2], c[3], element ? element : 1u, p);
}

// element: last word of the hit record -- for a hit on a curve geometry the piece of the segment that was hit (curve.cuh), else 0.
SELAS_D void surfaceFromHit(const SceneGpu& sc, V3 o, V3 d, float t, float u, float v, uint32_t primID, uint32_t geomID,
                            uint32_t instID, const float* ptexRgb, Surface& s, uint32_t element = 0u) {
    const MeshGpu mesh = sc.meshes[geomID];
    const InstanceGpu* inst = sc.instances + instID;
    const SelasMaterial* mat = sc.materials + mesh.materialIndex;
    const bool isCurve = mesh.indicesPerFace == 1;
    uint32_t ix[4] = {0u, 0u, 0u, 0u};
    if (!isCurve) {
        const uint32_t* ip = mesh.indices + (size_t)primID * mesh.indicesPerFace;
        ix[0] = __ldg(ip); ix[1] = __ldg(ip + 1); ix[2] = __ldg(ip + 2);
        ix[3] = mesh.indicesPerFace == 4 ? __ldg(ip + 3) : 0u;
    }
    const bool identity = inst->identity != 0;
    float m[16];
    if (!identity)
        for (int k = 0; k < 16; ++k) m[k] = inst->m[k];

    V3 normal;
    if (isCurve) {
        normal = curveSurfaceNormal(mesh, inst, primID, element, v3(o.x + t * d.x, o.y + t * d.y, o.z + t * d.z));
    } else if (mesh.normals) {
        normal = v3(interpAttr(mesh.normals, ix, mesh.indicesPerFace, 3, 0, u, v), interpAttr(mesh.normals, ix, mesh.indicesPerFace, 3, 1, u, v),
                    interpAttr(mesh.normals, ix, mesh.indicesPerFace, 3, 2, u, v));
    } else {
        // hit->normal = Embree's Ng: object-space cross(v1-v0, v2-v0) of the triangle that was hit
        uint32_t i0, i1, i2;
        if (mesh.indicesPerFace == 3) { i0 = ix[0]; i1 = ix[1]; i2 = ix[2]; }
        else if ((u + v) <= 1.0f) { i0 = ix[0]; i1 = ix[1]; i2 = ix[3]; }
        else { i0 = ix[2]; i1 = ix[3]; i2 = ix[1]; }
        V3 p0 = v3(mesh.positions + 3 * (size_t)i0), p1 = v3(mesh.positions + 3 * (size_t)i1), p2 = v3(mesh.positions + 3 * (size_t)i2);
        normal = cross(p1 - p0, p2 - p0);
    }
    // QUIRK SurfaceParameters.cpp:165-171: normals go through localToWorld as plain vectors (no inverse transpose)
    if (!identity && !isCurve) normal = mulVec44(normal, m);
    V3 n = normalize(normal);
    V3 tg, b;
    if (mesh.tangents && !isCurve) {
        V3 lt = v3(interpAttr(mesh.tangents, ix, mesh.indicesPerFace, 4, 0, u, v), interpAttr(mesh.tangents, ix, mesh.indicesPerFace, 4, 1, u, v),
                   interpAttr(mesh.tangents, ix, mesh.indicesPerFace, 4, 2, u, v));
        float w = interpAttr(mesh.tangents, ix, mesh.indicesPerFace, 4, 3, u, v);
        tg = identity ? lt : mulVec44(lt, m);
        b = cross(n, tg) * w;
    } else {  // MakeOrthogonalCoordinateSystem, GeometryLib/CoordinateSystem.h:14-21
        if (fabsf(n.x) > fabsf(n.y))
            tg = v3(-n.z, 0.f, n.x) * (1.0f / sqrtf(n.x * n.x + n.z * n.z));
        else
            tg = v3(0.f, n.z, -n.y) * (1.0f / sqrtf(n.y * n.y + n.z * n.z));
        b = cross(n, tg);
    }

    V3 base;
    bool constantBase = false;
    if (ptexRgb && (mat->flags & SELAS_MAT_USES_PTEX) && mesh.ptex) {
        // SurfaceParameters.cpp:196-208: filter->eval(faceid = primId, u, v, widths 0), then pow 2.2 only.  The lookup itself
        // was done by the Ptex stage that runs between the closest-hit and the shade kernels.
        base = v3(ptexRgb[0], ptexRgb[1], ptexRgb[2]);
    } else if (mesh.texture && !isCurve) {
        // SurfaceParameters.cpp:209-214: SampleTextureFloat3(baseColorTexture, uvs, sRGB = true, material baseColor); uvs = (0, 0)
        // without a uv buffer (:186-190)
        float tu = 0.0f, tv = 0.0f;
        if (mesh.uvs) {
            tu = interpAttr(mesh.uvs, ix, mesh.indicesPerFace, 2, 0, u, v);
            tv = interpAttr(mesh.uvs, ix, mesh.indicesPerFace, 2, 1, u, v);
        }
        base = sampleTextureFloat3(*mesh.texture, tu, tv, true);
    } else {
        base = v3(mat->baseColor);
        constantBase = true;
    }
    // QUIRK SurfaceParameters.cpp:204,:211-212: pow(x, 2.2) applied on top of everything.  For an untextured mesh that is a constant of
    // the material, made once at upload with this same powf (three powf were 4.3 % of the shade kernel's instructions on configs[2]).
    if (constantBase) {
        const float4 bp = __ldg(sc.materialBasePow + mesh.materialIndex);
        s.baseColor = v3(bp.x, bp.y, bp.z);
    } else {
        s.baseColor = v3(powf(base.x, 2.2f), powf(base.y, 2.2f), powf(base.z, 2.2f));
    }

    // worldToTangent = transpose(rows(t, n, b))
    s.w2t[0] = tg.x; s.w2t[1] = n.x; s.w2t[2] = b.x;
    s.w2t[3] = tg.y; s.w2t[4] = n.y; s.w2t[5] = b.y;
    s.w2t[6] = tg.z; s.w2t[7] = n.z; s.w2t[8] = b.z;
    // RayPick: position recomputed from the ray, error bound from |p| and t  (PathTracer.cpp:119-136)
    V3 pos = v3(o.x + t * d.x, o.y + t * d.y, o.z + t * d.z);
    s.position = pos;
    const float kErr = 32.0f * 1.19209e-07f;
    s.error = kErr * maxf_(maxf_(fabsf(pos.x), fabsf(pos.y)), maxf_(fabsf(pos.z), t));
    s.materialFlags = mat->flags;
    s.transmittance = v3(mat->transmittanceColor);
    s.sheen = mat->scalars[SELAS_SHEEN];
    s.sheenTint = mat->scalars[SELAS_SHEEN_TINT];
    s.clearcoat = mat->scalars[SELAS_CLEARCOAT];
    s.clearcoatGloss = mat->scalars[SELAS_CLEARCOAT_GLOSS];
    s.specTrans = saturate_(mat->scalars[SELAS_SPEC_TRANS]);
    s.diffTrans = mat->scalars[SELAS_DIFFUSE_TRANS] * 0.5f;  // QUIRK :229
    s.flatness = mat->scalars[SELAS_FLATNESS];
    s.anisotropic = mat->scalars[SELAS_ANISOTROPIC];
    s.specularTint = mat->scalars[SELAS_SPECULAR_TINT];
    s.roughness = mat->scalars[SELAS_ROUGHNESS];
    s.metallic = saturate_(mat->scalars[SELAS_METALLIC]);
    s.scatterDistance = mat->scalars[SELAS_SCATTER_DISTANCE];
    s.ior = mat->scalars[SELAS_IOR];
    s.lightSetIndex = mesh.lightSetIndex;
    s.shader = mat->shader;
    V3 view = v3(-d.x, -d.y, -d.z);
    s.relativeIOR = ((mat->flags & SELAS_MAT_TRANSPARENT) && dot(view, n) < 0.0f) ? s.ior : 1.0f / s.ior;
}

// Ptex stage of the wavefront: base-colour lookups for every hit of the queue whose material uses Ptex
// (SurfaceParameters.cpp:196-208), written as 3 floats per path for the shade kernel.
template <bool MIXED>
static __global__ void __launch_bounds__(128, SELAS_PTEX_MIN_CTAS) ptexStageKernel(SceneGpu sc, uint32_t n, const SelasHit* __restrict__ hits, float* __restrict__ rgbOut,
                                                                                  const uint32_t* __restrict__ nPtr = nullptr) {
    if (nPtr) n = min(n, *nPtr);  // wavefront queues: the length lives in device memory, n is the launch's upper bound
    ptexTileEval<128, 8, MIXED>(
        n,
        [&](size_t i, const PtexGpu*& t, int& f, float& uu, float& vv) {
            const uint4* hp = reinterpret_cast<const uint4*>(hits + i);
            uint4 h0 = __ldg(hp), h1 = __ldg(hp + 1);
            if (h1.x == SELAS_INVALID_ID) return false;
            const MeshGpu& mesh = sc.meshes[h1.x];
            if (!((sc.materials[mesh.materialIndex].flags & SELAS_MAT_USES_PTEX) && mesh.ptex)) return false;
            t = (const PtexGpu*)mesh.ptex;
            f = (int)h0.w;
            uu = __uint_as_float(h0.y);
            vv = __uint_as_float(h0.z);
            return true;
        },
        [&](size_t i, const float* c) {
            rgbOut[3 * i] = c[0];
            rgbOut[3 * i + 1] = c[1];
            rgbOut[3 * i + 2] = c[2];
        });
}

SELAS_D void storeSurface(const Surface& s, V3 view, SelasSurface& o) {
    store3(o.position, s.position);
    for (int i = 0; i < 9; ++i) o.worldToTangent[i] = s.w2t[i];
    o.error = s.error;
    store3(o.view, view);
    store3(o.baseColor, s.baseColor);
    store3(o.transmittanceColor, s.transmittance);
    o.sheen = s.sheen; o.sheenTint = s.sheenTint; o.clearcoat = s.clearcoat; o.clearcoatGloss = s.clearcoatGloss;
    o.metallic = s.metallic; o.specTrans = s.specTrans; o.diffTrans = s.diffTrans; o.flatness = s.flatness;
    o.anisotropic = s.anisotropic; o.relativeIOR = s.relativeIOR; o.specularTint = s.specularTint; o.roughness = s.roughness;
    o.scatterDistance = s.scatterDistance; o.ior = s.ior;
    o.shader = s.shader; o.materialFlags = s.materialFlags; o.lightSetIndex = s.lightSetIndex;
}

}  // namespace selas
