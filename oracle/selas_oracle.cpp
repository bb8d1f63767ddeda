// oracle/selas_oracle.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// C entry points (prefix orc_) over the CPU restatement in orc_shading.h / cpu_bvh.h, shaped like the
// product's C ABI (include/selas_cuda.h) so a parity test is "same inputs through both".  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library; the
// product (selas_b200/) never does.
//
// Restated here in addition to the headers:
//   * RayPick / OcclusionRay            Applications/Selas/Source/PathTracer.cpp:71-139
//   * CalculateSurfaceParams            Core/Shading/SurfaceParameters.cpp:144-247
//   * EvaluatePath / PathTracerKernel   Applications/Selas/Source/PathTracer.cpp:142-299
//   * BuildHashGrid / SearchHashGrid    Applications/Selas/Source/VCMHashGrid.cpp:15-155
//   * FramebufferWriter_Write           Core/TextureLib/Framebuffer.cpp:104-135 (film[y*w+x] += sample)
//   * VCM::GenerateImage and callees    Applications/Selas/Source/VCM.cpp (orc_vcm.h)
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <vector>

#include <xmmintrin.h>
#include <pmmintrin.h>

#include "cpu_bvh.h"
#include "orc_curves.h"
#include "orc_shading.h"

using namespace orc;

namespace {

struct Mesh {
    std::vector<float> positions, normals, tangents, uvs, texels;
    Texture texture;  // base colour texture (texels above), or texels == nullptr
    std::vector<uint32_t> indices;
    uint32_t faceCount, indicesPerFace, materialIndex, lightSetIndex;
    // Ptex base colour lookup of this mesh (orc_scene_set_ptex): the checker's stand-in for TextureCache + PtexFilter::eval,
    // SurfaceParameters.cpp:196-208.  Tests bind it to the reference's own Ptex library (oracle/_ref/libptex_ref.so).
    void (*ptexEval)(void* user, int faceid, float u, float v, float* rgb3) = nullptr;
    void* ptexUser = nullptr;
};
struct Instance {
    uint32_t mesh;
    float m[16];
};
struct Scene {
    std::vector<Mesh> meshes;
    std::vector<Instance> instances;
    std::vector<SelasMaterial> materials;
    std::vector<SelasLight> lights;
    std::vector<SelasLightSetRange> lightSets;
    float background[4];
    bool hasIbl = false;
    Ibl ibl;
    std::vector<float> iblMarginal, iblConditional, iblLight, iblMiss;
    Bvh bvh;
    // curve geometries ride in `meshes` after the triangle meshes (indicesPerFace = 1: positions = float4 control vertices, indices =
    // first control vertex per segment); their flattened world-space segments are tested exhaustively (orc_curves.h)
    std::vector<orc::CurveSeg> curves;

    // scene-level trace: the triangle BVH, then the curve segments in front of what it found.  Real = float reproduces the timed CPU
    // path of the integrators, double is the hit-ID oracle; brute = exhaustive over the triangles too.
    template <typename Real>
    bool trace(const float o[3], const float d[3], float tnear, float tfar, bool anyHit, RayHit& out, bool brute = false) const {
        bool hit = brute ? bvh.brute(o, d, tnear, tfar, out) : bvh.trace<Real>(o, d, tnear, tfar, anyHit, out);
        if (curves.empty() || (hit && anyHit)) return hit;
        uint32_t seg = 0;
        orc::CurveHitD ch = {};
        if (orc::traceCurves(curves, o, d, tnear, hit ? out.t : (double)tfar, anyHit, seg, ch)) {
            out.t = ch.t; out.u = ch.u; out.v = 0.0;
            out.tri = 0xFFFFFFFEu;
            out.curve = seg;
            for (int k = 0; k < 3; ++k) out.ng[k] = ch.ng[k];
            return true;
        }
        return hit;
    }
};

// Product RNG contract (DESIGN.md "Random numbers"): PCG-RXS-M-XS-32 stream per (seed, pixel, sample).
// The reference uses one std::mt19937 per worker thread (MathLib/Random.cpp:22-55) with dynamic pixel
// claiming, so its images are not reproducible run to run; the contract replaces it with a counter-keyed
// stream so the CUDA path and this oracle walk identical paths.
static inline uint32_t mix32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}
struct PathRng {
    uint32_t state;
    void init(uint32_t seed, uint32_t pixel, uint32_t sample) { state = mix32(seed ^ mix32(pixel ^ mix32(sample + 0x9e3779b9u))); }
    float next() {
        state = state * 747796405u + 2891336453u;
        uint32_t w = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
        w = (w >> 22u) ^ w;
        return (float)(w >> 8) * (1.0f / 16777216.0f);
    }
};

static inline V3 mulVec44(V3 v, const float* m) {  // MatrixMultiplyVector, FloatFuncs.cpp:367-380
    return V3{v.x * m[0] + v.y * m[4] + v.z * m[8], v.x * m[1] + v.y * m[5] + v.z * m[9],
              v.x * m[2] + v.y * m[6] + v.z * m[10]};
}
static inline V3 mulPoint44(V3 v, const float* m) {  // MatrixMultiplyPoint, FloatFuncs.cpp:383-396
    return V3{v.x * m[0] + v.y * m[4] + v.z * m[8] + m[12], v.x * m[1] + v.y * m[5] + v.z * m[9] + m[13],
              v.x * m[2] + v.y * m[6] + v.z * m[10] + m[14]};
}

static void setFtz() {  // main.cpp:84-85
    _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);
    _MM_SET_DENORMALS_ZERO_MODE(_MM_DENORMALS_ZERO_ON);
}

template <typename F>
static void parallelFor(size_t n, int nthreads, F fn) {
    if (nthreads <= 1 || n < 1024) {
        setFtz();
        fn(0, n, 0);
        return;
    }
    std::vector<std::thread> th;
    std::atomic<size_t> next(0);
    const size_t chunk = 4096;
    for (int t = 0; t < nthreads; ++t)
        th.emplace_back([&, t]() {
            setFtz();
            for (;;) {
                size_t b = next.fetch_add(chunk);
                if (b >= n) break;
                fn(b, std::min(n, b + chunk), t);
            }
        });
    for (auto& x : th) x.join();
}

static void hitToRecord(const Scene& sc, const RayHit& h, const SelasRay& r, SelasHit& out, float* Ng4) {
    if (h.tri == 0xFFFFFFFFu) {
        out.t = r.tfar;
        out.u = out.v = 0.f;
        out.primID = out.geomID = out.instID0 = out.instID1 = SELAS_INVALID_ID;
        out.pad = 0;
        if (Ng4) Ng4[0] = Ng4[1] = Ng4[2] = Ng4[3] = 0.f;
        return;
    }
    if (h.curve != 0xFFFFFFFFu) {  // round curve: u along the segment, v = 0, Ng = surface normal (RTC_GEOMETRY_TYPE_CURVE.3embree3)
        const orc::CurveSeg& S = sc.curves[h.curve];
        out.t = (float)h.t;
        out.u = (float)h.u;
        out.v = 0.f;
        out.primID = S.primID;
        out.geomID = S.geomID;
        out.instID0 = S.instID;
        out.instID1 = 0;
        out.pad = 0;
        if (Ng4) { Ng4[0] = (float)h.ng[0]; Ng4[1] = (float)h.ng[1]; Ng4[2] = (float)h.ng[2]; Ng4[3] = 0.f; }
        return;
    }
    const Tri& T = sc.bvh.tris[h.tri];
    out.t = (float)h.t;
    out.u = T.second ? (float)(1.0 - h.u) : (float)h.u;
    out.v = T.second ? (float)(1.0 - h.v) : (float)h.v;
    out.primID = T.primID;
    out.geomID = T.geomID;
    out.instID0 = T.instID;
    out.instID1 = 0;
    out.pad = 0;
    if (Ng4) {
        V3 e1 = v3(T.v1) - v3(T.v0), e2 = v3(T.v2) - v3(T.v0);
        V3 n = cross(e1, e2);
        Ng4[0] = n.x; Ng4[1] = n.y; Ng4[2] = n.z; Ng4[3] = 0.f;
    }
}

// CalculateSurfaceParams, SurfaceParameters.cpp:144-247 (constant base colour; Ptex via callback later)
static void surfaceParams(const Scene& sc, const SelasRay& r, const SelasHit& h, SelasSurface& s) {
    const Mesh& mesh = sc.meshes[h.geomID];
    const Instance& inst = sc.instances[h.instID0];
    const SelasMaterial& mat = sc.materials[mesh.materialIndex];
    const bool isCurve = mesh.indicesPerFace == 1;
    const uint32_t* idx = &mesh.indices[(size_t)h.primID * mesh.indicesPerFace];
    float u = h.u, v = h.v;

    auto interp3 = [&](const std::vector<float>& a, int comps, float* out) {
        for (int c = 0; c < comps; ++c) {
            if (mesh.indicesPerFace == 3) {
                // t_uv = (1-u-v)*t0 + u*t1 + v*t2           RTC_GEOMETRY_TYPE_TRIANGLE.3embree3:47
                float t0 = a[idx[0] * comps + c], t1 = a[idx[1] * comps + c], t2 = a[idx[2] * comps + c];
                out[c] = (1.0f - u - v) * t0 + u * t1 + v * t2;
            } else {
                // t_uv = (1-v)((1-u)*t0 + u*t1) + v*((1-u)*t3 + u*t2)   RTC_GEOMETRY_TYPE_QUAD.3embree3:52
                float t0 = a[idx[0] * comps + c], t1 = a[idx[1] * comps + c], t2 = a[idx[2] * comps + c],
                      t3 = a[idx[3] * comps + c];
                out[c] = (1.0f - v) * ((1.0f - u) * t0 + u * t1) + v * ((1.0f - u) * t3 + u * t2);
            }
        }
    };

    V3 normal;
    if (isCurve) {
        // hit->normal = Embree's Ng of a round curve = the surface normal; recovered by intersecting the segment again around the hit
        normal = v3(0.0f, 1.0f, 0.0f);
        for (const orc::CurveSeg& S : sc.curves) {
            if (S.geomID != h.geomID || S.primID != h.primID || S.instID != h.instID0) continue;
            const double o[3] = {r.ox, r.oy, r.oz}, d[3] = {r.dx, r.dy, r.dz};
            const double slack = 1e-3 * std::fabs((double)h.t) + 4.0 * std::fabs(S.c[1][3]) + 1e-6;
            orc::CurveHitD ch;
            if (orc::intersectCurveD(S, o, d, (double)h.t - slack, (double)h.t + slack, ch)) normal = v3((float)ch.ng[0], (float)ch.ng[1], (float)ch.ng[2]);
            break;
        }
    } else if (!mesh.normals.empty()) {
        float nn[3];
        interp3(mesh.normals, 3, nn);
        normal = mulVec44(v3(nn), inst.m);
    } else {
        // hit->normal = Embree Ng (object space, unnormalised)  PathTracer.cpp:123-125
        const float* P = mesh.positions.data();
        V3 ng;
        if (mesh.indicesPerFace == 3 || (u + v) <= 1.0f) {
            uint32_t i0 = idx[0], i1 = idx[1], i2 = idx[mesh.indicesPerFace == 3 ? 2 : 3];
            ng = cross(v3(P + 3 * i1) - v3(P + 3 * i0), v3(P + 3 * i2) - v3(P + 3 * i0));
        } else {
            uint32_t i0 = idx[2], i1 = idx[3], i2 = idx[1];
            ng = cross(v3(P + 3 * i1) - v3(P + 3 * i0), v3(P + 3 * i2) - v3(P + 3 * i0));
        }
        normal = mulVec44(ng, inst.m);
    }
    V3 n = normalize(normal);
    V3 t, b;
    if (!mesh.tangents.empty() && !isCurve) {
        float tt[4];
        interp3(mesh.tangents, 4, tt);
        t = mulVec44(v3(tt), inst.m);
        b = cross(n, t) * tt[3];
    } else {
        // MakeOrthogonalCoordinateSystem, GeometryLib/CoordinateSystem.h:14-21
        if (absf(n.x) > absf(n.y))
            t = v3(-n.z, 0, n.x) * (1.0f / sqrtf(n.x * n.x + n.z * n.z));
        else
            t = v3(0, n.z, -n.y) * (1.0f / sqrtf(n.y * n.y + n.z * n.z));
        b = cross(n, t);
    }
    V3 base = v3(mat.baseColor);
    if (mesh.ptexEval && (mat.flags & SELAS_MAT_USES_PTEX)) {  // SurfaceParameters.cpp:196-208: eval(faceid = primId, u, v, widths 0)
        float rgb[3] = {0.0f, 0.0f, 0.0f};
        mesh.ptexEval(mesh.ptexUser, (int)h.primID, u, v, rgb);
        base = v3(rgb);
    } else if (mesh.texture.texels && !isCurve) {  // SurfaceParameters.cpp:186-190, :209-214
        float uv[2] = {0.0f, 0.0f};
        if (!mesh.uvs.empty()) interp3(mesh.uvs, 2, uv);
        base = sampleTextureFloat3(mesh.texture, uv[0], uv[1], true);
    }
    base = v3(powf(base.x, 2.2f), powf(base.y, 2.2f), powf(base.z, 2.2f));
    store(s.baseColor, base);

    // worldToTangent = transpose(rows(t,n,b))  => rows (t.x,n.x,b.x),(t.y,n.y,b.y),(t.z,n.z,b.z)
    float* w = s.worldToTangent;
    w[0] = t.x; w[1] = n.x; w[2] = b.x;
    w[3] = t.y; w[4] = n.y; w[5] = b.y;
    w[6] = t.z; w[7] = n.z; w[8] = b.z;
    // RayPick, PathTracer.cpp:119-136
    V3 pos = v3(r.ox + h.t * r.dx, r.oy + h.t * r.dy, r.oz + h.t * r.dz);
    store(s.position, pos);
    const float kErr = 32.0f * 1.19209e-07f;
    s.error = kErr * maxf(maxf(absf(pos.x), absf(pos.y)), maxf(absf(pos.z), h.t));
    V3 view = v3(-r.dx, -r.dy, -r.dz);
    store(s.view, view);
    s.materialFlags = mat.flags;
    store(s.transmittanceColor, v3(mat.transmittanceColor));
    s.sheen = mat.scalars[SELAS_SHEEN];
    s.sheenTint = mat.scalars[SELAS_SHEEN_TINT];
    s.clearcoat = mat.scalars[SELAS_CLEARCOAT];
    s.clearcoatGloss = mat.scalars[SELAS_CLEARCOAT_GLOSS];
    s.specTrans = saturate(mat.scalars[SELAS_SPEC_TRANS]);
    s.diffTrans = mat.scalars[SELAS_DIFFUSE_TRANS] * 0.5f;
    s.flatness = mat.scalars[SELAS_FLATNESS];
    s.anisotropic = mat.scalars[SELAS_ANISOTROPIC];
    s.specularTint = mat.scalars[SELAS_SPECULAR_TINT];
    s.roughness = mat.scalars[SELAS_ROUGHNESS];
    s.metallic = saturate(mat.scalars[SELAS_METALLIC]);
    s.scatterDistance = mat.scalars[SELAS_SCATTER_DISTANCE];
    s.ior = mat.scalars[SELAS_IOR];
    s.lightSetIndex = mesh.lightSetIndex;
    s.shader = mat.shader;
    s.relativeIOR = ((mat.flags & SELAS_MAT_TRANSPARENT) && dot(view, n) < 0.0f) ? s.ior : 1.0f / s.ior;
}

}  // namespace

// Because BSDF sampling consumes a data-dependent number of uniforms, the integrator snapshots the PCG
// state, pre-draws SELAS_BSDF_RANDOMS values, lets the callee consume k of them, then re-advances the real
// state by exactly k.
template <typename Fn>
static inline void withStream(PathRng& rng, Fn fn) {
    PathRng tmp = rng;
    float buf[SELAS_BSDF_RANDOMS];
    for (int i = 0; i < SELAS_BSDF_RANDOMS; ++i) buf[i] = tmp.next();
    Stream st{buf, 0};
    fn(st);
    for (uint32_t i = 0; i < st.pos; ++i) rng.next();
}

static void evaluatePath(const Scene& sc, const SelasCamera& cam, const SelasRenderSettings& rs, uint32_t px, uint32_t py,
                         uint32_t sample, V3 Ld[2], uint64_t& closest, uint64_t& shadow) {
    PathRng rng;
    rng.init(rs.seed, py * rs.width + px, sample);
    // JitteredCameraRay, Camera.cpp:63-74
    float vx, vy;
    if (rs.pixelSampler == 0) {
        vx = (float)px + rng.next();
        vy = (float)py + rng.next();
    } else {  // JitteredCameraRay, CMJ overload: Camera.cpp:77-86 (s = sample index, p = pixel index)
        float j[2];
        cmj((int32_t)sample, (int32_t)(rs.pixelSampler & 0xffffu), (int32_t)(rs.pixelSampler >> 16), (int32_t)(py * rs.width + px), j);
        vx = (float)px + j[0];
        vy = (float)py + j[1];
    }
    SelasRay ray;
    V3 dir = imageToWorldDirection(cam, vx, vy);
    ray.ox = cam.position[0]; ray.oy = cam.position[1]; ray.oz = cam.position[2];
    ray.dx = dir.x; ray.dy = dir.y; ray.dz = dir.z;

    Ld[0] = Ld[1] = v3(0.0f);
    V3 throughput = v3(1.0f);
    Medium vacuum{SELAS_MEDIUM_VACUUM, v3(0.0f)};
    Medium medium = vacuum;
    bool isDeltaOnly = true;
    uint32_t bounce = 0;
    V3 background = v3(sc.background);

    while (bounce < rs.maxPathLength) {
        float rayDistance;
        withStream(rng, [&](Stream& st) { rayDistance = sampleDistance(st, medium); });
        ray.tnear = 0.0f;
        ray.tfar = rayDistance;
        RayHit rh;
        float o[3] = {ray.ox, ray.oy, ray.oz}, d[3] = {ray.dx, ray.dy, ray.dz};
        bool hit = sc.trace<float>(o, d, 0.0f, rayDistance, false, rh);
        ++closest;
        SelasHit h;
        hitToRecord(sc, rh, ray, h, nullptr);
        SelasSurface surf;
        if (hit) {
            surfaceParams(sc, ray, h, surf);
            rayDistance = length(v3(surf.position) - v3(ray.ox, ray.oy, ray.oz));
        }
        throughput = throughput * transmission(medium, rayDistance);

        if (hit) {
            SelasLightSample ls;
            memset(&ls, 0, sizeof(ls));
            uint32_t lightCount = 0;
            const SelasLight* lights = nullptr;
            if (surf.lightSetIndex < sc.lightSets.size()) {
                lightCount = sc.lightSets[surf.lightSetIndex].count;
                lights = sc.lights.data() + sc.lightSets[surf.lightSetIndex].start;
            }
            if (lightCount > 0)
                withStream(rng, [&](Stream& st) { nextEventEstimation(st, lights, lightCount, v3(surf.position), geometricNormal(surf), ls); });
            V3 view = v3(-ray.dx, -ray.dy, -ray.dz);
            if (dot(v3(ls.radiance), v3(1.0f)) > 0) {
                float f = 0, r = 0;
                V3 refl = evaluateBsdf(surf, view, v3(ls.direction), f, r);
                if (dot(refl, v3(1.0f)) > 0) {
                    // OcclusionRay, PathTracer.cpp:71-92
                    V3 org = offsetRayOrigin(surf, v3(ls.direction), 0.1f);
                    float so[3] = {org.x, org.y, org.z};
                    RayHit sh;
                    ++shadow;
                    bool occ = sc.trace<float>(so, ls.direction, surf.error, ls.distance, true, sh);
                    if (!occ) {
                        V3 smp = refl * v3(ls.radiance) * (1.0f / ls.pdfW);
                        Ld[1] = Ld[1] + smp * throughput;
                    }
                }
            }
            SelasBsdfSample bs;
            bool ok;
            withStream(rng, [&](Stream& st) { ok = sampleBsdf(st, surf, view, bs); });
            if (!ok) break;
            isDeltaOnly = isDeltaOnly && (bs.flags & SELAS_EVENT_DIRAC);
            if (bs.flags == SELAS_EVENT_TRANSMISSION) {
                if (medium.phase == SELAS_MEDIUM_VACUUM)
                    medium = Medium{bs.mediumPhase, v3(bs.mediumExtinction)};
                else
                    medium = vacuum;
            }
            float weight = 1.0f;  // PathTracer.cpp:215 (MIS disabled in the reference)
            if (rs.misMode == 1) {
                float lp = lightingPdf(lights, lightCount, ls.index, v3(surf.position), v3(bs.wi));
                weight = bs.forwardPdfW / (bs.forwardPdfW + lp);  // ImportanceSampling.cpp:19-22
            }
            throughput = weight * throughput * v3(bs.reflectance);
            V3 org = offsetRayOrigin(surf, v3(bs.wi), 1.0f);
            ray.ox = org.x; ray.oy = org.y; ray.oz = org.z;
            ray.dx = bs.wi[0]; ray.dy = bs.wi[1]; ray.dz = bs.wi[2];
            if (bounce == 0) Ld[0] = v3(bs.reflectance);
        } else if (medium.phase != SELAS_MEDIUM_VACUUM) {
            V3 org = v3(ray.ox, ray.oy, ray.oz) + rayDistance * v3(ray.dx, ray.dy, ray.dz);
            V3 nd;
            withStream(rng, [&](Stream& st) { nd = uniformSphere(st); });
            ray.ox = org.x; ray.oy = org.y; ray.oz = org.z;
            ray.dx = nd.x; ray.dy = nd.y; ray.dz = nd.z;
        } else {
            V3 smp = background;  // EvaluateBackground / EvaluateBackgroundMiss, AreaLighting.cpp:415-437
            if (sc.hasIbl) {
                float pdf;
                V3 dirv = v3(ray.dx, ray.dy, ray.dz);
                smp = isDeltaOnly ? iblLookupMiss(sc.ibl, dirv, pdf) : iblLookup(sc.ibl, dirv, pdf);
            }
            Ld[1] = Ld[1] + smp * throughput;
            break;
        }
        ++bounce;
        if (bounce > rs.rrStartBounce) {
            float p = maxf(maxf(throughput.x, throughput.y), throughput.z);
            if (rng.next() > p) break;
            throughput = throughput * (1.0f / p);
        }
    }
}

// DeferredPathTracer: GeneratePrimaryRays + TraceRayBatch + ShadeHitPosition + TraceOcclusionBatch
// (Applications/Selas/Source/DeferredPathTracer.cpp:61-148, :151-229, :232-290, :349-380) for one path.  The reference queues rays
// and hits in batches (PathTracingBatcher); paths are independent, so walking one path to its end is the same computation.
static void evaluatePathDeferred(const Scene& sc, const SelasCamera& cam, const SelasRenderSettings& rs, uint32_t px, uint32_t py,
                                 uint32_t sample, V3& Ld, uint64_t& closest, uint64_t& shadow) {
    PathRng rng;
    rng.init(rs.seed, py * rs.width + px, sample);
    float vx, vy;
    if (rs.pixelSampler == 0) {
        vx = (float)px + rng.next();
        vy = (float)py + rng.next();
    } else {  // JitteredCameraRay(camera, x, y, scan, SamplesPerPixelX_, SamplesPerPixelY_, index), :365-366
        float j[2];
        cmj((int32_t)sample, (int32_t)(rs.pixelSampler & 0xffffu), (int32_t)(rs.pixelSampler >> 16), (int32_t)(py * rs.width + px), j);
        vx = (float)px + j[0];
        vy = (float)py + j[1];
    }
    SelasRay ray;
    V3 dir = imageToWorldDirection(cam, vx, vy);
    ray.ox = cam.position[0]; ray.oy = cam.position[1]; ray.oz = cam.position[2];
    ray.dx = dir.x; ray.dy = dir.y; ray.dz = dir.z;
    Ld = v3(0.0f);
    V3 throughput = v3(1.0f);
    bool diracScatterOnly = true;
    uint32_t trackedBounces = 0, bounce = 0;
    const uint32_t kMaxTracked = 7;  // MaxTrackedBounces_, IntegratorContexts.h:43
    auto occluded = [&](const SelasSurface& surf, V3 d, float distance) {
        V3 org = offsetRayOrigin(surf, d, 0.1f);
        float so[3] = {org.x, org.y, org.z}, sd[3] = {d.x, d.y, d.z};
        RayHit sh;
        ++shadow;
        return sc.trace<float>(so, sd, 0.0f, distance, true, sh);  // ray.tnear = 0, :262
    };
    while (bounce < rs.maxPathLength) {
        ray.tnear = 0.0f;
        ray.tfar = kFloatMax;
        RayHit rh;
        float o[3] = {ray.ox, ray.oy, ray.oz}, d[3] = {ray.dx, ray.dy, ray.dz};
        bool hit = sc.trace<float>(o, d, 0.0f, kFloatMax, false, rh);
        ++closest;
        if (!hit) {  // :198-207
            V3 smp = v3(sc.background);
            if (sc.hasIbl) {
                float pdf;
                V3 dirv = v3(ray.dx, ray.dy, ray.dz);
                smp = diracScatterOnly ? iblLookupMiss(sc.ibl, dirv, pdf) : iblLookup(sc.ibl, dirv, pdf);
            }
            Ld = Ld + smp * throughput;
            break;
        }
        SelasHit h;
        hitToRecord(sc, rh, ray, h, nullptr);
        SelasSurface surf;
        surfaceParams(sc, ray, h, surf);
        V3 view = v3(-ray.dx, -ray.dy, -ray.dz);
        // ---- area lights, :68-89 (weight 1) ----
        SelasLightSample ls;
        memset(&ls, 0, sizeof(ls));
        uint32_t lightCount = 0;
        const SelasLight* lights = nullptr;
        if (surf.lightSetIndex < sc.lightSets.size()) {
            lightCount = sc.lightSets[surf.lightSetIndex].count;
            lights = sc.lights.data() + sc.lightSets[surf.lightSetIndex].start;
        }
        if (lightCount > 0)
            withStream(rng, [&](Stream& st) { nextEventEstimation(st, lights, lightCount, v3(surf.position), geometricNormal(surf), ls); });
        if (dot(v3(ls.radiance), v3(1.0f)) > 0) {
            float f = 0, r = 0;
            V3 refl = evaluateBsdf(surf, view, v3(ls.direction), f, r);
            float weight = 1.0f;
            V3 smp = weight * refl * v3(ls.radiance) * (1.0f / ls.pdfW);
            if (dot(smp, v3(1.0f)) > 0 && !occluded(surf, v3(ls.direction), ls.distance)) Ld = Ld + smp * throughput;
        }
        // ---- sky, :91-111: SampleBackground (AreaLighting.cpp:392-400) + balance heuristic ----
        SelasLightSample sky;
        memset(&sky, 0, sizeof(sky));
        withStream(rng, [&](Stream& st) {
            if (sc.hasIbl) {
                directIblLightSample(st, sc.ibl, sky);
            } else {  // BackgroundLightSample, AreaLighting.cpp:334-350
                V3 dd = uniformSphere(st);
                sky.distance = 1e36f;
                store(sky.direction, dd);
                store(sky.radiance, v3(sc.background));
                sky.pdfW = 1.0f;
            }
        });
        if (dot(v3(sky.radiance), v3(1.0f)) > 0) {
            float f = 0, r = 0;
            V3 refl = evaluateBsdf(surf, view, v3(sky.direction), f, r);
            float misWeight = (1 * sky.pdfW) / (1 * sky.pdfW + 1 * f);  // BalanceHeuristic, ImportanceSampling.cpp:19-22
            V3 smp = misWeight * refl * v3(sky.radiance) * (1.0f / sky.pdfW);
            if (dot(smp, v3(1.0f)) > 0 && !occluded(surf, v3(sky.direction), sky.distance)) Ld = Ld + smp * throughput;
        }
        // ---- BSDF sample, :113-146 ----
        SelasBsdfSample bs;
        bool ok;
        withStream(rng, [&](Stream& st) { ok = sampleBsdf(st, surf, view, bs); });
        if (!ok) break;
        float skyPdfW = sc.hasIbl ? iblPdf(sc.ibl, v3(bs.wi)) : kInv4Pi;  // BackgroundLightingPdf, AreaLighting.cpp:403-411
        float misWeight = (1 * bs.forwardPdfW) / (1 * bs.forwardPdfW + 1 * skyPdfW);
        throughput = misWeight * throughput * v3(bs.reflectance);
        if (dot(throughput, throughput) == 0.0f) break;
        if (trackedBounces >= kMaxTracked) {
            float p = maxf(maxf(throughput.x, throughput.y), throughput.z);
            if (rng.next() >= p) break;
            throughput = throughput * (1.0f / p);
        }
        V3 org = offsetRayOrigin(surf, v3(bs.wi), 1.0f);
        ray.ox = org.x; ray.oy = org.y; ray.oz = org.z;
        ray.dx = bs.wi[0]; ray.dy = bs.wi[1]; ray.dz = bs.wi[2];
        diracScatterOnly = diracScatterOnly && (bs.flags & SELAS_EVENT_DIRAC);
        trackedBounces = trackedBounces + 1 < kMaxTracked ? trackedBounces + 1 : kMaxTracked;
        ++bounce;
    }
}

extern "C" {

// Attach a Ptex lookup to mesh `meshIndex` (used when its material has SELAS_MAT_USES_PTEX).
int orc_scene_set_ptex(void* scene, uint32_t meshIndex, void (*fn)(void*, int, float, float, float*), void* user) {
    Scene* sc = (Scene*)scene;
    if (!sc || meshIndex >= sc->meshes.size()) return 1;
    sc->meshes[meshIndex].ptexEval = fn;
    sc->meshes[meshIndex].ptexUser = user;
    return 0;
}

int orc_scene_create(const SelasSceneDesc* d, void** out) {
    Scene* sc = new Scene();
    sc->meshes.resize(d->meshCount);
    for (uint32_t m = 0; m < d->meshCount; ++m) {
        const SelasMeshDesc& md = d->meshes[m];
        Mesh& M = sc->meshes[m];
        M.positions.assign(md.positions, md.positions + 3 * (size_t)md.vertexCount);
        if (md.normals) M.normals.assign(md.normals, md.normals + 3 * (size_t)md.vertexCount);
        if (md.tangents) M.tangents.assign(md.tangents, md.tangents + 4 * (size_t)md.vertexCount);
        if (md.uvs) M.uvs.assign(md.uvs, md.uvs + 2 * (size_t)md.vertexCount);
        if (md.baseColorTexture) {
            const SelasTextureDesc& t = *md.baseColorTexture;
            M.texels.assign(t.texels, t.texels + (size_t)t.width * t.height * (t.format + 1));
            M.texture.width = t.width;
            M.texture.height = t.height;
            M.texture.format = t.format;
        }
        M.indices.assign(md.indices, md.indices + (size_t)md.faceCount * md.indicesPerFace);
        M.faceCount = md.faceCount;
        M.indicesPerFace = md.indicesPerFace;
        M.materialIndex = md.materialIndex;
        M.lightSetIndex = md.lightSetIndex;
    }
    for (uint32_t c = 0; c < d->curveCount; ++c) {  // curve geometries: geomID = meshCount + c
        const SelasCurveDesc& cd = d->curves[c];
        Mesh M;
        M.positions.assign(cd.vertices, cd.vertices + 4 * (size_t)cd.vertexCount);
        M.indices.assign(cd.indices, cd.indices + cd.segmentCount);
        M.faceCount = cd.segmentCount;
        M.indicesPerFace = 1;
        M.materialIndex = cd.materialIndex;
        M.lightSetIndex = cd.lightSetIndex;
        sc->meshes.push_back(M);
    }
    if (d->instanceCount) {
        for (uint32_t i = 0; i < d->instanceCount; ++i) {
            Instance in;
            in.mesh = d->instances[i].meshIndex;
            memcpy(in.m, d->instances[i].localToWorld, sizeof(in.m));
            sc->instances.push_back(in);
        }
    } else {
        Instance in;
        in.mesh = 0xFFFFFFFFu;  // all meshes
        memset(in.m, 0, sizeof(in.m));
        in.m[0] = in.m[5] = in.m[10] = in.m[15] = 1.0f;
        sc->instances.push_back(in);
    }
    sc->materials.assign(d->materials, d->materials + d->materialCount);
    sc->lights.assign(d->lights, d->lights + d->lightCount);
    sc->lightSets.assign(d->lightSets, d->lightSets + d->lightSetCount);
    memcpy(sc->background, d->backgroundIntensity, sizeof(sc->background));
    if (d->ibl) {
        const SelasIblDesc& I = *d->ibl;
        sc->iblMarginal.assign(I.marginalDensityFunction, I.marginalDensityFunction + I.height);
        sc->iblConditional.assign(I.conditionalDensityFunctions, I.conditionalDensityFunctions + (size_t)I.width * I.height);
        sc->iblLight.assign(I.lightData, I.lightData + 3 * (size_t)I.width * I.height);
        sc->iblMiss.assign(I.missData, I.missData + 3 * (size_t)I.missWidth * I.missHeight);
        sc->ibl.width = I.width; sc->ibl.height = I.height; sc->ibl.missWidth = I.missWidth; sc->ibl.missHeight = I.missHeight;
        sc->ibl.marginal = sc->iblMarginal.data(); sc->ibl.conditional = sc->iblConditional.data();
        sc->ibl.light = sc->iblLight.data(); sc->ibl.miss = sc->iblMiss.data();
        sc->ibl.rotation = I.rotationRadians; sc->ibl.exposureScale = I.exposureScale;
        sc->hasIbl = true;
    }
    for (uint32_t ii = 0; ii < sc->instances.size(); ++ii) {
        const Instance& in = sc->instances[ii];
        for (uint32_t m = 0; m < sc->meshes.size(); ++m) {
            if (in.mesh != 0xFFFFFFFFu && in.mesh != m) continue;
            const Mesh& M = sc->meshes[m];
            bool identity = in.mesh == 0xFFFFFFFFu;
            if (M.indicesPerFace == 1) {  // curve geometry: world-space control vertices, radius scaled by the transform's cube-root volume change
                const double det = identity ? 1.0
                                            : (double)in.m[0] * ((double)in.m[5] * in.m[10] - (double)in.m[6] * in.m[9]) -
                                                  (double)in.m[1] * ((double)in.m[4] * in.m[10] - (double)in.m[6] * in.m[8]) +
                                                  (double)in.m[2] * ((double)in.m[4] * in.m[9] - (double)in.m[5] * in.m[8]);
                const double rscale = std::cbrt(std::fabs(det));
                for (uint32_t f = 0; f < M.faceCount; ++f) {
                    orc::CurveSeg S;
                    for (int k = 0; k < 4; ++k) {
                        const float* cv = &M.positions[4 * (size_t)(M.indices[f] + k)];
                        double p[3] = {cv[0], cv[1], cv[2]};
                        if (!identity) {
                            const double x = p[0], y = p[1], z = p[2];
                            for (int a = 0; a < 3; ++a) p[a] = x * in.m[a] + y * in.m[4 + a] + z * in.m[8 + a] + in.m[12 + a];
                        }
                        S.c[k][0] = p[0]; S.c[k][1] = p[1]; S.c[k][2] = p[2]; S.c[k][3] = cv[3] * rscale;
                    }
                    S.primID = f; S.geomID = m; S.instID = ii;
                    orc::curveBounds(S);
                    sc->curves.push_back(S);
                }
                continue;
            }
            auto P = [&](uint32_t vi) {
                V3 p = v3(&M.positions[3 * (size_t)vi]);
                return identity ? p : mulPoint44(p, in.m);
            };
            for (uint32_t f = 0; f < M.faceCount; ++f) {
                const uint32_t* ix = &M.indices[(size_t)f * M.indicesPerFace];
                Tri T;
                T.primID = f;
                T.geomID = m;
                T.instID = ii;
                T.second = 0;
                if (M.indicesPerFace == 3) {
                    store(T.v0, P(ix[0])); store(T.v1, P(ix[1])); store(T.v2, P(ix[2]));
                    sc->bvh.tris.push_back(T);
                } else {
                    store(T.v0, P(ix[0])); store(T.v1, P(ix[1])); store(T.v2, P(ix[3]));
                    sc->bvh.tris.push_back(T);
                    T.second = 1;
                    store(T.v0, P(ix[2])); store(T.v1, P(ix[3])); store(T.v2, P(ix[1]));
                    sc->bvh.tris.push_back(T);
                }
            }
        }
    }
    for (Mesh& M : sc->meshes)  // the meshes vector no longer moves: point the textures at their texel copies
        if (!M.texels.empty()) M.texture.texels = M.texels.data();
    sc->bvh.build();
    *out = sc;
    return 0;
}
int orc_texture_sample(const SelasTextureDesc* t, size_t n, const float* uv2, int sRGB, float* rgb3) {
    setFtz();
    Texture tex;
    tex.texels = t->texels; tex.width = t->width; tex.height = t->height; tex.format = t->format;
    for (size_t i = 0; i < n; ++i) store(rgb3 + 3 * i, sampleTextureFloat3(tex, uv2[2 * i], uv2[2 * i + 1], sRGB != 0));
    return 0;
}

// ---- TextureFiltering::Point / Triangle / Trilinear / EWA, TextureLib/TextureFiltering.h:54-234, restated over n components ----------
namespace {
struct MipTex {
    const SelasTextureMipsDesc* d;
    int comps;
    // Sample with WrapMode::Repeat (:35-52); `s % w` is an unsigned modulo there (int32 % uint32)
    const float* at(uint32_t level, uint32_t w, uint32_t h, int32_t s, int32_t t) const {
        const uint32_t ss = (uint32_t)s % w, ts = (uint32_t)t % h;
        return reinterpret_cast<const float*>(static_cast<const uint8_t*>(d->texels) + d->mipOffsets[level]) + ((size_t)ts * w + ss) * comps;
    }
};
float gEwaLut[128];
bool gEwaLutReady = false;
void filterTriangle(const MipTex& t, int32_t levelIn, float sx, float sy, float* r) {  // :77-104
    const uint32_t level = std::min<uint32_t>((uint32_t)levelIn, t.d->mipCount - 1u);
    const uint32_t w = t.d->mipWidths[level], h = t.d->mipHeights[level];
    const float s = sx * w - 0.5f, tt = sy * h - 0.5f;
    const int32_t s0 = (int32_t)floorf(s), t0 = (int32_t)floorf(tt);
    const float ds = s - s0, dt = tt - t0;
    const float w00 = (1 - ds) * (1 - dt), w01 = (1 - ds) * dt, w10 = ds * (1 - dt), w11 = ds * dt;
    const float *a = t.at(level, w, h, s0, t0), *b = t.at(level, w, h, s0, t0 + 1), *c = t.at(level, w, h, s0 + 1, t0), *e = t.at(level, w, h, s0 + 1, t0 + 1);
    for (int k = 0; k < t.comps; ++k) r[k] = w00 * a[k] + w01 * b[k] + w10 * c[k] + w11 * e[k];
}
void filterEwaLevel(const MipTex& t, int32_t reqLevel, float sx, float sy, float d0x, float d0y, float d1x, float d1y, float* r) {  // :107-170
    if (reqLevel >= (int32_t)t.d->mipCount) {
        const float* p = t.at(t.d->mipCount - 1u, 1u, 1u, 0, 0);
        for (int k = 0; k < t.comps; ++k) r[k] = p[k];
        return;
    }
    const uint32_t level = (uint32_t)std::max(reqLevel, 0);
    const uint32_t w = t.d->mipWidths[level], h = t.d->mipHeights[level];
    sx = sx * w - 0.5f;
    sy = sy * h - 0.5f;
    d0x *= w; d0y *= h; d1x *= w; d1y *= h;
    float A = d0y * d0y + d1y * d1y + 1;
    float B = -2 * (d0x * d0y + d1x * d1y);
    float C = d0x * d0x + d1x * d1x + 1;
    const float invF = 1 / (A * C - B * B * 0.25f);
    A *= invF; B *= invF; C *= invF;
    const float det = -B * B + 4 * A * C;
    const float invDet = 1 / det;
    const float uSqrt = sqrtf(det * C), vSqrt = sqrtf(A * det);
    const int32_t s0 = (int32_t)ceilf(sx - 2 * invDet * uSqrt), s1 = (int32_t)floorf(sx + 2 * invDet * uSqrt);
    const int32_t t0 = (int32_t)ceilf(sy - 2 * invDet * vSqrt), t1 = (int32_t)floorf(sy + 2 * invDet * vSqrt);
    float sum[4] = {0.f, 0.f, 0.f, 0.f};
    float sumWts = 0;
    for (int32_t it = t0; it <= t1; ++it) {
        const float tt = it - sy;
        for (int32_t is = s0; is <= s1; ++is) {
            const float ss = is - sx;
            const float r2 = A * ss * ss + B * ss * tt + C * tt * tt;
            if (r2 < 1) {
                const int32_t index = std::min<int32_t>((int32_t)(r2 * 128), 127);
                const float weight = gEwaLut[index];
                const float* x = t.at(level, w, h, is, it);
                for (int k = 0; k < t.comps; ++k) sum[k] += x[k] * weight;
                sumWts += weight;
            }
        }
    }
    const float inv = 1.0f / sumWts;
    for (int k = 0; k < t.comps; ++k) r[k] = sum[k] * inv;
}
void lerpN(const float* a, const float* b, float x, int n, float* r) {  // Lerp, FloatFuncs.h:369-387
    for (int k = 0; k < n; ++k) r[k] = (1.0f - x) * a[k] + x * b[k];
}
}  // namespace
int orc_texture_filter(const SelasTextureMipsDesc* d, int filter, int level, size_t n, const float* st2, const float* dst0, const float* dst1, float* out) {
    setFtz();
    if (!gEwaLutReady) {  // InitializeEWAFilterWeights, TextureFiltering.cpp:14-21
        for (int i = 0; i < 128; ++i) {
            const float alpha = 2.0f;
            const float r2 = float(i) / (128 - 1.0f);
            gEwaLut[i] = expf(-alpha * r2) - expf(-alpha);
        }
        gEwaLutReady = true;
    }
    MipTex t{d, (int)d->format + 1};
    for (size_t i = 0; i < n; ++i) {
        const float sx = st2[2 * i], sy = st2[2 * i + 1];
        float d0x = dst0 ? dst0[2 * i] : 0.f, d0y = dst0 ? dst0[2 * i + 1] : 0.f, d1x = dst1 ? dst1[2 * i] : 0.f, d1y = dst1 ? dst1[2 * i + 1] : 0.f;
        float* r = out + i * t.comps;
        float r0[4], r1[4];
        if (filter == SELAS_FILTER_POINT) {  // :55-74
            const uint32_t w = d->mipWidths[0], h = d->mipHeights[0];
            const float* p = t.at(0, w, h, (int32_t)floorf(sx * w), (int32_t)floorf(sy * h));
            for (int k = 0; k < t.comps; ++k) r[k] = p[k];
        } else if (filter == SELAS_FILTER_TRIANGLE) {
            filterTriangle(t, level, sx, sy, r);
        } else if (filter == SELAS_FILTER_TRILINEAR) {  // :173-194
            const float majorLength = sqrtf(d0x * d0x + d0y * d0y), minorLength = sqrtf(d1x * d1x + d1y * d1y);
            const float length = std::min(majorLength, minorLength);
            if (length == 0) { filterTriangle(t, 0, sx, sy, r); continue; }
            const float lod = std::max(0.0f, d->mipCount - 1.0f + log2f(length));
            const float ilod = floorf(lod);
            filterTriangle(t, (int32_t)ilod, sx, sy, r0);
            filterTriangle(t, (int32_t)ilod + 1, sx, sy, r1);
            lerpN(r0, r1, lod - ilod, t.comps, r);
        } else if (filter == SELAS_FILTER_EWA) {  // :197-234
            if (d0x * d0x + d0y * d0y < d1x * d1x + d1y * d1y) { std::swap(d0x, d1x); std::swap(d0y, d1y); }
            const float majorLength = sqrtf(d0x * d0x + d0y * d0y);
            float minorLength = sqrtf(d1x * d1x + d1y * d1y);
            const float maxAnisotropy = 16.0f;
            if (minorLength * maxAnisotropy < majorLength && minorLength > 0) {
                const float scale = majorLength / (minorLength * maxAnisotropy);
                d1x = d1x * scale; d1y = d1y * scale;
                minorLength *= scale;
            }
            if (minorLength == 0) { filterTriangle(t, 0, sx, sy, r); continue; }
            const float lod = std::max(0.0f, d->mipCount - 1.0f + log2f(minorLength));
            const float ilod = floorf(lod);
            filterEwaLevel(t, (int32_t)ilod, sx, sy, d0x, d0y, d1x, d1y, r0);
            filterEwaLevel(t, (int32_t)ilod + 1, sx, sy, d0x, d0y, d1x, d1y, r1);
            lerpN(r0, r1, lod - ilod, t.comps, r);
        } else {
            filterEwaLevel(t, level, sx, sy, d0x, d0y, d1x, d1y, r);
        }
    }
    return 0;
}

int orc_scene_destroy(void* s) {
    delete (Scene*)s;
    return 0;
}
size_t orc_scene_triangle_count(void* s) { return ((Scene*)s)->bvh.tris.size(); }

// Tree-quality tooling (scripts/tree_quality.py): one triangle soup under three builders -- builder 0 = binned SAH (the oracle's own),
// 1 = LBVH with per-axis Morton normalisation, 2 = LBVH on cubic cells (the product's device build), 3 = cubic Morton order with
// SAH-swept split positions, >= 16 = binned SAH for the top of the tree and cubic LBVH for ranges of at most that many triangles, < 0 = PLOC with that search radius -- with the SAH cost of each tree and
// the box / triangle tests a set of rays needs (closest hit, near child first).  out4 = {inner SAH, leaf SAH, box tests, tri tests} per ray.
int orc_tree_quality(size_t ntri, const float* tri9, int builder, size_t nrays, const SelasRay* rays, int nthreads, double* out4) {
    Bvh bvh;
    bvh.tris.resize(ntri);
    for (size_t i = 0; i < ntri; ++i) {
        memset(&bvh.tris[i], 0, sizeof(Tri));
        for (int k = 0; k < 3; ++k) {
            bvh.tris[i].v0[k] = tri9[9 * i + k];
            bvh.tris[i].v1[k] = tri9[9 * i + 3 + k];
            bvh.tris[i].v2[k] = tri9[9 * i + 6 + k];
        }
    }
    if (builder == 0) bvh.build();
    else if (builder < 0) bvh.buildPloc(-builder);      // PLOC with search radius -builder over the cubic Morton order
    else if (builder >= (1 << 24)) bvh.build(0, (uint32_t)(builder - (1 << 24)));  // cubic LBVH above ranges of that many triangles, binned SAH inside
    else if (builder >= 16) bvh.build((uint32_t)builder);  // binned SAH above ranges of `builder` triangles, cubic LBVH below
    else bvh.buildLbvh(builder >= 2, builder == 3);
    bvh.sahCost(out4[0], out4[1]);
    std::vector<uint64_t> boxes(nthreads > 0 ? nthreads : 1, 0), tris(nthreads > 0 ? nthreads : 1, 0);
    parallelFor(nrays, nthreads, [&](size_t b, size_t e, int t) {
        uint64_t bx = 0, tr = 0;
        for (size_t i = b; i < e; ++i) {
            const SelasRay& r = rays[i];
            float o[3] = {r.ox, r.oy, r.oz}, d[3] = {r.dx, r.dy, r.dz};
            bvh.traceCount(o, d, r.tfar, bx, tr);
        }
        boxes[t] += bx;
        tris[t] += tr;
    });
    uint64_t bx = 0, tr = 0;
    for (size_t t = 0; t < boxes.size(); ++t) { bx += boxes[t]; tr += tris[t]; }
    out4[2] = nrays ? (double)bx / (double)nrays : 0.0;
    out4[3] = nrays ? (double)tr / (double)nrays : 0.0;
    return 0;
}

// mode: 0 = float BVH (timed CPU baseline), 1 = double BVH (hit-ID oracle), 2 = double brute force
int orc_intersect(void* s, size_t n, const SelasRay* rays, SelasHit* hits, float* Ng4, int mode, int nthreads) {
    const Scene& sc = *(Scene*)s;
    parallelFor(n, nthreads, [&](size_t b, size_t e, int) {
        for (size_t i = b; i < e; ++i) {
            const SelasRay& r = rays[i];
            float o[3] = {r.ox, r.oy, r.oz}, d[3] = {r.dx, r.dy, r.dz};
            RayHit h;
            if (mode == 0) sc.trace<float>(o, d, r.tnear, r.tfar, false, h);
            else if (mode == 1) sc.trace<double>(o, d, r.tnear, r.tfar, false, h);
            else sc.trace<double>(o, d, r.tnear, r.tfar, false, h, true);
            hitToRecord(sc, h, r, hits[i], Ng4 ? Ng4 + 4 * i : nullptr);
        }
    });
    return 0;
}
int orc_occluded(void* s, size_t n, const SelasRay* rays, uint8_t* occluded, int mode, int nthreads) {
    const Scene& sc = *(Scene*)s;
    parallelFor(n, nthreads, [&](size_t b, size_t e, int) {
        for (size_t i = b; i < e; ++i) {
            const SelasRay& r = rays[i];
            float o[3] = {r.ox, r.oy, r.oz}, d[3] = {r.dx, r.dy, r.dz};
            RayHit h;
            bool hit;
            if (mode == 0) hit = sc.trace<float>(o, d, r.tnear, r.tfar, true, h);
            else if (mode == 1) hit = sc.trace<double>(o, d, r.tnear, r.tfar, true, h);
            else hit = sc.trace<double>(o, d, r.tnear, r.tfar, true, h, true);
            occluded[i] = hit ? 1 : 0;
        }
    });
    return 0;
}
// For a GPU-reported hit (geomID, primID, instID0): the exact double-precision t of the ray against that
// primitive (both halves for quads), or -1 when the ray misses it.  Used for the "ties count as agreement
// when both candidates' t differ by < 1e-5 relative" rule (hard parts).
int orc_prim_distance(void* s, size_t n, const SelasRay* rays, const SelasHit* hits, double* t_out) {
    const Scene& sc = *(Scene*)s;
    for (size_t i = 0; i < n; ++i) {
        t_out[i] = -1.0;
        const SelasHit& h = hits[i];
        if (h.geomID == SELAS_INVALID_ID) continue;
        const Mesh& M = sc.meshes[h.geomID];
        const Instance& in = sc.instances[h.instID0];
        bool identity = in.mesh == 0xFFFFFFFFu;
        const uint32_t* ix = &M.indices[(size_t)h.primID * M.indicesPerFace];
        auto P = [&](uint32_t vi) {
            V3 p = v3(&M.positions[3 * (size_t)vi]);
            return identity ? p : mulPoint44(p, in.m);
        };
        double o[3] = {rays[i].ox, rays[i].oy, rays[i].oz}, d[3] = {rays[i].dx, rays[i].dy, rays[i].dz};
        int halves = M.indicesPerFace == 3 ? 1 : 2;
        for (int k = 0; k < halves; ++k) {
            Tri T;
            if (M.indicesPerFace == 3) { store(T.v0, P(ix[0])); store(T.v1, P(ix[1])); store(T.v2, P(ix[2])); }
            else if (k == 0) { store(T.v0, P(ix[0])); store(T.v1, P(ix[1])); store(T.v2, P(ix[3])); }
            else { store(T.v0, P(ix[2])); store(T.v1, P(ix[3])); store(T.v2, P(ix[1])); }
            // infinite-plane distance with a slack barycentric test: edge hits are what we are adjudicating
            double e1[3], e2[3], p[3], sv[3], q[3];
            for (int c = 0; c < 3; ++c) { e1[c] = (double)T.v1[c] - T.v0[c]; e2[c] = (double)T.v2[c] - T.v0[c]; sv[c] = o[c] - T.v0[c]; }
            p[0] = d[1] * e2[2] - d[2] * e2[1]; p[1] = d[2] * e2[0] - d[0] * e2[2]; p[2] = d[0] * e2[1] - d[1] * e2[0];
            double det = e1[0] * p[0] + e1[1] * p[1] + e1[2] * p[2];
            if (det == 0) continue;
            double uu = (sv[0] * p[0] + sv[1] * p[1] + sv[2] * p[2]) / det;
            q[0] = sv[1] * e1[2] - sv[2] * e1[1]; q[1] = sv[2] * e1[0] - sv[0] * e1[2]; q[2] = sv[0] * e1[1] - sv[1] * e1[0];
            double vv = (d[0] * q[0] + d[1] * q[1] + d[2] * q[2]) / det;
            double tt = (e2[0] * q[0] + e2[1] * q[1] + e2[2] * q[2]) / det;
            const double slack = 1e-4;
            if (uu >= -slack && vv >= -slack && uu + vv <= 1 + slack) {
                if (t_out[i] < 0 || std::fabs(tt - h.t) < std::fabs(t_out[i] - h.t)) t_out[i] = tt;
            }
        }
    }
    return 0;
}

int orc_camera_init(const SelasCameraSettings* st, uint32_t w, uint32_t h, SelasCamera* out) {
    initCamera(*st, w, h, *out);
    return 0;
}
int orc_camera_rays(const SelasCamera* cam, size_t n, const float* px, const float* py, const float* jitter, SelasRay* rays) {
    for (size_t i = 0; i < n; ++i) {
        float jx = jitter ? jitter[2 * i] : 0.5f, jy = jitter ? jitter[2 * i + 1] : 0.5f;
        V3 d = imageToWorldDirection(*cam, px[i] + jx, py[i] + jy);
        rays[i].ox = cam->position[0]; rays[i].oy = cam->position[1]; rays[i].oz = cam->position[2];
        rays[i].tnear = 0.0f;
        rays[i].dx = d.x; rays[i].dy = d.y; rays[i].dz = d.z;
        rays[i].tfar = kFloatMax;
    }
    return 0;
}
int orc_bsdf_eval(size_t n, const SelasSurface* s, const float* v, const float* l, float* rgb, float* fwd, float* rev) {
    setFtz();
    for (size_t i = 0; i < n; ++i) {
        float f = 0.f, r = 0.f;
        V3 c = evaluateBsdf(s[i], v3(v + 3 * i), v3(l + 3 * i), f, r);
        store(rgb + 3 * i, c);
        fwd[i] = f;
        rev[i] = r;
    }
    return 0;
}
int orc_bsdf_sample(size_t n, const SelasSurface* s, const float* v, const float* u, SelasBsdfSample* out) {
    setFtz();
    for (size_t i = 0; i < n; ++i) {
        Stream st{u + SELAS_BSDF_RANDOMS * i, 0};
        sampleBsdf(st, s[i], v3(v + 3 * i), out[i]);
    }
    return 0;
}
int orc_light_sample(size_t n, const SelasLight* lights, uint32_t count, const float* pos, const float* nrm, const float* u3,
                     SelasLightSample* out) {
    setFtz();
    for (size_t i = 0; i < n; ++i) {
        Stream st{u3 + 3 * i, 0};
        nextEventEstimation(st, lights, count, v3(pos + 3 * i), v3(nrm + 3 * i), out[i]);
    }
    return 0;
}
int orc_light_pdf(size_t n, const SelasLight* lights, uint32_t count, const uint32_t* index, const float* pos, const float* wi,
                  float* pdf) {
    setFtz();
    for (size_t i = 0; i < n; ++i) pdf[i] = lightingPdf(lights, count, index[i], v3(pos + 3 * i), v3(wi + 3 * i));
    return 0;
}
int orc_surface_params(void* s, size_t n, const SelasRay* rays, const SelasHit* hits, SelasSurface* out) {
    setFtz();
    const Scene& sc = *(Scene*)s;
    for (size_t i = 0; i < n; ++i) {
        memset(&out[i], 0, sizeof(SelasSurface));
        if (hits[i].geomID != SELAS_INVALID_ID) surfaceParams(sc, rays[i], hits[i], out[i]);
    }
    return 0;
}
int orc_ibl_lookup(void* s, size_t n, const float* wi3, float* rgb3, float* pdf, float* miss3) {
    setFtz();
    const Scene& sc = *(Scene*)s;
    if (!sc.hasIbl) return 2;
    for (size_t i = 0; i < n; ++i) {
        float p = 0.f;
        V3 c = iblLookup(sc.ibl, v3(wi3 + 3 * i), p);
        if (rgb3) store(rgb3 + 3 * i, c);
        if (pdf) pdf[i] = p;
        if (miss3) store(miss3 + 3 * i, iblLookupMiss(sc.ibl, v3(wi3 + 3 * i), p));
    }
    return 0;
}
int orc_ibl_sample(void* s, size_t n, const float* u2, float* thetaPhiPdf3, uint32_t* xy2) {
    setFtz();
    const Scene& sc = *(Scene*)s;
    if (!sc.hasIbl) return 2;
    for (size_t i = 0; i < n; ++i) {
        float theta, phi, pdf;
        uint32_t x, y;
        iblSample(sc.ibl, u2[2 * i], u2[2 * i + 1], theta, phi, x, y, pdf);
        thetaPhiPdf3[3 * i] = theta; thetaPhiPdf3[3 * i + 1] = phi; thetaPhiPdf3[3 * i + 2] = pdf;
        xy2[2 * i] = x; xy2[2 * i + 1] = y;
    }
    return 0;
}
int orc_cmj(size_t n, const int32_t* s, int32_t m, int32_t nn, const int32_t* p, float* out2) {
    for (size_t i = 0; i < n; ++i) cmj(s[i], m, nn, p[i], out2 + 2 * i);
    return 0;
}
// u: 2 uniforms per item; out distance; and transmission for a given distance
int orc_medium(size_t n, const float* extinction3, const float* u2, const float* dist, float* sampled, float* trans3) {
    for (size_t i = 0; i < n; ++i) {
        Medium m{SELAS_MEDIUM_ISOTROPIC, v3(extinction3 + 3 * i)};
        Stream st{u2 + 2 * i, 0};
        sampled[i] = sampleDistance(st, m);
        store(trans3 + 3 * i, transmission(m, dist[i]));
    }
    return 0;
}

// PathTracer::GenerateImage (PathTracer.cpp:301-350) minus the final scale and .hdr save.
int orc_render_pathtracer(void* s, const SelasCamera* cam, const SelasRenderSettings* rs, float* const* film,
                          SelasRenderStats* stats, int nthreads) {
    const Scene& sc = *(Scene*)s;
    size_t pixels = (size_t)rs->width * rs->height;
    for (uint32_t l = 0; l < rs->layerCount; ++l) memset(film[l], 0, pixels * 3 * sizeof(float));
    std::atomic<uint64_t> closest(0), shadow(0);
    const size_t stride = rs->pixelStride > 1 ? rs->pixelStride : 1;
    parallelFor(pixels, nthreads, [&](size_t b, size_t e, int) {
        uint64_t c = 0, sh = 0;
        for (size_t p = b; p < e; ++p) {
            if (stride > 1 && (p % stride) != 0) continue;  // bounded sample of a workload (SelasRenderSettings.pixelStride)
            uint32_t py = (uint32_t)(p / rs->width), px = (uint32_t)(p - (size_t)py * rs->width);
            for (uint32_t k = 0; k < rs->samplesPerPixel; ++k) {
                V3 Ld[2];
                if (rs->integrator == SELAS_INTEGRATOR_DEFERRED) {
                    Ld[0] = v3(0.0f);
                    evaluatePathDeferred(sc, *cam, *rs, px, py, rs->firstSample + k, Ld[1], c, sh);
                } else
                evaluatePath(sc, *cam, *rs, px, py, rs->firstSample + k, Ld, c, sh);
                if (rs->layerCount == 1) {
                    film[0][3 * p] += Ld[1].x; film[0][3 * p + 1] += Ld[1].y; film[0][3 * p + 2] += Ld[1].z;
                } else {
                    for (uint32_t l = 0; l < 2; ++l) {
                        film[l][3 * p] += Ld[l].x; film[l][3 * p + 1] += Ld[l].y; film[l][3 * p + 2] += Ld[l].z;
                    }
                }
            }
        }
        closest += c;
        shadow += sh;
    });
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->closestRays = closest;
        stats->occlusionRays = shadow;
        stats->pathSamples = ((pixels + stride - 1) / stride) * rs->samplesPerPixel;
    }
    return 0;
}

// ---- VCM hash grid: VCMHashGrid.cpp ---------------------------------------------------------------------------
// The reference's `uint` is 64 bits wide (SystemLib/BasicTypes.h:47), so the int32 cell coordinates are
// sign-extended and the products / xor / modulo are evaluated in 64-bit arithmetic.
static inline uint32_t cellHash(int32_t x, int32_t y, int32_t z, uint32_t cellCount) {  // :15-22
    uint64_t ux = (uint64_t)(int64_t)x, uy = (uint64_t)(int64_t)y, uz = (uint64_t)(int64_t)z;
    return (uint32_t)(((ux * 73856093ull) ^ (uy * 19349663ull) ^ (uz * 83492791ull)) % (uint64_t)cellCount);
}
int orc_hashgrid_build(size_t n, const float* pos, float radius, uint32_t cellCount, uint32_t* cellRangeEnds,
                       uint32_t* cellIndices, float* aabb6) {  // :48-97
    float cellSize = 2.0f * radius;
    float inv = 1.0f / cellSize;
    float lo[3] = {kFloatMax, kFloatMax, kFloatMax}, hi[3] = {-kFloatMax, -kFloatMax, -kFloatMax};
    for (size_t i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) {
            lo[k] = minf(lo[k], pos[3 * i + k]);
            hi[k] = maxf(hi[k], pos[3 * i + k]);
        }
    memset(cellRangeEnds, 0, sizeof(uint32_t) * cellCount);
    auto cellOf = [&](size_t i) {  // :25-34
        int32_t x = (int32_t)floorf(inv * (pos[3 * i] - lo[0]));
        int32_t y = (int32_t)floorf(inv * (pos[3 * i + 1] - lo[1]));
        int32_t z = (int32_t)floorf(inv * (pos[3 * i + 2] - lo[2]));
        return cellHash(x, y, z, cellCount);
    };
    for (size_t i = 0; i < n; ++i) ++cellRangeEnds[cellOf(i)];
    uint32_t sum = 0;
    for (uint32_t c = 0; c < cellCount; ++c) {
        uint32_t k = cellRangeEnds[c];
        cellRangeEnds[c] = sum;
        sum += k;
    }
    for (size_t i = 0; i < n; ++i) cellIndices[cellRangeEnds[cellOf(i)]++] = (uint32_t)i;
    for (int k = 0; k < 3; ++k) {
        aabb6[k] = lo[k];
        aabb6[3 + k] = hi[k];
    }
    return 0;
}
int orc_hashgrid_query(size_t n, const float* pos, float radius, uint32_t cellCount, const uint32_t* cellRangeEnds,
                       const uint32_t* cellIndices, const float* aabb6, size_t m, const float* q, uint32_t* counts,
                       uint32_t* indices, uint32_t maxPerQuery) {  // :107-155
    (void)n;
    float inv = 1.0f / (2.0f * radius);
    float r2 = radius * radius;
    for (size_t i = 0; i < m; ++i) {
        counts[i] = 0;
        float dmin[3], dmax[3];
        bool outside = false;
        for (int k = 0; k < 3; ++k) {
            dmin[k] = q[3 * i + k] - aabb6[k];
            dmax[k] = aabb6[3 + k] - q[3 * i + k];
            if (dmin[k] < 0.0f || dmax[k] < 0.0f) outside = true;
        }
        if (outside) continue;
        int32_t base[3];
        for (int k = 0; k < 3; ++k) {
            float cp = inv * dmin[k];
            float fl = floorf(cp);
            base[k] = (int32_t)fl;
            if (cp - fl < 0.5f) --base[k];
        }
        for (int z = 0; z <= 1; ++z)
            for (int y = 0; y <= 1; ++y)
                for (int x = 0; x <= 1; ++x) {
                    uint32_t c = cellHash(base[0] + x, base[1] + y, base[2] + z, cellCount);
                    uint32_t b = c == 0 ? 0 : cellRangeEnds[c - 1], e = cellRangeEnds[c];
                    for (uint32_t sIdx = b; sIdx < e; ++sIdx) {
                        uint32_t pi = cellIndices[sIdx];
                        float dx = q[3 * i] - pos[3 * pi], dy = q[3 * i + 1] - pos[3 * pi + 1], dz = q[3 * i + 2] - pos[3 * pi + 2];
                        float d2 = (dx * dx) + (dy * dy) + (dz * dz);
                        if (d2 <= r2) {
                            if (indices && counts[i] < maxPerQuery) indices[(size_t)i * maxPerQuery + counts[i]] = pi;
                            ++counts[i];
                        }
                    }
                }
    }
    return 0;
}

// product RNG, exposed so tests can pin the device implementation bit for bit
int orc_rng(uint32_t seed, uint32_t pixel, uint32_t sample, size_t n, float* out) {
    PathRng r;
    r.init(seed, pixel, sample);
    for (size_t i = 0; i < n; ++i) out[i] = r.next();
    return 0;
}

}  // extern "C"

// ---- VCM integrator (Applications/Selas/Source/VCM.cpp) -----------------------------------------------------------
#include "orc_vcm.h"
