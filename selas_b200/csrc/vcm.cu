// vcm.cu -- the VCM (vertex connection and merging) integrator as a wavefront pipeline.
//
// Device restatement of Source/Applications/Selas/Source/VCM.cpp (VertexConnectionAndMerging, :664-697) and
// VCMCommon.cpp.  VCM.cpp is commented out in the reference tree (it predates the current LightDirectSample /
// CalculateSurfaceParams signatures); the fill-ins for the symbols that no longer exist are listed in DESIGN.md
// (section 4, VCM) and are the same here as in the CPU checker.  Compiled with -fmad=false -ftz=true -prec-div=true
// -prec-sqrt=true like shade.cu.
//
// One iteration (pass p, iterationIndex p + 1):
//   light pass    generate (EmitIblLightSample) -> [closest hit -> light shade]* : every hit stores a VCMVertex into the
//                 path's slot row, connects to the camera (shadow ray + film splat) and samples the BSDF.
//   compaction    scan of the per-path vertex counts = pathEnds (LightPathSet, VCM.cpp:57-62); slots -> dense arrays in the
//                 reference's order (path by path), plus the float3 positions the hash grid is built from.
//   hash grid     BuildHashGrid over the dense positions (hashgrid.cu), cellCount = vmCount.
//   camera pass   generate -> [closest hit -> camera shade -> connect -> merge -> shadow rays]* : camera shade handles the
//                 sky, the MIS update, the direct IBL connection and the BSDF sample; connect = one thread per (camera vertex,
//                 light vertex of the paired light path) (ConnectPathVertices); merge = one thread per camera vertex walking
//                 the 2x2x2 cells (SearchHashGrid + MergeVertices) in the reference's visiting order, so the per-vertex sum is
//                 accumulated in the same order as on the CPU.
//
// HBM layout (SoA, 16-byte records):
//   path queue     ray 32 B {o,tnear,d,tfar}; thrIdx 16 B {throughput, index:26|pathLength:5|isAreaMeasure:1}; mis 16 B {dVCM,dVC,dVM,rng}
//   light vertex   A 16 B {throughput, index:26|pathLength:5|transparent:1}; B 16 B {dVCM,dVC,dVM,continuationProbability};
//                  ray 32 B + hit 32 B (HitParameters is recomputed from them, RayPick VCM.cpp:123-166); position 12 B
//   shadow queue   ray 32 B + payload 16 B {contribution, pixel}
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "disney.cuh"
#include "hashgrid.cuh"
#include "internal.h"
#include "lights.cuh"
#include "surface.cuh"

namespace selas {

namespace {

inline uint32_t divUp(size_t a, uint32_t b) { return (uint32_t)((a + b - 1) / b); }

constexpr uint32_t kIndexMask = 0x3FFFFFFu;  // PathStateIndexBitCount_ 26, VCMCommon.h:16

struct VcmParams {
    uint32_t maxPathLength, flags, vmCount, stride;
    float vmWeight, vcWeight, vmNormalization, radius;
    float sphere[4];
    float pdfPosition;  // ConcentricDiscPdf() * (1 / R^2), AreaLighting.cpp:282,:326
    uint32_t seed, pass;
};

SELAS_D bool finite3(V3 c) { return isfinite(c.x) && isfinite(c.y) && isfinite(c.z); }
SELAS_D bool isZero3(V3 c) { return c.x == 0 && c.y == 0 && c.z == 0; }

// Math::SphericalToCartesian, MathLib/Projection.cpp:132-146
SELAS_D V3 sphericalToCartesian(float theta, float phi) {
    float st = sinf(theta), ct = cosf(theta), sp = sinf(phi), cp = cosf(phi);
    return v3(st * cp, ct, st * sp);
}
// MakeOrthogonalCoordinateSystem, GeometryLib/CoordinateSystem.h:14-21
SELAS_D void makeOrthogonalCoordinateSystem(V3 v1, V3& v2, V3& v3out) {
    if (fabsf(v1.x) > fabsf(v1.y))
        v2 = v3(-v1.z, 0.f, v1.x) * (1.0f / sqrtf(v1.x * v1.x + v1.z * v1.z));
    else
        v2 = v3(0.f, v1.z, -v1.y) * (1.0f / sqrtf(v1.y * v1.y + v1.z * v1.z));
    v3out = cross(v1, v2);
}
// SampleConcentricDisc, GeometryLib/Disc.cpp:13-56
SELAS_D void sampleConcentricDisc(float r0, float r1, float& dx, float& dy) {
    float phi, r;
    float a = 2.0f * r0 - 1.0f, b = 2.0f * r1 - 1.0f;
    if (a > -b) {
        if (a > b) { r = a; phi = (SELAS_PI / 4.0f) * (b / a); }
        else { r = b; phi = (SELAS_PI / 4.0f) * (2.0f - (a / b)); }
    } else {
        if (a < b) { r = -a; phi = (SELAS_PI / 4.0f) * (4.0f + (b / a)); }
        else {
            r = -b;
            if (fabsf(b) > 0.0f) phi = (SELAS_PI / 4.0f) * (6.0f - (a / b));
            else phi = 0.f;
        }
    }
    dx = r * cosf(phi);
    dy = r * sinf(phi);
}
// ContinuationProbability, Shading/SurfaceParameters.cpp:420-424
SELAS_D float continuationProbability(const Surface& s) { return saturate_(maxf_(maxf_(s.baseColor.x, s.baseColor.y), s.baseColor.z)); }
// WorldToImage (called at VCM.cpp:199, no longer present in GeometryLib/Camera.cpp): inverse of ImageToWorldDirection
// (Camera.cpp:56-60): dir ~ Z - cx X + cy Y with X, Y, Z mutually orthogonal; x = (cx + 1) w / 2, y = (1 - cy) h / 2, floor
SELAS_D void worldToImage(const SelasCamera& cam, V3 p, int32_t& ix, int32_t& iy) {
    V3 to = p - lightF3(cam.position);
    V3 X = lightF3(cam.cameraX), Y = lightF3(cam.cameraY), Z = lightF3(cam.cameraZ);
    float z = dot(to, Z) / dot(Z, Z);
    float cx = -(dot(to, X) / dot(X, X)) / z;
    float cy = (dot(to, Y) / dot(Y, Y)) / z;
    float x = (cx + 1.0f) * 0.5f * (float)(uint32_t)cam.width;
    float y = (1.0f - cy) * 0.5f * (float)(uint32_t)cam.height;
    ix = (int32_t)floorf(x);
    iy = (int32_t)floorf(y);
}

// ---- queue appends: one atomic per warp, ballot-ranked ---------------------------------------------------------------
SELAS_D uint32_t warpAppend(bool want, uint32_t* counter) {
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t mask = __ballot_sync(0xffffffffu, want);
    uint32_t base = 0;
    if (lane == 0 && mask) base = atomicAdd(counter, (uint32_t)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, 0);
    return base + __popc(mask & ((1u << lane) - 1u));
}
SELAS_D void storeRay(SelasRay* rays, uint32_t slot, V3 o, float tnear, V3 d, float tfar) {
    float4* r = reinterpret_cast<float4*>(rays + slot);
    r[0] = make_float4(o.x, o.y, o.z, tnear);
    r[1] = make_float4(d.x, d.y, d.z, tfar);
}

struct PathStateGpu {  // PathState, VCMCommon.h:48-60
    V3 position, direction, throughput;
    float dVCM, dVC, dVM;
    uint32_t index, pathLength, isAreaMeasure;
    PathRng rng;
};
SELAS_D void loadState(const SelasRay* rays, const float4* thrIdx, const float4* mis, uint32_t i, PathStateGpu& st) {
    const float4* rp = reinterpret_cast<const float4*>(rays + i);
    float4 r0 = __ldg(rp), r1 = __ldg(rp + 1), t = __ldg(thrIdx + i), m = __ldg(mis + i);
    st.position = v3(r0.x, r0.y, r0.z);
    st.direction = v3(r1.x, r1.y, r1.z);
    st.throughput = v3(t.x, t.y, t.z);
    uint32_t bits = __float_as_uint(t.w);
    st.index = bits & kIndexMask;
    st.pathLength = (bits >> 26) & 31u;
    st.isAreaMeasure = bits >> 31;
    st.dVCM = m.x; st.dVC = m.y; st.dVM = m.z;
    st.rng.state = __float_as_uint(m.w);
}
SELAS_D void storeState(const PathStateGpu& st, uint32_t slot, SelasRay* rays, float4* thrIdx, float4* mis) {
    // RayPick, VCM.cpp:123-166: tnear = 0.00001, tfar = FloatMax_, origin = state.position (no offset)
    storeRay(rays, slot, st.position, 0.00001f, st.direction, SELAS_FLOAT_MAX);
    uint32_t bits = (st.index & kIndexMask) | ((st.pathLength & 31u) << 26) | (st.isAreaMeasure << 31);
    thrIdx[slot] = make_float4(st.throughput.x, st.throughput.y, st.throughput.z, __uint_as_float(bits));
    mis[slot] = make_float4(st.dVCM, st.dVC, st.dVM, __uint_as_float(st.rng.state));
}

// SampleBsdfScattering, VCM.cpp:412-446
SELAS_D bool sampleBsdfScattering(const Surface& s, const DisneyFrame& frame, float vmWeight, float vcWeight, PathStateGpu& st) {
    BsdfResult bs;
    if (!sampleBsdf(st.rng, s, frame, -st.direction, bs)) return false;
    if (isZero3(bs.reflectance)) return false;
    float contProb = continuationProbability(s);
    float rr = st.rng.next();
    if (rr > contProb) return false;
    bs.forwardPdfW *= contProb;
    bs.reversePdfW *= contProb;
    float cosThetaBsdf = fabsf(dot(bs.wi, shadingNormal(s)));
    st.position = s.position;
    st.throughput = st.throughput * bs.reflectance;
    st.dVC = (cosThetaBsdf / bs.forwardPdfW) * (st.dVC * bs.reversePdfW + st.dVCM + vmWeight);
    st.dVM = (cosThetaBsdf / bs.forwardPdfW) * (st.dVM * bs.reversePdfW + st.dVCM * vcWeight + 1.0f);
    st.dVCM = 1.0f / bs.forwardPdfW;
    st.direction = bs.wi;
    ++st.pathLength;
    return true;
}

// ---- light pass -----------------------------------------------------------------------------------------------------
// GenerateLightSample (VCMCommon.cpp:37-63) over EmitIblLightSample (AreaLighting.cpp:246-290)
__global__ void __launch_bounds__(256) vcmGenerateLightKernel(IblGpu ibl, VcmParams prm, uint32_t count, SelasRay* __restrict__ rays,
                                                              float4* __restrict__ thrIdx, float4* __restrict__ mis) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    PathStateGpu st;
    st.rng.init(prm.seed, i, 2u * prm.pass);
    float r0 = st.rng.next(), r1 = st.rng.next(), r2 = st.rng.next(), r3 = st.rng.next();
    uint32_t x, y;
    float dirPhi, dirTheta, directionPdfA;
    iblSample(ibl, r0, r1, dirTheta, dirPhi, x, y, directionPdfA);
    V3 toIbl = sphericalToCartesian(dirTheta, dirPhi);
    const float* t = ibl.light + 3 * ((size_t)y * ibl.width + x);  // SampleIbl(ibl, x, y), ImageBasedLightResource.cpp:264-267
    V3 radiance = ibl.exposureScale * v3(__ldg(t), __ldg(t + 1), __ldg(t + 2));
    V3 dX, dZ;
    makeOrthogonalCoordinateSystem(toIbl, dX, dZ);
    float R = prm.sphere[3];
    V3 center = v3(prm.sphere[0], prm.sphere[1], prm.sphere[2]);
    float sx, sy;
    sampleConcentricDisc(r2, r3, sx, sy);
    V3 position = center + R * (toIbl + sx * dX + sy * dZ);
    float emissionPdfW = prm.pdfPosition * directionPdfA;
    float cosThetaLight = 1.0f;
    float lightSampleWeight = 1.0f;
    emissionPdfW *= lightSampleWeight;
    directionPdfA *= lightSampleWeight;
    st.position = position;
    st.direction = -toIbl;
    st.throughput = radiance * (1.0f / emissionPdfW);
    st.dVCM = directionPdfA / emissionPdfW;
    st.dVC = cosThetaLight / emissionPdfW;
    st.dVM = cosThetaLight / emissionPdfW * prm.vcWeight;
    st.pathLength = 1;
    st.isAreaMeasure = 0;
    st.index = i;
    storeState(st, i, rays, thrIdx, mis);
}

// EvaluateLightPaths loop body (VCM.cpp:503-561) for one dense queue of light sub-paths
__global__ void __launch_bounds__(128) vcmLightShadeKernel(SceneGpu sc, SelasCamera cam, VcmParams prm, const uint32_t* __restrict__ order, uint32_t n,
                                                           const SelasRay* __restrict__ rays,
                                                           const SelasHit* __restrict__ hits, const float4* __restrict__ thrIdx,
                                                           const float4* __restrict__ mis, const float* __restrict__ ptexRgb,
                                                           float4* __restrict__ slotA, float4* __restrict__ slotB, SelasRay* __restrict__ slotRay,
                                                           SelasHit* __restrict__ slotHit, float* __restrict__ slotPtex, uint32_t* __restrict__ lvCount,
                                                           SelasRay* __restrict__ raysOut, float4* __restrict__ thrIdxOut, float4* __restrict__ misOut,
                                                           SelasRay* __restrict__ shadowRays, float4* __restrict__ shadowPayload,
                                                           uint32_t* __restrict__ counters) {
    const uint32_t slotIndex = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t i = (order && slotIndex < n) ? __ldg(order + slotIndex) : slotIndex;  // material-sorted visiting order
    bool alive = false, wantShadow = false;
    PathStateGpu st;
    V3 shO = v3(0.f), shD = v3(0.f);
    float shNear = 0.f, shFar = 0.f;
    float4 shPayload = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n) {
        const uint4* hp = reinterpret_cast<const uint4*>(hits + i);
        uint4 h0 = __ldg(hp), h1 = __ldg(hp + 1);
        if (h1.x != SELAS_INVALID_ID) {
            loadState(rays, thrIdx, mis, i, st);
            Surface s;
            surfaceFromHit(sc, st.position, st.direction, __uint_as_float(h0.x), __uint_as_float(h0.y), __uint_as_float(h0.z), h0.w, h1.x, h1.y,
                           ptexRgb ? ptexRgb + 3 * (size_t)i : nullptr, s, h1.w);
            V3 view = -st.direction;
            V3 dp = st.position - s.position;
            float connectionLengthSqr = dot(dp, dp);
            float absDotNL = fabsf(dot(shadingNormal(s), view));
            if (st.pathLength > 1 || st.isAreaMeasure) st.dVCM *= connectionLengthSqr;
            st.dVCM *= (1.0f / absDotNL);
            st.dVC *= (1.0f / absDotNL);
            st.dVM *= (1.0f / absDotNL);
            const float contProb = continuationProbability(s);
            // lightPathSet->lightVertices.Add(vcmVertex), VCM.cpp:535-543: slot row of this path, compacted after the pass
            if (st.pathLength - 1 < prm.stride) {
                size_t slot = (size_t)st.index * prm.stride + (st.pathLength - 1);
                uint32_t bits = (st.index & kIndexMask) | (st.pathLength << 26) | ((s.materialFlags & SELAS_MAT_TRANSPARENT) ? 0x80000000u : 0u);
                slotA[slot] = make_float4(st.throughput.x, st.throughput.y, st.throughput.z, __uint_as_float(bits));
                slotB[slot] = make_float4(st.dVCM, st.dVC, st.dVM, contProb);
                const float4* rp = reinterpret_cast<const float4*>(rays + i);
                float4* ro = reinterpret_cast<float4*>(slotRay + slot);
                ro[0] = __ldg(rp);
                ro[1] = __ldg(rp + 1);
                uint4* ho = reinterpret_cast<uint4*>(slotHit + slot);
                ho[0] = h0;
                ho[1] = h1;
                if (slotPtex && ptexRgb) {  // the vertex keeps its Ptex base colour for ConnectPathVertices' CalculateSurfaceParams (VCM.cpp:314-317)
                    slotPtex[3 * slot] = ptexRgb[3 * (size_t)i];
                    slotPtex[3 * slot + 1] = ptexRgb[3 * (size_t)i + 1];
                    slotPtex[3 * slot + 2] = ptexRgb[3 * (size_t)i + 2];
                }
                lvCount[st.index] = st.pathLength;
            }
            DisneyFrame frame;
            makeDisneyFrame(s, s.shader == SELAS_SHADER_DISNEY_THIN, frame);
            // ---- ConnectLightPathToCamera, VCM.cpp:186-239 ----
            if (!(prm.flags & SELAS_VCM_NO_LIGHT_TRACING)) {
                V3 camPos = lightF3(cam.position), camFwd = lightF3(cam.forward);
                V3 toPosition = s.position - camPos;
                int32_t ix, iy;
                worldToImage(cam, s.position, ix, iy);
                if (dot(camFwd, toPosition) > 0.0f && ix >= 0 && ix < (int32_t)cam.width && iy >= 0 && iy < (int32_t)cam.height) {
                    float distance = length(toPosition);
                    toPosition = (1.0f / distance) * toPosition;
                    float cosThetaCamera = dot(camFwd, toPosition);
                    float imagePointToCameraDistance = cam.virtualImagePlaneDistance / cosThetaCamera;
                    float imageToSolidAngle = imagePointToCameraDistance * imagePointToCameraDistance / cosThetaCamera;
                    float imageToSurface = imageToSolidAngle * cosThetaCamera;
                    float surfaceToImage = 1.0f / imageToSurface;
                    float cameraPdfA = imageToSurface;
                    float fwd = 0.f, rev = 0.f;
                    V3 bsdf = evaluateBsdf(s, frame, view, -toPosition, fwd, rev);
                    if (!isZero3(bsdf)) {
                        fwd *= contProb;
                        rev *= contProb;
                        float lightPathCount = (float)prm.vmCount;
                        float lightPartialWeight = (cameraPdfA / lightPathCount) * (prm.vmWeight + st.dVCM + st.dVC * rev);
                        float misWeight = 1.0f / (lightPartialWeight + 1.0f);
                        V3 c = misWeight * st.throughput * bsdf * (1.0f / (lightPathCount * surfaceToImage));
                        if (!isZero3(c) && !((prm.flags & SELAS_VCM_SKIP_NONFINITE) && !finite3(c))) {
                            V3 dir = -toPosition;
                            shO = offsetRayOrigin(s, dir, 0.1f);  // OcclusionRay, VCM.cpp:71-92
                            shD = dir;
                            shNear = s.error;
                            shFar = distance;
                            shPayload = make_float4(c.x, c.y, c.z, __uint_as_float((uint32_t)iy * (uint32_t)cam.width + (uint32_t)ix));
                            wantShadow = true;
                        }
                    }
                }
            }
            // ---- advance the path ----
            if (sampleBsdfScattering(s, frame, prm.vmWeight, prm.vcWeight, st)) alive = st.pathLength + 2 < prm.maxPathLength;  // loop condition, :509
        }
    }
    uint32_t slotN = warpAppend(alive, counters + 0);
    if (alive) storeState(st, slotN, raysOut, thrIdxOut, misOut);
    uint32_t slotS = warpAppend(wantShadow, counters + 1);
    if (wantShadow) {
        storeRay(shadowRays, slotS, shO, shNear, shD, shFar);
        shadowPayload[slotS] = shPayload;
    }
}

// slots -> dense arrays in path order (the reference's lightVertices), and the positions the hash grid is built from
__global__ void __launch_bounds__(256) vcmCompactKernel(uint32_t vmCount, uint32_t stride, const uint32_t* __restrict__ lvCount,
                                                        const uint32_t* __restrict__ pathStarts, const float4* __restrict__ slotA,
                                                        const float4* __restrict__ slotB, const SelasRay* __restrict__ slotRay,
                                                        const SelasHit* __restrict__ slotHit, float4* __restrict__ lvA, float4* __restrict__ lvB,
                                                        SelasRay* __restrict__ lvRay, SelasHit* __restrict__ lvHit, float* __restrict__ lvPos,
                                                        const float* __restrict__ slotPtex, float* __restrict__ lvPtex) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)vmCount * stride) return;
    uint32_t path = (uint32_t)(t / stride), k = (uint32_t)(t - (size_t)path * stride);
    if (k >= __ldg(lvCount + path)) return;
    size_t dst = (size_t)__ldg(pathStarts + path) + k;
    lvA[dst] = slotA[t];
    lvB[dst] = slotB[t];
    const float4* ri = reinterpret_cast<const float4*>(slotRay + t);
    float4 r0 = ri[0], r1 = ri[1];
    float4* ro = reinterpret_cast<float4*>(lvRay + dst);
    ro[0] = r0;
    ro[1] = r1;
    const uint4* hi = reinterpret_cast<const uint4*>(slotHit + t);
    uint4 h0 = hi[0], h1 = hi[1];
    uint4* ho = reinterpret_cast<uint4*>(lvHit + dst);
    ho[0] = h0;
    ho[1] = h1;
    float tHit = __uint_as_float(h0.x);  // HitParameters::position = org + tfar * dir, VCM.cpp:149-151
    lvPos[3 * dst] = r0.x + tHit * r1.x;
    lvPos[3 * dst + 1] = r0.y + tHit * r1.y;
    lvPos[3 * dst + 2] = r0.z + tHit * r1.z;
    if (slotPtex) {
        lvPtex[3 * dst] = slotPtex[3 * t];
        lvPtex[3 * dst + 1] = slotPtex[3 * t + 1];
        lvPtex[3 * dst + 2] = slotPtex[3 * t + 2];
    }
}

__global__ void vcmPackVerticesKernel(size_t n, const float4* __restrict__ lvA, const float4* __restrict__ lvB, const SelasRay* __restrict__ lvRay,
                                      const SelasHit* __restrict__ lvHit, SelasVcmVertex* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 a = lvA[i], b = lvB[i];
    SelasVcmVertex v;
    v.throughput[0] = a.x; v.throughput[1] = a.y; v.throughput[2] = a.z;
    v.indexAndLength = __float_as_uint(a.w);
    v.dVCM = b.x; v.dVC = b.y; v.dVM = b.z; v.continuationProbability = b.w;
    v.ray = lvRay[i];
    v.hit = lvHit[i];
    out[i] = v;
}

// ---- camera pass ----------------------------------------------------------------------------------------------------
// GenerateCameraSample, VCMCommon.cpp:66-88 (x = scan - y * width: SURVEY.md 8(a) on VCM.cpp:569)
__global__ void __launch_bounds__(256) vcmGenerateCameraKernel(SelasCamera cam, VcmParams prm, uint32_t count, SelasRay* __restrict__ rays,
                                                               float4* __restrict__ thrIdx, float4* __restrict__ mis) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    PathStateGpu st;
    st.rng.init(prm.seed, i, 2u * prm.pass + 1u);
    uint32_t width = (uint32_t)cam.width;
    uint32_t y = i / width, x = i - y * width;
    float vx = (float)x + st.rng.next();  // JitteredCameraRay, Camera.cpp:63-74
    float vy = (float)y + st.rng.next();
    V3 dir = imageToWorldDirection(cam, vx, vy);
    float cosThetaCamera = dot(lightF3(cam.forward), dir);
    float imagePointToCameraDistance = cam.virtualImagePlaneDistance / cosThetaCamera;
    float invSolidAngleMeasure = imagePointToCameraDistance * imagePointToCameraDistance / cosThetaCamera;
    float revCameraPdfW = (1.0f / invSolidAngleMeasure);
    st.position = lightF3(cam.position);
    st.direction = dir;
    st.throughput = v3(1.0f);
    st.dVCM = (float)prm.vmCount * revCameraPdfW;
    st.dVC = 0.f;
    st.dVM = 0.f;
    st.pathLength = 1;
    st.isAreaMeasure = 1;
    st.index = y * (uint32_t)cam.viewportWidth + x;
    storeState(st, i, rays, thrIdx, mis);
}

SELAS_D void filmAdd(float* film, uint32_t pixel, V3 c) {
    atomicAdd(film + 3 * (size_t)pixel, c.x);
    atomicAdd(film + 3 * (size_t)pixel + 1, c.y);
    atomicAdd(film + 3 * (size_t)pixel + 2, c.z);
}

// EvaluateCameraPaths loop body (VCM.cpp:578-661) minus vertex connection and merging, which read what this kernel leaves in
// vertexMis (the MIS quantities *at* the vertex, before the BSDF sample advances them).
__global__ void __launch_bounds__(128) vcmCameraShadeKernel(SceneGpu sc, VcmParams prm, uint32_t n, const SelasRay* __restrict__ rays,
                                                            const SelasHit* __restrict__ hits, const float4* __restrict__ thrIdx,
                                                            const float4* __restrict__ mis, const float* __restrict__ ptexRgb,
                                                            float4* __restrict__ vertexMis, SelasRay* __restrict__ raysOut,
                                                            float4* __restrict__ thrIdxOut, float4* __restrict__ misOut,
                                                            SelasRay* __restrict__ shadowRays, float4* __restrict__ shadowPayload,
                                                            uint32_t* __restrict__ counters, float* __restrict__ film) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool alive = false, wantShadow = false;
    PathStateGpu st;
    V3 shO = v3(0.f), shD = v3(0.f);
    float shNear = 0.f, shFar = 0.f;
    float4 shPayload = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool skipNonFinite = (prm.flags & SELAS_VCM_SKIP_NONFINITE) != 0;
    if (i < n) {
        loadState(rays, thrIdx, mis, i, st);
        const uint4* hp = reinterpret_cast<const uint4*>(hits + i);
        uint4 h0 = __ldg(hp), h1 = __ldg(hp + 1);
        if (h1.x == SELAS_INVALID_ID) {
            // ConnectToSkyLight, VCM.cpp:168-183 over IblCalculateRadiance, AreaLighting.cpp:317-331
            if (!(prm.flags & SELAS_VCM_NO_SKY)) {
                float directPdfA = iblPdf(sc.ibl, st.direction);
                V3 radiance = iblLookup(sc.ibl, st.direction);
                float emissionPdfW = prm.pdfPosition * directPdfA;
                V3 c;
                if (st.pathLength == 1) {
                    c = st.throughput * radiance;
                } else {
                    float cameraWeight = directPdfA * st.dVCM + emissionPdfW * st.dVC;
                    float misWeight = 1.0f / (1.0f + cameraWeight);
                    c = st.throughput * (misWeight * radiance);
                }
                if (!(skipNonFinite && !finite3(c))) filmAdd(film, st.index, c);
            }
        } else {
            Surface s;
            surfaceFromHit(sc, st.position, st.direction, __uint_as_float(h0.x), __uint_as_float(h0.y), __uint_as_float(h0.z), h0.w, h1.x, h1.y,
                           ptexRgb ? ptexRgb + 3 * (size_t)i : nullptr, s, h1.w);
            V3 view = -st.direction;
            V3 dp = st.position - s.position;
            float connectionLengthSqr = dot(dp, dp);
            float absDotNL = fabsf(dot(shadingNormal(s), view));
            st.dVCM *= connectionLengthSqr;
            st.dVCM /= absDotNL;
            st.dVC /= absDotNL;
            st.dVM /= absDotNL;
            vertexMis[i] = make_float4(st.dVCM, st.dVC, st.dVM, __uint_as_float(st.pathLength));
            DisneyFrame frame;
            makeDisneyFrame(s, s.shader == SELAS_SHADER_DISNEY_THIN, frame);
            const float contProb = continuationProbability(s);
            // ---- ConnectCameraPathToLight, VCM.cpp:242-292 over DirectIblLightSample, AreaLighting.cpp:293-314 ----
            if (st.pathLength + 1 < prm.maxPathLength) {
                float r0 = st.rng.next(), r1 = st.rng.next();
                if (!(prm.flags & SELAS_VCM_NO_DIRECT)) {
                    uint32_t x, y;
                    float dirPhi, dirTheta, pdfW;
                    iblSample(sc.ibl, r0, r1, dirTheta, dirPhi, x, y, pdfW);
                    V3 toIbl = sphericalToCartesian(dirTheta, dirPhi);
                    const float* t = sc.ibl.light + 3 * ((size_t)y * sc.ibl.width + x);
                    V3 radiance = sc.ibl.exposureScale * v3(__ldg(t), __ldg(t + 1), __ldg(t + 2));
                    float directionPdfA = pdfW * 1.0f;
                    float emissionPdfW = prm.pdfPosition * pdfW;
                    float cosThetaLight = 1.0f;
                    float fwd = 0.f, rev = 0.f;
                    V3 bsdf = evaluateBsdf(s, frame, view, toIbl, fwd, rev);
                    if (!isZero3(bsdf)) {
                        fwd *= contProb;
                        rev *= contProb;
                        float cosThetaSurface = fabsf(dot(shadingNormal(s), toIbl));
                        float lightWeight = fwd / directionPdfA;
                        float cameraWeight = (emissionPdfW * cosThetaSurface / (directionPdfA * cosThetaLight)) * (prm.vmWeight + st.dVCM + st.dVC * rev);
                        float misWeight = 1.0f / (lightWeight + 1 + cameraWeight);
                        V3 pc = (misWeight * cosThetaSurface / directionPdfA) * radiance * bsdf;
                        if (!isZero3(pc)) {
                            V3 c = st.throughput * pc;
                            if (!(skipNonFinite && !finite3(c))) {
                                shO = offsetRayOrigin(s, toIbl, 0.1f);
                                shD = toIbl;
                                shNear = s.error;
                                shFar = 1e36f;
                                shPayload = make_float4(c.x, c.y, c.z, __uint_as_float(st.index));
                                wantShadow = true;
                            }
                        }
                    }
                }
            }
            // ---- advance the path ----
            if (sampleBsdfScattering(s, frame, prm.vmWeight, prm.vcWeight, st)) alive = st.pathLength < prm.maxPathLength;  // :585
        }
    }
    uint32_t slotN = warpAppend(alive, counters + 0);
    if (alive) storeState(st, slotN, raysOut, thrIdxOut, misOut);
    uint32_t slotS = warpAppend(wantShadow, counters + 1);
    if (wantShadow) {
        storeRay(shadowRays, slotS, shO, shNear, shD, shFar);
        shadowPayload[slotS] = shPayload;
    }
}

// camera vertex i of the queue as the connect / merge kernels see it
struct CameraVertex {
    Surface s;
    DisneyFrame frame;
    V3 view, throughput;
    float dVCM, dVC, dVM;
    uint32_t index, pathLength;
};
SELAS_D bool loadCameraVertex(const SceneGpu& sc, uint32_t i, const SelasRay* rays, const SelasHit* hits, const float4* thrIdx,
                              const float4* vertexMis, const float* ptexRgb, CameraVertex& cv) {
    const uint4* hp = reinterpret_cast<const uint4*>(hits + i);
    uint4 h1 = __ldg(hp + 1);
    if (h1.x == SELAS_INVALID_ID) return false;
    uint4 h0 = __ldg(hp);
    const float4* rp = reinterpret_cast<const float4*>(rays + i);
    float4 r0 = __ldg(rp), r1 = __ldg(rp + 1), t = __ldg(thrIdx + i), m = __ldg(vertexMis + i);
    V3 o = v3(r0.x, r0.y, r0.z), d = v3(r1.x, r1.y, r1.z);
    surfaceFromHit(sc, o, d, __uint_as_float(h0.x), __uint_as_float(h0.y), __uint_as_float(h0.z), h0.w, h1.x, h1.y,
                   ptexRgb ? ptexRgb + 3 * (size_t)i : nullptr, cv.s, h1.w);
    makeDisneyFrame(cv.s, cv.s.shader == SELAS_SHADER_DISNEY_THIN, cv.frame);
    cv.view = -d;
    cv.throughput = v3(t.x, t.y, t.z);
    cv.index = __float_as_uint(t.w) & kIndexMask;
    cv.dVCM = m.x; cv.dVC = m.y; cv.dVM = m.z;
    cv.pathLength = __float_as_uint(m.w);
    return true;
}

// Pairs for ConnectPathVertices: camera vertex i connects to the vertices of light path vcIndex = index (vcCount = 1, VCM.cpp:617-631).
// The path's vertices have pathLength 1, 2, ...: vertex k has length k + 1, so the reference's `break` at the first vertex that is too
// long (:624-626) bounds k by maxPathLength - 2 - cameraLength.  One thread per camera vertex, warp-scanned append of (i, li) pairs.
__global__ void __launch_bounds__(256) vcmConnectGatherKernel(VcmParams prm, uint32_t n, const SelasHit* __restrict__ hits,
                                                              const float4* __restrict__ thrIdx, const float4* __restrict__ vertexMis,
                                                              const uint32_t* __restrict__ pathStarts, const uint32_t* __restrict__ pathEnds,
                                                              uint2* __restrict__ pairs, uint32_t* __restrict__ pairCounter) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t count = 0, start = 0;
    if (i < n && __ldg(reinterpret_cast<const uint32_t*>(hits + i) + 4) != SELAS_INVALID_ID) {
        uint32_t index = __float_as_uint(__ldg(thrIdx + i).w) & kIndexMask;
        uint32_t cameraLength = __float_as_uint(__ldg(vertexMis + i).w);
        start = __ldg(pathStarts + index);
        count = __ldg(pathEnds + index) - start;
        uint32_t allowed = prm.maxPathLength > cameraLength + 1 ? prm.maxPathLength - 1 - cameraLength : 0u;
        count = count < allowed ? count : allowed;
    }
    uint32_t incl = count;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += y;
    }
    uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    uint32_t base = 0;
    if (lane == 31 && total) base = atomicAdd(pairCounter, total);
    base = __shfl_sync(0xffffffffu, base, 31) + incl - count;
    for (uint32_t k = 0; k < count; ++k) pairs[base + k] = make_uint2(i, start + k);
}

// ConnectPathVertices (VCM.cpp:295-359) for camera vertex i and light vertex li.  Returns false when the connection carries nothing;
// otherwise fills the shadow ray (VcOcclusionRay, VCM.cpp:95-120) and its payload {contribution, pixel}.
struct ShadowOut {
    V3 o, d;
    float tnear, tfar;
    float4 payload;
};
SELAS_D bool connectPathVertices(const SceneGpu& sc, const VcmParams& prm, uint32_t i, uint32_t li, const SelasRay* rays, const SelasHit* hits,
                                 const float4* thrIdx, const float4* vertexMis, const float* ptexRgb, const float4* lvA, const float4* lvB,
                                 const SelasRay* lvRay, const SelasHit* lvHit, const float* lvPtex, ShadowOut& out) {
    CameraVertex cv;
    loadCameraVertex(sc, i, rays, hits, thrIdx, vertexMis, ptexRgb, cv);
    float4 a = __ldg(lvA + li), b = __ldg(lvB + li);
    const float4* lrp = reinterpret_cast<const float4*>(lvRay + li);
    float4 lr0 = __ldg(lrp), lr1 = __ldg(lrp + 1);
    const uint4* lhp = reinterpret_cast<const uint4*>(lvHit + li);
    uint4 lh0 = __ldg(lhp), lh1 = __ldg(lhp + 1);
    Surface ls;  // CalculateSurfaceParams(context, &lightVertex.hit, lightSurface), :314-317
    V3 lo = v3(lr0.x, lr0.y, lr0.z), ld = v3(lr1.x, lr1.y, lr1.z);
    surfaceFromHit(sc, lo, ld, __uint_as_float(lh0.x), __uint_as_float(lh0.y), __uint_as_float(lh0.z), lh0.w, lh1.x, lh1.y,
                   lvPtex ? lvPtex + 3 * (size_t)li : nullptr, ls, lh1.w);
    V3 direction = ls.position - cv.s.position;
    float distanceSquared = dot(direction, direction);
    float distance = sqrtf(distanceSquared);
    direction = (1.0f / distance) * direction;
    float cFwd = 0.f, cRev = 0.f;
    V3 cameraBsdf = evaluateBsdf(cv.s, cv.frame, cv.view, direction, cFwd, cRev);
    if (isZero3(cameraBsdf)) return false;
    DisneyFrame lframe;
    makeDisneyFrame(ls, ls.shader == SELAS_SHADER_DISNEY_THIN, lframe);
    float lFwd = 0.f, lRev = 0.f;
    V3 lightView = -ld;
    V3 lightBsdf = evaluateBsdf(ls, lframe, -direction, lightView, lFwd, lRev);
    if (isZero3(lightBsdf)) return false;
    float cosThetaCamera = fabsf(dot(direction, shadingNormal(cv.s)));
    float cosThetaLight = fabsf(dot(-direction, shadingNormal(ls)));
    float geometryTerm = cosThetaLight * cosThetaCamera / distanceSquared;
    if (geometryTerm < 0.0f) return false;
    float cameraContProb = continuationProbability(cv.s);
    cFwd *= cameraContProb;
    cRev *= cameraContProb;
    float lightContProb = continuationProbability(ls);
    lFwd *= lightContProb;
    lRev *= lightContProb;
    float cameraBsdfPdfA = cFwd * fabsf(cosThetaLight) / distanceSquared;
    float lightBsdfPdfA = lFwd * fabsf(cosThetaCamera) / distanceSquared;
    float lightWeight = cameraBsdfPdfA * (prm.vmWeight + b.x + b.y * lRev);
    float cameraWeight = lightBsdfPdfA * (prm.vmWeight + cv.dVCM + cv.dVC * cRev);
    float misWeight = 1.0f / (lightWeight + 1.0f + cameraWeight);
    V3 pc = misWeight * geometryTerm * cameraBsdf * lightBsdf;
    if (isZero3(pc)) return false;
    V3 c = cv.throughput * v3(a.x, a.y, a.z) * pc;  // :628-629
    if ((prm.flags & SELAS_VCM_SKIP_NONFINITE) && !finite3(c)) return false;
    V3 nrm = shadingNormal(cv.s);
    float directionOffset = dot(direction, nrm) < 0.0f ? -1.0f : 1.0f;
    float biasDistance = directionOffset * cv.s.error * 0.1f;
    out.o = cv.s.position + biasDistance * nrm;
    out.d = direction;
    out.tnear = cv.s.error;
    out.tfar = distance - 16.0f * fabsf(biasDistance);
    out.payload = make_float4(c.x, c.y, c.z, __uint_as_float(cv.index));
    return true;
}

// One thread per gathered (camera vertex, light vertex) pair; the pair count lives in device memory, the grid strides over it.
__global__ void __launch_bounds__(128) vcmConnectKernel(SceneGpu sc, VcmParams prm, const uint32_t* __restrict__ pairCountPtr,
                                                        const uint2* __restrict__ pairs, const SelasRay* __restrict__ rays,
                                                        const SelasHit* __restrict__ hits, const float4* __restrict__ thrIdx,
                                                        const float4* __restrict__ vertexMis, const float* __restrict__ ptexRgb,
                                                        const float4* __restrict__ lvA, const float4* __restrict__ lvB,
                                                        const SelasRay* __restrict__ lvRay, const SelasHit* __restrict__ lvHit,
                                                        const float* __restrict__ lvPtex, SelasRay* __restrict__ shadowRays,
                                                        float4* __restrict__ shadowPayload, uint32_t* __restrict__ counters) {
    const uint32_t pairCount = __ldg(pairCountPtr);
    // whole warps iterate together: the queue append is warp-collective
    for (uint32_t t0 = blockIdx.x * blockDim.x; t0 < pairCount; t0 += gridDim.x * blockDim.x) {
        const uint32_t t = t0 + threadIdx.x;
        ShadowOut sh;
        bool wantShadow = false;
        if (t < pairCount) {
            const uint2 pr = __ldg(pairs + t);
            wantShadow = connectPathVertices(sc, prm, pr.x, pr.y, rays, hits, thrIdx, vertexMis, ptexRgb, lvA, lvB, lvRay, lvHit, lvPtex, sh);
        }
        uint32_t slotS = warpAppend(wantShadow, counters + 1);
        if (wantShadow) {
            storeRay(shadowRays, slotS, sh.o, sh.tnear, sh.d, sh.tfar);
            shadowPayload[slotS] = sh.payload;
        }
    }
}

// SearchHashGrid (VCMHashGrid.cpp:107-155) with MergeVertices (VCM.cpp:362-409) inlined: one thread per camera vertex, cells and
// candidates visited in the reference's order so `result` is summed in the same order as on the CPU.
__global__ void __launch_bounds__(128) vcmMergeKernel(SceneGpu sc, VcmParams prm, HashGridView grid, uint32_t n, const SelasRay* __restrict__ rays,
                                                      const SelasHit* __restrict__ hits, const float4* __restrict__ thrIdx,
                                                      const float4* __restrict__ vertexMis, const float* __restrict__ ptexRgb,
                                                      const float4* __restrict__ lvA, const float4* __restrict__ lvB,
                                                      const SelasRay* __restrict__ lvRay, float* __restrict__ film,
                                                      unsigned long long* __restrict__ statCounters) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    CameraVertex cv;
    if (!loadCameraVertex(sc, i, rays, hits, thrIdx, vertexMis, ptexRgb, cv)) return;
    const float px = cv.s.position.x, py = cv.s.position.y, pz = cv.s.position.z;
    float dminx = px - floatFromOrdered(__ldg(grid.bounds + 0)), dminy = py - floatFromOrdered(__ldg(grid.bounds + 1)),
          dminz = pz - floatFromOrdered(__ldg(grid.bounds + 2));
    float dmaxx = floatFromOrdered(__ldg(grid.bounds + 3)) - px, dmaxy = floatFromOrdered(__ldg(grid.bounds + 4)) - py,
          dmaxz = floatFromOrdered(__ldg(grid.bounds + 5)) - pz;
    if ((dminx < 0.0f || dminy < 0.0f || dminz < 0.0f) || (dmaxx < 0.0f || dmaxy < 0.0f || dmaxz < 0.0f)) return;
    const float inv = grid.inverseCellSize;
    float cx = inv * dminx, cy = inv * dminy, cz = inv * dminz;
    float fx = floorf(cx), fy = floorf(cy), fz = floorf(cz);
    int32_t bx = (int32_t)fx, by = (int32_t)fy, bz = (int32_t)fz;
    if (cx - fx < 0.5f) --bx;
    if (cy - fy < 0.5f) --by;
    if (cz - fz < 0.5f) --bz;
    const float cameraContProb = continuationProbability(cv.s);
    const bool cameraTransparent = (cv.s.materialFlags & SELAS_MAT_TRANSPARENT) != 0;
    V3 result = v3(0.0f);
    uint32_t candidates = 0, merges = 0;
#pragma unroll 1
    for (int c8 = 0; c8 < 8; ++c8) {
        uint32_t cell = cellHash(bx + (c8 & 1), by + ((c8 >> 1) & 1), bz + (c8 >> 2), grid.cellCount);
        uint32_t b = cell ? __ldg(grid.cellEnds + cell - 1) : 0u, e = __ldg(grid.cellEnds + cell);
        for (uint32_t sIdx = b; sIdx < e; ++sIdx) {
            float4 p = __ldg(grid.cellPoints + sIdx);
            ++candidates;
            float dx = px - p.x, dy = py - p.y, dz = pz - p.z;
            float d2 = (dx * dx) + (dy * dy) + (dz * dz);
            if (!(d2 <= grid.radiusSq)) continue;
            ++merges;
            uint32_t pi = __float_as_uint(p.w);
            float4 a = __ldg(lvA + pi);
            uint32_t bits = __float_as_uint(a.w);
            uint32_t lvLength = (bits >> 26) & 31u;
            if (cv.pathLength + lvLength > prm.maxPathLength) continue;   // :371-373
            if (cameraTransparent != ((bits & 0x80000000u) != 0)) continue;  // the eTransparent hack, :380-383
            float4 lb = __ldg(lvB + pi);
            float4 ld = __ldg(reinterpret_cast<const float4*>(lvRay + pi) + 1);
            V3 lightView = v3(-ld.x, -ld.y, -ld.z);  // lightSurface.view = hit.incDirection = -ray.direction
            float fwd = 0.f, rev = 0.f;
            V3 bsdf = evaluateBsdf(cv.s, cv.frame, cv.view, lightView, fwd, rev);
            if (isZero3(bsdf)) continue;
            fwd *= cameraContProb;
            rev *= lb.w;  // ContinuationProbability(lightSurface), stored with the vertex
            float lightWeight = lb.x * prm.vcWeight + lb.z * fwd;
            float cameraWeight = cv.dVCM * prm.vcWeight + cv.dVM * rev;
            float misWeight = 1.0f / (lightWeight + 1.0f + cameraWeight);
            result = result + misWeight * bsdf * v3(a.x, a.y, a.z);
        }
    }
    V3 c = cv.throughput * prm.vmNormalization * result;  // :649-650
    if (!((prm.flags & SELAS_VCM_SKIP_NONFINITE) && !finite3(c))) filmAdd(film, cv.index, c);
    if (candidates) atomicAdd(statCounters + 0, (unsigned long long)candidates);
    if (merges) atomicAdd(statCounters + 1, (unsigned long long)merges);
}

// ---- merging in two phases (the default) ------------------------------------------------------------------------------
// vcmMergeKernel above keeps the reference's serial order per camera vertex, but its per-thread candidate loops are badly
// balanced: the caustic concentrates light vertices, so a few lanes walk hundreds of candidates and evaluate hundreds of BSDFs
// while the rest of the warp idles (ncu, profiles/r1_vcm_*: 77 % of an iteration, 12 of 32 lanes active).  The two-phase form:
//   gather  warp-cooperative SearchHashGrid: each lane prepares one camera vertex (position, AABB test, eight cell ranges), then
//           the warp serves its 32 vertices in turn, 32 consecutive candidates per step (coalesced 16-byte reads); accepted
//           (camera vertex, light vertex) pairs are appended to a dense pair list.
//   eval    one thread per pair: MergeVertices; contributions of the same camera vertex are contiguous in the list, so a
//           segmented warp reduction leaves one film atomic per run.
// The per-vertex sum is then accumulated in a different order than on the CPU (fp32 reassociation only).
constexpr int kGatherWarps = 4;
__global__ void __launch_bounds__(kGatherWarps * 32) vcmMergeGatherKernel(HashGridView grid, uint32_t n, const SelasRay* __restrict__ rays,
                                                                          const SelasHit* __restrict__ hits, uint2* __restrict__ pairs,
                                                                          uint32_t pairCapacity, uint32_t* __restrict__ pairCounter,
                                                                          unsigned long long* __restrict__ statCounters) {
    __shared__ float sPos[kGatherWarps][32][3];
    __shared__ uint32_t sBegin[kGatherWarps][32][8];
    __shared__ uint32_t sCum[kGatherWarps][32][8];
    // accepted pairs are staged per warp and appended to the global list kStage at a time: one atomic on the list's counter per
    // flush instead of one per 32-candidate step (ncu: 80 % of this kernel's stall samples sat on the shuffle that waits for that
    // atomic's return, profiles/r1_vcm_kernel_metrics.csv)
    constexpr uint32_t kStage = 224;
    __shared__ uint2 sStage[kGatherWarps][kStage + 32];
    uint32_t staged = 0;  // warp-uniform
    const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
    auto flush = [&]() {
        uint32_t slot = 0;
        if (lane == 0) slot = atomicAdd(pairCounter, staged);
        slot = __shfl_sync(0xffffffffu, slot, 0);
        for (uint32_t k = lane; k < staged; k += 32)
            if (slot + k < pairCapacity) pairs[slot + k] = sStage[wib][k];
        __syncwarp();
        staged = 0;
    };
    const uint32_t ltMask = (1u << lane) - 1u;
    const uint32_t warpCount = gridDim.x * kGatherWarps;
    const float minx = floatFromOrdered(__ldg(grid.bounds + 0)), miny = floatFromOrdered(__ldg(grid.bounds + 1)),
                minz = floatFromOrdered(__ldg(grid.bounds + 2));
    const float maxx = floatFromOrdered(__ldg(grid.bounds + 3)), maxy = floatFromOrdered(__ldg(grid.bounds + 4)),
                maxz = floatFromOrdered(__ldg(grid.bounds + 5));
    const float inv = grid.inverseCellSize, r2 = grid.radiusSq;
    unsigned long long candidates = 0;
    for (uint32_t base = (blockIdx.x * kGatherWarps + wib) * 32u; base < n; base += warpCount * 32u) {
        const uint32_t qi = base + lane;
        uint32_t total = 0;
        if (qi < n) {
            const uint4* hp = reinterpret_cast<const uint4*>(hits + qi);
            uint4 h0 = __ldg(hp), h1 = __ldg(hp + 1);
            if (h1.x != SELAS_INVALID_ID) {
                const float4* rp = reinterpret_cast<const float4*>(rays + qi);
                float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
                float t = __uint_as_float(h0.x);
                float px = r0.x + t * r1.x, py = r0.y + t * r1.y, pz = r0.z + t * r1.z;  // surface.position, RayPick
                sPos[wib][lane][0] = px; sPos[wib][lane][1] = py; sPos[wib][lane][2] = pz;
                float dminx = px - minx, dminy = py - miny, dminz = pz - minz;
                float dmaxx = maxx - px, dmaxy = maxy - py, dmaxz = maxz - pz;
                bool outside = (dminx < 0.0f || dminy < 0.0f || dminz < 0.0f) || (dmaxx < 0.0f || dmaxy < 0.0f || dmaxz < 0.0f);
                if (!outside) {
                    float cx = inv * dminx, cy = inv * dminy, cz = inv * dminz;
                    float fx = floorf(cx), fy = floorf(cy), fz = floorf(cz);
                    int32_t bx = (int32_t)fx, by = (int32_t)fy, bz = (int32_t)fz;
                    if (cx - fx < 0.5f) --bx;
                    if (cy - fy < 0.5f) --by;
                    if (cz - fz < 0.5f) --bz;
#pragma unroll
                    for (int c8 = 0; c8 < 8; ++c8) {
                        uint32_t cell = cellHash(bx + (c8 & 1), by + ((c8 >> 1) & 1), bz + (c8 >> 2), grid.cellCount);
                        uint32_t b = cell ? __ldg(grid.cellEnds + cell - 1) : 0u, e = __ldg(grid.cellEnds + cell);
                        total += e - b;
                        sBegin[wib][lane][c8] = b;
                        sCum[wib][lane][c8] = total;
                    }
                }
            }
        }
        __syncwarp();
        uint32_t pending = __ballot_sync(0xffffffffu, total != 0u);
        while (pending) {
            const uint32_t L = __ffs(pending) - 1u;
            pending &= pending - 1u;
            const float px = sPos[wib][L][0], py = sPos[wib][L][1], pz = sPos[wib][L][2];
            uint32_t cum[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) cum[k] = sCum[wib][L][k];
            const uint32_t T = cum[7];
            if (lane == 0) candidates += T;
            for (uint32_t i0 = 0; i0 < T; i0 += 32) {
                const uint32_t i = i0 + lane;
                bool accept = false;
                uint32_t index = 0;
                if (i < T) {
                    uint32_t c = 0;
#pragma unroll
                    for (int k = 0; k < 7; ++k) c += (i >= cum[k]) ? 1u : 0u;
                    const uint32_t before = c ? sCum[wib][L][c - 1] : 0u;
                    const float4 p = __ldg(grid.cellPoints + sBegin[wib][L][c] + (i - before));
                    float dx = px - p.x, dy = py - p.y, dz = pz - p.z;
                    float d2 = (dx * dx) + (dy * dy) + (dz * dz);
                    accept = d2 <= r2;
                    index = __float_as_uint(p.w);
                }
                const uint32_t mask = __ballot_sync(0xffffffffu, accept);
                if (mask) {
                    if (accept) sStage[wib][staged + __popc(mask & ltMask)] = make_uint2(base + L, index);
                    staged += (uint32_t)__popc(mask);
                    __syncwarp();
                    if (staged >= kStage) flush();
                }
            }
        }
        __syncwarp();
    }
    if (staged) flush();
    if (lane == 0 && candidates) atomicAdd(statCounters + 0, candidates);
}

// MergeVertices (VCM.cpp:362-409) for one accepted (camera vertex i, light vertex pi) pair: the film contribution
// throughput * vmNormalization * misWeight * bsdf * lightVertex.throughput (:405, :649-650), or zero.
SELAS_D V3 mergeVertices(const SceneGpu& sc, const VcmParams& prm, uint32_t i, uint32_t pi, const SelasRay* rays, const SelasHit* hits,
                         const float4* thrIdx, const float4* vertexMis, const float* ptexRgb, const float4* lvA, const float4* lvB,
                         const SelasRay* lvRay) {
    float4 a = __ldg(lvA + pi);
    uint32_t bits = __float_as_uint(a.w);
    uint32_t lvLength = (bits >> 26) & 31u;
    uint32_t cameraLength = __float_as_uint(__ldg(vertexMis + i).w);
    if (cameraLength + lvLength > prm.maxPathLength) return v3(0.0f);  // :371-373
    CameraVertex cv;
    loadCameraVertex(sc, i, rays, hits, thrIdx, vertexMis, ptexRgb, cv);
    if (((cv.s.materialFlags & SELAS_MAT_TRANSPARENT) != 0) != ((bits & 0x80000000u) != 0)) return v3(0.0f);  // :380-383
    float4 lb = __ldg(lvB + pi);
    float4 ld = __ldg(reinterpret_cast<const float4*>(lvRay + pi) + 1);
    V3 lightView = v3(-ld.x, -ld.y, -ld.z);
    float fwd = 0.f, rev = 0.f;
    V3 bsdf = evaluateBsdf(cv.s, cv.frame, cv.view, lightView, fwd, rev);
    if (isZero3(bsdf)) return v3(0.0f);
    fwd *= continuationProbability(cv.s);
    rev *= lb.w;
    float lightWeight = lb.x * prm.vcWeight + lb.z * fwd;
    float cameraWeight = cv.dVCM * prm.vcWeight + cv.dVM * rev;
    float misWeight = 1.0f / (lightWeight + 1.0f + cameraWeight);
    V3 c = cv.throughput * prm.vmNormalization * (misWeight * bsdf * v3(a.x, a.y, a.z));
    if ((prm.flags & SELAS_VCM_SKIP_NONFINITE) && !finite3(c)) return v3(0.0f);
    return c;
}

__global__ void __launch_bounds__(128) vcmMergeEvalKernel(SceneGpu sc, VcmParams prm, const uint32_t* __restrict__ pairCountPtr, uint32_t pairCapacity,
                                                          const uint2* __restrict__ pairs, const SelasRay* __restrict__ rays,
                                                          const SelasHit* __restrict__ hits, const float4* __restrict__ thrIdx,
                                                          const float4* __restrict__ vertexMis, const float* __restrict__ ptexRgb,
                                                          const float4* __restrict__ lvA, const float4* __restrict__ lvB,
                                                          const SelasRay* __restrict__ lvRay, float* __restrict__ film) {
    const uint32_t pairCount = __ldg(pairCountPtr);
    if (pairCount > pairCapacity) return;  // the list overflowed: the host runs the serial kernel for this bounce instead
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t p0 = blockIdx.x * blockDim.x; p0 < pairCount; p0 += gridDim.x * blockDim.x) {
        const uint32_t p = p0 + threadIdx.x;
        V3 c = v3(0.0f);
        uint32_t key = 0xFFFFFFFFu;  // film pixel of this pair's camera vertex
        if (p < pairCount) {
            const uint2 pr = __ldg(pairs + p);
            key = __float_as_uint(__ldg(thrIdx + pr.x).w) & kIndexMask;
            c = mergeVertices(sc, prm, pr.x, pr.y, rays, hits, thrIdx, vertexMis, ptexRgb, lvA, lvB, lvRay);
        }
        // segmented reduction over runs of equal pixel (pairs of one camera vertex are mostly contiguous), one atomic per run.  Runs
        // are numbered by their head so that the same pixel appearing in two separate runs of a warp is not merged across the gap.
        uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
        const bool head = lane == 0 || prev != key;
        const uint32_t heads = __ballot_sync(0xffffffffu, head);
        const uint32_t run = __popc(heads & ((2u << lane) - 1u));
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            float tx = __shfl_down_sync(0xffffffffu, c.x, o), ty = __shfl_down_sync(0xffffffffu, c.y, o), tz = __shfl_down_sync(0xffffffffu, c.z, o);
            uint32_t r = __shfl_down_sync(0xffffffffu, run, o);
            if (lane + o < 32 && r == run) {
                c.x += tx; c.y += ty; c.z += tz;
            }
        }
        if (head && key != 0xFFFFFFFFu && !isZero3(c)) filmAdd(film, key, c);
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------
struct VcmBuffers {
    uint32_t capacity = 0, stride = 0;
    SelasRay* rays[2] = {nullptr, nullptr};
    float4* thrIdx[2] = {nullptr, nullptr};
    float4* mis[2] = {nullptr, nullptr};
    SelasHit* hits = nullptr;
    float4* vertexMis = nullptr;
    float4 *slotA = nullptr, *slotB = nullptr, *lvA = nullptr, *lvB = nullptr;
    SelasRay *slotRay = nullptr, *lvRay = nullptr;
    SelasHit *slotHit = nullptr, *lvHit = nullptr;
    float* lvPos = nullptr;
    uint32_t *lvCount = nullptr, *pathEnds = nullptr, *pathStarts = nullptr, *scanSums = nullptr;
    SelasRay* shadowRays = nullptr;
    float4* shadowPayload = nullptr;
    uint32_t shadowCapacity = 0;
    float* ptexRgb = nullptr;
    float *slotPtex = nullptr, *lvPtex = nullptr;  // Ptex base colour of every light vertex (scenes with Ptex only)
    uint32_t* counters = nullptr;  // 4 per launch slot: next paths, shadow rays, closest fetch, shadow fetch
    uint32_t counterSlots = 0;
    uint32_t* order = nullptr;  // material-sorted visiting order of the light queue
    uint32_t* bins = nullptr;   // 2 * kMaterialBinWords words per bounce, zeroed with the counters
    unsigned long long* statCounters = nullptr;  // merge candidates, merges, connections
    uint2* pairs = nullptr;  // (camera vertex, light vertex) pairs accepted by the range query
    uint32_t pairCapacity = 0;
    uint32_t* pairCounter = nullptr;  // 2 per bounce: merge pairs, connect pairs
    uint2* connPairs = nullptr;
    uint32_t* h_counts = nullptr;
    SelasHashGrid* grid = nullptr;
    std::vector<cudaEvent_t> events;
    cudaEvent_t evCounts = nullptr;
    std::vector<void*> owned;
};

template <typename T>
int vcmAlloc(SelasContext* ctx, VcmBuffers* w, T** p, size_t count) {
    SELAS_CUDA_CHECK(ctx, cudaMalloc((void**)p, sizeof(T) * (count ? count : 1)));
    w->owned.push_back(*p);
    return SELAS_OK;
}

// Light sub-paths start in random directions from the sky and camera sub-paths are incoherent after the first hit: those launches, and all
// shadow rays, go through the kernel with the warp-cooperative triangle stage (trace_coop.cuh); SELAS_VCM_COOP=0 keeps traceKernel.
static int vcmCoop() {
    static const int on = (getenv("SELAS_VCM_COOP") && atoi(getenv("SELAS_VCM_COOP")) == 0) ? 0 : 1;
    return on;
}

static int allocateVcm(SelasScene* scene, uint32_t vmCount, uint32_t maxPathLength) {
    SelasContext* ctx = scene->ctx;
    VcmBuffers* w = (VcmBuffers*)scene->vcm;
    const uint32_t stride = maxPathLength > 3 ? maxPathLength - 3 : 0;  // light vertices have pathLength 1 .. max-3 (VCM.cpp:509)
    const uint32_t slots = 4 * (2 * maxPathLength + 4);
    if (w && w->capacity == vmCount && w->stride == stride && w->counterSlots == slots) return SELAS_OK;
    if (w) freeVcm(scene);
    w = new VcmBuffers();
    scene->vcm = w;
    w->capacity = vmCount;
    w->stride = stride;
    w->counterSlots = slots;
    const size_t n = vmCount, nv = (size_t)vmCount * std::max(1u, stride);
#define VCM_TRY(x)              \
    do {                        \
        int rc__ = (x);         \
        if (rc__) return rc__;  \
    } while (0)
    for (int k = 0; k < 2; ++k) {
        VCM_TRY(vcmAlloc(ctx, w, &w->rays[k], n));
        VCM_TRY(vcmAlloc(ctx, w, &w->thrIdx[k], n));
        VCM_TRY(vcmAlloc(ctx, w, &w->mis[k], n));
    }
    VCM_TRY(vcmAlloc(ctx, w, &w->hits, n));
    VCM_TRY(vcmAlloc(ctx, w, &w->vertexMis, n));
    VCM_TRY(vcmAlloc(ctx, w, &w->slotA, nv));
    VCM_TRY(vcmAlloc(ctx, w, &w->slotB, nv));
    VCM_TRY(vcmAlloc(ctx, w, &w->slotRay, nv));
    VCM_TRY(vcmAlloc(ctx, w, &w->slotHit, nv));
    VCM_TRY(vcmAlloc(ctx, w, &w->lvA, nv));
    VCM_TRY(vcmAlloc(ctx, w, &w->lvB, nv));
    VCM_TRY(vcmAlloc(ctx, w, &w->lvRay, nv));
    VCM_TRY(vcmAlloc(ctx, w, &w->lvHit, nv));
    VCM_TRY(vcmAlloc(ctx, w, &w->lvPos, 3 * nv));
    VCM_TRY(vcmAlloc(ctx, w, &w->lvCount, n));
    VCM_TRY(vcmAlloc(ctx, w, &w->pathEnds, n));
    VCM_TRY(vcmAlloc(ctx, w, &w->pathStarts, n));
    VCM_TRY(vcmAlloc(ctx, w, &w->scanSums, (size_t)scanScratchWords(vmCount)));
    VCM_TRY(vcmAlloc(ctx, w, &w->shadowRays, n + nv));       // one direct connection + up to `stride` vertex connections per vertex
    VCM_TRY(vcmAlloc(ctx, w, &w->shadowPayload, n + nv));
    w->shadowCapacity = (uint32_t)std::min<size_t>(n + nv, 0xFFFFFFFFull);
    if (scene->usesPtex) {
        VCM_TRY(vcmAlloc(ctx, w, &w->ptexRgb, 3 * n));
        VCM_TRY(vcmAlloc(ctx, w, &w->slotPtex, 3 * nv));
        VCM_TRY(vcmAlloc(ctx, w, &w->lvPtex, 3 * nv));
    }
    VCM_TRY(vcmAlloc(ctx, w, &w->counters, (size_t)slots));
    VCM_TRY(vcmAlloc(ctx, w, &w->order, n));
    VCM_TRY(vcmAlloc(ctx, w, &w->bins, (size_t)2 * kMaterialBinWords * std::max(1u, stride)));
    VCM_TRY(vcmAlloc(ctx, w, &w->statCounters, (size_t)8));
    w->pairCapacity = (uint32_t)std::min<size_t>(std::max<size_t>(n * 8, (size_t)1 << 20), (size_t)1 << 28);
    VCM_TRY(vcmAlloc(ctx, w, &w->pairs, (size_t)w->pairCapacity));
    VCM_TRY(vcmAlloc(ctx, w, &w->pairCounter, (size_t)(2 * maxPathLength + 2)));
    VCM_TRY(vcmAlloc(ctx, w, &w->connPairs, nv));
#undef VCM_TRY
    SELAS_CUDA_CHECK(ctx, cudaMallocHost((void**)&w->h_counts, sizeof(uint32_t) * 8));
    SELAS_CUDA_CHECK(ctx, cudaEventCreateWithFlags(&w->evCounts, cudaEventDisableTiming));
    // the hash grid is rebuilt every iteration over a varying number of vertices: size it once for the worst case
    return hashGridReserve(ctx, &w->grid, nv, vmCount);
}

// A failed allocation half-way (2 M paths x 7 vertex slots is 3 GB) must not leave a half-null buffer set behind that the next call would
// take for "already the right size": whatever was allocated is released and the scene has no VCM scratch until a call succeeds.
int ensureVcm(SelasScene* scene, uint32_t vmCount, uint32_t maxPathLength) {
    const int rc = allocateVcm(scene, vmCount, maxPathLength);
    if (rc) {
        freeVcm(scene);
        cudaGetLastError();
    }
    return rc;
}

struct EventSpans {
    VcmBuffers* w;
    cudaStream_t st;
    size_t used = 0;
    struct Span { int kind; size_t a, b; };
    std::vector<Span> spans;
    size_t mark() {
        if (used == w->events.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            w->events.push_back(e);
        }
        cudaEventRecord(w->events[used], st);
        return used++;
    }
    void span(int kind, size_t a) { spans.push_back({kind, a, mark()}); }
    void collect(SelasVcmStats& s) {
        for (const Span& sp : spans) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, w->events[sp.a], w->events[sp.b]);
            if (sp.kind == 0) s.lightPassMs += ms;
            else if (sp.kind == 1) s.gridBuildMs += ms;
            else if (sp.kind == 2) s.cameraPassMs += ms;
            else if (sp.kind == 3) s.mergeMs += ms;
            else if (sp.kind == 4) s.connectMs += ms;
            else s.renderMs += ms;
        }
    }
};

int checkSettings(SelasScene* scene, const SelasVcmSettings& vs) {
    SelasContext* ctx = scene->ctx;
    if (!scene->gpu.ibl.present) {
        setError(ctx, "VCM emits from the image based light only (VCMCommon.cpp:37-63): the scene has none");
        return SELAS_ERR_ARGUMENT;
    }
    if (vs.width == 0 || vs.height == 0 || (uint64_t)vs.width * vs.height >= (1ull << 26) || vs.maxPathLength > 31) {
        setError(ctx, "VCM settings out of range: width*height < 2^26 (PathStateIndexBitCount_), maxPathLength <= 31 (PathStatePathLengthBitCount_)");
        return SELAS_ERR_ARGUMENT;
    }
    if (!(vs.boundingSphere[3] > 0.0f) || !(vs.radiusFactor > 0.0f)) {
        setError(ctx, "VCM needs the scene's bounding sphere (radius > 0) and radiusFactor > 0");
        return SELAS_ERR_ARGUMENT;
    }
    return SELAS_OK;
}

VcmParams makeParams(const SelasVcmSettings& vs, uint32_t pass, uint32_t stride) {
    const uint32_t vmCount = vs.width * vs.height, vcCount = 1;
    float out[5];
    selas_cuda_vcm_iteration_constants(vmCount, vcCount, vs.radiusFactor * vs.boundingSphere[3], vs.radiusAlpha, (float)pass + 1.0f, out);
    VcmParams p;
    p.maxPathLength = vs.maxPathLength;
    p.flags = vs.flags;
    p.vmCount = vmCount;
    p.stride = stride;
    p.radius = out[0];
    p.vmNormalization = out[2];
    p.vmWeight = out[3];
    p.vcWeight = out[4];
    for (int k = 0; k < 4; ++k) p.sphere[k] = vs.boundingSphere[k];
    const float R = vs.boundingSphere[3];
    p.pdfPosition = SELAS_INV_PI * (1.0f / (R * R));
    p.seed = vs.seed;
    p.pass = pass;
    return p;
}

// light pass + compaction of one iteration; *total = number of light vertices
int lightPass(SelasScene* scene, const SelasCamera& cam, const VcmParams& prm, float* d_film, SelasVcmStats& st, uint32_t* total,
              cudaStream_t stream) {
    SelasContext* ctx = scene->ctx;
    VcmBuffers* w = (VcmBuffers*)scene->vcm;
    const uint32_t vmCount = prm.vmCount;
    SELAS_CUDA_CHECK(ctx, cudaMemsetAsync(w->counters, 0, sizeof(uint32_t) * w->counterSlots, stream));
    SELAS_CUDA_CHECK(ctx, cudaMemsetAsync(w->lvCount, 0, sizeof(uint32_t) * vmCount, stream));
    SELAS_CUDA_CHECK(ctx, cudaMemsetAsync(w->bins, 0, sizeof(uint32_t) * 2 * kMaterialBinWords * std::max(1u, w->stride), stream));
    static const bool sortLight = !getenv("SELAS_VCM_NO_LIGHT_SORT");
    vcmGenerateLightKernel<<<divUp(vmCount, 256), 256, 0, stream>>>(scene->gpu.ibl, prm, vmCount, w->rays[0], w->thrIdx[0], w->mis[0]);
    ++st.kernelLaunches;
    uint32_t n = prm.stride ? vmCount : 0;
    int cur = 0;
    for (uint32_t it = 0; it < prm.stride && n > 0; ++it) {
        uint32_t* ctr = w->counters + 4 * (size_t)it;
        int rc = launchIntersectIndirect(scene, nullptr, n, w->rays[cur], w->hits, ctr + 2, nullptr, stream, nullptr, vcmCoop());
        if (rc) return rc;
        st.closestRays += n;
        if (scene->usesPtex) {
            unsigned grid = (unsigned)std::min<uint32_t>(divUp(n, 1024), (uint32_t)ctx->smCount * 8);
            if (scene->ptexMixedRes) ptexStageKernel<true><<<grid, 128, 0, stream>>>(scene->gpu, n, w->hits, w->ptexRgb);
            else ptexStageKernel<false><<<grid, 128, 0, stream>>>(scene->gpu, n, w->hits, w->ptexRgb);
            ++st.kernelLaunches;
        }
        // light sub-paths leave the environment in random directions: a warp holds misses, diffuse and glass hits side by side (7.5 of
        // 32 lanes per instruction measured).  Visiting the queue in material order, as the path tracer does, gives warps one code path.
        const uint32_t* order = nullptr;
        if (sortLight && n >= 4096) {
            rc = launchMaterialOrder(scene, n, w->hits, w->bins + 2 * kMaterialBinWords * (size_t)it, w->order, stream);
            if (rc) return rc;
            order = w->order;
            st.kernelLaunches += 3;
        }
        vcmLightShadeKernel<<<divUp(n, 128), 128, 0, stream>>>(scene->gpu, cam, prm, order, n, w->rays[cur], w->hits, w->thrIdx[cur], w->mis[cur],
                                                                scene->usesPtex ? w->ptexRgb : nullptr, w->slotA, w->slotB, w->slotRay, w->slotHit,
                                                                w->slotPtex, w->lvCount, w->rays[cur ^ 1], w->thrIdx[cur ^ 1], w->mis[cur ^ 1], w->shadowRays,
                                                                w->shadowPayload, ctr);
        SELAS_CUDA_CHECK(ctx, cudaGetLastError());
        // counts go back for the next bounce's launch shapes; the shadow launch reads its count on the device and is enqueued before
        // the host blocks, so the GPU stays busy while the host wakes up
        SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(w->h_counts, ctr, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        SELAS_CUDA_CHECK(ctx, cudaEventRecord(w->evCounts, stream));
        st.kernelLaunches += 2;
        if (!(prm.flags & SELAS_VCM_NO_LIGHT_TRACING)) {
            rc = launchShadowAccumulate(scene, ctr + 1, n, w->shadowRays, w->shadowPayload, d_film, ctr + 3, nullptr, stream, vcmCoop());
            if (rc) return rc;
            ++st.kernelLaunches;
        }
        SELAS_CUDA_CHECK(ctx, cudaEventSynchronize(w->evCounts));
        uint32_t nNext = w->h_counts[0];
        st.occlusionRays += w->h_counts[1];
        n = nNext;
        cur ^= 1;
    }
    // pathEnds = inclusive scan of the per-path vertex counts; slots -> dense
    *total = 0;
    int rc = launchScanU32(ctx, w->lvCount, vmCount, w->scanSums, w->pathEnds, w->pathStarts, stream);
    if (rc) return rc;
    st.kernelLaunches += 3;
    if (prm.stride) {
        vcmCompactKernel<<<divUp((size_t)vmCount * prm.stride, 256), 256, 0, stream>>>(vmCount, prm.stride, w->lvCount, w->pathStarts, w->slotA, w->slotB,
                                                                                        w->slotRay, w->slotHit, w->lvA, w->lvB, w->lvRay, w->lvHit,
                                                                                        w->lvPos, w->slotPtex, w->lvPtex);
        ++st.kernelLaunches;
    }
    SELAS_CUDA_CHECK(ctx, cudaGetLastError());
    SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(w->h_counts, w->pathEnds + (vmCount - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(stream));
    *total = w->h_counts[0];
    return SELAS_OK;
}

}  // namespace

void freeVcm(SelasScene* scene) {
    VcmBuffers* w = (VcmBuffers*)scene->vcm;
    if (!w) return;
    for (void* p : w->owned) cudaFree(p);
    if (w->h_counts) cudaFreeHost(w->h_counts);
    for (cudaEvent_t e : w->events) cudaEventDestroy(e);
    if (w->evCounts) cudaEventDestroy(w->evCounts);
    if (w->grid) selas_cuda_hashgrid_destroy(w->grid);
    delete w;
    scene->vcm = nullptr;
}

int renderVcm(SelasScene* scene, const SelasCamera& cam, const SelasVcmSettings& vs, float* d_film, SelasVcmStats* stats, cudaStream_t stream) {
    SelasContext* ctx = scene->ctx;
    int rc = checkSettings(scene, vs);
    if (rc) return rc;
    const uint32_t vmCount = vs.width * vs.height;
    SelasVcmStats st{};
    rc = ensureVcm(scene, vmCount, vs.maxPathLength);
    if (rc) return rc;
    VcmBuffers* w = (VcmBuffers*)scene->vcm;
    SELAS_CUDA_CHECK(ctx, cudaMemsetAsync(w->statCounters, 0, 8 * sizeof(unsigned long long), stream));
    EventSpans ev{w, stream};
    // SELAS_VCM_SERIAL_MERGE=1: one thread per camera vertex, candidates in the reference's order (A/B and parity of the summation order)
    static const bool serialMerge = getenv("SELAS_VCM_SERIAL_MERGE") && atoi(getenv("SELAS_VCM_SERIAL_MERGE")) != 0;
    unsigned long long pairsAccepted = 0, connectionsMade = 0;
    const size_t tAll = ev.mark();
    for (uint32_t itn = 0; itn < vs.iterations; ++itn) {
        const uint32_t pass = vs.firstIteration + itn;
        const VcmParams prm = makeParams(vs, pass, vs.maxPathLength > 3 ? vs.maxPathLength - 3 : 0);
        st.lastRadius = prm.radius;
        // ---- light paths ----
        size_t t0 = ev.mark();
        uint32_t total = 0;
        rc = lightPass(scene, cam, prm, d_film, st, &total, stream);
        if (rc) return rc;
        ev.span(0, t0);
        st.lightPaths += vmCount;
        st.lightVertices += total;
        // ---- BuildHashGrid(&hashGrid, vmCount, vmSearchRadius, lightVertices), VCM.cpp:689 ----
        const bool merging = !(vs.flags & SELAS_VCM_NO_MERGING) && total > 0;
        if (merging) {
            t0 = ev.mark();
            rc = selas_cuda_hashgrid_build_device(ctx, total, w->lvPos, st.lastRadius, vmCount, &w->grid, stream);
            if (rc) return rc;
            ev.span(1, t0);
            st.kernelLaunches += 11;
        }
        // ---- camera paths ----
        t0 = ev.mark();
        SELAS_CUDA_CHECK(ctx, cudaMemsetAsync(w->counters, 0, sizeof(uint32_t) * w->counterSlots, stream));
        SELAS_CUDA_CHECK(ctx, cudaMemsetAsync(w->pairCounter, 0, sizeof(uint32_t) * (2 * (size_t)vs.maxPathLength + 2), stream));
        vcmGenerateCameraKernel<<<divUp(vmCount, 256), 256, 0, stream>>>(cam, prm, vmCount, w->rays[0], w->thrIdx[0], w->mis[0]);
        ++st.kernelLaunches;
        uint32_t n = vmCount;
        int cur = 0;
        for (uint32_t it = 0; it + 1 < vs.maxPathLength && n > 0; ++it) {  // pathLength = it + 1 < maxPathLength
            uint32_t* ctr = w->counters + 4 * (size_t)it;
            rc = launchIntersectIndirect(scene, nullptr, n, w->rays[cur], w->hits, ctr + 2, nullptr, stream, nullptr, it >= 1 ? vcmCoop() : 0);
            if (rc) return rc;
            st.closestRays += n;
            const float* ptexRgb = nullptr;
            if (scene->usesPtex) {
                unsigned grid = (unsigned)std::min<uint32_t>(divUp(n, 1024), (uint32_t)ctx->smCount * 8);
                if (scene->ptexMixedRes) ptexStageKernel<true><<<grid, 128, 0, stream>>>(scene->gpu, n, w->hits, w->ptexRgb);
                else ptexStageKernel<false><<<grid, 128, 0, stream>>>(scene->gpu, n, w->hits, w->ptexRgb);
                ptexRgb = w->ptexRgb;
                ++st.kernelLaunches;
            }
            vcmCameraShadeKernel<<<divUp(n, 128), 128, 0, stream>>>(scene->gpu, prm, n, w->rays[cur], w->hits, w->thrIdx[cur], w->mis[cur], ptexRgb,
                                                                     w->vertexMis, w->rays[cur ^ 1], w->thrIdx[cur ^ 1], w->mis[cur ^ 1],
                                                                     w->shadowRays, w->shadowPayload, ctr, d_film);
            st.kernelLaunches += 2;
            // Everything below is enqueued before the one host synchronisation of the bounce; the pair and shadow-ray counts stay
            // in device memory (grid-stride kernels, the persistent trace kernel reads its ray count through a pointer).
            uint32_t* pairCtr = w->pairCounter + 2 * (size_t)it;  // [0] merge pairs, [1] connect pairs (zeroed with the queue counters)
            const bool connecting = !(vs.flags & SELAS_VCM_NO_CONNECTIONS) && prm.stride && total > 0;
            // persistent grids sized to what is actually resident (a grid of 8 CTAs per SM ran in two uneven waves for the 78-register
            // merge kernel: 2.39 instead of 1.40 ms on the first bounce, profiles/r1_vcm_launches_summary.csv)
            static int connectCtas = 0, mergeCtas = 0;
            if (!connectCtas) {
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&connectCtas, vcmConnectKernel, 128, 0);
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&mergeCtas, vcmMergeEvalKernel, 128, 0);
                connectCtas = std::max(1, connectCtas);
                mergeCtas = std::max(1, mergeCtas);
            }
            const unsigned connectGrid = (unsigned)(ctx->smCount * connectCtas), mergeGrid = (unsigned)(ctx->smCount * mergeCtas);
            if (connecting) {
                size_t tc = ev.mark();
                vcmConnectGatherKernel<<<divUp(n, 256), 256, 0, stream>>>(prm, n, w->hits, w->thrIdx[cur], w->vertexMis, w->pathStarts, w->pathEnds,
                                                                          w->connPairs, pairCtr + 1);
                vcmConnectKernel<<<connectGrid, 128, 0, stream>>>(scene->gpu, prm, pairCtr + 1, w->connPairs, w->rays[cur], w->hits, w->thrIdx[cur],
                                                                w->vertexMis, ptexRgb, w->lvA, w->lvB, w->lvRay, w->lvHit, w->lvPtex,
                                                                w->shadowRays, w->shadowPayload, ctr);
                ev.span(4, tc);
                st.kernelLaunches += 2;
            }
            if (merging && serialMerge) {
                size_t tm = ev.mark();
                vcmMergeKernel<<<divUp(n, 128), 128, 0, stream>>>(scene->gpu, prm, hashGridView(w->grid), n, w->rays[cur], w->hits, w->thrIdx[cur],
                                                                   w->vertexMis, ptexRgb, w->lvA, w->lvB, w->lvRay, d_film, w->statCounters);
                ev.span(3, tm);
                ++st.kernelLaunches;
            } else if (merging) {
                size_t tm = ev.mark();
                unsigned grid = (unsigned)std::min<uint32_t>(divUp(n, kGatherWarps * 32), (uint32_t)ctx->smCount * 16);
                vcmMergeGatherKernel<<<grid, kGatherWarps * 32, 0, stream>>>(hashGridView(w->grid), n, w->rays[cur], w->hits, w->pairs, w->pairCapacity,
                                                                             pairCtr, w->statCounters);
                vcmMergeEvalKernel<<<mergeGrid, 128, 0, stream>>>(scene->gpu, prm, pairCtr, w->pairCapacity, w->pairs, w->rays[cur], w->hits,
                                                                  w->thrIdx[cur], w->vertexMis, ptexRgb, w->lvA, w->lvB, w->lvRay, d_film);
                ev.span(3, tm);
                st.kernelLaunches += 2;
            }
            // shadow rays of the direct connection and of the vertex connections: count read on the device
            rc = launchShadowAccumulate(scene, ctr + 1, w->shadowCapacity, w->shadowRays, w->shadowPayload, d_film, ctr + 3, nullptr, stream, vcmCoop());
            if (rc) return rc;
            ++st.kernelLaunches;
            SELAS_CUDA_CHECK(ctx, cudaGetLastError());
            SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(w->h_counts, ctr, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
            SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(w->h_counts + 2, pairCtr, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
            SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(stream));
            const uint32_t nNext = w->h_counts[0];
            st.occlusionRays += w->h_counts[1];
            connectionsMade += w->h_counts[3];
            if (merging && !serialMerge) {
                const uint32_t pairCount = w->h_counts[2];
                if (pairCount > w->pairCapacity) {
                    // more accepted pairs than the list holds (the eval kernel then did nothing): this bounce is merged by the serial
                    // kernel instead, before the queue buffers are reused
                    size_t tm = ev.mark();
                    vcmMergeKernel<<<divUp(n, 128), 128, 0, stream>>>(scene->gpu, prm, hashGridView(w->grid), n, w->rays[cur], w->hits, w->thrIdx[cur],
                                                                       w->vertexMis, ptexRgb, w->lvA, w->lvB, w->lvRay, d_film, w->statCounters + 4);
                    ev.span(3, tm);
                    ++st.kernelLaunches;
                }
                pairsAccepted += pairCount;
            }
            n = nNext;
            cur ^= 1;
        }
        ev.span(2, t0);
        st.cameraPaths += vmCount;
    }
    ev.span(5, tAll);
    unsigned long long h[4] = {0, 0, 0, 0};
    SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(h, w->statCounters, sizeof(h), cudaMemcpyDeviceToHost, stream));
    SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(stream));
    ev.collect(st);
    st.mergeCandidates = h[0];
    st.mergesAccepted = serialMerge ? h[1] : pairsAccepted;
    st.connections = connectionsMade;
    if (stats) *stats = st;
    return SELAS_OK;
}

int vcmLightPassInspect(SelasScene* scene, const SelasCamera& cam, const SelasVcmSettings& vs, SelasVcmVertex* vertices, size_t maxVertices,
                        uint32_t* pathEnds, size_t* vertexCount) {
    SelasContext* ctx = scene->ctx;
    int rc = checkSettings(scene, vs);
    if (rc) return rc;
    const uint32_t vmCount = vs.width * vs.height;
    rc = ensureVcm(scene, vmCount, vs.maxPathLength);
    if (rc) return rc;
    VcmBuffers* w = (VcmBuffers*)scene->vcm;
    cudaStream_t stream = ctx->stream;
    float* d_film = nullptr;  // the light-tracing splats of this pass go to a scratch film
    SELAS_CUDA_CHECK(ctx, cudaMalloc((void**)&d_film, sizeof(float) * 3 * (size_t)vmCount));
    SELAS_CUDA_CHECK(ctx, cudaMemsetAsync(d_film, 0, sizeof(float) * 3 * (size_t)vmCount, stream));
    const VcmParams prm = makeParams(vs, vs.firstIteration, vs.maxPathLength > 3 ? vs.maxPathLength - 3 : 0);
    SelasVcmStats st{};
    uint32_t total = 0;
    rc = lightPass(scene, cam, prm, d_film, st, &total, stream);
    cudaFree(d_film);
    if (rc) return rc;
    if (vertexCount) *vertexCount = total;
    if (pathEnds) SELAS_CUDA_CHECK(ctx, cudaMemcpy(pathEnds, w->pathEnds, sizeof(uint32_t) * vmCount, cudaMemcpyDeviceToHost));
    size_t nOut = std::min<size_t>(maxVertices, total);
    if (vertices && nOut) {
        SelasVcmVertex* d_out = nullptr;
        SELAS_CUDA_CHECK(ctx, cudaMalloc((void**)&d_out, sizeof(SelasVcmVertex) * nOut));
        vcmPackVerticesKernel<<<divUp(nOut, 256), 256, 0, stream>>>(nOut, w->lvA, w->lvB, w->lvRay, w->lvHit, d_out);
        cudaError_t e = cudaMemcpyAsync(vertices, d_out, sizeof(SelasVcmVertex) * nOut, cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        cudaFree(d_out);
        SELAS_CUDA_CHECK(ctx, e);
    }
    return SELAS_OK;
}

}  // namespace selas

using namespace selas;

extern "C" {

// VCMCommon::SearchRadius / CalculateIterationConstants, VCMCommon.cpp:91-112 (host arithmetic, as in the reference)
int selas_cuda_vcm_iteration_constants(uint32_t vmCount, uint32_t vcCount, float baseRadius, float radiusAlpha, float iterationIndex, float* out5) {
    if (!out5 || vmCount == 0 || vcCount == 0) return SELAS_ERR_ARGUMENT;
    const float pi = 3.141592653589793f;
    float vmSearchRadius = baseRadius / powf(iterationIndex, 0.5f * (1.0f - radiusAlpha));
    float r2 = vmSearchRadius * vmSearchRadius;
    out5[0] = vmSearchRadius;
    out5[1] = r2;
    out5[2] = 1.0f / (pi * r2 * vmCount);
    out5[3] = pi * r2 * vmCount / vcCount;
    out5[4] = vcCount / (pi * r2 * vmCount);
    return SELAS_OK;
}

}  // extern "C"
