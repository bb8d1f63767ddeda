// oracle/orc_vcm.h -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Included by selas_oracle.cpp.
//
// CPU restatement of the reference's VCM integrator, Applications/Selas/Source/VCM.cpp.  That file is entirely
// commented out in the reference tree (it predates the current LightDirectSample / CalculateSurfaceParams
// signatures and no longer compiles), so parity for the *orchestration* below is UNPINNED: it follows the commented
// source line by line.  Its building blocks that are still live code are pinned bit for bit against the compiled
// reference objects (tests/test_oracle_vs_ref.py): EmitIblLightSample / DirectIblLightSample / IblCalculateRadiance
// (Shading/AreaLighting.cpp:246-331), CalculateIterationConstants (VCMCommon.cpp:97-112), the hash grid
// (VCMHashGrid.cpp), EvaluateBsdf / SampleBsdfFunction, the camera.
//
// Where the dead code refers to things that no longer exist, the restatement fills in as follows (also in DESIGN.md):
//   * LightDirectSample::directionPdfA / emissionPdfW / cosThetaLight (VCM.cpp:255-274): directionPdfA = pdfW of
//     DirectIblLightSample, emissionPdfW = pdfPosition * pdfW (as EmitIblLightSample and IblCalculateRadiance define it),
//     cosThetaLight = 1 (AreaLighting.cpp:289 "not used for ibl light sources").
//   * surface.perturbedNormal: SurfaceParameters has no such field any more; GeometricNormal(surface) (the shading
//     normal in column 1 of worldToTangent) is used, as the camera pass already does (VCM.cpp:609).
//   * WorldToImage: see orc_shading.h.
//   * x = scan - y * height (VCM.cpp:569) is a bug; width is used, so index == scan.
//   * random numbers: one counter-keyed PCG stream per (seed, path index, 2 * pass + {0 light, 1 camera}) instead of one
//     mt19937 per worker thread (VCM.cpp:712), the same contract as the path tracer.
//   * IntegrationSeconds_ (VCM.cpp:31) becomes an iteration count; pass p has iterationIndex p + 1 (VCM.cpp:727-728).

namespace {

struct VcmPathState {  // PathState, VCMCommon.h:48-60
    V3 position, direction, throughput;
    float dVCM, dVC, dVM;
    uint32_t index, pathLength, isAreaMeasure;
};

struct VcmCounters {
    uint64_t closest = 0, shadow = 0, mergeCandidates = 0, merges = 0, connections = 0;
};

static inline void atomicAddFloat(float* p, float v) {
    uint32_t* u = reinterpret_cast<uint32_t*>(p);
    uint32_t old = __atomic_load_n(u, __ATOMIC_RELAXED);
    for (;;) {
        float f;
        memcpy(&f, &old, 4);
        f += v;
        uint32_t nu;
        memcpy(&nu, &f, 4);
        if (__atomic_compare_exchange_n(u, &old, nu, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) return;
    }
}
static inline bool finite3(V3 c) { return std::isfinite(c.x) && std::isfinite(c.y) && std::isfinite(c.z); }

// RayPick, VCM.cpp:123-166 (tnear = 0.00001, tfar = FloatMax_)
static bool vcmRayPick(const Scene& sc, V3 o, V3 d, SelasRay& ray, SelasHit& h, VcmCounters& ctr) {
    ray.ox = o.x; ray.oy = o.y; ray.oz = o.z; ray.tnear = 0.00001f;
    ray.dx = d.x; ray.dy = d.y; ray.dz = d.z; ray.tfar = kFloatMax;
    float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    RayHit rh;
    ++ctr.closest;
    bool hit = sc.trace<float>(oo, dd, ray.tnear, ray.tfar, false, rh);
    hitToRecord(sc, rh, ray, h, nullptr);
    return hit;
}
// OcclusionRay, VCM.cpp:71-92: true = unoccluded
static bool vcmOcclusionRay(const Scene& sc, const SelasSurface& surf, V3 direction, float distance, VcmCounters& ctr) {
    V3 org = offsetRayOrigin(surf, direction, 0.1f);
    float so[3] = {org.x, org.y, org.z}, sd[3] = {direction.x, direction.y, direction.z};
    RayHit sh;
    ++ctr.shadow;
    return !sc.trace<float>(so, sd, surf.error, distance, true, sh);
}
// VcOcclusionRay, VCM.cpp:95-120
static bool vcmVcOcclusionRay(const Scene& sc, const SelasSurface& surf, V3 direction, float distance, VcmCounters& ctr) {
    float bias;
    V3 org = offsetRayOriginBias(surf, direction, 0.1f, bias);
    float so[3] = {org.x, org.y, org.z}, sd[3] = {direction.x, direction.y, direction.z};
    RayHit sh;
    ++ctr.shadow;
    return !sc.trace<float>(so, sd, surf.error, distance - 16.0f * absf(bias), true, sh);
}

// SampleBsdfScattering, VCM.cpp:412-446
static bool vcmSampleBsdfScattering(PathRng& rng, const SelasSurface& surf, float vmWeight, float vcWeight, VcmPathState& st) {
    SelasBsdfSample bs;
    bool ok;
    withStream(rng, [&](Stream& s) { ok = sampleBsdf(s, surf, -st.direction, bs); });
    if (!ok) return false;
    if (bs.reflectance[0] == 0.0f && bs.reflectance[1] == 0.0f && bs.reflectance[2] == 0.0f) return false;
    float contProb = continuationProbability(surf);
    float rr = rng.next();
    if (rr > contProb) return false;
    bs.forwardPdfW *= contProb;
    bs.reversePdfW *= contProb;
    float cosThetaBsdf = absf(dot(v3(bs.wi), geometricNormal(surf)));
    st.position = v3(surf.position);
    st.throughput = st.throughput * v3(bs.reflectance);
    st.dVC = (cosThetaBsdf / bs.forwardPdfW) * (st.dVC * bs.reversePdfW + st.dVCM + vmWeight);
    st.dVM = (cosThetaBsdf / bs.forwardPdfW) * (st.dVM * bs.reversePdfW + st.dVCM * vcWeight + 1.0f);
    st.dVCM = 1.0f / bs.forwardPdfW;
    st.direction = v3(bs.wi);
    ++st.pathLength;
    return true;
}

// ConnectLightPathToCamera, VCM.cpp:186-239
static void vcmConnectLightPathToCamera(const Scene& sc, const SelasCamera& cam, const VcmPathState& st, const SelasSurface& surf,
                                        float vmWeight, float lightPathCount, bool skipNonFinite, float* film, VcmCounters& ctr) {
    V3 toPosition = v3(surf.position) - v3(cam.position);
    if (dot(v3(cam.forward), toPosition) <= 0.0f) return;
    int32_t ix, iy;
    worldToImage(cam, v3(surf.position), ix, iy);
    if (ix < 0 || ix >= (int32_t)cam.width || iy < 0 || iy >= (int32_t)cam.height) return;
    float distance = length(toPosition);
    toPosition = (1.0f / distance) * toPosition;
    float cosThetaCamera = dot(v3(cam.forward), toPosition);
    float imagePointToCameraDistance = cam.virtualImagePlaneDistance / cosThetaCamera;
    float imageToSolidAngle = imagePointToCameraDistance * imagePointToCameraDistance / cosThetaCamera;
    float imageToSurface = imageToSolidAngle * cosThetaCamera;
    float surfaceToImage = 1.0f / imageToSurface;
    float cameraPdfA = imageToSurface;
    float fwd = 0.f, rev = 0.f;
    V3 bsdf = evaluateBsdf(surf, -st.direction, -toPosition, fwd, rev);
    if (bsdf.x == 0 && bsdf.y == 0 && bsdf.z == 0) return;
    float contProb = continuationProbability(surf);
    fwd *= contProb;
    rev *= contProb;
    float lightPartialWeight = (cameraPdfA / lightPathCount) * (vmWeight + st.dVCM + st.dVC * rev);
    float misWeight = 1.0f / (lightPartialWeight + 1.0f);
    V3 c = misWeight * st.throughput * bsdf * (1.0f / (lightPathCount * surfaceToImage));
    if (c.x == 0 && c.y == 0 && c.z == 0) return;
    if (skipNonFinite && !finite3(c)) return;
    if (vcmOcclusionRay(sc, surf, -toPosition, distance, ctr)) {
        size_t p = (size_t)iy * (size_t)cam.width + (size_t)ix;
        atomicAddFloat(film + 3 * p, c.x);
        atomicAddFloat(film + 3 * p + 1, c.y);
        atomicAddFloat(film + 3 * p + 2, c.z);
    }
}

// GenerateLightSample, VCMCommon.cpp:37-63
static void vcmGenerateLightSample(PathRng& rng, const Scene& sc, const float sphere[4], float vcWeight, uint32_t index, VcmPathState& st) {
    LightEmission e;
    withStream(rng, [&](Stream& s) { emitIblLightSample(s, sc.ibl, sphere, e); });
    float lightSampleWeight = 1.0f;
    e.emissionPdfW *= lightSampleWeight;
    e.directionPdfA *= lightSampleWeight;
    st.position = e.position;
    st.direction = e.direction;
    st.throughput = e.radiance * (1.0f / e.emissionPdfW);
    st.dVCM = e.directionPdfA / e.emissionPdfW;
    st.dVC = e.cosThetaLight / e.emissionPdfW;
    st.dVM = e.cosThetaLight / e.emissionPdfW * vcWeight;
    st.pathLength = 1;
    st.isAreaMeasure = 0;
    st.index = index;
}
// GenerateCameraSample, VCMCommon.cpp:66-88
static void vcmGenerateCameraSample(PathRng& rng, const SelasCamera& cam, uint32_t x, uint32_t y, float lightPathCount, VcmPathState& st) {
    float vx = (float)x + rng.next();  // JitteredCameraRay, Camera.cpp:63-74
    float vy = (float)y + rng.next();
    V3 dir = imageToWorldDirection(cam, vx, vy);
    float cosThetaCamera = dot(v3(cam.forward), dir);
    float imagePointToCameraDistance = cam.virtualImagePlaneDistance / cosThetaCamera;
    float invSolidAngleMeasure = imagePointToCameraDistance * imagePointToCameraDistance / cosThetaCamera;
    float revCameraPdfW = (1.0f / invSolidAngleMeasure);
    st.position = v3(cam.position);
    st.direction = dir;
    st.throughput = v3(1.0f);
    st.dVCM = lightPathCount * revCameraPdfW;
    st.dVC = 0;
    st.dVM = 0;
    st.pathLength = 1;
    st.isAreaMeasure = 1;
    st.index = y * (uint32_t)cam.viewportWidth + x;
}

// one light path: EvaluateLightPaths loop body, VCM.cpp:503-561
static void vcmLightPath(const Scene& sc, const SelasCamera& cam, const SelasVcmSettings& vs, const VcmConstants& k, uint32_t pass,
                         uint32_t index, std::vector<SelasVcmVertex>& out, float* film, VcmCounters& ctr) {
    PathRng rng;
    rng.init(vs.seed, index, 2u * pass);
    VcmPathState st;
    vcmGenerateLightSample(rng, sc, vs.boundingSphere, k.vcWeight, index, st);
    while (st.pathLength + 2 < vs.maxPathLength) {
        SelasRay ray;
        SelasHit h;
        if (!vcmRayPick(sc, st.position, st.direction, ray, h, ctr)) break;
        SelasSurface surf;
        surfaceParams(sc, ray, h, surf);
        float connectionLengthSqr = dot(st.position - v3(surf.position), st.position - v3(surf.position));
        float absDotNL = absf(dot(geometricNormal(surf), v3(surf.view)));
        if (st.pathLength > 1 || st.isAreaMeasure) st.dVCM *= connectionLengthSqr;
        st.dVCM *= (1.0f / absDotNL);
        st.dVC *= (1.0f / absDotNL);
        st.dVM *= (1.0f / absDotNL);
        SelasVcmVertex v;
        store(v.throughput, st.throughput);
        v.indexAndLength = (st.index & 0x3FFFFFFu) | (st.pathLength << 26) | ((surf.materialFlags & SELAS_MAT_TRANSPARENT) ? 0x80000000u : 0u);
        v.dVCM = st.dVCM;
        v.dVC = st.dVC;
        v.dVM = st.dVM;
        v.continuationProbability = continuationProbability(surf);
        v.ray = ray;
        v.hit = h;
        out.push_back(v);
        if (!(vs.flags & SELAS_VCM_NO_LIGHT_TRACING))
            vcmConnectLightPathToCamera(sc, cam, st, surf, k.vmWeight, (float)k.vmCount, (vs.flags & SELAS_VCM_SKIP_NONFINITE) != 0, film, ctr);
        if (!vcmSampleBsdfScattering(rng, surf, k.vmWeight, k.vcWeight, st)) break;
    }
}

// ConnectToSkyLight, VCM.cpp:168-183
static V3 vcmConnectToSkyLight(const Scene& sc, const float sphere[4], const VcmPathState& st) {
    float directPdfA, emissionPdfW;
    V3 radiance = iblCalculateRadiance(sc.ibl, sphere, st.direction, directPdfA, emissionPdfW);
    if (st.pathLength == 1) return radiance;
    float cameraWeight = directPdfA * st.dVCM + emissionPdfW * st.dVC;
    float misWeight = 1.0f / (1.0f + cameraWeight);
    return misWeight * radiance;
}
// ConnectCameraPathToLight, VCM.cpp:242-292
static V3 vcmConnectCameraPathToLight(PathRng& rng, const Scene& sc, const float sphere[4], const VcmPathState& st, const SelasSurface& surf,
                                      float vmWeight, VcmCounters& ctr) {
    SelasLightSample ls;
    withStream(rng, [&](Stream& s) { directIblLightSample(s, sc.ibl, ls); });
    float directionPdfA = ls.pdfW * 1.0f;
    float R = sphere[3];
    float emissionPdfW = (kInvPi * (1.0f / (R * R))) * ls.pdfW;
    float cosThetaLight = 1.0f;
    float fwd = 0.f, rev = 0.f;
    V3 dir = v3(ls.direction);
    V3 bsdf = evaluateBsdf(surf, -st.direction, dir, fwd, rev);
    if (bsdf.x == 0 && bsdf.y == 0 && bsdf.z == 0) return v3(0.0f);
    float contProb = continuationProbability(surf);
    fwd *= contProb;
    rev *= contProb;
    float cosThetaSurface = absf(dot(geometricNormal(surf), dir));
    float lightWeight = fwd / directionPdfA;
    float cameraWeight = (emissionPdfW * cosThetaSurface / (directionPdfA * cosThetaLight)) * (vmWeight + st.dVCM + st.dVC * rev);
    float misWeight = 1.0f / (lightWeight + 1 + cameraWeight);
    V3 c = (misWeight * cosThetaSurface / directionPdfA) * v3(ls.radiance) * bsdf;
    if (c.x == 0 && c.y == 0 && c.z == 0) return v3(0.0f);
    if (vcmOcclusionRay(sc, surf, dir, ls.distance, ctr)) return c;
    return v3(0.0f);
}
// ConnectPathVertices, VCM.cpp:295-359
static V3 vcmConnectPathVertices(const Scene& sc, const SelasSurface& surf, const VcmPathState& cs, const SelasVcmVertex& lv, float vmWeight,
                                 VcmCounters& ctr) {
    ++ctr.connections;
    SelasSurface ls;
    surfaceParams(sc, lv.ray, lv.hit, ls);
    V3 direction = v3(ls.position) - v3(surf.position);
    float distanceSquared = dot(direction, direction);
    float distance = sqrtf(distanceSquared);
    direction = (1.0f / distance) * direction;
    float cFwd = 0.f, cRev = 0.f;
    V3 cameraBsdf = evaluateBsdf(surf, -cs.direction, direction, cFwd, cRev);
    if (cameraBsdf.x == 0 && cameraBsdf.y == 0 && cameraBsdf.z == 0) return v3(0.0f);
    float lFwd = 0.f, lRev = 0.f;
    V3 lightBsdf = evaluateBsdf(ls, -direction, v3(ls.view), lFwd, lRev);
    if (lightBsdf.x == 0 && lightBsdf.y == 0 && lightBsdf.z == 0) return v3(0.0f);
    float cosThetaCamera = absf(dot(direction, geometricNormal(surf)));
    float cosThetaLight = absf(dot(-direction, geometricNormal(ls)));
    float geometryTerm = cosThetaLight * cosThetaCamera / distanceSquared;
    if (geometryTerm < 0.0f) return v3(0.0f);
    float cameraContProb = continuationProbability(surf);
    cFwd *= cameraContProb;
    cRev *= cameraContProb;
    float lightContProb = continuationProbability(ls);
    lFwd *= lightContProb;
    lRev *= lightContProb;
    float cameraBsdfPdfA = cFwd * absf(cosThetaLight) / distanceSquared;
    float lightBsdfPdfA = lFwd * absf(cosThetaCamera) / distanceSquared;
    float lightWeight = cameraBsdfPdfA * (vmWeight + lv.dVCM + lv.dVC * lRev);
    float cameraWeight = lightBsdfPdfA * (vmWeight + cs.dVCM + cs.dVC * cRev);
    float misWeight = 1.0f / (lightWeight + 1.0f + cameraWeight);
    V3 c = misWeight * geometryTerm * cameraBsdf * lightBsdf;
    if (c.x == 0 && c.y == 0 && c.z == 0) return v3(0.0f);
    if (vcmVcOcclusionRay(sc, surf, direction, distance, ctr)) return c;
    return v3(0.0f);
}

struct VcmLightSet {  // LightPathSet, VCM.cpp:57-62, + the hash grid arrays
    std::vector<SelasVcmVertex> vertices;
    std::vector<uint32_t> pathEnds;
    std::vector<float> positions;  // hit.position of every vertex
    std::vector<uint32_t> cellRangeEnds, cellIndices;
    float aabb[6];
};

static void vcmLightPass(const Scene& sc, const SelasCamera& cam, const SelasVcmSettings& vs, const VcmConstants& k, uint32_t pass,
                         float* film, VcmLightSet& set, VcmCounters& total, int nthreads) {
    const uint32_t vmCount = k.vmCount;
    struct Chunk { size_t b; std::vector<SelasVcmVertex> v; std::vector<uint32_t> counts; };
    std::vector<Chunk> chunks;
    std::mutex mu;
    parallelFor(vmCount, nthreads, [&](size_t b, size_t e, int) {
        Chunk c;
        c.b = b;
        VcmCounters ctr;
        for (size_t i = b; i < e; ++i) {
            size_t before = c.v.size();
            vcmLightPath(sc, cam, vs, k, pass, (uint32_t)i, c.v, film, ctr);
            c.counts.push_back((uint32_t)(c.v.size() - before));
        }
        std::lock_guard<std::mutex> g(mu);
        total.closest += ctr.closest;
        total.shadow += ctr.shadow;
        chunks.push_back(std::move(c));
    });
    std::sort(chunks.begin(), chunks.end(), [](const Chunk& a, const Chunk& b) { return a.b < b.b; });
    set.vertices.clear();
    set.pathEnds.clear();
    for (Chunk& c : chunks) {
        set.vertices.insert(set.vertices.end(), c.v.begin(), c.v.end());
        for (uint32_t n : c.counts) set.pathEnds.push_back((set.pathEnds.empty() ? 0u : set.pathEnds.back()) + n);
    }
    set.positions.resize(3 * set.vertices.size());
    for (size_t i = 0; i < set.vertices.size(); ++i) {
        const SelasVcmVertex& v = set.vertices[i];
        set.positions[3 * i] = v.ray.ox + v.hit.t * v.ray.dx;  // HitParameters::position, VCM.cpp:149-151
        set.positions[3 * i + 1] = v.ray.oy + v.hit.t * v.ray.dy;
        set.positions[3 * i + 2] = v.ray.oz + v.hit.t * v.ray.dz;
    }
}

// one camera path: EvaluateCameraPaths loop body, VCM.cpp:578-661
static V3 vcmCameraPath(const Scene& sc, const SelasCamera& cam, const SelasVcmSettings& vs, const VcmConstants& k, uint32_t pass,
                        uint32_t scan, const VcmLightSet& set, VcmCounters& ctr) {
    PathRng rng;
    rng.init(vs.seed, scan, 2u * pass + 1u);
    uint32_t y = scan / (uint32_t)cam.width, x = scan - y * (uint32_t)cam.width;
    VcmPathState cs;
    vcmGenerateCameraSample(rng, cam, x, y, (float)k.vmCount, cs);
    V3 color = v3(0.0f);
    const bool skipNonFinite = (vs.flags & SELAS_VCM_SKIP_NONFINITE) != 0;
    auto add = [&](V3 c) {
        if (skipNonFinite && !finite3(c)) return;
        color = color + c;
    };
    const float inv = 1.0f / (2.0f * k.vmSearchRadius);
    while (cs.pathLength < vs.maxPathLength) {
        SelasRay ray;
        SelasHit h;
        if (!vcmRayPick(sc, cs.position, cs.direction, ray, h, ctr)) {
            if (!(vs.flags & SELAS_VCM_NO_SKY)) add(cs.throughput * vcmConnectToSkyLight(sc, vs.boundingSphere, cs));
            break;
        }
        SelasSurface surf;
        surfaceParams(sc, ray, h, surf);
        float connectionLengthSqr = dot(cs.position - v3(surf.position), cs.position - v3(surf.position));
        float absDotNL = absf(dot(geometricNormal(surf), v3(surf.view)));
        cs.dVCM *= connectionLengthSqr;
        cs.dVCM /= absDotNL;
        cs.dVC /= absDotNL;
        cs.dVM /= absDotNL;
        if (cs.pathLength + 1 < vs.maxPathLength && !(vs.flags & SELAS_VCM_NO_DIRECT))
            add(cs.throughput * vcmConnectCameraPathToLight(rng, sc, vs.boundingSphere, cs, surf, k.vmWeight, ctr));
        else if (cs.pathLength + 1 < vs.maxPathLength) {  // keep the stream position independent of the flag
            rng.next();
            rng.next();
        }
        if (!(vs.flags & SELAS_VCM_NO_CONNECTIONS)) {
            uint32_t vcIndex = cs.index;  // vcCount == 1
            uint32_t pathStart = vcIndex == 0 ? 0 : set.pathEnds[vcIndex - 1];
            uint32_t pathEnd = set.pathEnds[vcIndex];
            for (uint32_t li = pathStart; li < pathEnd; ++li) {
                const SelasVcmVertex& lv = set.vertices[li];
                uint32_t lvLength = (lv.indexAndLength >> 26) & 31u;
                if (lvLength + 1 + cs.pathLength > vs.maxPathLength) break;
                add(cs.throughput * v3(lv.throughput) * vcmConnectPathVertices(sc, surf, cs, lv, k.vmWeight, ctr));
            }
        }
        if (!(vs.flags & SELAS_VCM_NO_MERGING) && !set.vertices.empty()) {
            // SearchHashGrid, VCMHashGrid.cpp:107-155, with MergeVertices (VCM.cpp:362-409) as the callback
            V3 result = v3(0.0f);
            const float* q = surf.position;
            float dmin[3], dmax[3];
            bool outside = false;
            for (int a = 0; a < 3; ++a) {
                dmin[a] = q[a] - set.aabb[a];
                dmax[a] = set.aabb[3 + a] - q[a];
                if (dmin[a] < 0.0f || dmax[a] < 0.0f) outside = true;
            }
            if (!outside) {
                int32_t base[3];
                for (int a = 0; a < 3; ++a) {
                    float cp = inv * dmin[a];
                    float fl = floorf(cp);
                    base[a] = (int32_t)fl;
                    if (cp - fl < 0.5f) --base[a];
                }
                for (int zz = 0; zz <= 1; ++zz)
                    for (int yy = 0; yy <= 1; ++yy)
                        for (int xx = 0; xx <= 1; ++xx) {
                            uint32_t c = cellHash(base[0] + xx, base[1] + yy, base[2] + zz, k.vmCount);
                            uint32_t b = c == 0 ? 0 : set.cellRangeEnds[c - 1], e = set.cellRangeEnds[c];
                            for (uint32_t sIdx = b; sIdx < e; ++sIdx) {
                                uint32_t pi = set.cellIndices[sIdx];
                                ++ctr.mergeCandidates;
                                float dx = q[0] - set.positions[3 * pi], dy = q[1] - set.positions[3 * pi + 1], dz = q[2] - set.positions[3 * pi + 2];
                                float d2 = (dx * dx) + (dy * dy) + (dz * dz);
                                if (!(d2 <= k.vmSearchRadiusSqr)) continue;
                                ++ctr.merges;
                                const SelasVcmVertex& lv = set.vertices[pi];
                                uint32_t lvLength = (lv.indexAndLength >> 26) & 31u;
                                if (cs.pathLength + lvLength > vs.maxPathLength) continue;
                                bool lvTransparent = (lv.indexAndLength & 0x80000000u) != 0;
                                if (((surf.materialFlags & SELAS_MAT_TRANSPARENT) != 0) != lvTransparent) continue;  // VCM.cpp:381
                                float fwd = 0.f, rev = 0.f;
                                V3 lightView = v3(-lv.ray.dx, -lv.ray.dy, -lv.ray.dz);  // lightSurface.view = hit.incDirection
                                V3 bsdf = evaluateBsdf(surf, -cs.direction, lightView, fwd, rev);
                                if (bsdf.x == 0 && bsdf.y == 0 && bsdf.z == 0) continue;
                                fwd *= continuationProbability(surf);
                                rev *= lv.continuationProbability;
                                float lightWeight = lv.dVCM * k.vcWeight + lv.dVM * fwd;
                                float cameraWeight = cs.dVCM * k.vcWeight + cs.dVM * rev;
                                float misWeight = 1.0f / (lightWeight + 1.0f + cameraWeight);
                                result = result + misWeight * bsdf * v3(lv.throughput);
                            }
                        }
            }
            add(cs.throughput * k.vmNormalization * result);
        }
        if (!vcmSampleBsdfScattering(rng, surf, k.vmWeight, k.vcWeight, cs)) break;
    }
    return color;
}

static int vcmCheck(const Scene& sc, const SelasVcmSettings& vs) {
    if (!sc.hasIbl) return SELAS_ERR_ARGUMENT;
    if (vs.width == 0 || vs.height == 0 || vs.maxPathLength > 31 || (uint64_t)vs.width * vs.height >= (1u << 26)) return SELAS_ERR_ARGUMENT;
    if (!(vs.boundingSphere[3] > 0.0f) || !(vs.radiusFactor > 0.0f)) return SELAS_ERR_ARGUMENT;
    return SELAS_OK;
}

}  // namespace

extern "C" {

int orc_vcm_iteration_constants(uint32_t vmCount, uint32_t vcCount, float baseRadius, float alpha, float iteration, float* out5) {
    VcmConstants c = vcmIterationConstants(vmCount, vcCount, baseRadius, alpha, iteration);
    out5[0] = c.vmSearchRadius; out5[1] = c.vmSearchRadiusSqr; out5[2] = c.vmNormalization; out5[3] = c.vmWeight; out5[4] = c.vcWeight;
    return 0;
}

// scripted-stream forms of the IBL emission functions, for pinning against the reference objects
int orc_ibl_emit(void* s, const float* sphere4, size_t n, const float* u4, float* posDirRad9, float* pdfs3) {
    const Scene& sc = *(Scene*)s;
    if (!sc.hasIbl) return SELAS_ERR_ARGUMENT;
    setFtz();
    for (size_t i = 0; i < n; ++i) {
        Stream st{u4 + 4 * i, 0};
        LightEmission e;
        emitIblLightSample(st, sc.ibl, sphere4, e);
        store(posDirRad9 + 9 * i, e.position);
        store(posDirRad9 + 9 * i + 3, e.direction);
        store(posDirRad9 + 9 * i + 6, e.radiance);
        pdfs3[3 * i] = e.emissionPdfW; pdfs3[3 * i + 1] = e.directionPdfA; pdfs3[3 * i + 2] = e.cosThetaLight;
    }
    return 0;
}
int orc_ibl_direct(void* s, size_t n, const float* u2, SelasLightSample* out) {
    const Scene& sc = *(Scene*)s;
    if (!sc.hasIbl) return SELAS_ERR_ARGUMENT;
    setFtz();
    for (size_t i = 0; i < n; ++i) {
        Stream st{u2 + 2 * i, 0};
        directIblLightSample(st, sc.ibl, out[i]);
    }
    return 0;
}
int orc_ibl_radiance(void* s, const float* sphere4, size_t n, const float* dir3, float* rgb3, float* pdfs2) {
    const Scene& sc = *(Scene*)s;
    if (!sc.hasIbl) return SELAS_ERR_ARGUMENT;
    setFtz();
    for (size_t i = 0; i < n; ++i) {
        float a, b;
        store(rgb3 + 3 * i, iblCalculateRadiance(sc.ibl, sphere4, v3(dir3 + 3 * i), a, b));
        pdfs2[2 * i] = a; pdfs2[2 * i + 1] = b;
    }
    return 0;
}
int orc_world_to_image(const SelasCamera* cam, size_t n, const float* p3, int32_t* xy2) {
    for (size_t i = 0; i < n; ++i) worldToImage(*cam, v3(p3 + 3 * i), xy2[2 * i], xy2[2 * i + 1]);
    return 0;
}

// light pass of one iteration (inspection form of selas_cuda_vcm_light_pass)
int orc_vcm_light_pass(void* s, const SelasCamera* cam, const SelasVcmSettings* vs, SelasVcmVertex* vertices, size_t maxVertices,
                       uint32_t* pathEnds, size_t* vertexCount, float* film, int nthreads) {
    const Scene& sc = *(Scene*)s;
    int rc = vcmCheck(sc, *vs);
    if (rc) return rc;
    const uint32_t vmCount = vs->width * vs->height;
    VcmConstants k = vcmIterationConstants(vmCount, 1, vs->radiusFactor * vs->boundingSphere[3], vs->radiusAlpha, (float)vs->firstIteration + 1.0f);
    std::vector<float> scratch;
    if (!film) {
        scratch.assign((size_t)vmCount * 3, 0.f);
        film = scratch.data();
    }
    VcmLightSet set;
    VcmCounters ctr;
    vcmLightPass(sc, *cam, *vs, k, vs->firstIteration, film, set, ctr, nthreads);
    if (vertexCount) *vertexCount = set.vertices.size();
    if (vertices) memcpy(vertices, set.vertices.data(), sizeof(SelasVcmVertex) * std::min(maxVertices, set.vertices.size()));
    if (pathEnds) memcpy(pathEnds, set.pathEnds.data(), sizeof(uint32_t) * vmCount);
    return 0;
}

// VCM::GenerateImage (VCM.cpp:738-807) minus FrameBuffer_Normalize and the .hdr save; film = sum over iterations
int orc_render_vcm(void* s, const SelasCamera* cam, const SelasVcmSettings* vs, float* film, SelasVcmStats* stats, int nthreads) {
    const Scene& sc = *(Scene*)s;
    int rc = vcmCheck(sc, *vs);
    if (rc) return rc;
    const uint32_t vmCount = vs->width * vs->height;
    memset(film, 0, sizeof(float) * 3 * (size_t)vmCount);
    SelasVcmStats st;
    memset(&st, 0, sizeof(st));
    for (uint32_t it = 0; it < vs->iterations; ++it) {
        const uint32_t pass = vs->firstIteration + it;
        VcmConstants k = vcmIterationConstants(vmCount, 1, vs->radiusFactor * vs->boundingSphere[3], vs->radiusAlpha, (float)pass + 1.0f);
        VcmLightSet set;
        VcmCounters ctr;
        vcmLightPass(sc, *cam, *vs, k, pass, film, set, ctr, nthreads);
        set.cellRangeEnds.resize(vmCount);
        set.cellIndices.resize(set.vertices.size() ? set.vertices.size() : 1);
        orc_hashgrid_build(set.vertices.size(), set.positions.data(), k.vmSearchRadius, vmCount, set.cellRangeEnds.data(), set.cellIndices.data(),
                           set.aabb);
        std::mutex mu;
        parallelFor(vmCount, nthreads, [&](size_t b, size_t e, int) {
            VcmCounters c;
            for (size_t i = b; i < e; ++i) {
                V3 color = vcmCameraPath(sc, *cam, *vs, k, pass, (uint32_t)i, set, c);
                // FramebufferWriter_Write(&writer, &color, 1, cameraPathState.index), VCM.cpp:659: light-tracing splats of this pass
                // land in the same film, hence the atomic
                atomicAddFloat(film + 3 * i, color.x);
                atomicAddFloat(film + 3 * i + 1, color.y);
                atomicAddFloat(film + 3 * i + 2, color.z);
            }
            std::lock_guard<std::mutex> g(mu);
            ctr.closest += c.closest; ctr.shadow += c.shadow; ctr.mergeCandidates += c.mergeCandidates; ctr.merges += c.merges;
            ctr.connections += c.connections;
        });
        st.lightPaths += vmCount;
        st.cameraPaths += vmCount;
        st.lightVertices += set.vertices.size();
        st.closestRays += ctr.closest;
        st.occlusionRays += ctr.shadow;
        st.mergeCandidates += ctr.mergeCandidates;
        st.mergesAccepted += ctr.merges;
        st.connections += ctr.connections;
        st.lastRadius = k.vmSearchRadius;
    }
    if (stats) *stats = st;
    return 0;
}

}  // extern "C"
