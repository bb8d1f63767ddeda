// oracle/ref_harness.h -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
// Pieces shared by the two harness translation units that sit on top of the UNMODIFIED reference objects of oracle/_ref:
// ref_harness.cpp (operators, restated loops, blob writers) and ref_pathtracer.cpp (the reference's own PathTracer.cpp +
// SurfaceParameters.cpp behind an Embree-API shim).  The scripted uniform source, the scene the harness keeps, and a
// GIIntegratorContext carved out of zeroed storage.
#pragma once
#include <cstdint>
#include <cstring>
#include <random>
#include <vector>

#include <xmmintrin.h>
#include <pmmintrin.h>

#include "Shading/IntegratorContexts.h"
#include "Shading/AreaLighting.h"
#include "SceneLib/SceneResource.h"
#include "SceneLib/ImageBasedLightResource.h"
#include "MathLib/FloatStructs.h"

#include "../include/selas_cuda.h"
#include "cpu_bvh.h"

namespace refh {
using namespace Selas;

struct DrawSource {
    const float* script;  // scripted values (bsdf/light unit tests) ...
    uint32_t pos, count;
    uint32_t pcg;  // ... or the product's PCG stream (integrator)
    bool usePcg;
};
extern thread_local DrawSource g_src;  // defined in ref_harness.cpp
extern thread_local std::mt19937* g_mt;

inline uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
inline float pcgNext(uint32_t& state) {
    state = state * 747796405u + 2891336453u;
    uint32_t w = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
    w = (w >> 22u) ^ w;
    return (float)(w >> 8) * (1.0f / 16777216.0f);
}
inline void setFtz() {
    _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);
    _MM_SET_DENORMALS_ZERO_MODE(_MM_DENORMALS_ZERO_ON);
}


// A GIIntegratorContext + SceneResource carved out of zeroed storage: their constructors live in TUs that
// need Embree, and only `sampler`, `scene->lightSets`, `scene->data->backgroundIntensity`, `scene->iblResource`
// are read by the functions called here.
struct FakeContext {
    alignas(16) unsigned char ctxBytes[sizeof(GIIntegratorContext)];
    alignas(16) unsigned char sceneBytes[sizeof(SceneResource)];
    alignas(16) unsigned char dataBytes[sizeof(SceneResourceData)];
    std::vector<SceneLight> lights;
    std::vector<SceneLightSet> sets;
    alignas(16) unsigned char iblResBytes[sizeof(ImageBasedLightResource)];
    ImageBasedLightResourceData iblData;
    void setIbl(const ImageBasedLightResourceData* src) {  // borrowed arrays; the resource struct is never destructed
        if (!src) return;
        iblData = *src;
        ImageBasedLightResource* res = reinterpret_cast<ImageBasedLightResource*>(iblResBytes);
        res->data = &iblData;
        scene()->iblResource = res;
    }
    GIIntegratorContext* ctx() { return reinterpret_cast<GIIntegratorContext*>(ctxBytes); }
    SceneResource* scene() { return reinterpret_cast<SceneResource*>(sceneBytes); }
    SceneResourceData* data() { return reinterpret_cast<SceneResourceData*>(dataBytes); }

    void init(const SelasLight* L, uint32_t lightCount, const SelasLightSetRange* ranges, uint32_t setCount,
              const float* background) {
        memset(ctxBytes, 0, sizeof(ctxBytes));
        memset(sceneBytes, 0, sizeof(sceneBytes));
        memset(dataBytes, 0, sizeof(dataBytes));
        lights.resize(lightCount);
        if (lightCount) memcpy(lights.data(), L, sizeof(SceneLight) * lightCount);
        sets.clear();
        for (uint32_t i = 0; i < setCount; ++i) {
            SceneLightSet s;
            s.count = ranges[i].count;
            s.lights = lights.data() + ranges[i].start;
            sets.push_back(s);
        }
        // CArray<SceneLightSet> laid out as {flags,pad,count,capacity,data} (ContainersLib/CArray.h:57-62)
        struct RawArray { uint32_t flags, pad; uint64_t count, capacity; void* data; };
        static_assert(sizeof(RawArray) == sizeof(CArray<SceneLightSet>), "CArray layout");
        RawArray raw = {0, 0, sets.size(), sets.size(), sets.data()};
        memcpy(&scene()->lightSets, &raw, sizeof(raw));
        if (background) data()->backgroundIntensity = float4(background[0], background[1], background[2], background[3]);
        scene()->data = data();
        scene()->iblResource = nullptr;
        ctx()->scene = scene();
    }
    ~FakeContext() {
        // make the borrowed CArray look empty so nothing is freed
        memset(&scene()->lightSets, 0, sizeof(scene()->lightSets));
    }
};

inline float3 f3(const float* p) { return float3(p[0], p[1], p[2]); }
inline void st3(float* p, float3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

struct RefIbl {
    std::vector<float> marginal, conditional, light, miss;
    ImageBasedLightResourceData data;
    bool present = false;
    void set(const SelasIblDesc* I) {
        if (!I) return;
        marginal.assign(I->marginalDensityFunction, I->marginalDensityFunction + I->height);
        conditional.assign(I->conditionalDensityFunctions, I->conditionalDensityFunctions + (size_t)I->width * I->height);
        // the reference clamps texel coordinates to [0, width] / [0, height] (appendix B #23) and can read one row past the image;
        // pad one row (a copy of the last texel, which is what the restatement and the CUDA path read) so that stays defined
        light.assign(I->lightData, I->lightData + 3 * (size_t)I->width * I->height);
        miss.assign(I->missData, I->missData + 3 * (size_t)I->missWidth * I->missHeight);
        auto pad = [](std::vector<float>& img, size_t w) {
            float last[3] = {img[img.size() - 3], img[img.size() - 2], img[img.size() - 1]};
            for (size_t i = 0; i < w + 1; ++i) img.insert(img.end(), last, last + 3);
        };
        pad(light, I->width);
        pad(miss, I->missWidth);
        float lastC = conditional.back();
        conditional.resize(conditional.size() + I->width + 1, lastC);
        marginal.push_back(marginal.back());
        memset(&data, 0, sizeof(data));
        data.densityfunctions.width = I->width;
        data.densityfunctions.height = I->height;
        data.densityfunctions.marginalDensityFunction = marginal.data();
        data.densityfunctions.conditionalDensityFunctions = conditional.data();
        data.missWidth = I->missWidth;
        data.missHeight = I->missHeight;
        data.rotationRadians = I->rotationRadians;
        data.exposureScale = I->exposureScale;
        data.lightData = reinterpret_cast<float3*>(light.data());
        data.missData = reinterpret_cast<float3*>(miss.data());
        present = true;
    }
};

struct RefScene {
    RefIbl ibl;
    struct Mesh {
        std::vector<float> positions, normals, tangents, uvs;
        std::vector<uint32_t> indices;
        uint32_t indicesPerFace, materialIndex, lightSetIndex;
    };
    std::vector<Mesh> meshes;
    std::vector<SelasMaterial> materials;
    std::vector<SelasLight> lights;
    std::vector<SelasLightSetRange> sets;
    float background[4];
    orc::Bvh bvh;
};

}  // namespace refh
