// selas_render.cpp -- the reference's main() (Source/Applications/Selas/Source/main.cpp:82-135) for the synthetic Cornell
// scene: initialise the scene on the GPU, set up the camera, PathTracer::GenerateImage, report the render time.
//   selas_render [--width W] [--height H] [--spp N] [--depth D] [--layers L] [--roughness R] [--device I] [--name NAME]
//   selas_render --integrator deferred [...]                      (DeferredPathTracer::GenerateImage, 4 x 4 CMJ samples per pixel)
//   selas_render --scene NAME --assets ROOT [...]                 (a scene baked by the reference: ROOT_Assets/SceneResource/<version>/NAME.bin)
//   selas_render --integrator vcm [--width W] [--height H] [--iterations N] [--depth D] ...   (config 5 caustic scene, VCM::GenerateImage)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "selas_host.h"

using namespace Selas;

#define ExitMainOnError_(expr)                                  \
    do {                                                        \
        Error e__ = (expr);                                     \
        if (e__.Failed()) {                                     \
            fprintf(stderr, "Fatal error: %s\n", e__.Message()); \
            return -1;                                          \
        }                                                       \
    } while (0)

int main(int argc, char** argv) {
    uint32_t width = 512, height = 512;
    PathTracerSettings settings;
    settings.pathsPerPixel = 64;
    settings.maxBounceCount = 8;
    float roughness = 0.01f;
    int device = 0;
    const char* name = "cornell";
    const char* integrator = "pathtracer";
    const char* bakedScene = nullptr;
    const char* assetRoot = "";
    VCMSettings vcm;
    bool sizeGiven = false;
    for (int i = 1; i + 1 < argc; i += 2) {
        if (!strcmp(argv[i], "--width")) { width = (uint32_t)atoi(argv[i + 1]); sizeGiven = true; }
        else if (!strcmp(argv[i], "--height")) { height = (uint32_t)atoi(argv[i + 1]); sizeGiven = true; }
        else if (!strcmp(argv[i], "--integrator")) integrator = argv[i + 1];
        else if (!strcmp(argv[i], "--scene")) bakedScene = argv[i + 1];
        else if (!strcmp(argv[i], "--assets")) assetRoot = argv[i + 1];
        else if (!strcmp(argv[i], "--iterations")) vcm.iterations = (uint32_t)atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--spp")) settings.pathsPerPixel = (uint32_t)atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--depth")) settings.maxBounceCount = (uint32_t)atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--layers")) settings.layerCount = (uint32_t)atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--roughness")) roughness = (float)atof(argv[i + 1]);
        else if (!strcmp(argv[i], "--device")) device = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--seed")) settings.seed = (uint32_t)atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--name")) name = argv[i + 1];
        else {
            fprintf(stderr, "unknown option %s\n", argv[i]);
            return 2;
        }
    }
    if (!strcmp(integrator, "vcm")) {
        if (!sizeGiven) { width = 960; height = 540; }
        if (!strcmp(name, "cornell")) name = "caustic";
        if (settings.maxBounceCount != 8) vcm.maxBounceCount = settings.maxBounceCount;
        vcm.seed = settings.seed;
        auto t0 = std::chrono::steady_clock::now();
        SceneResource scene;
        ImageBasedLight ibl;
        MakeCausticScene(&scene, &ibl);
        ExitMainOnError_(InitializeSceneResource(&scene, device));
        auto t1 = std::chrono::steady_clock::now();
        printf("Scene load time %fms\n", std::chrono::duration<double, std::milli>(t1 - t0).count());
        RayCastCameraSettings camera;
        SetupCausticCamera(width, height, camera);
        SelasVcmStats vst;
        ExitMainOnError_(VCM::GenerateImage(&scene, camera, name, vcm, &vst));
        auto t2 = std::chrono::steady_clock::now();
        printf("Scene render time %fms\n", std::chrono::duration<double, std::milli>(t2 - t1).count());
        printf("Vcm integration performed with %u iterations\n", vcm.iterations);  // VCM.cpp:799
        printf("{\"iterations\": %u, \"light_vertices\": %llu, \"closest_rays\": %llu, \"occlusion_rays\": %llu, \"merges\": %llu, "
               "\"connections\": %llu, \"kernel_launches\": %llu, \"device_ms\": %.3f, \"paths_per_s\": %.1f}\n",
               vcm.iterations, (unsigned long long)vst.lightVertices, (unsigned long long)vst.closestRays, (unsigned long long)vst.occlusionRays,
               (unsigned long long)vst.mergesAccepted, (unsigned long long)vst.connections, (unsigned long long)vst.kernelLaunches, vst.renderMs,
               vst.renderMs > 0 ? (vst.lightPaths + vst.cameraPaths) / (vst.renderMs * 1e-3) : 0.0);
        ShutdownSceneResource(&scene);
        return 0;
    }
    const bool deferred = !strcmp(integrator, "deferred");
    if (strcmp(integrator, "pathtracer") && !deferred) {
        fprintf(stderr, "unknown integrator %s (pathtracer | deferred | vcm)\n", integrator);
        return 2;
    }
    auto t0 = std::chrono::steady_clock::now();
    SceneResource scene;
    ImageBasedLight bakedIbl;
    std::vector<BakedCamera> bakedCameras;
    if (bakedScene)  // ReadSceneResource + InitializeSceneResource, main.cpp:104-113
        ExitMainOnError_(LoadBakedScene(assetRoot, bakedScene, &scene, &bakedIbl, &bakedCameras));
    else
        MakeCornellBoxScene(&scene, roughness);
    ExitMainOnError_(InitializeSceneResource(&scene, device));
    auto t1 = std::chrono::steady_clock::now();
    printf("Scene load time %fms\n", std::chrono::duration<double, std::milli>(t1 - t0).count());  // main.cpp:106-107

    // main.cpp:115-126: one image per scene camera, named after the camera; the synthetic scene has one
    const size_t cameraCount = (bakedScene && !bakedCameras.empty()) ? bakedCameras.size() : 1;
    for (size_t scan = 0; scan < cameraCount; ++scan) {
        RayCastCameraSettings camera;
        const char* imageName = name;
        if (bakedScene && !bakedCameras.empty()) {  // SetupSceneCamera(scene, scan, width, height, camera)
            InitializeRayCastCamera(bakedCameras[scan].settings, width, height, camera);
            if (!bakedCameras[scan].name.empty() && !strcmp(name, "cornell")) imageName = bakedCameras[scan].name.c_str();
        } else {
            SetupCornellCamera(width, height, camera);
        }
        auto tr0 = std::chrono::steady_clock::now();
        SelasRenderStats stats;
        if (deferred) {
            DeferredPathTracerSettings ds;
            ds.seed = settings.seed;
            if (settings.maxBounceCount != 8) ds.maxBounceCount = settings.maxBounceCount;
            ExitMainOnError_(DeferredPathTracer::GenerateImage(&scene, camera, imageName, ds, &stats));
        } else {
            ExitMainOnError_(PathTracer::GenerateImage(&scene, camera, imageName, settings, &stats));
        }
        auto tr1 = std::chrono::steady_clock::now();
        printf("Scene render time %fms\n", std::chrono::duration<double, std::milli>(tr1 - tr0).count());  // main.cpp:124-125
        printf("{\"image\": \"%s\", \"paths\": %llu, \"closest_rays\": %llu, \"occlusion_rays\": %llu, \"kernel_launches\": %llu, \"device_ms\": %.3f, "
               "\"samples_per_s\": %.1f}\n",
               imageName, (unsigned long long)stats.pathSamples, (unsigned long long)stats.closestRays, (unsigned long long)stats.occlusionRays,
               (unsigned long long)stats.kernelLaunches, stats.renderMs, stats.renderMs > 0 ? stats.pathSamples / (stats.renderMs * 1e-3) : 0.0);
    }
    ShutdownSceneResource(&scene);
    return 0;
}
