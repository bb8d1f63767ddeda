// selas_render.cpp -- the reference's main() (Source/Applications/Selas/Source/main.cpp:82-135) for the synthetic Cornell
// scene: initialise the scene on the GPU, set up the camera, PathTracer::GenerateImage, report the render time.
//   selas_render [--width W] [--height H] [--spp N] [--depth D] [--layers L] [--roughness R] [--device I] [--name NAME]
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
    for (int i = 1; i + 1 < argc; i += 2) {
        if (!strcmp(argv[i], "--width")) width = (uint32_t)atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--height")) height = (uint32_t)atoi(argv[i + 1]);
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
    auto t0 = std::chrono::steady_clock::now();
    SceneResource scene;
    MakeCornellBoxScene(&scene, roughness);
    ExitMainOnError_(InitializeSceneResource(&scene, device));
    auto t1 = std::chrono::steady_clock::now();
    printf("Scene load time %fms\n", std::chrono::duration<double, std::milli>(t1 - t0).count());  // main.cpp:106-107

    RayCastCameraSettings camera;
    SetupCornellCamera(width, height, camera);
    SelasRenderStats stats;
    ExitMainOnError_(PathTracer::GenerateImage(&scene, camera, name, settings, &stats));
    auto t2 = std::chrono::steady_clock::now();
    printf("Scene render time %fms\n", std::chrono::duration<double, std::milli>(t2 - t1).count());  // main.cpp:124-125
    printf("{\"paths\": %llu, \"closest_rays\": %llu, \"occlusion_rays\": %llu, \"kernel_launches\": %llu, \"device_ms\": %.3f, "
           "\"samples_per_s\": %.1f}\n",
           (unsigned long long)stats.pathSamples, (unsigned long long)stats.closestRays, (unsigned long long)stats.occlusionRays,
           (unsigned long long)stats.kernelLaunches, stats.renderMs, stats.renderMs > 0 ? stats.pathSamples / (stats.renderMs * 1e-3) : 0.0);
    ShutdownSceneResource(&scene);
    return 0;
}
