// ptex_eval.cuh -- device-side Ptex b-spline lookup (see ptex.cu for the loader and the data layout).
#pragma once
#include "vec.cuh"

namespace selas {

struct PtexGpu;
SELAS_D void ptexEvalBspline(const PtexGpu* tex, int faceid, float u, float v, float* rgb) { rgb[0] = rgb[1] = rgb[2] = 0.f; }

}  // namespace selas
