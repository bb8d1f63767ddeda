// vec.cuh -- float3 helpers in the reference's conventions (row vectors, y-up tangent space).
// Operation order follows Source/Core/MathLib/FloatFuncs.h so that, compiled with -fmad=false, the device
// arithmetic matches the CPU reference to the last bit wherever libm is not involved.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SELAS_HD __host__ __device__ __forceinline__
#define SELAS_D __device__ __forceinline__
// Shared out-of-line helpers: the fused shade kernel is instruction-cache bound when everything is inlined
// (ncu: 71 % of its stall samples were "no instruction", profiles/r1a), so the larger leaf functions are kept as calls.
#ifndef SELAS_OUTLINE
#define SELAS_OUTLINE 1
#endif
#if SELAS_OUTLINE
#define SELAS_HDN __host__ __device__ __noinline__
#else
#define SELAS_HDN __host__ __device__ __forceinline__
#endif

namespace selas {

struct V3 {
    float x, y, z;
};
SELAS_HD V3 v3(float x, float y, float z) { return V3{x, y, z}; }
SELAS_HD V3 v3(float s) { return V3{s, s, s}; }
SELAS_HD V3 v3(const float* p) { return V3{p[0], p[1], p[2]}; }
SELAS_HD void store3(float* p, V3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
SELAS_HD V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
SELAS_HD V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
SELAS_HD V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
SELAS_HD V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
SELAS_HD V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
SELAS_HD V3 operator*(float s, V3 a) { return V3{a.x * s, a.y * s, a.z * s}; }
SELAS_HD V3 operator/(V3 a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }  // FloatFuncs.h:111-115
// FloatFuncs.h:147-151 defines float/float3 with the operands swapped (rhs / dividend); kept as a named
// function so the quirk is visible at its one call site (Disney.cpp:342).
SELAS_HD V3 refScalarOverVec(float s, V3 a) { return V3{a.x / s, a.y / s, a.z / s}; }
SELAS_HD float dot(V3 a, V3 b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }  // FloatFuncs.h:213
SELAS_HD V3 cross(V3 a, V3 b) {                                                      // FloatFuncs.h:198
    return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
SELAS_HD float length(V3 a) { return sqrtf(dot(a, a)); }
SELAS_HD V3 normalize(V3 a) {  // FloatFuncs.h:401-407: multiply by 1/length
    float inv = 1.f / length(a);
    return V3{a.x * inv, a.y * inv, a.z * inv};
}
SELAS_HD float sq(float x) { return x * x; }
SELAS_HD float maxf_(float a, float b) { return (a > b) ? a : b; }  // SystemLib/MinMax.h:16 (NaN-order kept)
SELAS_HD float minf_(float a, float b) { return (a < b) ? a : b; }
SELAS_HD float clampf_(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
SELAS_HD float saturate_(float x) { return x < 0.f ? 0.f : (x > 1.f ? 1.f : x); }
SELAS_HD float lerpf(float a, float b, float t) { return (1.0f - t) * a + t * b; }  // FloatFuncs.h:369
SELAS_HD V3 lerp3(V3 a, V3 b, float t) { return (1.0f - t) * a + t * b; }
SELAS_HD V3 sqrt3(V3 a) { return V3{sqrtf(a.x), sqrtf(a.y), sqrtf(a.z)}; }
SELAS_HD float signf_(float x) { return x < 0.f ? -1.f : (x > 0.f ? 1.f : 0.f); }  // Trigonometric.cpp:121

#define SELAS_PI 3.141592653589793f
#define SELAS_TWO_PI (SELAS_PI * 2.0f)
#define SELAS_INV_PI (1.0f / SELAS_PI)
#define SELAS_INV_4PI (1.0f / (4.0f * SELAS_PI))
#define SELAS_FLOAT_MAX 3.402823466e+38F

// v * M for a 3x3 stored as rows r0,r1,r2 (FloatFuncs.cpp:356-364) and v * transpose(M)
SELAS_HD V3 mul33(V3 v, const float* m) {
    return V3{v.x * m[0] + v.y * m[3] + v.z * m[6], v.x * m[1] + v.y * m[4] + v.z * m[7],
              v.x * m[2] + v.y * m[5] + v.z * m[8]};
}
SELAS_HD V3 mul33t(V3 v, const float* m) {
    return V3{v.x * m[0] + v.y * m[1] + v.z * m[2], v.x * m[3] + v.y * m[4] + v.z * m[5],
              v.x * m[6] + v.y * m[7] + v.z * m[8]};
}
// MatrixMultiplyVector / MatrixMultiplyPoint with a row-major 4x4 (FloatFuncs.cpp:367-396)
SELAS_HD V3 mulVec44(V3 v, const float* m) {
    return V3{v.x * m[0] + v.y * m[4] + v.z * m[8], v.x * m[1] + v.y * m[5] + v.z * m[9],
              v.x * m[2] + v.y * m[6] + v.z * m[10]};
}
SELAS_HD V3 mulPoint44(V3 v, const float* m) {
    return V3{v.x * m[0] + v.y * m[4] + v.z * m[8] + m[12], v.x * m[1] + v.y * m[5] + v.z * m[9] + m[13],
              v.x * m[2] + v.y * m[6] + v.z * m[10] + m[14]};
}
SELAS_HD V3 reflect_(V3 n, V3 l) { return 2.0f * dot(n, l) * n - l; }  // FloatFuncs.h:348
SELAS_HD bool transmit_(V3 wm, V3 wi, float n, V3& wo) {               // FloatFuncs.h:353-367
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

// tangent-space trigonometry, y up (MathLib/Trigonometric.h:51-108)
SELAS_HD float cosTheta(V3 w) { return w.y; }
SELAS_HD float cos2Theta(V3 w) { return w.y * w.y; }
SELAS_HD float absCosTheta(V3 w) { return fabsf(w.y); }
SELAS_HD float sin2Theta(V3 w) { return maxf_(0.0f, 1.0f - cos2Theta(w)); }
SELAS_HD float sinTheta(V3 w) { return sqrtf(sin2Theta(w)); }
SELAS_HD float tanTheta(V3 w) { return sinTheta(w) / cosTheta(w); }
SELAS_HD float cosPhi(V3 w) {
    float s = sinTheta(w);
    return (s == 0) ? 1.0f : clampf_(w.x / s, -1.0f, 1.0f);
}
SELAS_HD float sinPhi(V3 w) {
    float s = sinTheta(w);
    return (s == 0) ? 1.0f : clampf_(w.z / s, -1.0f, 1.0f);
}

}  // namespace selas
