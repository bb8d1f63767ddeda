// texture.cu -- the reference's texture filters over a mip chain (TextureLib/TextureFiltering.h:54-234): Point, Triangle at
// a level, Trilinear, EWA.  Compiled like the shading units (-fmad=false, IEEE division / sqrt) and written in the reference's operation
// order, so everything except log2f matches the CPU bit for bit; the EWA weight table comes from the host (the C library's expf, as in
// InitializeEWAFilterWeights, TextureFiltering.cpp:14-21).
#include "internal.h"
#include "vec.cuh"

namespace selas {

__constant__ float cEwaLut[128];

struct TextureMipsGpu {
    const uint8_t* texels;
    uint32_t mipCount, comps;
    uint32_t w[SELAS_TEXTURE_MAX_MIPS], h[SELAS_TEXTURE_MAX_MIPS];
    uint64_t offset[SELAS_TEXTURE_MAX_MIPS];
};

struct Texel4 {
    float c[4];
};

// TextureFiltering::Sample with WrapMode::Repeat (:35-52).  QUIRK: `s % w` mixes int32 and uint32, so a negative coordinate is reduced as
// a 32-bit unsigned number (-1 wraps to w - 1 only when w is a power of two).
SELAS_D Texel4 mipSample(const TextureMipsGpu& t, uint32_t level, uint32_t w, uint32_t h, int32_t s, int32_t tt) {
    const uint32_t ss = (uint32_t)s % w, ts = (uint32_t)tt % h;
    const float* p = reinterpret_cast<const float*>(t.texels + t.offset[level]) + ((size_t)ts * w + ss) * t.comps;
    Texel4 r;
#pragma unroll
    for (int c = 0; c < 4; ++c) r.c[c] = c < (int)t.comps ? __ldg(p + c) : 0.0f;
    return r;
}

// Point (:55-74)
SELAS_D Texel4 filterPoint(const TextureMipsGpu& t, float sx, float sy) {
    const uint32_t w = t.w[0], h = t.h[0];
    const float s = sx * w, tt = sy * h;
    return mipSample(t, 0, w, h, (int32_t)floorf(s), (int32_t)floorf(tt));
}

// Triangle (:77-104): level = Min<uint32>(level, mipCount - 1); four taps weighted (1-ds)(1-dt), (1-ds)dt, ds(1-dt), ds dt, summed in that order
SELAS_D Texel4 filterTriangle(const TextureMipsGpu& t, int32_t levelIn, float sx, float sy) {
    const uint32_t level = min((uint32_t)levelIn, t.mipCount - 1u);
    const uint32_t w = t.w[level], h = t.h[level];
    const float s = sx * w - 0.5f, tt = sy * h - 0.5f;
    const int32_t s0 = (int32_t)floorf(s), t0 = (int32_t)floorf(tt);
    const float ds = s - s0, dt = tt - t0;
    const float w00 = (1 - ds) * (1 - dt), w01 = (1 - ds) * dt, w10 = ds * (1 - dt), w11 = ds * dt;
    const Texel4 a = mipSample(t, level, w, h, s0, t0), b = mipSample(t, level, w, h, s0, t0 + 1), c = mipSample(t, level, w, h, s0 + 1, t0),
                 d = mipSample(t, level, w, h, s0 + 1, t0 + 1);
    Texel4 r;
#pragma unroll
    for (int k = 0; k < 4; ++k) r.c[k] = w00 * a.c[k] + w01 * b.c[k] + w10 * c.c[k] + w11 * d.c[k];
    return r;
}

// EWA at one level (:107-170)
SELAS_D Texel4 filterEwaLevel(const TextureMipsGpu& t, int32_t reqLevel, float sx, float sy, float d0x, float d0y, float d1x, float d1y) {
    if (reqLevel >= (int32_t)t.mipCount) return mipSample(t, t.mipCount - 1u, 1u, 1u, 0, 0);
    // QUIRK: a negative level is not guarded in the reference (it would index mipOffsets[-1]); callers of the full EWA never produce one
    const uint32_t level = (uint32_t)max(reqLevel, 0);
    const uint32_t w = t.w[level], h = t.h[level];
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
                const int32_t index = min((int32_t)(r2 * 128), 127);
                const float weight = cEwaLut[index];
                const Texel4 x = mipSample(t, level, w, h, is, it);
#pragma unroll
                for (int k = 0; k < 4; ++k) sum[k] += x.c[k] * weight;
                sumWts += weight;
            }
        }
    }
    const float inv = 1.0f / sumWts;
    Texel4 r;
#pragma unroll
    for (int k = 0; k < 4; ++k) r.c[k] = sum[k] * inv;
    return r;
}

SELAS_D Texel4 lerpTexel(const Texel4& a, const Texel4& b, float x) {  // Lerp, FloatFuncs.h:369-387: (1 - t) a + t b
    Texel4 r;
#pragma unroll
    for (int k = 0; k < 4; ++k) r.c[k] = (1.0f - x) * a.c[k] + x * b.c[k];
    return r;
}

// Trilinear (:173-194)
SELAS_D Texel4 filterTrilinear(const TextureMipsGpu& t, float sx, float sy, float d0x, float d0y, float d1x, float d1y) {
    const float majorLength = sqrtf(d0x * d0x + d0y * d0y), minorLength = sqrtf(d1x * d1x + d1y * d1y);
    const float length = fminf(majorLength, minorLength);
    if (length == 0) return filterTriangle(t, 0, sx, sy);
    const float lod = fmaxf(0.0f, t.mipCount - 1.0f + log2f(length));
    const float ilod = floorf(lod);
    const Texel4 r0 = filterTriangle(t, (int32_t)ilod, sx, sy), r1 = filterTriangle(t, (int32_t)ilod + 1, sx, sy);
    return lerpTexel(r0, r1, lod - ilod);
}

// EWA (:197-234)
SELAS_D Texel4 filterEwa(const TextureMipsGpu& t, float sx, float sy, float d0x, float d0y, float d1x, float d1y) {
    if (d0x * d0x + d0y * d0y < d1x * d1x + d1y * d1y) {
        float a = d0x, b = d0y;
        d0x = d1x; d0y = d1y;
        d1x = a; d1y = b;
    }
    const float majorLength = sqrtf(d0x * d0x + d0y * d0y);
    float minorLength = sqrtf(d1x * d1x + d1y * d1y);
    const float maxAnisotropy = 16.0f;
    if (minorLength * maxAnisotropy < majorLength && minorLength > 0) {
        const float scale = majorLength / (minorLength * maxAnisotropy);
        d1x = d1x * scale;
        d1y = d1y * scale;
        minorLength *= scale;
    }
    if (minorLength == 0) return filterTriangle(t, 0, sx, sy);
    const float lod = fmaxf(0.0f, t.mipCount - 1.0f + log2f(minorLength));
    const float ilod = floorf(lod);
    const Texel4 r0 = filterEwaLevel(t, (int32_t)ilod, sx, sy, d0x, d0y, d1x, d1y), r1 = filterEwaLevel(t, (int32_t)ilod + 1, sx, sy, d0x, d0y, d1x, d1y);
    return lerpTexel(r0, r1, lod - ilod);
}

__global__ void textureFilterKernel(TextureMipsGpu tex, int filter, int level, size_t n, const float* __restrict__ st2, const float* __restrict__ dst0,
                                    const float* __restrict__ dst1, float* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float sx = st2[2 * i], sy = st2[2 * i + 1];
    const float d0x = dst0 ? dst0[2 * i] : 0.f, d0y = dst0 ? dst0[2 * i + 1] : 0.f, d1x = dst1 ? dst1[2 * i] : 0.f, d1y = dst1 ? dst1[2 * i + 1] : 0.f;
    Texel4 r;
    switch (filter) {
        case SELAS_FILTER_POINT: r = filterPoint(tex, sx, sy); break;
        case SELAS_FILTER_TRIANGLE: r = filterTriangle(tex, level, sx, sy); break;
        case SELAS_FILTER_TRILINEAR: r = filterTrilinear(tex, sx, sy, d0x, d0y, d1x, d1y); break;
        case SELAS_FILTER_EWA: r = filterEwa(tex, sx, sy, d0x, d0y, d1x, d1y); break;
        default: r = filterEwaLevel(tex, level, sx, sy, d0x, d0y, d1x, d1y); break;
    }
    for (uint32_t c = 0; c < tex.comps; ++c) out[i * tex.comps + c] = r.c[c];
}

int launchTextureFilter(SelasContext* ctx, const SelasTextureMipsDesc& desc, const void* d_texels, const float* ewaLut128, int filter, int level, size_t n,
                        const float* d_st2, const float* d_dst0, const float* d_dst1, float* d_out) {
    if (!n) return SELAS_OK;
    TextureMipsGpu g;
    g.texels = static_cast<const uint8_t*>(d_texels);
    g.mipCount = desc.mipCount;
    g.comps = desc.format + 1u;
    for (uint32_t l = 0; l < SELAS_TEXTURE_MAX_MIPS; ++l) {
        g.w[l] = l < desc.mipCount ? desc.mipWidths[l] : 1u;
        g.h[l] = l < desc.mipCount ? desc.mipHeights[l] : 1u;
        g.offset[l] = l < desc.mipCount ? desc.mipOffsets[l] : 0u;
    }
    SELAS_CUDA_CHECK(ctx, cudaMemcpyToSymbolAsync(cEwaLut, ewaLut128, sizeof(float) * 128, 0, cudaMemcpyHostToDevice, ctx->stream));
    textureFilterKernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(g, filter, level, n, d_st2, d_dst0, d_dst1, d_out);
    SELAS_CUDA_CHECK(ctx, cudaGetLastError());
    return SELAS_OK;
}

}  // namespace selas
