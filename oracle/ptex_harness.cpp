// oracle/ptex_harness.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// extern "C" wrappers around the Ptex v2.2 library vendored by the reference (Middleware/ptex/Include,
// compiled in place by oracle/Makefile into oracle/_ref/libptex_ref.so).  Two jobs:
//   * ptexref_write_grid  -- author .ptx fixtures with the reference's own PtexWriter;
//   * ptexref_eval        -- the exact lookup the reference's shading performs:
//       PtexCache::create(128, size, premultiply=true)            TextureCache.cpp:49-57
//       PtexFilter::getFilter(tex, Options(f_bspline))            SurfaceParameters.cpp:199-200
//       filter->eval(&rgb, 0, 3, primId, u, v, 0, 0, 0, 0)        SurfaceParameters.cpp:203
#include <cstdint>
#include <cstring>
#include <vector>

#include "Ptexture.h"

static inline uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

extern "C" {

// Quad grid of nu x nv faces with full edge adjacency.  Face (i,j) has resolution
// 2^(ulog2 + ((i+j)%3==0 ? resJitter : 0)) x 2^(vlog2 + ((i*7+j)%4==0 ? resJitter : 0)) so that neighbouring
// faces can differ in resolution when resJitter != 0.  Texel = hash(seed, face, x, y, channel).
// dtype: 0 uint8, 1 uint16, 2 half (unsupported here), 3 float.  constantEvery > 0 makes every n-th face
// a constant face.
int ptexref_write_grid(const char* path, int nu, int nv, int ulog2, int vlog2, int resJitter, int dtype, int nchan,
                       uint32_t seed, int genmipmaps, int constantEvery) {
    Ptex::String err;
    Ptex::DataType dt = dtype == 0 ? Ptex::dt_uint8 : dtype == 1 ? Ptex::dt_uint16 : Ptex::dt_float;
    int nfaces = nu * nv;
    PtexWriter* w = PtexWriter::open(path, Ptex::mt_quad, dt, nchan, -1, nfaces, err, genmipmaps != 0);
    if (!w) return 1;
    int bytes = dtype == 0 ? 1 : dtype == 1 ? 2 : 4;
    std::vector<uint8_t> buf;
    for (int j = 0; j < nv; ++j)
        for (int i = 0; i < nu; ++i) {
            int f = j * nu + i;
            int ul = ulog2 + (((i + j) % 3 == 0) ? resJitter : 0);
            int vl = vlog2 + (((i * 7 + j) % 4 == 0) ? resJitter : 0);
            Ptex::Res res((int8_t)ul, (int8_t)vl);
            int adjf[4] = {j > 0 ? f - nu : -1, i < nu - 1 ? f + 1 : -1, j < nv - 1 ? f + nu : -1, i > 0 ? f - 1 : -1};
            int adje[4] = {2, 3, 0, 1};
            Ptex::FaceInfo info(res, adjf, adje);
            int ures = 1 << ul, vres = 1 << vl;
            bool constant = constantEvery > 0 && (f % constantEvery) == 0;
            buf.resize((size_t)ures * vres * nchan * bytes);
            for (int y = 0; y < vres; ++y)
                for (int x = 0; x < ures; ++x)
                    for (int c = 0; c < nchan; ++c) {
                        uint32_t h = constant ? mix32(seed ^ mix32((uint32_t)f * 4u + c))
                                              : mix32(seed ^ mix32((uint32_t)f ^ mix32((uint32_t)(y * 4096 + x) * 4u + c)));
                        size_t o = ((size_t)(y * ures + x) * nchan + c) * bytes;
                        if (dtype == 0) buf[o] = (uint8_t)(h >> 24);
                        else if (dtype == 1) { uint16_t v = (uint16_t)(h >> 16); memcpy(&buf[o], &v, 2); }
                        else { float v = (h >> 8) * (1.0f / 16777216.0f); memcpy(&buf[o], &v, 4); }
                    }
            bool ok = constant ? w->writeConstantFace(f, info, buf.data()) : w->writeFace(f, info, buf.data());
            if (!ok) { w->release(); return 2; }
        }
    bool ok = w->close(err);
    w->release();
    return ok ? 0 : 3;
}

int ptexref_eval(const char* path, size_t n, const int32_t* faceid, const float* u, const float* v, float* rgb) {
    Ptex::String err;
    PtexCache* cache = PtexCache::create(128, (size_t)1 << 30, true, nullptr, nullptr);
    PtexTexture* tex = cache->get(path, err);
    if (!tex) { cache->release(); return 1; }
    PtexFilter::Options opts(PtexFilter::f_bspline);
    PtexFilter* filter = PtexFilter::getFilter(tex, opts);
    for (size_t i = 0; i < n; ++i) {
        float s[3] = {0, 0, 0};
        filter->eval(s, 0, 3, faceid[i], u[i], v[i], 0, 0, 0, 0);
        rgb[3 * i] = s[0]; rgb[3 * i + 1] = s[1]; rgb[3 * i + 2] = s[2];
    }
    filter->release();
    tex->release();
    cache->release();
    return 0;
}

// Persistent form for the CPU integrator's Ptex hook: one PtexCache + texture per handle (TextureCache.cpp:49-57, :99-126);
// every lookup makes its own filter exactly as the reference's shading does (SurfaceParameters.cpp:199-206), which is also
// what makes it callable from several threads at once.
struct PtexRefHandle {
    PtexCache* cache;
    PtexTexture* tex;
};

void* ptexref_open(const char* path) {
    Ptex::String err;
    PtexCache* cache = PtexCache::create(128, (size_t)1 << 30, true, nullptr, nullptr);
    PtexTexture* tex = cache->get(path, err);
    if (!tex) { cache->release(); return nullptr; }
    return new PtexRefHandle{cache, tex};
}

void ptexref_eval1(void* handle, int faceid, float u, float v, float* rgb3) {
    PtexRefHandle* h = (PtexRefHandle*)handle;
    PtexFilter::Options opts(PtexFilter::f_bspline);
    PtexFilter* filter = PtexFilter::getFilter(h->tex, opts);
    rgb3[0] = rgb3[1] = rgb3[2] = 0.0f;
    filter->eval(rgb3, 0, 3, faceid, u, v, 0, 0, 0, 0);
    filter->release();
}

void ptexref_close(void* handle) {
    PtexRefHandle* h = (PtexRefHandle*)handle;
    if (!h) return;
    h->tex->release();
    h->cache->release();
    delete h;
}

int ptexref_info(const char* path, int* out8) {
    Ptex::String err;
    PtexTexture* tex = PtexTexture::open(path, err, true);
    if (!tex) return 1;
    out8[0] = tex->numFaces();
    out8[1] = tex->numChannels();
    out8[2] = (int)tex->dataType();
    out8[3] = (int)tex->meshType();
    out8[4] = tex->alphaChannel();
    out8[5] = (int)tex->uBorderMode();
    out8[6] = (int)tex->vBorderMode();
    out8[7] = (int)tex->edgeFilterMode();
    tex->release();
    return 0;
}

}  // extern "C"
