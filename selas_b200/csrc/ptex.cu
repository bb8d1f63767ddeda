// ptex.cu -- Ptex (.ptx) loader and the batch lookup kernel.
//
// Replaces TextureCache::LoadTexturePtex / FetchPtex (Source/Core/TextureLib/TextureCache.cpp:99-126,:168-185), i.e.
// Ptex::PtexCache::get + PtexReader of the vendored Ptex v2.2 (Middleware/ptex/Include), with a load-everything design:
// a 180 GB GPU keeps the whole texture resident, so the file is parsed once on the host, every face's texels are
// inflated into one pool, and the pool goes to HBM.  No cache, no on-demand zlib in the shading loop.
//
// File format followed (PtexIO.h:42-103; PtexReader.cpp:137-208 open, :277-331 face info / const data / level info,
// :563-590 readZipBlock, :595-625 readLevel, :641-710 readFaceData; PtexUtils.cpp:160-195 interleave, :255-272
// decodeDifference, :275-375 reduce/reduceu/reducev, :548-582 multalpha, :624-641 genRfaceids):
//   Header 64 B | ExtHeader | zip(FaceInfo[nfaces] 20 B) | zip(constdata) | LevelInfo[nlevels] 16 B | level data ...
//   level: zip(FaceDataHeader[nfaces]) then the face blocks; encodings constant / zipped / diffzipped / tiled;
//   texels are stored planar per channel and interleaved on load; premultiplied by alpha (TextureCache.cpp:55
//   creates the cache with premultiply = true) at level 0 and in the constant block.
// Reduced-resolution data (PtexReader::getData(faceid, res), PtexReader.cpp:785-883): stored mip level when the
// reduction is symmetric and present, otherwise generated with the library's alternating reduceu / reducev rule.
// The loader pre-builds exactly the reductions a kernel arriving from a face's 2-ring can request.
//
// Not supported (rejected with SELAS_ERR_UNSUPPORTED): triangle-mesh textures (the reference's PtexTriangleFilter path),
// files with edit blocks, tangent-vector edge filter mode.
#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <new>
#include <set>
#include <vector>

#include "internal.h"
#include "ptex_eval.cuh"

struct SelasPtex {
    SelasContext* ctx = nullptr;
    selas::PtexGpu host{};         // device pointers + scalars
    selas::PtexGpu* d_self = nullptr;  // the same struct in device memory (what MeshGpu::ptex points at)
    std::vector<void*> allocations;
    uint64_t deviceBytes = 0;
};

namespace selas {

const void* ptexDevicePointer(SelasPtex* tex) { return tex ? tex->d_self : nullptr; }
bool ptexMixedResolutions(SelasPtex* tex) { return tex && tex->host.mixedRes != 0; }

namespace {

#pragma pack(push, 1)
struct FileHeader {
    uint32_t magic, version, meshtype, datatype;
    int32_t alphachan;
    uint16_t nchannels, nlevels;
    uint32_t nfaces, extheadersize, faceinfosize, constdatasize, levelinfosize, minorversion;
    uint64_t leveldatasize;
    uint32_t metadatazipsize, metadatamemsize;
};
struct FileExtHeader {
    uint16_t ubordermode, pad, vbordermode, edgefiltermode;
    uint32_t lmdheaderzipsize, lmdheadermemsize;
    uint64_t lmddatasize, editdatasize, editdatapos;
};
struct FileFaceInfo {
    int8_t ulog2, vlog2;
    uint8_t adjedges, flags;
    int32_t adjfaces[4];
};
struct FileLevelInfo {
    uint64_t leveldatasize;
    uint32_t levelheadersize, nfaces;
};
#pragma pack(pop)
static_assert(sizeof(FileHeader) == 64 && sizeof(FileExtHeader) == 40 && sizeof(FileFaceInfo) == 20 && sizeof(FileLevelInfo) == 16,
              "Ptex on-disk structs");

enum { kEncConstant = 0, kEncZipped = 1, kEncDiffZipped = 2, kEncTiled = 3 };
enum { kFlagConstant = 1, kFlagHasEdits = 2, kFlagNbConstant = 4, kFlagSubface = 8 };

inline float halfToFloat(uint16_t h) {
    uint32_t s = (h >> 15) & 1, e = (h >> 10) & 0x1f, m = h & 0x3ff, out;
    if (e == 0) {
        if (m == 0) out = s << 31;
        else {
            int sh = 0;
            while (!(m & 0x400)) { m <<= 1; ++sh; }
            out = (s << 31) | ((uint32_t)(127 - 15 - sh + 1) << 23) | ((m & 0x3ff) << 13);
        }
    } else if (e == 31) out = (s << 31) | 0x7f800000u | (m << 13);
    else out = (s << 31) | ((e + 112) << 23) | (m << 13);
    float f;
    memcpy(&f, &out, 4);
    return f;
}
inline uint16_t floatToHalf(float f) {  // round to nearest even
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t s = (x >> 16) & 0x8000u;
    int32_t e = (int32_t)((x >> 23) & 0xff) - 127 + 15;
    uint32_t m = x & 0x7fffffu;
    if (((x >> 23) & 0xff) == 0xff) return (uint16_t)(s | 0x7c00u | (m ? 0x200u : 0u));
    if (e >= 31) return (uint16_t)(s | 0x7c00u);
    if (e <= 0) {
        if (e < -10) return (uint16_t)s;
        m |= 0x800000u;
        int shift = 14 - e;
        uint32_t r = m >> shift, rem = m & ((1u << shift) - 1), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (r & 1))) ++r;
        return (uint16_t)(s | r);
    }
    uint32_t r = (uint32_t)(e << 10) | (m >> 13), rem = m & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (r & 1))) ++r;
    return (uint16_t)(s | r);
}

struct Loader {
    SelasContext* ctx;
    const uint8_t* file;
    size_t size;
    FileHeader hdr{};
    FileExtHeader ext{};
    int dataSize = 1, pixelSize = 1;
    std::vector<FileFaceInfo> faces;
    std::vector<uint32_t> rfaceids;
    std::vector<uint8_t> constdata;
    std::vector<FileLevelInfo> levelInfo;
    std::vector<uint64_t> levelPos;
    std::vector<std::vector<uint32_t>> levelFdh;      // per level, per (r)face
    std::vector<std::vector<uint64_t>> levelOffsets;  // file offset of every face block
    z_stream zs{};
    bool zinit = false;
    std::string error;

    ~Loader() {
        if (zinit) inflateEnd(&zs);
    }
    bool fail(const char* msg) {
        if (error.empty()) error = msg;
        return false;
    }
    bool inRange(uint64_t pos, uint64_t len) const { return pos <= size && len <= size - pos; }

    bool unzip(uint64_t pos, uint32_t zipsize, void* out, size_t outsize) {  // readZipBlock
        if (!inRange(pos, zipsize)) return fail("ptex: zip block runs past the end of the file");
        if (!zinit) {
            memset(&zs, 0, sizeof(zs));
            if (inflateInit(&zs) != Z_OK) return fail("ptex: inflateInit failed");
            zinit = true;
        }
        zs.next_in = const_cast<Bytef*>(file + pos);
        zs.avail_in = zipsize;
        zs.next_out = (Bytef*)out;
        zs.avail_out = (uInt)outsize;
        int zr = inflate(&zs, Z_FINISH);
        size_t total = zs.total_out;
        inflateReset(&zs);
        if (zr != Z_STREAM_END && zr != Z_OK && zr != Z_BUF_ERROR) return fail("ptex: unzip failed, file corrupt");
        if (total != outsize) return fail("ptex: unzip produced the wrong size, file corrupt");
        return true;
    }

    void multAlpha(uint8_t* data, size_t npixels) {  // PtexUtils::multalpha, premultiply = true
        const int nchan = hdr.nchannels, ac = hdr.alphachan;
        if (!(ac >= 0 && ac < nchan)) return;
        int first = ac == 0 ? 1 : 0, count = ac == 0 ? nchan - 1 : ac;
        for (size_t p = 0; p < npixels; ++p) {
            uint8_t* px = data + p * pixelSize;
            switch (hdr.datatype) {
                case 0: {
                    float a = (1.f / 255.f) * (float)px[ac];
                    for (int i = 0; i < count; ++i) px[first + i] = (uint8_t)((float)px[first + i] * a);
                } break;
                case 1: {
                    uint16_t* q = (uint16_t*)px;
                    float a = (1.f / 65535.f) * (float)q[ac];
                    for (int i = 0; i < count; ++i) q[first + i] = (uint16_t)((float)q[first + i] * a);
                } break;
                case 2: {
                    uint16_t* q = (uint16_t*)px;
                    float a = halfToFloat(q[ac]);
                    for (int i = 0; i < count; ++i) q[first + i] = floatToHalf(halfToFloat(q[first + i]) * a);
                } break;
                default: {
                    float* q = (float*)px;
                    float a = q[ac];
                    for (int i = 0; i < count; ++i) q[first + i] = q[first + i] * a;
                }
            }
        }
    }

    bool parseHeaders() {
        if (size < sizeof(FileHeader)) return fail("ptex: file too small");
        memcpy(&hdr, file, sizeof(hdr));
        if (hdr.magic != (uint32_t)('P' | ('t' << 8) | ('e' << 16) | ('x' << 24))) return fail("ptex: not a ptex file");
        if (hdr.version != 1) return fail("ptex: unsupported file version");
        if (hdr.datatype > 3 || hdr.nchannels == 0) return fail("ptex: bad data type / channel count");
        static const int sizes[4] = {1, 2, 2, 4};
        dataSize = sizes[hdr.datatype];
        pixelSize = dataSize * hdr.nchannels;
        memset(&ext, 0, sizeof(ext));
        size_t extn = std::min<size_t>(sizeof(ext), hdr.extheadersize);
        if (!inRange(sizeof(hdr), extn)) return fail("ptex: truncated ext header");
        memcpy(&ext, file + sizeof(hdr), extn);
        uint64_t pos = sizeof(hdr) + hdr.extheadersize;
        uint64_t faceinfopos = pos; pos += hdr.faceinfosize;
        uint64_t constdatapos = pos; pos += hdr.constdatasize;
        uint64_t levelinfopos = pos; pos += hdr.levelinfosize;
        uint64_t leveldatapos = pos;
        faces.resize(hdr.nfaces);
        if (hdr.nfaces && !unzip(faceinfopos, hdr.faceinfosize, faces.data(), sizeof(FileFaceInfo) * (size_t)hdr.nfaces)) return false;
        constdata.resize((size_t)pixelSize * hdr.nfaces);
        if (hdr.nfaces && !unzip(constdatapos, hdr.constdatasize, constdata.data(), constdata.size())) return false;
        multAlpha(constdata.data(), hdr.nfaces);
        levelInfo.resize(hdr.nlevels);
        if (!inRange(levelinfopos, sizeof(FileLevelInfo) * (size_t)hdr.nlevels)) return fail("ptex: truncated level info");
        memcpy(levelInfo.data(), file + levelinfopos, sizeof(FileLevelInfo) * (size_t)hdr.nlevels);
        levelPos.resize(hdr.nlevels);
        uint64_t lp = leveldatapos;
        for (int i = 0; i < hdr.nlevels; ++i) {
            levelPos[i] = lp;
            lp += levelInfo[i].leveldatasize;
        }
        levelFdh.resize(hdr.nlevels);
        levelOffsets.resize(hdr.nlevels);
        for (int l = 0; l < hdr.nlevels; ++l) {
            const FileLevelInfo& li = levelInfo[l];
            levelFdh[l].resize(li.nfaces);
            if (li.nfaces && !unzip(levelPos[l], li.levelheadersize, levelFdh[l].data(), 4 * (size_t)li.nfaces)) return false;
            levelOffsets[l].resize(li.nfaces);
            uint64_t p = levelPos[l] + li.levelheadersize;
            for (uint32_t i = 0; i < li.nfaces; ++i) {
                levelOffsets[l][i] = p;
                p += levelFdh[l][i] & 0x3fffffffu;
            }
        }
        // genRfaceids: stable sort by min(ulog2, vlog2) descending, constant faces count as 1
        std::vector<uint32_t> order(hdr.nfaces);
        for (uint32_t i = 0; i < hdr.nfaces; ++i) order[i] = i;
        auto key = [&](uint32_t f) { return (faces[f].flags & kFlagConstant) ? 1 : (int)std::min(faces[f].ulog2, faces[f].vlog2); };
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return key(a) > key(b); });
        rfaceids.resize(hdr.nfaces);
        for (uint32_t i = 0; i < hdr.nfaces; ++i) rfaceids[order[i]] = i;
        return true;
    }

    // one packed block (zipped / diffzipped / constant) -> interleaved pixels of ures x vres
    bool readPacked(uint64_t pos, uint32_t fdh, int ures, int vres, bool level0, std::vector<uint8_t>& out) {
        const uint32_t enc = (fdh >> 30) & 3, blocksize = fdh & 0x3fffffffu;
        const size_t npix = (size_t)ures * vres, unpacked = npix * pixelSize;
        out.resize(unpacked);
        if (enc == kEncConstant) {
            if (!inRange(pos, pixelSize)) return fail("ptex: truncated constant face");
            std::vector<uint8_t> px(file + pos, file + pos + pixelSize);
            if (level0) multAlpha(px.data(), 1);
            for (size_t p = 0; p < npix; ++p) memcpy(&out[p * pixelSize], px.data(), pixelSize);
            return true;
        }
        if (enc == kEncTiled) return fail("ptex: nested tiling is not valid");
        std::vector<uint8_t> tmp(unpacked);
        if (!unzip(pos, blocksize, tmp.data(), unpacked)) return false;
        if (enc == kEncDiffZipped) {
            if (hdr.datatype == 0) {
                uint8_t prev = 0;
                for (size_t i = 0; i < unpacked; ++i) { tmp[i] = (uint8_t)(tmp[i] + prev); prev = tmp[i]; }
            } else if (hdr.datatype == 1) {
                uint16_t* p = (uint16_t*)tmp.data();
                uint16_t prev = 0;
                for (size_t i = 0; i < unpacked / 2; ++i) { p[i] = (uint16_t)(p[i] + prev); prev = p[i]; }
            }
        }
        // planar (channel-major) -> interleaved
        const int nchan = hdr.nchannels;
        for (int c = 0; c < nchan; ++c)
            for (size_t p = 0; p < npix; ++p) memcpy(&out[(p * nchan + c) * dataSize], &tmp[((size_t)c * npix + p) * dataSize], dataSize);
        if (level0) multAlpha(out.data(), npix);
        return true;
    }

    // PtexReader::readFaceData: the face block of (level, index) as ures x vres interleaved pixels
    bool readFaceBlock(int level, uint32_t index, int ulog2, int vlog2, std::vector<uint8_t>& out) {
        const uint32_t fdh = levelFdh[level][index];
        const uint64_t pos = levelOffsets[level][index];
        const int ures = 1 << ulog2, vres = 1 << vlog2;
        if (((fdh >> 30) & 3) != kEncTiled) return readPacked(pos, fdh, ures, vres, level == 0, out);
        if (!inRange(pos, 6)) return fail("ptex: truncated tile header");
        int8_t tul = (int8_t)file[pos], tvl = (int8_t)file[pos + 1];
        uint32_t tileheadersize;
        memcpy(&tileheadersize, file + pos + 2, 4);
        if (tul < 0 || tvl < 0 || tul > ulog2 || tvl > vlog2) return fail("ptex: bad tile resolution");
        const int tu = 1 << tul, tv = 1 << tvl, ntu = ures / tu, ntv = vres / tv, ntiles = ntu * ntv;
        std::vector<uint32_t> tfdh(ntiles);
        if (!unzip(pos + 6, tileheadersize, tfdh.data(), 4 * (size_t)ntiles)) return false;
        uint64_t p = pos + 6 + tileheadersize;
        out.resize((size_t)ures * vres * pixelSize);
        std::vector<uint8_t> tile;
        for (int t = 0; t < ntiles; ++t) {
            if (!readPacked(p, tfdh[t], tu, tv, level == 0, tile)) return false;
            p += tfdh[t] & 0x3fffffffu;
            int ty = t / ntu, tx = t % ntu;
            for (int y = 0; y < tv; ++y)
                memcpy(&out[(((size_t)ty * tv + y) * ures + (size_t)tx * tu) * pixelSize], &tile[(size_t)y * tu * pixelSize], (size_t)tu * pixelSize);
        }
        return true;
    }

    // PtexUtils::reduceu / reducev on interleaved pixels (integer types halve by shifting, float/half by 0.5f)
    void reduceAxis(const std::vector<uint8_t>& src, int ures, int vres, bool alongU, std::vector<uint8_t>& dst) {
        const int nchan = hdr.nchannels;
        const int du = alongU ? ures / 2 : ures, dv = alongU ? vres : vres / 2;
        dst.resize((size_t)du * dv * pixelSize);
        auto at = [&](int x, int y, int c) { return ((size_t)y * ures + x) * nchan + c; };
        for (int y = 0; y < dv; ++y)
            for (int x = 0; x < du; ++x)
                for (int c = 0; c < nchan; ++c) {
                    size_t a = alongU ? at(2 * x, y, c) : at(x, 2 * y, c);
                    size_t b = alongU ? at(2 * x + 1, y, c) : at(x, 2 * y + 1, c);
                    size_t o = ((size_t)y * du + x) * nchan + c;
                    switch (hdr.datatype) {
                        case 0: dst[o] = (uint8_t)(((int)src[a] + (int)src[b]) >> 1); break;
                        case 1: ((uint16_t*)dst.data())[o] = (uint16_t)(((int)((const uint16_t*)src.data())[a] + (int)((const uint16_t*)src.data())[b]) >> 1); break;
                        case 2: {
                            // PtexHalf + PtexHalf goes through float; halve(PtexHalf) = 0.5f * val, rounded back to half
                            float s = halfToFloat(((const uint16_t*)src.data())[a]) + halfToFloat(((const uint16_t*)src.data())[b]);
                            ((uint16_t*)dst.data())[o] = floatToHalf(0.5f * halfToFloat(floatToHalf(s)));
                        } break;
                        default: ((float*)dst.data())[o] = 0.5f * (((const float*)src.data())[a] + ((const float*)src.data())[b]);
                    }
                }
    }

    // PtexReader::getData(faceid, res) for a non-constant face and res != 0, memoised per face
    typedef std::map<uint32_t, std::vector<uint8_t>> FaceCache;
    const std::vector<uint8_t>* getData(uint32_t faceid, int ul, int vl, FaceCache& cache) {
        uint32_t key = (uint32_t)(ul & 0xff) | ((uint32_t)(vl & 0xff) << 8);
        auto it = cache.find(key);
        if (it != cache.end()) return &it->second;
        const FileFaceInfo& fi = faces[faceid];
        int redu = fi.ulog2 - ul, redv = fi.vlog2 - vl;
        std::vector<uint8_t> data;
        if (redu < 0 || redv < 0 || ul < 0 || vl < 0) { fail("ptex: invalid reduction request"); return nullptr; }
        if (redu == 0 && redv == 0) {
            if (hdr.nlevels == 0 || faceid >= levelFdh[0].size()) { fail("ptex: face missing from level 0"); return nullptr; }
            if (!readFaceBlock(0, faceid, ul, vl, data)) return nullptr;
        } else {
            bool done = false;
            if (redu == redv && !(fi.flags & kFlagHasEdits) && redu < (int)hdr.nlevels) {
                uint32_t rf = rfaceids[faceid];
                if (rf < levelFdh[redu].size()) {
                    if (!readFaceBlock(redu, rf, ul, vl, data)) return nullptr;
                    done = true;
                }
            }
            if (!done) {
                bool blendu = (redu == redv) ? ((ul & 1) != 0) : (redu > redv);
                const std::vector<uint8_t>* src = blendu ? getData(faceid, ul + 1, vl, cache) : getData(faceid, ul, vl + 1, cache);
                if (!src) return nullptr;
                reduceAxis(*src, 1 << (blendu ? ul + 1 : ul), 1 << (blendu ? vl : vl + 1), blendu, data);
            }
        }
        return &(cache[key] = std::move(data));
    }
};

template <bool MIXED>
__global__ void __launch_bounds__(128, SELAS_PTEX_MIN_CTAS) ptexEvalKernel(const PtexGpu* __restrict__ tex, size_t n, const int32_t* __restrict__ faceid,
                                                      const float* __restrict__ u, const float* __restrict__ v, float* __restrict__ rgb) {
    ptexTileEval<128, 8, MIXED>(
        n,
        [&](size_t i, const PtexGpu*& t, int& f, float& uu, float& vv) {
            t = tex;
            f = faceid[i];
            uu = u[i];
            vv = v[i];
            return true;
        },
        [&](size_t i, const float* c) {
            rgb[3 * i] = c[0];
            rgb[3 * i + 1] = c[1];
            rgb[3 * i + 2] = c[2];
        });
}

template <typename T>
int uploadVec(SelasPtex* tex, const T** out, const void* host, size_t bytes) {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes + 16);
    if (e != cudaSuccess) {
        setError(tex->ctx, "cudaMalloc(%zu bytes) for a ptex texture failed: %s", bytes, cudaGetErrorString(e));
        return SELAS_ERR_MEMORY;
    }
    tex->allocations.push_back(p);
    tex->deviceBytes += bytes;
    if (bytes) {
        e = cudaMemcpy(p, host, bytes, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            setError(tex->ctx, "upload of a ptex texture failed: %s", cudaGetErrorString(e));
            return SELAS_ERR_CUDA;
        }
    }
    *out = (const T*)p;
    return SELAS_OK;
}

}  // namespace
}  // namespace selas

using namespace selas;

extern "C" {

int selas_cuda_ptex_destroy(SelasPtex* tex) {
    if (!tex) return SELAS_OK;
    cudaSetDevice(tex->ctx->device);
    cudaStreamSynchronize(tex->ctx->stream);
    for (void* p : tex->allocations) cudaFree(p);
    delete tex;
    return SELAS_OK;
}

int selas_cuda_ptex_load_memory(SelasContext* ctx, const void* bytes, size_t size, SelasPtex** out) {
    if (!ctx || !bytes || !out) {
        setError(ctx, "selas_cuda_ptex_load_memory: null argument");
        return SELAS_ERR_ARGUMENT;
    }
    *out = nullptr;
    SELAS_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    Loader L;
    L.ctx = ctx;
    L.file = (const uint8_t*)bytes;
    L.size = size;
    if (!L.parseHeaders()) {
        setError(ctx, "%s", L.error.c_str());
        return SELAS_ERR_FORMAT;
    }
    if (L.hdr.meshtype != 1) {
        setError(ctx, "ptex: triangle-mesh textures are not supported (quad textures only)");
        return SELAS_ERR_UNSUPPORTED;
    }
    if (L.ext.editdatasize != 0) {
        setError(ctx, "ptex: files with edit blocks are not supported");
        return SELAS_ERR_UNSUPPORTED;
    }
    if (L.ext.edgefiltermode != 0) {
        setError(ctx, "ptex: tangent-vector edge filter mode is not supported");
        return SELAS_ERR_UNSUPPORTED;
    }
    const uint32_t nfaces = L.hdr.nfaces;
    for (uint32_t f = 0; f < nfaces; ++f) {
        const FileFaceInfo& fi = L.faces[f];
        if (fi.ulog2 < 0 || fi.vlog2 < 0 || fi.ulog2 > 24 || fi.vlog2 > 24) {
            setError(ctx, "ptex: face %u has an invalid resolution", f);
            return SELAS_ERR_FORMAT;
        }
        for (int e = 0; e < 4; ++e)
            if (fi.adjfaces[e] >= (int32_t)nfaces) {
                setError(ctx, "ptex: face %u names a neighbour outside the file", f);
                return SELAS_ERR_FORMAT;
            }
    }

    // ---- which reduced resolutions can a neighbour's kernel ask of each face? ----
    bool anySubface = false, uniform = true;
    for (uint32_t f = 0; f < nfaces; ++f) {
        if (L.faces[f].flags & kFlagSubface) anySubface = true;
        if (L.faces[f].ulog2 != L.faces[0].ulog2 || L.faces[f].vlog2 != L.faces[0].vlog2 || L.faces[f].ulog2 != L.faces[f].vlog2) uniform = false;
    }
    const bool needReductions = !(uniform && !anySubface);

    std::vector<uint8_t> pool;
    std::vector<uint4> faceRecords(2 * (size_t)nfaces);
    std::vector<PtexRed> reds;
    auto appendBlock = [&](const std::vector<uint8_t>& data) -> uint32_t {
        size_t off = (pool.size() + 15) & ~(size_t)15;
        pool.resize(off + data.size());
        if (!data.empty()) memcpy(&pool[off], data.data(), data.size());
        return (uint32_t)(off / 16);
    };
    for (uint32_t f = 0; f < nfaces; ++f) {
        const FileFaceInfo& fi = L.faces[f];
        Loader::FaceCache cache;
        uint32_t dataOffset = 0, redStart = (uint32_t)reds.size(), redCount = 0;
        const bool constant = (fi.flags & kFlagConstant) != 0 || (fi.ulog2 == 0 && fi.vlog2 == 0);
        if (!constant) {
            const std::vector<uint8_t>* native = L.getData(f, fi.ulog2, fi.vlog2, cache);
            if (!native) {
                setError(ctx, "%s (face %u)", L.error.c_str(), f);
                return SELAS_ERR_FORMAT;
            }
            dataOffset = appendBlock(*native);
            if (needReductions) {
                std::set<uint32_t> want;
                auto consider = [&](int su, int sv) {
                    for (int d = anySubface ? -1 : 0; d <= (anySubface ? 1 : 0); ++d) {
                        int ru = std::min(std::max(su + d, 0), (int)fi.ulog2), rv = std::min(std::max(sv + d, 0), (int)fi.vlog2);
                        if ((ru == fi.ulog2 && rv == fi.vlog2) || (ru == 0 && rv == 0)) continue;
                        want.insert((uint32_t)ru | ((uint32_t)rv << 8));
                    }
                };
                auto considerFace = [&](int s) {
                    const FileFaceInfo& sf = L.faces[s];
                    consider(sf.ulog2, sf.vlog2);
                    consider(sf.vlog2, sf.ulog2);
                    int m = std::min(sf.ulog2, sf.vlog2);
                    consider(m, m);
                };
                for (int e = 0; e < 4; ++e) {
                    int a = fi.adjfaces[e];
                    if (a < 0) continue;
                    considerFace(a);
                    for (int e2 = 0; e2 < 4; ++e2) {
                        int b = L.faces[a].adjfaces[e2];
                        if (b < 0) continue;
                        considerFace(b);
                        if (anySubface)
                            for (int e3 = 0; e3 < 4; ++e3)
                                if (L.faces[b].adjfaces[e3] >= 0) considerFace(L.faces[b].adjfaces[e3]);
                    }
                }
                for (uint32_t key : want) {
                    const std::vector<uint8_t>* r = L.getData(f, (int)(key & 0xff), (int)((key >> 8) & 0xff), cache);
                    if (!r) {
                        setError(ctx, "%s (face %u reduction)", L.error.c_str(), f);
                        return SELAS_ERR_FORMAT;
                    }
                    PtexRed pr;
                    pr.res = key;
                    pr.offset16 = appendBlock(*r);
                    reds.push_back(pr);
                    ++redCount;
                }
            }
        }
        if (pool.size() / 16 > 0xFFFFFFF0ull) {
            setError(ctx, "ptex: texture exceeds the 64 GB texel pool addressing");
            return SELAS_ERR_UNSUPPORTED;
        }
        uint32_t w0 = (uint32_t)(uint8_t)fi.ulog2 | ((uint32_t)(uint8_t)fi.vlog2 << 8) | ((uint32_t)fi.adjedges << 16) | ((uint32_t)fi.flags << 24);
        faceRecords[2 * (size_t)f] = make_uint4(w0, (uint32_t)fi.adjfaces[0], (uint32_t)fi.adjfaces[1], (uint32_t)fi.adjfaces[2]);
        faceRecords[2 * (size_t)f + 1] = make_uint4((uint32_t)fi.adjfaces[3], dataOffset, redStart, redCount);
    }

    SelasPtex* tex = new (std::nothrow) SelasPtex();
    if (!tex) return SELAS_ERR_MEMORY;
    tex->ctx = ctx;
    int rc;
    PtexGpu& g = tex->host;
    if ((rc = uploadVec(tex, &g.faces, faceRecords.data(), faceRecords.size() * sizeof(uint4))) ||
        (rc = uploadVec(tex, &g.reds, reds.data(), reds.size() * sizeof(PtexRed))) ||
        (rc = uploadVec(tex, &g.texels, pool.data(), pool.size())) ||
        (rc = uploadVec(tex, &g.constData, L.constdata.data(), L.constdata.size()))) {
        selas_cuda_ptex_destroy(tex);
        return rc;
    }
    g.numFaces = (int)nfaces;
    g.nchan = L.hdr.nchannels;
    g.dataType = (int)L.hdr.datatype;
    g.pixelSize = L.pixelSize;
    g.uMode = L.ext.ubordermode;
    g.vMode = L.ext.vbordermode;
    g.edgeFilterMode = L.ext.edgefiltermode;
    g.mixedRes = uniform ? 0 : 1;
    const PtexGpu* d_self = nullptr;
    if ((rc = uploadVec(tex, &d_self, &g, sizeof(PtexGpu)))) {
        selas_cuda_ptex_destroy(tex);
        return rc;
    }
    tex->d_self = const_cast<PtexGpu*>(d_self);
    *out = tex;
    return SELAS_OK;
}

int selas_cuda_ptex_load(SelasContext* ctx, const char* path, SelasPtex** out) {
    if (!ctx || !path || !out) {
        setError(ctx, "selas_cuda_ptex_load: null argument");
        return SELAS_ERR_ARGUMENT;
    }
    *out = nullptr;
    FILE* fp = fopen(path, "rb");
    if (!fp) {
        setError(ctx, "Can't open ptex file: %s", path);
        return SELAS_ERR_FORMAT;
    }
    std::vector<uint8_t> bytes;
    fseek(fp, 0, SEEK_END);
    long sz = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    if (sz > 0) {
        bytes.resize((size_t)sz);
        if (fread(bytes.data(), 1, bytes.size(), fp) != bytes.size()) {
            fclose(fp);
            setError(ctx, "short read on ptex file: %s", path);
            return SELAS_ERR_FORMAT;
        }
    }
    fclose(fp);
    return selas_cuda_ptex_load_memory(ctx, bytes.data(), bytes.size(), out);
}

int selas_cuda_ptex_eval_device(SelasPtex* tex, size_t n, const int32_t* d_faceid, const float* d_u, const float* d_v, float* d_rgb3,
                                void* stream) {
    if (!tex) return SELAS_ERR_ARGUMENT;
    SelasContext* ctx = tex->ctx;
    SELAS_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    if (n == 0) return SELAS_OK;
    if (!d_faceid || !d_u || !d_v || !d_rgb3) {
        setError(ctx, "selas_cuda_ptex_eval: null argument");
        return SELAS_ERR_ARGUMENT;
    }
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    unsigned grid = (unsigned)std::min<size_t>((n + 1023) / 1024, (size_t)ctx->smCount * 8);
    if (tex->host.mixedRes) ptexEvalKernel<true><<<grid, 128, 0, st>>>(tex->d_self, n, d_faceid, d_u, d_v, d_rgb3);
    else ptexEvalKernel<false><<<grid, 128, 0, st>>>(tex->d_self, n, d_faceid, d_u, d_v, d_rgb3);
    SELAS_CUDA_CHECK(ctx, cudaGetLastError());
    return SELAS_OK;
}

int selas_cuda_ptex_eval(SelasPtex* tex, size_t n, const int32_t* faceid, const float* u, const float* v, float* rgb3) {
    if (!tex) return SELAS_ERR_ARGUMENT;
    SelasContext* ctx = tex->ctx;
    SELAS_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    if (n == 0) return SELAS_OK;
    if (!faceid || !u || !v || !rgb3) {
        setError(ctx, "selas_cuda_ptex_eval: null argument");
        return SELAS_ERR_ARGUMENT;
    }
    int32_t* d_f = nullptr;
    float *d_u = nullptr, *d_v = nullptr, *d_o = nullptr;
    int rc = SELAS_OK;
    auto chk = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && !rc) {
            setError(ctx, "%s failed: %s", what, cudaGetErrorString(e));
            rc = SELAS_ERR_CUDA;
        }
    };
    chk(cudaMalloc((void**)&d_f, n * 4), "cudaMalloc");
    chk(cudaMalloc((void**)&d_u, n * 4), "cudaMalloc");
    chk(cudaMalloc((void**)&d_v, n * 4), "cudaMalloc");
    chk(cudaMalloc((void**)&d_o, n * 12), "cudaMalloc");
    if (!rc) {
        chk(cudaMemcpyAsync(d_f, faceid, n * 4, cudaMemcpyHostToDevice, ctx->stream), "upload");
        chk(cudaMemcpyAsync(d_u, u, n * 4, cudaMemcpyHostToDevice, ctx->stream), "upload");
        chk(cudaMemcpyAsync(d_v, v, n * 4, cudaMemcpyHostToDevice, ctx->stream), "upload");
    }
    if (!rc) rc = selas_cuda_ptex_eval_device(tex, n, d_f, d_u, d_v, d_o, ctx->stream);
    if (!rc) chk(cudaMemcpyAsync(rgb3, d_o, n * 12, cudaMemcpyDeviceToHost, ctx->stream), "download");
    chk(cudaStreamSynchronize(ctx->stream), "synchronize");
    cudaFree(d_f);
    cudaFree(d_u);
    cudaFree(d_v);
    cudaFree(d_o);
    return rc;
}

}  // extern "C"
