// ptex_eval.cuh -- device-side Ptex lookup: PtexFilter::eval with the f_bspline filter, exactly as the reference's
// shading calls it (SurfaceParameters.cpp:203: eval(rgb, 0, 3, primId, u, v, 0, 0, 0, 0), width 1, blur 0).
//
// Restates, for the GPU (paths relative to Middleware/ptex/Include):
//   PtexSeparableFilter::eval / splitAndApply / applyAcrossEdge / applyToCorner / applyToCornerFace / apply
//                                                     PtexSeparableFilter.cpp:47-132, :135-213, :215-255, :257-330, :333-345, :348-401
//   PtexWidth4Filter::buildKernelAxis                 PtexFilters.cpp:141-248 (lerp = false, as PtexFilter::Options defaults)
//   PtexBicubicFilter coefficients + kernelFn         PtexFilters.cpp:256-284 (sharpness 0 => B = 1, the cubic B-spline)
//   PtexSeparableKernel (split/merge/rotate/downres/upres/adjust*/makeSymmetric/apply)   PtexSeparableKernel.h:55-457,
//                                                     PtexSeparableKernel.cpp:45-75 (row-wise accumulate, then rows * kv)
//   PtexReader::getData(faceid, res) selection rules  PtexReader.cpp:785-830 (constant face or res 0 => per-face constant;
//                                                     otherwise native or reduced data, here pre-built by the loader in ptex.cu)
//
// Data layout in HBM (built by ptex.cu):
//   faces[f]    32 B = { ulog2 | vlog2<<8 | adjedges<<16 | flags<<24, adjfaces[4], dataOffset16, redStart, redCount }
//   reds[i]      8 B = { ulog2 | vlog2<<8, dataOffset16 }   reduced-resolution copies a neighbour's kernel can ask for
//   texels      native data type (u8 / u16 / half / float), nchan interleaved, row-major u fastest, every block 16-B aligned
//   constData   one pixel per face (the file's constdata block)
// Algorithmic bytes per lookup: 32 B face record + 16 texels x pixel size (x2 when the 4x4 straddles an edge).
//
// The one structural change against the library: the recursion splitAndApply -> applyAcrossEdge -> splitAndApply (taken
// only when a kernel crosses onto a subface) is unrolled into a small work list, so partial sums may be added in a
// different order (fp32 reassociation, far inside the 1e-4 parity bound).
#pragma once
#include <cuda_fp16.h>

#include "vec.cuh"

namespace selas {

struct PtexRed {
    uint32_t res;       // ulog2 | vlog2 << 8
    uint32_t offset16;  // texel block offset in 16-byte units
};

struct PtexGpu {
    const uint4* faces;  // 2 per face
    const PtexRed* reds;
    const uint8_t* texels;
    const uint8_t* constData;
    int numFaces;
    int nchan;       // channels in the texture
    int dataType;    // 0 u8, 1 u16, 2 half, 3 float
    int pixelSize;   // bytes
    int uMode, vMode;  // BorderMode: 0 clamp, 1 black, 2 periodic
    int edgeFilterMode;
    int mixedRes;  // 0: every face has the same square resolution (no kernel ever meets another resolution)
};

enum { kPtexFlagConstant = 1, kPtexFlagNbConstant = 4, kPtexFlagSubface = 8 };
enum { kEdgeBottom = 0, kEdgeRight = 1, kEdgeTop = 2, kEdgeLeft = 3 };

struct PtexFace {
    int ulog2, vlog2;
    uint32_t adjedges, flags;
    int adjfaces[4];
    uint32_t dataOffset16, redStart, redCount;
    // selects instead of a runtime index: an indexed read puts the record in local memory (4 STL + 1 LDL per use, measured)
    SELAS_D int adjface(int e) const {
        const int lo = (e & 1) ? adjfaces[1] : adjfaces[0], hi = (e & 1) ? adjfaces[3] : adjfaces[2];
        return (e & 2) ? hi : lo;
    }
    SELAS_D int adjedge(int e) const { return (int)((adjedges >> (2 * e)) & 3u); }
    SELAS_D bool isSubface() const { return (flags & kPtexFlagSubface) != 0; }
    SELAS_D bool isConstant() const { return (flags & kPtexFlagConstant) != 0; }
};

SELAS_D PtexFace ptexLoadFace(const PtexGpu& tx, int faceid) {
    uint4 a = __ldg(tx.faces + 2 * (size_t)faceid), b = __ldg(tx.faces + 2 * (size_t)faceid + 1);
    PtexFace f;
    f.ulog2 = (int)(int8_t)(a.x & 0xff);
    f.vlog2 = (int)(int8_t)((a.x >> 8) & 0xff);
    f.adjedges = (a.x >> 16) & 0xff;
    f.flags = (a.x >> 24) & 0xff;
    f.adjfaces[0] = (int)a.y; f.adjfaces[1] = (int)a.z; f.adjfaces[2] = (int)a.w; f.adjfaces[3] = (int)b.x;
    f.dataOffset16 = b.y; f.redStart = b.z; f.redCount = b.w;
    return f;
}

#define PTEX_KMAX 10

// PtexSeparableKernel.  ku/kv are offsets into the fixed buffers instead of moving pointers.
struct PtexKernel {
    int ulog2, vlog2;
    int u, v, uw, vw;
    int ku0, kv0;  // start of the live weights inside kub / kvb
    int rot;
    float kub[PTEX_KMAX], kvb[PTEX_KMAX];

    SELAS_D int resu() const { return 1 << ulog2; }
    SELAS_D int resv() const { return 1 << vlog2; }
    SELAS_D float& ku(int i) { return kub[ku0 + i]; }
    SELAS_D float& kv(int i) { return kvb[kv0 + i]; }
    SELAS_D float kuc(int i) const { return kub[ku0 + i]; }
    SELAS_D float kvc(int i) const { return kvb[kv0 + i]; }

    __device__ __noinline__ void set(int ul, int vl, int u_, int v_, int uw_, int vw_, const float* kuSrc, const float* kvSrc, int rot_) {
        ulog2 = ul; vlog2 = vl; u = u_; v = v_; uw = uw_; vw = vw_;
        for (int i = 0; i < uw_; ++i) kub[i] = kuSrc[i];
        for (int i = 0; i < vw_; ++i) kvb[i] = kvSrc[i];
        ku0 = 0; kv0 = 0; rot = rot_;
    }
    SELAS_D void copyFrom(const PtexKernel& k) { set(k.ulog2, k.vlog2, k.u, k.v, k.uw, k.vw, k.kub + k.ku0, k.kvb + k.kv0, k.rot); }

    SELAS_D static float accumulate(const float* p, int n) {
        float r = 0;
        for (int i = 0; i < n; ++i) r += p[i];
        return r;
    }
    SELAS_D float weight() const { return accumulate(kub + ku0, uw) * accumulate(kvb + kv0, vw); }

    __device__ __noinline__ void stripZeros() {  // :97-104
        while (ku(0) == 0) { ku0++; u++; uw--; }
        while (ku(uw - 1) == 0) { uw--; }
        while (kv(0) == 0) { kv0++; v++; vw--; }
        while (kv(vw - 1) == 0) { vw--; }
    }
    // merge*: fold the part hanging over a border without neighbour back onto the edge texel (:111-147)
    __device__ __noinline__ void mergeL(int mode) {
        int w = -u;
        if (mode != 1) ku(w) += accumulate(kub + ku0, w);
        ku0 += w; uw -= w; u = 0;
    }
    __device__ __noinline__ void mergeR(int mode) {
        int w = uw + u - resu();
        int kp = uw - w;
        if (mode != 1) ku(kp - 1) += accumulate(kub + ku0 + kp, w);
        uw -= w;
    }
    __device__ __noinline__ void mergeB(int mode) {
        int w = -v;
        if (mode != 1) kv(w) += accumulate(kvb + kv0, w);
        kv0 += w; vw -= w; v = 0;
    }
    __device__ __noinline__ void mergeT(int mode) {
        int w = vw + v - resv();
        int kp = vw - w;
        if (mode != 1) kv(kp - 1) += accumulate(kvb + kv0 + kp, w);
        vw -= w;
    }
    // split*: cut the overhanging piece off into k (:149-237)
    __device__ __noinline__ void splitL(PtexKernel& k) {
        int w = -u;
        if (w < uw) {
            k.set(ulog2, vlog2, resu() - w, v, w, vw, kub + ku0, kvb + kv0, 0);
            u = 0; uw -= w; ku0 += w;
        } else {
            k.copyFrom(*this);
            k.u += resu();
            u = 0; uw = 0;
        }
    }
    __device__ __noinline__ void splitR(PtexKernel& k) {
        int w = u + uw - resu();
        if (w < uw) {
            k.set(ulog2, vlog2, 0, v, w, vw, kub + ku0 + uw - w, kvb + kv0, 0);
            uw -= w;
        } else {
            k.copyFrom(*this);
            k.u -= resu();
            u = 0; uw = 0;
        }
    }
    __device__ __noinline__ void splitB(PtexKernel& k) {
        int w = -v;
        if (w < vw) {
            k.set(ulog2, vlog2, u, resv() - w, uw, w, kub + ku0, kvb + kv0, 0);
            v = 0; vw -= w; kv0 += w;
        } else {
            k.copyFrom(*this);
            k.v += resv();
            v = 0; vw = 0;
        }
    }
    __device__ __noinline__ void splitT(PtexKernel& k) {
        int w = v + vw - resv();
        if (w < vw) {
            k.set(ulog2, vlog2, u, 0, uw, w, kub + ku0, kvb + kv0 + vw - w, 0);
            vw -= w;
        } else {
            k.copyFrom(*this);
            k.v -= resv();
            v = 0; vw = 0;
        }
    }
    __device__ __noinline__ void flipU() {
        u = resu() - u - uw;
        for (int i = 0; i < uw / 2; ++i) { float t = ku(i); ku(i) = ku(uw - 1 - i); ku(uw - 1 - i) = t; }
    }
    __device__ __noinline__ void flipV() {
        v = resv() - v - vw;
        for (int i = 0; i < vw / 2; ++i) { float t = kv(i); kv(i) = kv(vw - 1 - i); kv(vw - 1 - i) = t; }
    }
    __device__ __noinline__ void swapUV() {
        const int oldUw = uw, oldVw = vw;
        float tmpU[PTEX_KMAX], tmpV[PTEX_KMAX];
        for (int i = 0; i < oldUw; ++i) tmpU[i] = kub[ku0 + i];
        for (int i = 0; i < oldVw; ++i) tmpV[i] = kvb[kv0 + i];
        for (int i = 0; i < oldVw; ++i) kub[i] = tmpV[i];
        for (int i = 0; i < oldUw; ++i) kvb[i] = tmpU[i];
        ku0 = 0; kv0 = 0;
        int t = ulog2; ulog2 = vlog2; vlog2 = t;
        t = u; u = v; v = t;
        uw = oldVw; vw = oldUw;
    }
    __device__ __noinline__ void rotate(int r) {  // :260-270, 'r' steps counter-clockwise
        switch (r & 3) {
            default: return;
            case 1: flipU(); swapUV(); break;
            case 2: flipU(); flipV(); break;
            case 3: flipV(); swapUV(); break;
        }
        rot = (rot + r) & 3;
    }
    SELAS_D void compactU() {
        if (ku0) { for (int i = 0; i < uw; ++i) kub[i] = kub[ku0 + i]; ku0 = 0; }
    }
    SELAS_D void compactV() {
        if (kv0) { for (int i = 0; i < vw; ++i) kvb[i] = kvb[kv0 + i]; kv0 = 0; }
    }
    __device__ __noinline__ void downresU() {  // :313-339
        int src = 0, dst = 0;
        if (u & 1) { dst++; src++; uw--; }
        for (int i = uw / 2; i > 0; i--) { ku(dst++) = ku(src) + ku(src + 1); src += 2; }
        if (uw & 1) { ku(dst++) = ku(src++); }
        u /= 2;
        uw = dst;
        ulog2--;
    }
    __device__ __noinline__ void downresV() {
        int src = 0, dst = 0;
        if (v & 1) { dst++; src++; vw--; }
        for (int i = vw / 2; i > 0; i--) { kv(dst++) = kv(src) + kv(src + 1); src += 2; }
        if (vw & 1) { kv(dst++) = kv(src++); }
        v /= 2;
        vw = dst;
        vlog2--;
    }
    __device__ __noinline__ void upresU() {  // :369-379 (only reached with uw <= 5: the result fits the buffer)
        compactU();
        for (int i = uw - 1; i >= 0; --i) { float h = kub[i] / 2; kub[2 * i] = h; kub[2 * i + 1] = h; }
        uw *= 2; u *= 2; ulog2++;
    }
    __device__ __noinline__ void upresV() {
        compactV();
        for (int i = vw - 1; i >= 0; --i) { float h = kvb[i] / 2; kvb[2 * i] = h; kvb[2 * i + 1] = h; }
        vw *= 2; v *= 2; vlog2++;
    }
    __device__ __noinline__ bool adjustMainToSubface(int eid) {  // :272-303
        if (ulog2 == 0) upresU();
        if (vlog2 == 0) upresV();
        if (ulog2 > 0) ulog2--;
        if (vlog2 > 0) vlog2--;
        bool primary = false;
        int ru = resu(), rv = resv();
        switch (eid & 3) {
            case kEdgeBottom: primary = (u < ru); v -= rv; if (!primary) u -= ru; break;
            case kEdgeRight: primary = (v < rv); if (!primary) v -= rv; break;
            case kEdgeTop: primary = (u >= ru); if (primary) u -= ru; break;
            case kEdgeLeft: primary = (v >= rv); u -= ru; if (primary) v -= rv; break;
        }
        return primary;
    }
    SELAS_D void adjustSubfaceToMain(int eid) {  // :305-313
        switch (eid & 3) {
            case kEdgeBottom: v += resv(); break;
            case kEdgeRight: break;
            case kEdgeTop: u += resu(); break;
            case kEdgeLeft: u += resu(); v += resv(); break;
        }
        ulog2++; vlog2++;
    }
    __device__ __noinline__ float makeSymmetric(float initialWeight) {  // :397-455
        if (ulog2 > vlog2) { do { downresU(); } while (ulog2 > vlog2); }
        else if (vlog2 > ulog2) { do { downresV(); } while (vlog2 > ulog2); }
        uw = vw = (uw < vw ? uw : vw);
        float newWeight = 0;
        for (int i = 0; i < uw; i++) {
            float sum = ku(i) + kv(i);
            ku(i) = sum; kv(i) = sum;
            newWeight += sum;
        }
        newWeight *= newWeight;
        float scale = newWeight == 0 ? 1.f : initialWeight / newWeight;
        if (scale >= 1) {
        } else if (scale < -1) {
            for (int i = 0; i < uw; i++) ku(i) *= -1;
            newWeight = -newWeight;
        } else {
            for (int i = 0; i < uw; i++) ku(i) *= scale;
            newWeight = initialWeight;
        }
        return newWeight;
    }
};

SELAS_D float ptexTexel(const uint8_t* base, int dataType, size_t idx) {
    switch (dataType) {
        case 0: return (float)__ldg(base + idx);
        case 1: return (float)__ldg(reinterpret_cast<const uint16_t*>(base) + idx);
        case 2: return __half2float(__ushort_as_half(__ldg(reinterpret_cast<const uint16_t*>(base) + idx)));
        default: return __ldg(reinterpret_cast<const float*>(base) + idx);
    }
}

struct PtexEvalState {
    float result[3];
    float weight;
    int nchan;  // channels evaluated (<= 3)
};

// PtexSeparableKernel::applyConst (PtexUtils.cpp:643-651): result += pixel * weight
SELAS_D void ptexApplyConst(const PtexGpu& tx, PtexEvalState& st, const uint8_t* pixel, float w) {
    for (int c = 0; c < st.nchan; ++c) st.result[c] += ptexTexel(pixel, tx.dataType, c) * w;
}

// PtexSeparableFilter::apply, :348-401
static __device__ __noinline__ void ptexApply(const PtexGpu& tx, PtexEvalState& st, PtexKernel& k, int faceid, const PtexFace& f) {
    if (k.uw <= 0 || k.vw <= 0) return;
    while (k.ulog2 > f.ulog2) k.downresU();
    while (k.vlog2 > f.vlog2) k.downresV();
    // PtexReader::getData(faceid, k.res), PtexReader.cpp:785-830
    const uint8_t* constPixel = tx.constData + (size_t)faceid * tx.pixelSize;
    if (f.isConstant() || (k.ulog2 == 0 && k.vlog2 == 0)) {
        ptexApplyConst(tx, st, constPixel, k.weight());
        return;
    }
    const uint8_t* data = nullptr;
    if (k.ulog2 == f.ulog2 && k.vlog2 == f.vlog2) {
        data = tx.texels + (size_t)f.dataOffset16 * 16;
    } else {
        uint32_t want = (uint32_t)(k.ulog2 & 0xff) | ((uint32_t)(k.vlog2 & 0xff) << 8);
        for (uint32_t i = 0; i < f.redCount; ++i) {
            PtexRed r = tx.reds[f.redStart + i];
            if (r.res == want) { data = tx.texels + (size_t)r.offset16 * 16; break; }
        }
        if (!data) return;  // a reduction the loader did not foresee: contributes nothing (PtexReader errorData is zero too)
    }
    // PtexSeparableKernel.cpp:45-75 Apply<T,nChan> / ApplyS: per row sum_u texel*ku, then result += row * kv
    const int resu = k.resu();
    const int nt = tx.nchan;
    for (int y = 0; y < k.vw; ++y) {
        float row[3] = {0.f, 0.f, 0.f};
        size_t rowBase = ((size_t)(k.v + y) * resu + k.u) * nt;
        for (int c = 0; c < st.nchan; ++c) row[c] = ptexTexel(data, tx.dataType, rowBase + c) * k.kuc(0);
        for (int x = 1; x < k.uw; ++x) {
            float w = k.kuc(x);
            for (int c = 0; c < st.nchan; ++c) row[c] += ptexTexel(data, tx.dataType, rowBase + (size_t)x * nt + c) * w;
        }
        float wv = k.kvc(y);
        for (int c = 0; c < st.nchan; ++c) st.result[c] += row[c] * wv;
    }
}

// PtexBicubicFilter::kernelFn with B = 1 (PtexFilters.cpp:256-284)
SELAS_D float ptexBsplineFn(float x) {
    const float B = 1.0f;
    const float c0 = 1.5f - B, c1 = 1.5f * B - 2.5f, c2 = 1.0f - float(1.0 / 3.0) * B;
    const float c3 = float(1.0 / 3.0) * B - 0.5f, c4 = 2.5f - 1.5f * B, c5 = 2.0f * B - 4.0f, c6 = 2.0f - float(2.0 / 3.0) * B;
    x = fabsf(x);
    if (x < 1.0f) return (c0 * x + c1) * x * x + c2;
    else if (x < 2.0f) return ((c3 * x + c4) * x + c5) * x + c6;
    return 0.0f;
}
SELAS_D float ptexBlur(float x) {  // PtexFilters.cpp:133-139
    x = fabsf(x);
    return x < 1.0f ? (2.0f * x - 3.0f) * x * x + 1.0f : 0.0f;
}

// PtexWidth4Filter::buildKernelAxis with options.lerp == false (PtexFilters.cpp:141-248)
static __device__ __noinline__ void ptexBuildKernelAxis(int& kReslog2, int& kU, int& kUw, float* ku, float u, float uw, int fReslog2) {
    float minw = __uint_as_float((uint32_t)(127 - fReslog2) << 23);  // reciprocalPow2
    uw = uw > minw ? uw : minw;                                       // PtexUtils::max(uw, 1/res)
    kReslog2 = 127 - (int)((__float_as_uint(uw) >> 23) & 0xff);      // calcResFromWidth
    int resu = 1 << kReslog2;
    if (uw >= .25f) {
        if (uw < .5f) {
            kReslog2 = 2;
            float upix = u * 4.0f - 0.5f;
            int u1 = (int)ceilf(upix - 2), u2 = (int)ceilf(upix + 2);
            u1 = u1 & ~1;
            u2 = (u2 + 1) & ~1;
            kU = u1;
            kUw = u2 - u1;
            float x1 = (float)u1 - upix;
            for (int i = 0; i < kUw; i += 2) {
                float xa = x1 + (float)i, xb = xa + 1.0f;
                ku[i] = ptexBsplineFn(xa) * 1.0f;  // lerp1 = 1, lerp2 = 0: "ka*lerp1 + kc*lerp2"
                ku[i + 1] = ptexBsplineFn(xb) * 1.0f;
            }
            return;
        } else if (uw < 1) {
            kReslog2 = 1;
            float upix = u * 2.0f - 0.5f;
            kU = (int)floorf(u - .5f) * 2;
            kUw = 4;
            float x1 = (float)kU - upix;
            float s = 1.0f / (uw * 1.5f + .5f);
            for (int i = 0; i < kUw; i += 2) {
                float xa = x1 + (float)i, xb = xa + 1.0f;
                ku[i] = ptexBlur(xa * s);
                ku[i + 1] = ptexBlur(xb * s);
            }
            return;
        } else {
            kReslog2 = 0;
            float upix = u - .5f;
            kUw = 2;
            float ui = floorf(upix);
            kU = (int)ui;
            ku[0] = ptexBlur(upix - ui);
            ku[1] = 1 - ku[0];
            return;
        }
    }
    float upix = u * (float)resu - 0.5f;
    float uwpix = uw * (float)resu;
    float dupix = 2.0f * uwpix;
    int u1 = (int)ceilf(upix - dupix), u2 = (int)ceilf(upix + dupix);
    kU = u1;
    kUw = u2 - u1;
    float x1 = ((float)u1 - upix) / uwpix, step = 1.0f / uwpix;
    for (int i = 0; i < kUw; i++) ku[i] = ptexBsplineFn(x1 + (float)i * step);
}

struct PtexWork {
    PtexKernel k;
    int faceid;
};
#define PTEX_WORK_MAX 6

struct PtexWorkList {
    PtexWork items[PTEX_WORK_MAX];
    int count;
    SELAS_D void push(const PtexKernel& k, int faceid) {
        if (count < PTEX_WORK_MAX) {
            items[count].k.copyFrom(k);
            items[count].faceid = faceid;
            ++count;
        }
    }
};

// PtexSeparableFilter::applyToCornerFace, :333-345
static __device__ __noinline__ void ptexApplyToCornerFace(const PtexGpu& tx, PtexEvalState& st, PtexWorkList& wl, PtexKernel& k, const PtexFace& f, int eid,
                                   int cfid, const PtexFace& cf, int ceid) {
    bool fIsSubface = f.isSubface(), cfIsSubface = cf.isSubface();
    if (fIsSubface != cfIsSubface) {
        if (cfIsSubface) k.adjustMainToSubface(eid + 3);
        else k.adjustSubfaceToMain(eid + 3);
    }
    k.rotate(eid - ceid + 2);
    if (cfIsSubface) wl.push(k, cfid);
    else ptexApply(tx, st, k, cfid, cf);
}

// PtexSeparableFilter::applyToCorner, :257-330
static __device__ __noinline__ void ptexApplyToCorner(const PtexGpu& tx, PtexEvalState& st, PtexWorkList& wl, PtexKernel& k, int faceid, const PtexFace& f,
                               int eid) {
    int afid = faceid, aeid = eid;
    PtexFace af = f;
    bool prevIsSubface = af.isSubface();
    const int MaxValence = 10;
    int cfaceId[MaxValence];
    int cedgeId[MaxValence];
    int numCorners = 0;
    for (int i = 0; i < MaxValence; i++) {
        int prevFace = afid;
        int nextFace = af.adjface(aeid);
        aeid = (af.adjedge(aeid) + 1) % 4;
        afid = nextFace;
        if (afid < 0 || (afid == faceid && aeid == eid)) {
            numCorners = i - 2;
            break;
        }
        af = ptexLoadFace(tx, afid);
        cfaceId[i] = afid;
        cedgeId[i] = aeid;
        bool isSubface = af.isSubface();
        if (prevIsSubface && !isSubface && af.adjface((aeid + 3) % 4) == prevFace) {
            // subface "tee": the corner is the middle of a main-face edge
            bool primary = (i == 1);
            k.adjustSubfaceToMain(eid + primary * 2);
            k.rotate(eid - aeid + 3 - primary);
            wl.push(k, afid);
            return;
        }
        prevIsSubface = isSubface;
    }
    if (numCorners == 1) {
        PtexFace cf = ptexLoadFace(tx, cfaceId[1]);
        ptexApplyToCornerFace(tx, st, wl, k, f, eid, cfaceId[1], cf, cedgeId[1]);
    } else if (numCorners > 1) {
        // valence 5+: make the kernel symmetric and apply it equally to each corner face
        k.rotate(eid + 2);
        float initialWeight = k.weight();
        float newWeight = k.makeSymmetric(initialWeight);
        for (int i = 1; i <= numCorners; i++) {
            PtexKernel kc;
            kc.copyFrom(k);
            PtexFace cf = ptexLoadFace(tx, cfaceId[i]);
            ptexApplyToCornerFace(tx, st, wl, kc, f, 2, cfaceId[i], cf, cedgeId[i]);
        }
        st.weight += newWeight * (float)numCorners - initialWeight;
    } else {
        // valence 2 or 3: ignore the corner face, just take its weight out
        st.weight -= k.weight();
    }
}

// PtexSeparableFilter::applyAcrossEdge, :215-255
static __device__ __noinline__ void ptexApplyAcrossEdge(const PtexGpu& tx, PtexEvalState& st, PtexWorkList& wl, PtexKernel& k, int faceid, const PtexFace& f,
                                 int eid) {
    int afid = f.adjface(eid), aeid = f.adjedge(eid);
    PtexFace af = ptexLoadFace(tx, afid);
    int rot = eid - aeid + 2;
    bool fIsSubface = f.isSubface(), afIsSubface = af.isSubface();
    if (fIsSubface != afIsSubface) {
        if (afIsSubface) {
            bool primary = k.adjustMainToSubface(eid);
            if (!primary) {
                int neid = (aeid + 3) % 4;
                int nfid = af.adjface(neid);
                aeid = af.adjedge(neid);
                afid = nfid;
                af = ptexLoadFace(tx, afid);
                rot += neid - aeid + 2;
            }
        } else {
            bool primary = (af.adjface(aeid) == faceid);
            k.adjustSubfaceToMain(eid - primary);
        }
    }
    k.rotate(rot);
    if (afIsSubface) wl.push(k, afid);
    else ptexApply(tx, st, k, afid, af);
}

// PtexSeparableFilter::splitAndApply, :135-213 (noedgeblend = false)
static __device__ __noinline__ void ptexSplitAndApply(const PtexGpu& tx, PtexEvalState& st, PtexWorkList& wl, PtexKernel& k, int faceid, const PtexFace& f) {
    bool splitR = (k.u + k.uw > k.resu()), splitL = (k.u < 0);
    bool splitT = (k.v + k.vw > k.resv()), splitB = (k.v < 0);
    if (splitR || splitL || splitT || splitB) {
        PtexKernel ka, kc;
        if (splitR) {
            if (f.adjface(kEdgeRight) >= 0) {
                k.splitR(ka);
                if (splitT) {
                    if (f.adjface(kEdgeTop) >= 0) { ka.splitT(kc); ptexApplyToCorner(tx, st, wl, kc, faceid, f, kEdgeTop); }
                    else ka.mergeT(tx.vMode);
                }
                if (splitB) {
                    if (f.adjface(kEdgeBottom) >= 0) { ka.splitB(kc); ptexApplyToCorner(tx, st, wl, kc, faceid, f, kEdgeRight); }
                    else ka.mergeB(tx.vMode);
                }
                ptexApplyAcrossEdge(tx, st, wl, ka, faceid, f, kEdgeRight);
            } else k.mergeR(tx.uMode);
        }
        if (splitL) {
            if (f.adjface(kEdgeLeft) >= 0) {
                k.splitL(ka);
                if (splitT) {
                    if (f.adjface(kEdgeTop) >= 0) { ka.splitT(kc); ptexApplyToCorner(tx, st, wl, kc, faceid, f, kEdgeLeft); }
                    else ka.mergeT(tx.vMode);
                }
                if (splitB) {
                    if (f.adjface(kEdgeBottom) >= 0) { ka.splitB(kc); ptexApplyToCorner(tx, st, wl, kc, faceid, f, kEdgeBottom); }
                    else ka.mergeB(tx.vMode);
                }
                ptexApplyAcrossEdge(tx, st, wl, ka, faceid, f, kEdgeLeft);
            } else k.mergeL(tx.uMode);
        }
        if (splitT) {
            if (f.adjface(kEdgeTop) >= 0) { k.splitT(ka); ptexApplyAcrossEdge(tx, st, wl, ka, faceid, f, kEdgeTop); }
            else k.mergeT(tx.vMode);
        }
        if (splitB) {
            if (f.adjface(kEdgeBottom) >= 0) { k.splitB(ka); ptexApplyAcrossEdge(tx, st, wl, ka, faceid, f, kEdgeBottom); }
            else k.mergeB(tx.vMode);
        }
    }
    ptexApply(tx, st, k, faceid, f);
}


// ---------------------------------------------------------------------------------------------------------------
// Register-resident fast path.  The general evaluator above keeps its kernels in local memory (dynamic indexing of the
// weight windows); measured at 1.1 G lookups/s it is far from the 104-B/lookup roofline.  The overwhelmingly common
// cases -- a 4x4 kernel that stays inside its face, or crosses exactly one edge onto an ordinary neighbour of matching
// resolution (or onto the mesh border) -- are handled here with compile-time indexed 4-element arrays.  Every sum is
// formed in the same order as the general path, so the two agree bit for bit; anything else (valence != 4 corners, subfaces,
// faces below 8 texels, kernels that are not 4 wide) returns false and takes the general path.
// ---------------------------------------------------------------------------------------------------------------
struct PtexPiece {  // a (sub)kernel: window origin, extent and weights compacted to index 0
    int u, v, uw, vw;
    float ku[4], kv[4];
};

SELAS_D void ptexReverse4(float* k, int n) {
    if (n == 2) { float t = k[0]; k[0] = k[1]; k[1] = t; }
    else if (n == 3) { float t = k[0]; k[0] = k[2]; k[2] = t; }
    else if (n == 4) { float t = k[0]; k[0] = k[3]; k[3] = t; t = k[1]; k[1] = k[2]; k[2] = t; }
}
// PtexSeparableKernel::rotate on a piece (resu/resv = resolution the piece currently lives in)
SELAS_D void ptexRotatePiece(PtexPiece& p, int rot, int& ulog2, int& vlog2) {
    rot &= 3;
    if (rot == 0) return;
    if (rot == 1 || rot == 2) {  // flipU
        p.u = (1 << ulog2) - p.u - p.uw;
        ptexReverse4(p.ku, p.uw);
    }
    if (rot == 2 || rot == 3) {  // flipV
        p.v = (1 << vlog2) - p.v - p.vw;
        ptexReverse4(p.kv, p.vw);
    }
    if (rot == 1 || rot == 3) {  // swapUV
        int t = ulog2; ulog2 = vlog2; vlog2 = t;
        t = p.u; p.u = p.v; p.v = t;
        t = p.uw; p.uw = p.vw; p.vw = t;
#pragma unroll
        for (int i = 0; i < 4; ++i) { float w = p.ku[i]; p.ku[i] = p.kv[i]; p.kv[i] = w; }
    }
}
SELAS_D float ptexPick4(const float* k, int i) { return i == 0 ? k[0] : (i == 1 ? k[1] : (i == 2 ? k[2] : k[3])); }
// k'[i] = k[i + w] (0 past the end): the "ku += w" of the library on a register array
SELAS_D void ptexShift4(float* k, int w) {
    float t0 = ptexPick4(k, w), t1 = w + 1 < 4 ? ptexPick4(k, w + 1) : 0.f, t2 = w + 2 < 4 ? ptexPick4(k, w + 2) : 0.f;
    float t3 = w + 3 < 4 ? ptexPick4(k, w + 3) : 0.f;
    k[0] = t0; k[1] = t1; k[2] = t2; k[3] = t3;
}
SELAS_D float ptexSum4(const float* k, int n) {  // accumulate(): sequential from 0
    float r = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (i < n) r += k[i];
    return r;
}
template <int DT>
SELAS_D float ptexTexelT(const uint8_t* base, size_t idx) {
    if (DT == 0) return (float)__ldg(base + idx);
    if (DT == 1) return (float)__ldg(reinterpret_cast<const uint16_t*>(base) + idx);
    if (DT == 2) return __half2float(__ushort_as_half(__ldg(reinterpret_cast<const uint16_t*>(base) + idx)));
    return __ldg(reinterpret_cast<const float*>(base) + idx);
}
// PtexSeparableKernel::Apply on native-resolution data of a face (or applyConst for a constant face).  The data type is a template
// parameter so that the per-texel switch is decided once per lookup; NT = 3 is the three-channel texture read with every texel at
// a constant offset from one row pointer (the generic form spent 28 % of the kernel's instructions on 64-bit index arithmetic).
// `data` / `ulog2`: the texel block the piece is applied to and its row length -- the face's native data, or the reduced copy
// PtexReader::getData(faceid, k.res) hands a kernel whose resolution is below the face's (PtexReader.cpp:785-830); `constant`: the
// face's constant pixel instead (constant face, or a kernel reduced to 1 x 1).
template <int DT, int NT>
SELAS_D void ptexApplyPiece(const PtexGpu& tx, bool constant, int faceid, const uint8_t* data, int ulog2, const PtexPiece& p, int nchan,
                            float* result) {
    if (p.uw <= 0 || p.vw <= 0) return;
    if (constant) {
        float w = ptexSum4(p.ku, p.uw) * ptexSum4(p.kv, p.vw);
        const uint8_t* px = tx.constData + (size_t)faceid * tx.pixelSize;
        for (int c = 0; c < nchan; ++c) result[c] += ptexTexelT<DT>(px, c) * w;
        return;
    }
    constexpr int kTexelBytes = DT == 0 ? 1 : (DT == 3 ? 4 : 2);
    const int resu = 1 << ulog2, nt = NT ? NT : tx.nchan;
    const uint8_t* rowPtr = data + ((size_t)p.v * resu + p.u) * (size_t)(nt * kTexelBytes);
    const size_t rowStride = (size_t)resu * (size_t)(nt * kTexelBytes);
#pragma unroll
    for (int y = 0; y < 4; ++y) {
        if (y < p.vw) {
            float row[3] = {0.f, 0.f, 0.f};
            // u8 RGB rows: the 12 bytes of the four texels come in as four aligned 32-bit words and three funnel shifts instead of
            // twelve single-byte loads (profiles/r1_ptex_kernel_metrics.csv: 48 byte loads per piece); the pool is padded by 16 bytes
            // and every face block starts 16-byte aligned, so the words around the row are always inside the allocation
            uint32_t packed[3] = {0u, 0u, 0u};
            if (DT == 0 && NT == 3) {
                const uint32_t a = (uint32_t)(reinterpret_cast<uintptr_t>(rowPtr) & 3u);
                const uint32_t* wp = reinterpret_cast<const uint32_t*>(rowPtr - a);
                const uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2), w3 = __ldg(wp + 3);
                packed[0] = __funnelshift_r(w0, w1, a * 8u);
                packed[1] = __funnelshift_r(w1, w2, a * 8u);
                packed[2] = __funnelshift_r(w2, w3, a * 8u);
            }
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                if (x < p.uw) {
                    float w = p.ku[x];
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        if (NT == 3 || c < nchan) {
                            float t;
                            if (DT == 0 && NT == 3) {
                                const int k = 3 * x + c;
                                t = (float)((packed[k >> 2] >> (8 * (k & 3))) & 0xffu);
                            } else {
                                t = ptexTexelT<DT>(rowPtr, (size_t)(x * nt + c));
                            }
                            row[c] = (x == 0) ? t * w : row[c] + t * w;
                        }
                    }
                }
            }
            float wv = p.kv[y];
#pragma unroll
            for (int c = 0; c < 3; ++c)
                if (NT == 3 || c < nchan) result[c] += row[c] * wv;
        }
        rowPtr += rowStride;
    }
}

// One axis of the kernel for a face of >= 8 texels at filter width 0: four b-spline weights starting at texel k0.
SELAS_D bool ptexFastAxis(float u, int reslog2, int& k0, int& kw, float* k) {
    const float uw = __uint_as_float((uint32_t)(127 - reslog2) << 23);
    const float resf = (float)(1 << reslog2);
    float upix = u * resf - 0.5f;
    float uwpix = uw * resf;  // == 1
    float dupix = 2.0f * uwpix;
    int u1 = (int)ceilf(upix - dupix), u2 = (int)ceilf(upix + dupix);
    if (u2 - u1 != 4) return false;
    float x1 = ((float)u1 - upix) / uwpix, step = 1.0f / uwpix;
#pragma unroll
    for (int i = 0; i < 4; ++i) k[i] = ptexBsplineFn(x1 + (float)i * step);
    // stripZeros
    int a = 0, b = 3;
    if (k[0] == 0) { a = 1; if (k[1] == 0) return false; }
    if (k[3] == 0) { b = 2; if (k[2] == 0) return false; }
    if (a == 1) { k[0] = k[1]; k[1] = k[2]; k[2] = k[3]; k[3] = 0.f; }
    k0 = u1 + a;
    kw = b - a + 1;
    return true;
}

// One axis of a piece, as (origin, extent, weights compacted to index 0).  The helpers below are the library's
// splitL/R/B/T and mergeL/R/B/T (PtexSeparableKernel.h:149-237) on that representation, with every array index a
// compile-time constant so the weights stay in registers.
//   split: the w texels hanging over the low / high side leave as (o0, on, ko); the rest stays in (k0, kn, kk)
SELAS_D void ptexSplitAxis(int& k0, int& kn, float* kk, int res, bool lowSide, int& o0, int& on, float* ko) {
    const int w = lowSide ? -k0 : k0 + kn - res;
#pragma unroll
    for (int i = 0; i < 4; ++i) ko[i] = kk[i];
    if (!lowSide) ptexShift4(ko, kn - w);
    o0 = lowSide ? res - w : 0;
    on = w;
    if (lowSide) { ptexShift4(kk, w); k0 = 0; }
    kn -= w;
}
//   merge (mesh border): the overhang folds onto the edge texel; border mode black (1) drops it
SELAS_D void ptexMergeAxis(int& k0, int& kn, float* kk, int res, bool lowSide, int mode) {
    const int w = lowSide ? -k0 : k0 + kn - res;
    if (lowSide) {
        float acc = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i < w) acc += kk[i];
        float edge = ptexPick4(kk, w);
        if (mode != 1) edge += acc;
        ptexShift4(kk, w);
        kk[0] = edge;
        k0 = 0;
        kn -= w;
    } else {
        const int newKn = kn - w;
        float acc = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i >= newKn && i < kn) acc += kk[i];
        if (mode != 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (i == newKn - 1) kk[i] += acc;
        }
        kn = newKn;
    }
}

// PtexSeparableKernel::downresU / downresV (PtexSeparableKernel.h:313-367) on one axis of a piece: neighbouring weights are added in
// pairs aligned to the coarser grid (an odd origin leaves the first weight alone), in the library's operand order.
SELAS_D void ptexDownresAxis(int& k0, int& kn, float* kk) {
    const float e0 = kk[0], e1 = kn > 1 ? kk[1] : 0.f, e2 = kn > 2 ? kk[2] : 0.f, e3 = kn > 3 ? kk[3] : 0.f;
    if (k0 & 1) {
        kk[0] = e0;
        kk[1] = kn > 2 ? e1 + e2 : e1;
        kk[2] = e3;
        kk[3] = 0.f;
        kn = 1 + (kn >> 1);  // 1 + ceil((kn - 1) / 2)
    } else {
        kk[0] = kn > 1 ? e0 + e1 : e0;
        kk[1] = kn > 3 ? e2 + e3 : e2;
        kk[2] = 0.f;
        kk[3] = 0.f;
        kn = (kn + 1) >> 1;
    }
    k0 >>= 1;
}

// splitAndApply (PtexSeparableFilter.cpp:135-213) for a 4x4 kernel on an ordinary face of >= 8x8 texels.  The library applies up to
// four pieces in a fixed order -- across the corner, across the u edge, across the v edge, the face itself -- and this walks the same
// four slots in ONE loop: per axis a slot's piece is the part that leaves the face, the part that stays, the whole window, or (at a
// mesh border) the window with the overhang merged onto the edge texel.  One copy of the code serves every slot, a warp holding
// interior, edge and corner lookups stays converged (slots a lane does not need are predicated off), and the sums are formed in
// the library's order, so the result matches the general path bit for bit.  Neighbours of another resolution are served here too
// since round 2 (the piece is reduced, or the neighbour's reduced copy is read).  Anything else -- subfaces, valence 5+ corners,
// faces below 8 texels, kernels that are not 4 wide -- returns false and takes the general path.  Written
// with status flags instead of early exits: with `return` in the middle the compiler moved the warp's reconvergence point to the
// end of the lookup and the variants ran one after the other (measured 17.7 ms for 16 M mixed lookups).
#ifndef SELAS_PTEX_FAST_MIXED
#define SELAS_PTEX_FAST_MIXED 1
#endif
enum { kAxisWhole = 0, kAxisLeaving = 1, kAxisStaying = 2, kAxisMerged = 3 };
SELAS_D void ptexDeriveAxis(int k0, int kn, const float* kk, int res, bool lowSide, int how, int mode, int& o0, int& on, float* ko) {
    float t[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { t[i] = kk[i]; ko[i] = kk[i]; }
    o0 = k0;
    on = kn;
    if (how == kAxisLeaving) {
        int tk0 = k0, tkn = kn;
        ptexSplitAxis(tk0, tkn, t, res, lowSide, o0, on, ko);
    } else if (how == kAxisStaying) {
        int d0, dn;
        ptexSplitAxis(o0, on, ko, res, lowSide, d0, dn, t);
    } else if (how == kAxisMerged) {
        ptexMergeAxis(o0, on, ko, res, lowSide, mode);
    }
}

template <int DT, int NT, bool MIXED>
SELAS_D bool ptexEvalFastT(const PtexGpu& tx, const PtexFace& f, int faceid, float u, float v, int nchan, float oneValue, float* rgb) {
    PtexPiece k;
    bool ok = !(f.isSubface() || f.isConstant() || f.ulog2 < 3 || f.vlog2 < 3);
    ok = ok && ptexFastAxis(u, f.ulog2, k.u, k.uw, k.ku);
    ok = ok && ptexFastAxis(v, f.vlog2, k.v, k.vw, k.kv);
    const int resu = 1 << f.ulog2, resv = 1 << f.vlog2;
    const bool splitR = ok && (k.u + k.uw > resu), splitL = ok && (k.u < 0), splitT = ok && (k.v + k.vw > resv), splitB = ok && (k.v < 0);
    ok = ok && !((splitR && splitL) || (splitT && splitB));
    const bool splitU = splitR || splitL, splitV = splitT || splitB;
    const int eU = splitR ? kEdgeRight : kEdgeLeft, eV = splitT ? kEdgeTop : kEdgeBottom;
    const bool adjU = f.adjface(eU) >= 0, adjV = f.adjface(eV) >= 0;
    float weight = ok ? ptexSum4(k.ku, k.uw) * ptexSum4(k.kv, k.vw) : 1.0f;
    float result[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
    for (int s = 0; s < 4; ++s) {
        const bool viaU = (s == 0 || s == 1), viaV = (s == 0 || s == 2);  // the slot's piece leaves through the u / v edge
        const bool active = ok && (!viaU || (splitU && adjU)) && (!viaV || (splitV && adjV));
        if (active) {
            PtexPiece p;
            const int howU = !splitU ? kAxisWhole : (adjU ? (viaU ? kAxisLeaving : kAxisStaying) : kAxisMerged);
            const int howV = !splitV ? kAxisWhole : (adjV ? (viaV ? kAxisLeaving : kAxisStaying) : kAxisMerged);
            ptexDeriveAxis(k.u, k.uw, k.ku, resu, splitL, howU, tx.uMode, p.u, p.uw, p.ku);
            ptexDeriveAxis(k.v, k.vw, k.kv, resv, splitB, howV, tx.vMode, p.v, p.vw, p.kv);
            int tfid = faceid, rot = 0;
            PtexFace tf = f;
            bool apply = true;
            if (s == 0) {
                // applyToCorner (:257-330) without subfaces: walk around the corner; an ordinary valence-4 corner applies the piece to
                // the diagonal face, a walk that hits the mesh border or closes early (valence 2, 3) only takes the weight out
                const int eid = splitR ? (splitT ? kEdgeTop : kEdgeRight) : (splitT ? kEdgeLeft : kEdgeBottom);
                int afid = faceid, aeid = eid, numCorners = 0, state = 0;  // state: 0 walking, 1 finished, 2 general path
                PtexFace af = f;
#pragma unroll 1
                for (int i = 0; i < 4; ++i) {
                    if (state == 0) {
                        const int next = af.adjface(aeid);
                        aeid = (af.adjedge(aeid) + 1) & 3;
                        afid = next;
                        if (afid < 0 || (afid == faceid && aeid == eid)) {
                            numCorners = i - 2;
                            state = 1;
                        } else if (i == 3) {
                            state = 2;  // valence 5+
                        } else {
                            af = ptexLoadFace(tx, afid);
                            if (af.isSubface()) state = 2;
                            if (i == 1) { tfid = afid; tf = af; rot = eid - aeid + 2; }
                        }
                    }
                }
                ok = state == 1;
                if (ok && numCorners != 1) {
                    weight -= ptexSum4(p.ku, p.uw) * ptexSum4(p.kv, p.vw);
                    apply = false;
                }
            } else if (s < 3) {
                // applyAcrossEdge (:215-255) between two ordinary faces
                const int e = s == 1 ? eU : eV;
                tfid = f.adjface(e);
                tf = ptexLoadFace(tx, tfid);
                rot = e - f.adjedge(e) + 2;
            }
            if (ok && apply) {
                int ul = f.ulog2, vl = f.vlog2;
                ptexRotatePiece(p, rot, ul, vl);
                ok = !tf.isSubface();  // else: the subface transitions of applyAcrossEdge / applyToCornerFace
                // !MIXED: the instantiation for textures whose faces all share one resolution -- without the reduction code the kernel keeps
                // its registers (with it, uniform textures lose 6-10 %: profiles/r2_ptex_classes_mixed.json); a mismatch cannot occur there
                if (!MIXED || !SELAS_PTEX_FAST_MIXED) ok = ok && ul == tf.ulog2 && vl == tf.vlog2;
                if (ok && !MIXED) ptexApplyPiece<DT, NT>(tx, tf.isConstant(), tfid, tx.texels + (size_t)tf.dataOffset16 * 16, ul, p, nchan, result);
                if (ok && MIXED) {
                    // PtexSeparableFilter::apply (:346-401): a coarser face gets the piece reduced to its resolution, a finer one
                    // is read through its reduced copy at the piece's resolution
#pragma unroll 1
                    while (ul > tf.ulog2) { ptexDownresAxis(p.u, p.uw, p.ku); --ul; }
#pragma unroll 1
                    while (vl > tf.vlog2) { ptexDownresAxis(p.v, p.vw, p.kv); --vl; }
                    const bool constant = tf.isConstant() || (ul == 0 && vl == 0);
                    const uint8_t* data = tx.texels + (size_t)tf.dataOffset16 * 16;
                    bool found = true;
                    if (!constant && (ul != tf.ulog2 || vl != tf.vlog2)) {
                        const uint32_t want = (uint32_t)(ul & 0xff) | ((uint32_t)(vl & 0xff) << 8);
                        found = false;
                        for (uint32_t i = 0; i < tf.redCount && !found; ++i) {
                            const PtexRed r = tx.reds[tf.redStart + i];
                            if (r.res == want) { data = tx.texels + (size_t)r.offset16 * 16; found = true; }
                        }
                    }
                    // a reduction the loader did not foresee contributes nothing (as in ptexApply)
                    if (found) ptexApplyPiece<DT, NT>(tx, constant, tfid, data, ul, p, nchan, result);
                }
            }
        }
    }
    if (ok) {
        float scale = 1.0f / (weight * oneValue);
        for (int c = 0; c < nchan; ++c) rgb[c] = result[c] * scale;
    }
    return ok;
}

template <bool MIXED>
SELAS_D bool ptexEvalFast(const PtexGpu& tx, const PtexFace& f, int faceid, float u, float v, int nchan, float oneValue, float* rgb) {
    if (tx.nchan == 3) {  // RGB texture, all three channels evaluated
        switch (tx.dataType) {
            case 0: return ptexEvalFastT<0, 3, MIXED>(tx, f, faceid, u, v, nchan, oneValue, rgb);
            case 1: return ptexEvalFastT<1, 3, MIXED>(tx, f, faceid, u, v, nchan, oneValue, rgb);
            case 2: return ptexEvalFastT<2, 3, MIXED>(tx, f, faceid, u, v, nchan, oneValue, rgb);
            default: return ptexEvalFastT<3, 3, MIXED>(tx, f, faceid, u, v, nchan, oneValue, rgb);
        }
    }
    switch (tx.dataType) {
        case 0: return ptexEvalFastT<0, 0, MIXED>(tx, f, faceid, u, v, nchan, oneValue, rgb);
        case 1: return ptexEvalFastT<1, 0, MIXED>(tx, f, faceid, u, v, nchan, oneValue, rgb);
        case 2: return ptexEvalFastT<2, 0, MIXED>(tx, f, faceid, u, v, nchan, oneValue, rgb);
        default: return ptexEvalFastT<3, 0, MIXED>(tx, f, faceid, u, v, nchan, oneValue, rgb);
    }
}

// PtexSeparableFilter::eval, :47-132, with firstChan 0, nChannels 3, all widths 0, width 1, blur 0, split in three steps:
//   ptexPrologue   early outs (bad face id, constant neighbourhood, border modes); returns true when a filter is needed
//   ptexEvalFast   register-resident common cases (above)
//   ptexGeneral    everything else (the library's full splitAndApply logic)
// Channels the texture does not have are returned as 0 (the reference leaves them uninitialised).
SELAS_D bool ptexPrologue(const PtexGpu& tx, int faceid, float& u, float& v, PtexFace& f, float* rgb) {
    rgb[0] = rgb[1] = rgb[2] = 0.f;
    if (faceid < 0 || faceid >= tx.numFaces) return false;
    const int nchan = tx.nchan < 3 ? tx.nchan : 3;
    f = ptexLoadFace(tx, faceid);
    if (f.flags & kPtexFlagNbConstant) {  // :62-69, ConvertToFloat of the face's constant pixel
        const float oneValueInv = tx.dataType == 0 ? 1.f / 255.f : (tx.dataType == 1 ? 1.f / 65535.f : 1.f);
        const uint8_t* px = tx.constData + (size_t)faceid * tx.pixelSize;
        for (int c = 0; c < nchan; ++c) {
            float t = ptexTexel(px, tx.dataType, c);
            rgb[c] = tx.dataType == 3 ? t : t * oneValueInv;
        }
        return false;
    }
    bool black = false;
    switch (tx.uMode) {
        case 0: u = clampf_(u, 0.0f, 1.0f); break;
        case 2: u = u - floorf(u); break;
        case 1: if (u <= -1.0f || u >= 2.0f) black = true; break;
    }
    switch (tx.vMode) {
        case 0: v = clampf_(v, 0.0f, 1.0f); break;
        case 2: v = v - floorf(v); break;
        case 1: if (v <= -1.0f || v >= 2.0f) black = true; break;
    }
    return !black;
}

static __device__ __noinline__ void ptexGeneral(const PtexGpu* texPtr, int faceid, float u, float v, float* rgb) {
    const PtexGpu tx = *texPtr;
    PtexEvalState st;
    st.nchan = tx.nchan < 3 ? tx.nchan : 3;
    st.result[0] = st.result[1] = st.result[2] = 0.f;
    const float oneValue = tx.dataType == 0 ? 255.f : (tx.dataType == 1 ? 65535.f : 1.f);
    PtexFace f = ptexLoadFace(tx, faceid);
    float uw = 0.f, vw = 0.f;
    PtexWorkList wl;
    wl.count = 0;
    PtexKernel k;
    k.ku0 = k.kv0 = 0;
    k.rot = 0;
    if (f.isSubface()) {
        uw = uw * 1.0f + 0.0f * 2.0f;
        vw = vw * 1.0f + 0.0f * 2.0f;
        ptexBuildKernelAxis(k.ulog2, k.u, k.uw, k.kub, u * .5f, uw * .5f, f.ulog2 + 1);
        ptexBuildKernelAxis(k.vlog2, k.v, k.vw, k.kvb, v * .5f, vw * .5f, f.vlog2 + 1);
        if (k.ulog2 == 0) k.upresU();
        if (k.vlog2 == 0) k.upresV();
        k.ulog2--;
        k.vlog2--;
    } else {
        ptexBuildKernelAxis(k.ulog2, k.u, k.uw, k.kub, u, uw, f.ulog2);
        ptexBuildKernelAxis(k.vlog2, k.v, k.vw, k.kvb, v, vw, f.vlog2);
    }
    k.stripZeros();
    st.weight = k.weight();
    ptexSplitAndApply(tx, st, wl, k, faceid, f);
    // kernels that crossed onto a subface are re-split there (the library recurses; see the header comment)
    for (int guard = 0; wl.count > 0 && guard < 32; ++guard) {
        --wl.count;
        PtexKernel kk;
        kk.copyFrom(wl.items[wl.count].k);
        int fid = wl.items[wl.count].faceid;
        PtexFace ff = ptexLoadFace(tx, fid);
        ptexSplitAndApply(tx, st, wl, kk, fid, ff);
    }
    float scale = 1.0f / (st.weight * oneValue);
    rgb[0] = rgb[1] = rgb[2] = 0.f;
    for (int c = 0; c < st.nchan; ++c) rgb[c] = st.result[c] * scale;
}

// Resident CTAs per SM the lookup kernels are compiled for: 4 -> 128 registers (the four-slot loop wants 158 uncapped; 5 -> 96 spills).
#ifndef SELAS_PTEX_MIN_CTAS
#define SELAS_PTEX_MIN_CTAS 4
#endif

// Tile-cooperative lookups: a CTA of BLOCK threads walks tiles of BLOCK * PER_THREAD lookups.  Phase 1 runs the early
// outs and the register-resident fast path for every lookup of the tile; the ones it cannot take (corners, subfaces, mixed
// resolutions: ~14 % on 8x8 faces) are queued in shared memory.  Phase 2 spreads the queue over all threads, so the long
// general path runs in dense warps and its latency is amortised over the whole tile instead of being paid once per warp.
//   fetch(i, tex, faceid, u, v) -> bool   describes lookup i (false: nothing to do for i)
//   store(i, rgb)                         receives the result
template <int BLOCK, int PER_THREAD, bool MIXED, typename Fetch, typename Store>
SELAS_D void ptexTileEval(size_t n, Fetch fetch, Store store) {
    struct Item {
        const PtexGpu* tex;
        uint32_t index;
        int faceid;
        float u, v;
    };
    constexpr int TILE = BLOCK * PER_THREAD;
    __shared__ int failCount;
    __shared__ Item failItem[TILE];
    for (size_t base = (size_t)blockIdx.x * TILE; base < n; base += (size_t)gridDim.x * TILE) {
        if (threadIdx.x == 0) failCount = 0;
        __syncthreads();
#pragma unroll 1
        for (int k = 0; k < PER_THREAD; ++k) {
            size_t i = base + (size_t)k * BLOCK + threadIdx.x;
            const PtexGpu* texPtr = nullptr;
            int faceid = 0;
            float u = 0.f, v = 0.f;
            // flags instead of `continue`, so that the warp reconverges after every step
            const bool live = i < n && fetch(i, texPtr, faceid, u, v);
            float rgb[3] = {0.f, 0.f, 0.f};
            bool done = live;
            if (live) {
                const PtexGpu tx = *texPtr;
                PtexFace f;
                if (ptexPrologue(tx, faceid, u, v, f, rgb)) {
                    const float oneValue = tx.dataType == 0 ? 255.f : (tx.dataType == 1 ? 65535.f : 1.f);
                    if (!ptexEvalFast<MIXED>(tx, f, faceid, u, v, tx.nchan < 3 ? tx.nchan : 3, oneValue, rgb)) {
                        int slot = atomicAdd(&failCount, 1);
                        failItem[slot] = Item{texPtr, (uint32_t)(i - base), faceid, u, v};
                        done = false;
                    }
                }
            }
            if (done) store(i, rgb);
        }
        __syncthreads();
        const int nf = failCount;
        for (int j = threadIdx.x; j < nf; j += BLOCK) {
            Item it = failItem[j];
            float rgb[3];
            ptexGeneral(it.tex, it.faceid, it.u, it.v, rgb);
            store(base + it.index, rgb);
        }
        __syncthreads();
    }
}

}  // namespace selas
