// oracle/cpu_bvh.h -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// CPU restatement of the ray-cast semantics the reference gets from Embree 3.2.0 (third-party, not under
// /root/reference; pinned by Middleware/embree-latest/include/embree3/rtcore_version.h:17-21).  The
// semantics followed are the documented ones:
//   * closest hit with t in [tnear, tfar]           share/man/man3/rtcIntersect1.3embree3:24
//   * Ng unnormalised, object space, = cross(v1-v0, v2-v0)  rtcIntersect1.3embree3:53,
//     RTC_GEOMETRY_TYPE_TRIANGLE.3embree3:52
//   * triangle (u,v) are the weights of v1 and v2   RTC_GEOMETRY_TYPE_TRIANGLE.3embree3:47
//   * quad = triangles (v0,v1,v3) and (v2,v3,v1); on the second u = 1-u', v = 1-v'
//                                                   RTC_GEOMETRY_TYPE_QUAD.3embree3:38-41
//   * no back-face culling; any-hit reports only "something in [tnear,tfar]"  rtcOccluded1.3embree3:43
// "parity unpinned": the reference holds no golden vectors for this boundary and Embree itself cannot be
// built here, so hit-ID parity is pinned only against the double-precision brute force below.
//
// A binned-SAH BVH2 with a scalar stack traversal; the triangle test is Moller-Trumbore evaluated in
// Real (double for the hit-ID oracle, float for the timed CPU baseline).
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace orc {

struct Tri {
    float v0[3], v1[3], v2[3];
    uint32_t primID, geomID, instID;
    uint32_t second;  // 1 = second half of a quad (u,v flipped on output)
};

struct RayHit {
    double t, u, v;
    uint32_t tri;  // index into tris, 0xFFFFFFFF = miss
    // a hit on a curve segment (orc_curves.h; set by the scene-level trace of selas_oracle.cpp): index into the scene's segment list,
    // and the unnormalised surface normal there
    uint32_t curve = 0xFFFFFFFFu;
    double ng[3] = {0.0, 0.0, 0.0};
};

template <typename Real>
static inline bool intersect_tri(const Tri& T, const Real o[3], const Real d[3], Real tnear, Real tfar, Real& t,
                                 Real& u, Real& v) {
    Real e1[3], e2[3], p[3], s[3], q[3];
    for (int k = 0; k < 3; ++k) {
        e1[k] = (Real)T.v1[k] - (Real)T.v0[k];
        e2[k] = (Real)T.v2[k] - (Real)T.v0[k];
    }
    p[0] = d[1] * e2[2] - d[2] * e2[1];
    p[1] = d[2] * e2[0] - d[0] * e2[2];
    p[2] = d[0] * e2[1] - d[1] * e2[0];
    Real det = e1[0] * p[0] + e1[1] * p[1] + e1[2] * p[2];
    if (det == (Real)0) return false;
    Real inv = (Real)1 / det;
    for (int k = 0; k < 3; ++k) s[k] = o[k] - (Real)T.v0[k];
    Real uu = (s[0] * p[0] + s[1] * p[1] + s[2] * p[2]) * inv;
    if (uu < (Real)0 || uu > (Real)1) return false;
    q[0] = s[1] * e1[2] - s[2] * e1[1];
    q[1] = s[2] * e1[0] - s[0] * e1[2];
    q[2] = s[0] * e1[1] - s[1] * e1[0];
    Real vv = (d[0] * q[0] + d[1] * q[1] + d[2] * q[2]) * inv;
    if (vv < (Real)0 || uu + vv > (Real)1) return false;
    Real tt = (e2[0] * q[0] + e2[1] * q[1] + e2[2] * q[2]) * inv;
    if (!(tt >= tnear && tt <= tfar)) return false;
    t = tt;
    u = uu;
    v = vv;
    return true;
}

struct Node {
    float lo[3], hi[3];
    uint32_t left;   // inner: index of left child (right = left+1); leaf: first triangle
    uint32_t count;  // 0 = inner
};

class Bvh {
public:
    std::vector<Tri> tris;
    std::vector<Node> nodes;

    // lbvhBelow > 0 (tree-quality tooling): ranges of at most that many triangles are finished as an LBVH over cubic Morton codes, i.e. the
    // binned SAH only shapes the top of the tree.
    // lbvhAbove > 0 is the opposite: Morton splits for ranges larger than that, binned SAH inside them.
    void build(uint32_t lbvhBelow = 0, uint32_t lbvhAbove = 0) {
        const size_t n = tris.size();
        nodes.clear();
        if (n == 0) return;
        nodes.reserve(2 * n);
        std::vector<float> cen(3 * n), blo(3 * n), bhi(3 * n);
        order_.resize(n);
        for (size_t i = 0; i < n; ++i) {
            order_[i] = (uint32_t)i;
            for (int k = 0; k < 3; ++k) {
                float a = tris[i].v0[k], b = tris[i].v1[k], c = tris[i].v2[k];
                float lo = std::min(a, std::min(b, c)), hi = std::max(a, std::max(b, c));
                blo[3 * i + k] = lo;
                bhi[3 * i + k] = hi;
                cen[3 * i + k] = 0.5f * (lo + hi);
            }
        }
        std::vector<uint64_t> codes;
        if (lbvhBelow || lbvhAbove) {
            float slo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, shi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
            for (size_t i = 0; i < n; ++i)
                for (int k = 0; k < 3; ++k) { slo[k] = std::min(slo[k], blo[3 * i + k]); shi[k] = std::max(shi[k], bhi[3 * i + k]); }
            const float longest = std::max(shi[0] - slo[0], std::max(shi[1] - slo[1], shi[2] - slo[2]));
            codes.resize(n);
            for (size_t i = 0; i < n; ++i) {
                uint64_t q[3];
                for (int k = 0; k < 3; ++k) {
                    float nrm = longest > 0.f ? (cen[3 * i + k] - slo[k]) / longest : 0.f;
                    q[k] = (uint64_t)std::min(std::max(nrm * 2097152.0f, 0.0f), 2097151.0f);
                }
                codes[i] = (expandBits21(q[0]) << 2) | (expandBits21(q[1]) << 1) | expandBits21(q[2]);
            }
        }
        nodes.push_back(Node());
        struct Job { uint32_t node, first, count; bool lbvh; };
        std::vector<Job> stack;
        if (lbvhAbove)
            std::sort(order_.begin(), order_.end(), [&](uint32_t a, uint32_t b) { return codes[a] < codes[b] || (codes[a] == codes[b] && a < b); });
        stack.push_back({0, 0, (uint32_t)n, lbvhAbove && n > lbvhAbove});
        while (!stack.empty()) {
            Job j = stack.back();
            stack.pop_back();
            if (lbvhAbove) j.lbvh = j.count > lbvhAbove;
            if (lbvhBelow && !j.lbvh && j.count <= lbvhBelow) {  // from here down: Morton order, highest-differing-bit splits
                std::sort(order_.begin() + j.first, order_.begin() + j.first + j.count,
                          [&](uint32_t a, uint32_t b) { return codes[a] < codes[b] || (codes[a] == codes[b] && a < b); });
                j.lbvh = true;
            }
            float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
            float clo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, chi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
            for (uint32_t i = j.first; i < j.first + j.count; ++i) {
                uint32_t t = order_[i];
                for (int k = 0; k < 3; ++k) {
                    lo[k] = std::min(lo[k], blo[3 * t + k]);
                    hi[k] = std::max(hi[k], bhi[3 * t + k]);
                    clo[k] = std::min(clo[k], cen[3 * t + k]);
                    chi[k] = std::max(chi[k], cen[3 * t + k]);
                }
            }
            Node nd;
            for (int k = 0; k < 3; ++k) {
                // pad one ulp-ish so that the double-precision triangle test never loses a hit to the box test
                float pad = 4.0f * FLT_EPSILON * std::max(std::fabs(lo[k]), std::fabs(hi[k])) + FLT_MIN;
                nd.lo[k] = lo[k] - pad;
                nd.hi[k] = hi[k] + pad;
            }
            if (j.count <= 4) {
                nd.left = j.first;
                nd.count = j.count;
                nodes[j.node] = nd;
                continue;
            }
            if (j.lbvh) {
                uint64_t a = codes[order_[j.first]], b = codes[order_[j.first + j.count - 1]];
                uint32_t mid = j.first + j.count / 2;
                if (a != b) {
                    int bit = 63 - __builtin_clzll(a ^ b);
                    uint64_t prefix = (b >> bit) << bit;
                    auto it = std::lower_bound(order_.begin() + j.first, order_.begin() + j.first + j.count, prefix,
                                               [&](uint32_t t, uint64_t key) { return codes[t] < key; });
                    mid = (uint32_t)(it - order_.begin());
                }
                nd.count = 0;
                nd.left = (uint32_t)nodes.size();
                nodes[j.node] = nd;
                nodes.push_back(Node());
                nodes.push_back(Node());
                stack.push_back({nd.left, j.first, mid - j.first, true});
                stack.push_back({nd.left + 1, mid, j.first + j.count - mid, true});
                continue;
            }
            // binned SAH, 16 bins per axis
            const int B = 16;
            int bestAxis = -1, bestSplit = 0;
            float bestCost = FLT_MAX;
            for (int ax = 0; ax < 3; ++ax) {
                float ext = chi[ax] - clo[ax];
                if (!(ext > 0.0f)) continue;
                float scale = B / ext;
                int cnt[B] = {0};
                float bl[B][3], bh[B][3];
                for (int b = 0; b < B; ++b)
                    for (int k = 0; k < 3; ++k) { bl[b][k] = FLT_MAX; bh[b][k] = -FLT_MAX; }
                for (uint32_t i = j.first; i < j.first + j.count; ++i) {
                    uint32_t t = order_[i];
                    int b = std::min(B - 1, (int)((cen[3 * t + ax] - clo[ax]) * scale));
                    cnt[b]++;
                    for (int k = 0; k < 3; ++k) {
                        bl[b][k] = std::min(bl[b][k], blo[3 * t + k]);
                        bh[b][k] = std::max(bh[b][k], bhi[3 * t + k]);
                    }
                }
                float rarea[B];
                int rcnt[B];
                float al[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, ah[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
                int c = 0;
                for (int b = B - 1; b > 0; --b) {
                    c += cnt[b];
                    for (int k = 0; k < 3; ++k) { al[k] = std::min(al[k], bl[b][k]); ah[k] = std::max(ah[k], bh[b][k]); }
                    rarea[b] = c ? half_area(al, ah) : 0.0f;
                    rcnt[b] = c;
                }
                for (int k = 0; k < 3; ++k) { al[k] = FLT_MAX; ah[k] = -FLT_MAX; }
                c = 0;
                for (int b = 0; b < B - 1; ++b) {
                    c += cnt[b];
                    for (int k = 0; k < 3; ++k) { al[k] = std::min(al[k], bl[b][k]); ah[k] = std::max(ah[k], bh[b][k]); }
                    if (c == 0 || rcnt[b + 1] == 0) continue;
                    float cost = half_area(al, ah) * c + rarea[b + 1] * rcnt[b + 1];
                    if (cost < bestCost) { bestCost = cost; bestAxis = ax; bestSplit = b; }
                }
            }
            uint32_t mid;
            if (bestAxis < 0) {
                mid = j.first + j.count / 2;  // all centroids coincide: median split
            } else {
                float scale = B / (chi[bestAxis] - clo[bestAxis]);
                float c0 = clo[bestAxis];
                int ax = bestAxis, sp = bestSplit;
                auto it = std::partition(order_.begin() + j.first, order_.begin() + j.first + j.count, [&](uint32_t t) {
                    int b = std::min(B - 1, (int)((cen[3 * t + ax] - c0) * scale));
                    return b <= sp;
                });
                mid = (uint32_t)(it - order_.begin());
                if (mid == j.first || mid == j.first + j.count) mid = j.first + j.count / 2;
            }
            nd.count = 0;
            nd.left = (uint32_t)nodes.size();
            nodes[j.node] = nd;
            nodes.push_back(Node());
            nodes.push_back(Node());
            stack.push_back({nd.left, j.first, mid - j.first, false});
            stack.push_back({nd.left + 1, mid, j.first + j.count - mid, false});
        }
        std::vector<Tri> sorted(n);
        for (size_t i = 0; i < n; ++i) sorted[i] = tris[order_[i]];
        tris.swap(sorted);
        order_.clear();
        order_.shrink_to_fit();
    }

    // Closest hit (anyHit=false) or any hit (anyHit=true) with t in [tnear,tfar].
    template <typename Real>
    bool trace(const float of[3], const float df[3], float tnear, float tfar, bool anyHit, RayHit& out) const {
        out.tri = 0xFFFFFFFFu;
        out.t = tfar;
        out.u = out.v = 0;
        if (nodes.empty()) return false;
        // non-finite rays are misses (Embree behaviour); without this they would pass every slab test
        for (int k = 0; k < 3; ++k)
            if (!std::isfinite(of[k]) || !std::isfinite(df[k])) return false;
        if (!(tnear <= tfar)) return false;
        Real o[3] = {(Real)of[0], (Real)of[1], (Real)of[2]}, d[3] = {(Real)df[0], (Real)df[1], (Real)df[2]};
        Real inv[3];
        for (int k = 0; k < 3; ++k) inv[k] = (Real)1 / d[k];  // +-inf on zero components is fine for the slab test
        Real best = (Real)tfar, tn = (Real)tnear;
        uint32_t stack[128];
        int sp = 0;
        stack[sp++] = 0;
        while (sp) {
            const Node& nd = nodes[stack[--sp]];
            if (!slab(nd, o, inv, tn, best)) continue;
            if (nd.count) {
                for (uint32_t i = nd.left; i < nd.left + nd.count; ++i) {
                    Real t, u, v;
                    if (intersect_tri<Real>(tris[i], o, d, tn, best, t, u, v)) {
                        // strict '<' keeps the first-found on exact ties unless this one is closer
                        if (out.tri == 0xFFFFFFFFu || t < best || (t == best && i < out.tri)) {
                            best = t;
                            out.t = (double)t;
                            out.u = (double)u;
                            out.v = (double)v;
                            out.tri = i;
                            if (anyHit) return true;
                        }
                    }
                }
            } else {
                stack[sp++] = nd.left;
                stack[sp++] = nd.left + 1;
            }
        }
        return out.tri != 0xFFFFFFFFu;
    }

    // Exhaustive double-precision reference used to pin the BVH itself.
    bool brute(const float of[3], const float df[3], float tnear, float tfar, RayHit& out) const {
        double o[3] = {of[0], of[1], of[2]}, d[3] = {df[0], df[1], df[2]};
        out.tri = 0xFFFFFFFFu;
        out.t = tfar;
        out.u = out.v = 0;
        double best = tfar;
        for (size_t i = 0; i < tris.size(); ++i) {
            double t, u, v;
            if (intersect_tri<double>(tris[i], o, d, (double)tnear, best, t, u, v)) {
                if (out.tri == 0xFFFFFFFFu || t < best) {
                    best = t;
                    out.t = t; out.u = u; out.v = v;
                    out.tri = (uint32_t)i;
                }
            }
        }
        return out.tri != 0xFFFFFFFFu;
    }

    // ---- tree-quality tooling (scripts/tree_quality.py): the same scene under different builders ---------------------------------
    // LBVH: 63-bit Morton codes of the triangle-box centres (21 bits per axis, x most significant, as the product's device build),
    // split top-down at the highest differing bit; cubic = all axes quantised on the grid of the longest one.  Leaves <= 4 triangles.
    // sweep = true keeps the Morton order of the triangles but places every split where the surface area heuristic is lowest along
    // that order (a sweep over the node's range) instead of at the highest differing bit.
    void buildLbvh(bool cubic, bool sweep = false) {
        const size_t n = tris.size();
        nodes.clear();
        if (n == 0) return;
        std::vector<float> blo(3 * n), bhi(3 * n);
        float slo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, shi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (size_t i = 0; i < n; ++i)
            for (int k = 0; k < 3; ++k) {
                float a = tris[i].v0[k], b = tris[i].v1[k], c = tris[i].v2[k];
                blo[3 * i + k] = std::min(a, std::min(b, c));
                bhi[3 * i + k] = std::max(a, std::max(b, c));
                slo[k] = std::min(slo[k], blo[3 * i + k]);
                shi[k] = std::max(shi[k], bhi[3 * i + k]);
            }
        float ext[3] = {shi[0] - slo[0], shi[1] - slo[1], shi[2] - slo[2]};
        if (cubic) ext[0] = ext[1] = ext[2] = std::max(ext[0], std::max(ext[1], ext[2]));
        auto expand = [](uint64_t x) { return expandBits21(x); };
        std::vector<std::pair<uint64_t, uint32_t>> keyed(n);
        for (size_t i = 0; i < n; ++i) {
            uint64_t q[3];
            for (int k = 0; k < 3; ++k) {
                float c = 0.5f * (blo[3 * i + k] + bhi[3 * i + k]);
                float nrm = ext[k] > 0.f ? (c - slo[k]) / ext[k] : 0.f;
                q[k] = (uint64_t)std::min(std::max(nrm * 2097152.0f, 0.0f), 2097151.0f);
            }
            keyed[i] = {(expand(q[0]) << 2) | (expand(q[1]) << 1) | expand(q[2]), (uint32_t)i};
        }
        std::sort(keyed.begin(), keyed.end());
        nodes.reserve(2 * n);
        nodes.push_back(Node());
        struct Job { uint32_t node, first, count; };
        std::vector<Job> stack;
        stack.push_back({0, 0, (uint32_t)n});
        while (!stack.empty()) {
            Job j = stack.back();
            stack.pop_back();
            Node nd;
            for (int k = 0; k < 3; ++k) { nd.lo[k] = FLT_MAX; nd.hi[k] = -FLT_MAX; }
            for (uint32_t i = j.first; i < j.first + j.count; ++i) {
                uint32_t t = keyed[i].second;
                for (int k = 0; k < 3; ++k) {
                    nd.lo[k] = std::min(nd.lo[k], blo[3 * t + k]);
                    nd.hi[k] = std::max(nd.hi[k], bhi[3 * t + k]);
                }
            }
            if (j.count <= 4) {
                nd.left = j.first;
                nd.count = j.count;
                nodes[j.node] = nd;
                continue;
            }
            uint64_t a = keyed[j.first].first, b = keyed[j.first + j.count - 1].first;
            uint32_t mid = j.first + j.count / 2;
            if (sweep) {
                std::vector<float>& suffix = sweepScratch_;
                suffix.resize(j.count + 1);
                float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
                suffix[j.count] = 0.f;
                for (uint32_t i = j.count; i-- > 0;) {
                    uint32_t t = keyed[j.first + i].second;
                    for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], blo[3 * t + k]); hi[k] = std::max(hi[k], bhi[3 * t + k]); }
                    suffix[i] = half_area(lo, hi);
                }
                for (int k = 0; k < 3; ++k) { lo[k] = FLT_MAX; hi[k] = -FLT_MAX; }
                float bestCost = FLT_MAX;
                for (uint32_t i = 0; i + 1 < j.count; ++i) {
                    uint32_t t = keyed[j.first + i].second;
                    for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], blo[3 * t + k]); hi[k] = std::max(hi[k], bhi[3 * t + k]); }
                    float cost = half_area(lo, hi) * (float)(i + 1) + suffix[i + 1] * (float)(j.count - i - 1);
                    if (cost < bestCost) { bestCost = cost; mid = j.first + i + 1; }
                }
            } else if (a != b) {
                int bit = 63 - __builtin_clzll(a ^ b);
                uint64_t prefix = (b >> bit) << bit;  // first code with that bit set (codes below share the higher bits)
                auto it = std::lower_bound(keyed.begin() + j.first, keyed.begin() + j.first + j.count, std::make_pair(prefix, (uint32_t)0));
                mid = (uint32_t)(it - keyed.begin());
            }
            nd.count = 0;
            nd.left = (uint32_t)nodes.size();
            nodes[j.node] = nd;
            nodes.push_back(Node());
            nodes.push_back(Node());
            stack.push_back({nd.left, j.first, mid - j.first});
            stack.push_back({nd.left + 1, mid, j.first + j.count - mid});
        }
        std::vector<Tri> sorted(n);
        for (size_t i = 0; i < n; ++i) sorted[i] = tris[keyed[i].second];
        tris.swap(sorted);
    }

    // PLOC (parallel locally-ordered clustering, Meister & Bittner 2018) written serially: clusters start as the triangles in cubic Morton
    // order; every round each cluster looks `radius` places to either side for the partner with the smallest merged surface area, mutual
    // choices merge, the array is compacted.  Subtrees of at most 4 triangles become leaves.
    void buildPloc(int radius) {
        const size_t n = tris.size();
        nodes.clear();
        if (n == 0) return;
        struct T { float lo[3], hi[3]; uint32_t left, right, count; };  // leaf: left = triangle id, right = 0xFFFFFFFF
        std::vector<T> t;
        t.reserve(2 * n);
        float slo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, shi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (size_t i = 0; i < n; ++i) {
            T leaf;
            for (int k = 0; k < 3; ++k) {
                float a = tris[i].v0[k], b = tris[i].v1[k], c = tris[i].v2[k];
                leaf.lo[k] = std::min(a, std::min(b, c));
                leaf.hi[k] = std::max(a, std::max(b, c));
                slo[k] = std::min(slo[k], leaf.lo[k]);
                shi[k] = std::max(shi[k], leaf.hi[k]);
            }
            leaf.left = (uint32_t)i;
            leaf.right = 0xFFFFFFFFu;
            leaf.count = 1;
            t.push_back(leaf);
        }
        const float longest = std::max(shi[0] - slo[0], std::max(shi[1] - slo[1], shi[2] - slo[2]));
        std::vector<std::pair<uint64_t, uint32_t>> keyed(n);
        for (size_t i = 0; i < n; ++i) {
            uint64_t q[3];
            for (int k = 0; k < 3; ++k) {
                float c = 0.5f * (t[i].lo[k] + t[i].hi[k]);
                float nrm = longest > 0.f ? (c - slo[k]) / longest : 0.f;
                q[k] = (uint64_t)std::min(std::max(nrm * 2097152.0f, 0.0f), 2097151.0f);
            }
            keyed[i] = {(expandBits21(q[0]) << 2) | (expandBits21(q[1]) << 1) | expandBits21(q[2]), (uint32_t)i};
        }
        std::sort(keyed.begin(), keyed.end());
        std::vector<uint32_t> cur(n), next, nearest;
        for (size_t i = 0; i < n; ++i) cur[i] = keyed[i].second;
        auto mergedArea = [&](uint32_t a, uint32_t b) {
            float lo[3], hi[3];
            for (int k = 0; k < 3; ++k) { lo[k] = std::min(t[a].lo[k], t[b].lo[k]); hi[k] = std::max(t[a].hi[k], t[b].hi[k]); }
            return half_area(lo, hi);
        };
        while (cur.size() > 1) {
            const int m = (int)cur.size();
            nearest.assign(m, 0);
            for (int i = 0; i < m; ++i) {
                float best = FLT_MAX;
                int bj = i == 0 ? 1 : i - 1;
                for (int j = std::max(0, i - radius); j <= std::min(m - 1, i + radius); ++j) {
                    if (j == i) continue;
                    float a = mergedArea(cur[i], cur[j]);
                    if (a < best) { best = a; bj = j; }
                }
                nearest[i] = (uint32_t)bj;
            }
            next.clear();
            for (int i = 0; i < m; ++i) {
                int j = (int)nearest[i];
                if ((int)nearest[j] == i) {
                    if (i < j) {
                        T nd;
                        for (int k = 0; k < 3; ++k) { nd.lo[k] = std::min(t[cur[i]].lo[k], t[cur[j]].lo[k]); nd.hi[k] = std::max(t[cur[i]].hi[k], t[cur[j]].hi[k]); }
                        nd.left = cur[i];
                        nd.right = cur[j];
                        nd.count = t[cur[i]].count + t[cur[j]].count;
                        t.push_back(nd);
                        next.push_back((uint32_t)t.size() - 1);
                    }
                } else {
                    next.push_back(cur[i]);
                }
            }
            cur.swap(next);
        }
        // lay the tree out in this class's format (inner: children adjacent; leaf: a contiguous triangle range)
        std::vector<Tri> sorted;
        sorted.reserve(n);
        nodes.reserve(2 * n);
        nodes.push_back(Node());
        struct Job { uint32_t node, src; };
        std::vector<Job> stack;
        stack.push_back({0, cur[0]});
        std::vector<uint32_t> walk;
        while (!stack.empty()) {
            Job j = stack.back();
            stack.pop_back();
            const T& src = t[j.src];
            Node nd;
            for (int k = 0; k < 3; ++k) { nd.lo[k] = src.lo[k]; nd.hi[k] = src.hi[k]; }
            if (src.count <= 4) {
                nd.left = (uint32_t)sorted.size();
                nd.count = src.count;
                walk.clear();
                walk.push_back(j.src);
                while (!walk.empty()) {
                    uint32_t w = walk.back();
                    walk.pop_back();
                    if (t[w].right == 0xFFFFFFFFu) sorted.push_back(tris[t[w].left]);
                    else { walk.push_back(t[w].right); walk.push_back(t[w].left); }
                }
                nodes[j.node] = nd;
                continue;
            }
            nd.count = 0;
            nd.left = (uint32_t)nodes.size();
            nodes[j.node] = nd;
            nodes.push_back(Node());
            nodes.push_back(Node());
            stack.push_back({nd.left, src.left});
            stack.push_back({nd.left + 1, src.right});
        }
        tris.swap(sorted);
    }

    // Surface-area-heuristic cost relative to the root: sum of inner-node areas, sum of leaf area x triangle count.
    void sahCost(double& inner, double& leaf) const {
        inner = leaf = 0.0;
        if (nodes.empty()) return;
        const double root = half_area(nodes[0].lo, nodes[0].hi);
        for (const Node& nd : nodes) {
            double a = half_area(nd.lo, nd.hi) / root;
            if (nd.count) leaf += a * nd.count;
            else inner += a;
        }
    }

    // Closest hit with the near child visited first; counts box tests and triangle tests (what a tree costs a ray).
    void traceCount(const float of[3], const float df[3], float tfar, uint64_t& boxTests, uint64_t& triTests) const {
        if (nodes.empty()) return;
        double o[3] = {of[0], of[1], of[2]}, d[3] = {df[0], df[1], df[2]}, inv[3];
        for (int k = 0; k < 3; ++k) inv[k] = 1.0 / d[k];
        double best = tfar;
        uint32_t stack[128];
        int sp = 0;
        stack[sp++] = 0;
        ++boxTests;
        if (!slab(nodes[0], o, inv, 0.0, best)) return;
        while (sp) {
            const Node& nd = nodes[stack[--sp]];
            if (nd.count) {
                for (uint32_t i = nd.left; i < nd.left + nd.count; ++i) {
                    double t, u, v;
                    ++triTests;
                    if (intersect_tri<double>(tris[i], o, d, 0.0, best, t, u, v)) best = t;
                }
                continue;
            }
            double tn[2];
            bool hit[2];
            for (int c = 0; c < 2; ++c) {
                const Node& ch = nodes[nd.left + c];
                ++boxTests;
                double t0 = 0.0, t1 = best;
                for (int k = 0; k < 3; ++k) {
                    double x = ((double)ch.lo[k] - o[k]) * inv[k], y = ((double)ch.hi[k] - o[k]) * inv[k];
                    if (x != x || y != y) continue;
                    if (x > y) std::swap(x, y);
                    t0 = std::max(t0, x);
                    t1 = std::min(t1, y);
                }
                hit[c] = t0 <= t1;
                tn[c] = t0;
            }
            if (hit[0] && hit[1]) {
                int nearC = tn[0] <= tn[1] ? 0 : 1;
                stack[sp++] = nd.left + (1 - nearC);
                stack[sp++] = nd.left + nearC;
            } else if (hit[0]) stack[sp++] = nd.left;
            else if (hit[1]) stack[sp++] = nd.left + 1;
        }
    }

private:
    static uint64_t expandBits21(uint64_t x) {
        x &= 0x1fffffull;
        x = (x | x << 32) & 0x1f00000000ffffull;
        x = (x | x << 16) & 0x1f0000ff0000ffull;
        x = (x | x << 8) & 0x100f00f00f00f00full;
        x = (x | x << 4) & 0x10c30c30c30c30c3ull;
        x = (x | x << 2) & 0x1249249249249249ull;
        return x;
    }
    std::vector<uint32_t> order_;
    std::vector<float> sweepScratch_;
    static float half_area(const float lo[3], const float hi[3]) {
        float x = hi[0] - lo[0], y = hi[1] - lo[1], z = hi[2] - lo[2];
        return x * y + y * z + z * x;
    }
    template <typename Real>
    static bool slab(const Node& nd, const Real o[3], const Real inv[3], Real tn, Real tf) {
        Real t0 = tn, t1 = tf;
        for (int k = 0; k < 3; ++k) {
            Real a = ((Real)nd.lo[k] - o[k]) * inv[k], b = ((Real)nd.hi[k] - o[k]) * inv[k];
            if (a != a || b != b) continue;  // 0*inf: ray lies in the slab plane, no constraint
            if (a > b) std::swap(a, b);
            // widen a hair so rounding in the slab test never rejects a box whose triangle is hit
            t0 = std::max(t0, a - std::fabs(a) * (Real)1e-6);
            t1 = std::min(t1, b + std::fabs(b) * (Real)1e-6);
        }
        return t0 <= t1;
    }
};

}  // namespace orc
