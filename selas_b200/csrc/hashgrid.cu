// hashgrid.cu -- VCM light-vertex hash grid on the GPU: build (count / scan / scatter) and the 2x2x2-cell
// range query.  Replaces BuildHashGrid and SearchHashGrid
// (Source/Applications/Selas/Source/VCMHashGrid.cpp:48-97 and :107-155).
//
// Build pipeline, all device-side, no host synchronisation:
//   1. boundsKernel      AABB of the points (unpadded, as the reference leaves it -- VCMHashGrid.cpp:66-72)
//   2. countKernel       cell = hash(floor((p - min) / (2r)));  atomicAdd on the per-cell counter; the cell id is kept
//   3. scan (3 kernels)  inclusive prefix sum -> cellRangeEnds (exclusive end of every cell, the reference's final
//                        array) and a cursor array holding every cell's start
//   4. scatterKernel     slot = atomicAdd(cursor[cell]); cellIndices[slot] = point
//   5. orderCellsKernel  the serial reference inserts points in ascending index order, so each cell's slice is
//                        sorted ascending: insertion sort per cell (cells hold ~1 point on average), cells with
//                        more than 16 points go to a list that bigCellsKernel sorts with a per-CTA bitonic network
//   6. gatherKernel      cellPoints[slot] = {position, index}: queries then stream 16-byte candidates that are
//                        contiguous per cell instead of chasing cellIndices into the 108-byte VCMVertex array
//
// Hash: the reference's `uint` is 64-bit (SystemLib/BasicTypes.h:47): int32 coordinates are sign-extended, the
// products, xor and modulo are 64-bit (VCMHashGrid.cpp:15-22).
//
// Compiled with -fmad=false so that floor(inv * (p - min)) and dx*dx + dy*dy + dz*dz round like the CPU code
// (the d2 <= r2 accept test is otherwise not reproducible at the boundary).
//
// Algorithmic bytes: build 36 B/point + 8 B/cell; query 64 B + 16 B/candidate.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>

#include "hashgrid.cuh"
#include "internal.h"
#include "vec.cuh"

struct SelasHashGrid {
    SelasContext* ctx = nullptr;
    size_t n = 0, capacityN = 0;
    uint32_t cellCount = 0, capacityCells = 0;
    float radius = 0.f, radiusSq = 0.f, inverseCellSize = 0.f;
    uint32_t* d_bounds = nullptr;       // 6 order-preserving uints: min xyz, max xyz
    uint32_t* d_cellEnds = nullptr;     // cellRangeEnds[cellCount]
    uint32_t* d_cursor = nullptr;       // cellCount
    uint32_t* d_cellIndices = nullptr;  // n
    uint32_t* d_pointCell = nullptr;    // n
    float4* d_cellPoints = nullptr;     // n
    uint32_t* d_blockSums = nullptr;    // scan partials
    uint32_t* d_bigCells = nullptr;     // cellCount (worst case) entries + counter at [-1]
    uint32_t scanBlocks = 0;
};

namespace selas {

namespace {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;  // per thread
constexpr int kScanTile = kScanThreads * kScanItems;

struct GridParams {
    float minx, miny, minz, maxx, maxy, maxz;
    float inv, r2;
    uint32_t cellCount;
};

__global__ void boundsKernel(const float* __restrict__ pos, size_t n, uint32_t* __restrict__ bounds) {
    float lo[3] = {SELAS_FLOAT_MAX, SELAS_FLOAT_MAX, SELAS_FLOAT_MAX}, hi[3] = {-SELAS_FLOAT_MAX, -SELAS_FLOAT_MAX, -SELAS_FLOAT_MAX};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        for (int k = 0; k < 3; ++k) {
            float p = __ldg(pos + 3 * i + k);
            lo[k] = minf_(lo[k], p);  // AxisAlignedBox IncludePosition uses Min/Max (MinMax.h)
            hi[k] = maxf_(hi[k], p);
        }
    }
    for (int k = 0; k < 3; ++k) {
        float a = lo[k], b = hi[k];
        for (int s = 16; s > 0; s >>= 1) {
            a = fminf(a, __shfl_xor_sync(0xffffffffu, a, s));
            b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, s));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(bounds + k, orderedFromFloat(a));
            atomicMax(bounds + 3 + k, orderedFromFloat(b));
        }
    }
}

__global__ void countKernel(const float* __restrict__ pos, size_t n, const uint32_t* __restrict__ bounds, float inv, uint32_t cellCount,
                            uint32_t* __restrict__ counts, uint32_t* __restrict__ pointCell) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float mx = floatFromOrdered(bounds[0]), my = floatFromOrdered(bounds[1]), mz = floatFromOrdered(bounds[2]);
    // CalculateCellIndex(point), VCMHashGrid.cpp:25-34
    int32_t x = (int32_t)floorf(inv * (__ldg(pos + 3 * i) - mx));
    int32_t y = (int32_t)floorf(inv * (__ldg(pos + 3 * i + 1) - my));
    int32_t z = (int32_t)floorf(inv * (__ldg(pos + 3 * i + 2) - mz));
    uint32_t c = cellHash(x, y, z, cellCount);
    pointCell[i] = c;
    atomicAdd(counts + c, 1u);
}

// ---- inclusive scan: tile sums -> scan of sums -> tile scan with offsets ------------------------------------------
__global__ void __launch_bounds__(kScanThreads) scanTileSums(const uint32_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ sums) {
    __shared__ uint32_t warpSums[kScanThreads / 32];
    uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
        if (base + k < n) s += in[base + k];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) warpSums[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < kScanThreads / 32; ++w) t += warpSums[w];
        sums[blockIdx.x] = t;
    }
}
// single CTA: exclusive scan of the tile sums in place (tile count is small: n / 2048)
__global__ void __launch_bounds__(1024) scanSums(uint32_t* __restrict__ sums, uint32_t count) {
    __shared__ uint32_t warpTotals[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < count; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < count ? sums[i] : 0u;
        uint32_t x = v;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= (uint32_t)o) x += y;
        }
        if ((threadIdx.x & 31) == 31) warpTotals[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = warpTotals[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
                if (threadIdx.x >= (uint32_t)o) w += y;
            }
            warpTotals[threadIdx.x] = w;
        }
        __syncthreads();
        uint32_t warpOffset = (threadIdx.x >> 5) ? warpTotals[(threadIdx.x >> 5) - 1] : 0u;
        uint32_t incl = x + warpOffset + carry;
        if (i < count) sums[i] = incl - v;  // exclusive
        __syncthreads();
        if (threadIdx.x == 1023) carry = incl;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(kScanThreads) scanTiles(const uint32_t* in, uint32_t n, const uint32_t* __restrict__ sums,
                                                          uint32_t* __restrict__ ends, uint32_t* starts /* may alias in */) {
    __shared__ uint32_t warpTotals[kScanThreads / 32];
    uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0u;
        s += v[k];
    }
    uint32_t x = s;
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) >= (uint32_t)o) x += y;
    }
    if ((threadIdx.x & 31) == 31) warpTotals[threadIdx.x >> 5] = x;
    __syncthreads();
    uint32_t off = sums[blockIdx.x];
    for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w) off += warpTotals[w];
    uint32_t run = off + x - s;  // exclusive prefix of this thread's first item
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) {
            starts[base + k] = run;
            run += v[k];
            ends[base + k] = run;
        }
    }
}

__global__ void scatterKernel(const uint32_t* __restrict__ pointCell, size_t n, uint32_t* __restrict__ cursor,
                              uint32_t* __restrict__ cellIndices) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t slot = atomicAdd(cursor + pointCell[i], 1u);
    cellIndices[slot] = (uint32_t)i;
}

// cursor[c] now equals ends[c]; start of cell c is ends[c-1]
__global__ void orderCellsKernel(const uint32_t* __restrict__ ends, uint32_t cellCount, uint32_t* __restrict__ cellIndices,
                                 uint32_t* __restrict__ bigCells, uint32_t* __restrict__ bigCount) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cellCount) return;
    uint32_t b = c ? ends[c - 1] : 0u, e = ends[c];
    uint32_t k = e - b;
    if (k < 2) return;
    if (k > 16) {
        bigCells[atomicAdd(bigCount, 1u)] = c;
        return;
    }
    uint32_t v[16];
    for (uint32_t i = 0; i < k; ++i) v[i] = cellIndices[b + i];
    for (uint32_t i = 1; i < k; ++i) {
        uint32_t x = v[i];
        uint32_t j = i;
        while (j > 0 && v[j - 1] > x) {
            v[j] = v[j - 1];
            --j;
        }
        v[j] = x;
    }
    for (uint32_t i = 0; i < k; ++i) cellIndices[b + i] = v[i];
}

// One CTA per crowded cell: ascending bitonic network (mirror step + half-cleaners), elements beyond k are virtual
// +inf that never move, so any k works in place in global memory.
__global__ void __launch_bounds__(256) bigCellsKernel(const uint32_t* __restrict__ ends, const uint32_t* __restrict__ bigCells,
                                                      const uint32_t* __restrict__ bigCount, uint32_t* __restrict__ cellIndices) {
    const uint32_t count = *bigCount;
    for (uint32_t bi = blockIdx.x; bi < count; bi += gridDim.x) {
        uint32_t c = bigCells[bi];
        uint32_t b = c ? ends[c - 1] : 0u, e = ends[c];
        uint32_t k = e - b;
        uint32_t* a = cellIndices + b;
        uint32_t P = 1;
        while (P < k) P <<= 1;
        const uint32_t pairs = P >> 1;
        for (uint32_t size = 2; size <= P; size <<= 1) {
            uint32_t half = size >> 1;
            for (uint32_t t = threadIdx.x; t < pairs; t += blockDim.x) {
                uint32_t blk = t / half, off = t - blk * half;
                uint32_t i = blk * size + off, j = blk * size + size - 1 - off;
                if (j < k) {
                    uint32_t x = a[i], y = a[j];
                    if (x > y) { a[i] = y; a[j] = x; }
                }
            }
            __syncthreads();
            for (uint32_t stride = half >> 1; stride > 0; stride >>= 1) {
                for (uint32_t t = threadIdx.x; t < pairs; t += blockDim.x) {
                    uint32_t i = (t / stride) * 2 * stride + (t % stride), j = i + stride;
                    if (j < k) {
                        uint32_t x = a[i], y = a[j];
                        if (x > y) { a[i] = y; a[j] = x; }
                    }
                }
                __syncthreads();
            }
        }
    }
}

__global__ void gatherKernel(const float* __restrict__ pos, const uint32_t* __restrict__ cellIndices, size_t n, float4* __restrict__ cellPoints) {
    size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    uint32_t i = cellIndices[s];
    cellPoints[s] = make_float4(__ldg(pos + 3 * (size_t)i), __ldg(pos + 3 * (size_t)i + 1), __ldg(pos + 3 * (size_t)i + 2), __uint_as_float(i));
}

// SearchHashGrid, VCMHashGrid.cpp:107-155: one thread per query.
__global__ void __launch_bounds__(128) queryKernel(const uint32_t* __restrict__ bounds, float inv, float r2, uint32_t cellCount,
                                                   const uint32_t* __restrict__ ends, const float4* __restrict__ cellPoints, size_t m,
                                                   const float* __restrict__ q, uint32_t* __restrict__ counts, uint32_t* __restrict__ indices,
                                                   uint32_t maxPerQuery) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    float px = __ldg(q + 3 * i), py = __ldg(q + 3 * i + 1), pz = __ldg(q + 3 * i + 2);
    float dminx = px - floatFromOrdered(bounds[0]), dminy = py - floatFromOrdered(bounds[1]), dminz = pz - floatFromOrdered(bounds[2]);
    float dmaxx = floatFromOrdered(bounds[3]) - px, dmaxy = floatFromOrdered(bounds[4]) - py, dmaxz = floatFromOrdered(bounds[5]) - pz;
    uint32_t found = 0;
    // written as the reference's early-outs: a NaN coordinate passes them, as it does on the CPU
    bool outside = (dminx < 0.0f || dminy < 0.0f || dminz < 0.0f) || (dmaxx < 0.0f || dmaxy < 0.0f || dmaxz < 0.0f);
    if (!outside) {
        float cx = inv * dminx, cy = inv * dminy, cz = inv * dminz;
        float fx = floorf(cx), fy = floorf(cy), fz = floorf(cz);
        int32_t bx = (int32_t)fx, by = (int32_t)fy, bz = (int32_t)fz;
        if (cx - fx < 0.5f) --bx;
        if (cy - fy < 0.5f) --by;
        if (cz - fz < 0.5f) --bz;
        uint32_t* out = indices ? indices + i * (size_t)maxPerQuery : nullptr;
#pragma unroll 1
        for (int c8 = 0; c8 < 8; ++c8) {
            uint32_t cell = cellHash(bx + (c8 & 1), by + ((c8 >> 1) & 1), bz + (c8 >> 2), cellCount);
            uint32_t b = cell ? __ldg(ends + cell - 1) : 0u, e = __ldg(ends + cell);
            for (uint32_t s = b; s < e; ++s) {
                float4 p = __ldg(cellPoints + s);
                float dx = px - p.x, dy = py - p.y, dz = pz - p.z;
                float d2 = (dx * dx) + (dy * dy) + (dz * dz);
                if (d2 <= r2) {
                    if (out && found < maxPerQuery) out[found] = __float_as_uint(p.w);
                    ++found;
                }
            }
        }
    }
    counts[i] = found;
}

// SearchHashGrid, warp-cooperative form.  Every lane first prepares one query (AABB test, the eight cell ranges); the
// warp then walks the 32 queries one after the other, all lanes testing 32 consecutive candidates of the query's
// flattened candidate list (cell 0 .. cell 7, slots in order) per step.  Candidate loads are coalesced 16-byte reads,
// all lanes stay busy even when the cell occupancy is heavy-tailed, and the ballot prefix keeps the accepted indices in
// the reference's visiting order.
constexpr int kQueryWarps = 4;
__global__ void __launch_bounds__(kQueryWarps * 32) queryWarpKernel(const uint32_t* __restrict__ bounds, float inv, float r2, uint32_t cellCount,
                                                                    const uint32_t* __restrict__ ends, const float4* __restrict__ cellPoints,
                                                                    size_t m, const float* __restrict__ q, uint32_t* __restrict__ counts,
                                                                    uint32_t* __restrict__ indices, uint32_t maxPerQuery) {
    __shared__ float sPos[kQueryWarps][32][3];
    __shared__ uint32_t sBegin[kQueryWarps][32][8];
    __shared__ uint32_t sCum[kQueryWarps][32][8];  // inclusive prefix of the eight range lengths
    const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
    const uint32_t ltMask = (1u << lane) - 1u;
    const size_t warpCount = (size_t)gridDim.x * kQueryWarps;
    const float minx = floatFromOrdered(bounds[0]), miny = floatFromOrdered(bounds[1]), minz = floatFromOrdered(bounds[2]);
    const float maxx = floatFromOrdered(bounds[3]), maxy = floatFromOrdered(bounds[4]), maxz = floatFromOrdered(bounds[5]);
    for (size_t base = ((size_t)blockIdx.x * kQueryWarps + wib) * 32; base < m; base += warpCount * 32) {
        // ---- phase 1: one query per lane ----
        const size_t qi = base + lane;
        uint32_t total = 0;
        if (qi < m) {
            float px = __ldg(q + 3 * qi), py = __ldg(q + 3 * qi + 1), pz = __ldg(q + 3 * qi + 2);
            sPos[wib][lane][0] = px; sPos[wib][lane][1] = py; sPos[wib][lane][2] = pz;
            float dminx = px - minx, dminy = py - miny, dminz = pz - minz;
            float dmaxx = maxx - px, dmaxy = maxy - py, dmaxz = maxz - pz;
            bool outside = (dminx < 0.0f || dminy < 0.0f || dminz < 0.0f) || (dmaxx < 0.0f || dmaxy < 0.0f || dmaxz < 0.0f);
            if (!outside) {
                float cx = inv * dminx, cy = inv * dminy, cz = inv * dminz;
                float fx = floorf(cx), fy = floorf(cy), fz = floorf(cz);
                int32_t bx = (int32_t)fx, by = (int32_t)fy, bz = (int32_t)fz;
                if (cx - fx < 0.5f) --bx;
                if (cy - fy < 0.5f) --by;
                if (cz - fz < 0.5f) --bz;
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) {
                    uint32_t cell = cellHash(bx + (c8 & 1), by + ((c8 >> 1) & 1), bz + (c8 >> 2), cellCount);
                    uint32_t b = cell ? __ldg(ends + cell - 1) : 0u, e = __ldg(ends + cell);
                    total += e - b;
                    sBegin[wib][lane][c8] = b;
                    sCum[wib][lane][c8] = total;
                }
            }
        }
        __syncwarp();
        // ---- phase 2: the warp serves its 32 queries in turn ----
        uint32_t pending = __ballot_sync(0xffffffffu, total != 0u);
        uint32_t myCount = 0;  // lane L ends up holding the count of query L
        while (pending) {
            const uint32_t L = __ffs(pending) - 1u;
            pending &= pending - 1u;
            const float px = sPos[wib][L][0], py = sPos[wib][L][1], pz = sPos[wib][L][2];
            uint32_t cum[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) cum[k] = sCum[wib][L][k];
            const uint32_t T = cum[7];
            uint32_t* out = indices ? indices + (base + L) * (size_t)maxPerQuery : nullptr;
            uint32_t found = 0;
            for (uint32_t i0 = 0; i0 < T; i0 += 32) {
                const uint32_t i = i0 + lane;
                bool accept = false;
                uint32_t index = 0;
                if (i < T) {
                    uint32_t c = 0;
#pragma unroll
                    for (int k = 0; k < 7; ++k) c += (i >= cum[k]) ? 1u : 0u;
                    const uint32_t before = c ? sCum[wib][L][c - 1] : 0u;
                    const float4 p = __ldg(cellPoints + sBegin[wib][L][c] + (i - before));
                    float dx = px - p.x, dy = py - p.y, dz = pz - p.z;
                    float d2 = (dx * dx) + (dy * dy) + (dz * dz);
                    accept = d2 <= r2;
                    index = __float_as_uint(p.w);
                }
                const uint32_t mask = __ballot_sync(0xffffffffu, accept);
                if (accept && out) {
                    uint32_t pos = found + __popc(mask & ltMask);
                    if (pos < maxPerQuery) out[pos] = index;
                }
                found += __popc(mask);
            }
            if (lane == L) myCount = found;
        }
        if (qi < m) counts[qi] = myCount;
        __syncwarp();
    }
}

inline uint32_t divUp(size_t a, uint32_t b) { return (uint32_t)((a + b - 1) / b); }

void releaseBuffers(SelasHashGrid* g) {
    cudaFree(g->d_bounds);
    cudaFree(g->d_cellEnds);
    cudaFree(g->d_cursor);
    cudaFree(g->d_cellIndices);
    cudaFree(g->d_pointCell);
    cudaFree(g->d_cellPoints);
    cudaFree(g->d_blockSums);
    cudaFree(g->d_bigCells);
    g->d_bounds = g->d_cellEnds = g->d_cursor = g->d_cellIndices = g->d_pointCell = g->d_blockSums = g->d_bigCells = nullptr;
    g->d_cellPoints = nullptr;
    g->capacityN = 0;
    g->capacityCells = 0;
}

int reserve(SelasHashGrid* g, size_t n, uint32_t cellCount) {
    SelasContext* ctx = g->ctx;
    if (g->capacityN >= n && g->capacityCells >= cellCount && g->d_bounds) return SELAS_OK;
    releaseBuffers(g);
    size_t nn = n ? n : 1;
    uint32_t tiles = divUp(cellCount, kScanTile);
    SELAS_CUDA_CHECK(ctx, cudaMalloc((void**)&g->d_bounds, 8 * sizeof(uint32_t)));
    SELAS_CUDA_CHECK(ctx, cudaMalloc((void**)&g->d_cellEnds, sizeof(uint32_t) * (size_t)cellCount));
    SELAS_CUDA_CHECK(ctx, cudaMalloc((void**)&g->d_cursor, sizeof(uint32_t) * (size_t)cellCount));
    SELAS_CUDA_CHECK(ctx, cudaMalloc((void**)&g->d_cellIndices, sizeof(uint32_t) * nn));
    SELAS_CUDA_CHECK(ctx, cudaMalloc((void**)&g->d_pointCell, sizeof(uint32_t) * nn));
    SELAS_CUDA_CHECK(ctx, cudaMalloc((void**)&g->d_cellPoints, sizeof(float4) * nn));
    SELAS_CUDA_CHECK(ctx, cudaMalloc((void**)&g->d_blockSums, sizeof(uint32_t) * (size_t)(tiles + 1)));
    // a cell is "big" when it holds > 16 points, so at most n/17 of them
    SELAS_CUDA_CHECK(ctx, cudaMalloc((void**)&g->d_bigCells, sizeof(uint32_t) * (nn / 17 + 2)));
    g->capacityN = n;
    g->capacityCells = cellCount;
    return SELAS_OK;
}

int buildOnDevice(SelasHashGrid* g, size_t n, const float* d_pos, float radius, uint32_t cellCount, cudaStream_t st) {
    SelasContext* ctx = g->ctx;
    int rc = reserve(g, n, cellCount);
    if (rc) return rc;
    g->n = n;
    g->cellCount = cellCount;
    g->radius = radius;
    g->radiusSq = radius * radius;            // VCMHashGrid.cpp:50-52
    g->inverseCellSize = 1.0f / (2.0f * radius);
    g->scanBlocks = divUp(cellCount, kScanTile);
    // MakeInvalid (AxisAlignedBox.cpp): min = +FLT_MAX, max = -FLT_MAX; in the order-preserving encoding
    uint32_t init[8] = {0xFF7FFFFFu, 0xFF7FFFFFu, 0xFF7FFFFFu, 0x00800000u, 0x00800000u, 0x00800000u, 0u, 0u};
    SELAS_CUDA_CHECK(ctx, cudaMemcpyAsync(g->d_bounds, init, sizeof(init), cudaMemcpyHostToDevice, st));
    SELAS_CUDA_CHECK(ctx, cudaMemsetAsync(g->d_cursor, 0, sizeof(uint32_t) * (size_t)cellCount, st));
    if (n) {
        int grid = std::min<size_t>((n + 255) / 256, (size_t)ctx->smCount * 8);
        boundsKernel<<<grid, 256, 0, st>>>(d_pos, n, g->d_bounds);
        countKernel<<<divUp(n, 256), 256, 0, st>>>(d_pos, n, g->d_bounds, g->inverseCellSize, cellCount, g->d_cursor, g->d_pointCell);
    }
    scanTileSums<<<g->scanBlocks, kScanThreads, 0, st>>>(g->d_cursor, cellCount, g->d_blockSums);
    scanSums<<<1, 1024, 0, st>>>(g->d_blockSums, g->scanBlocks);
    scanTiles<<<g->scanBlocks, kScanThreads, 0, st>>>(g->d_cursor, cellCount, g->d_blockSums, g->d_cellEnds, g->d_cursor);
    if (n) {
        scatterKernel<<<divUp(n, 256), 256, 0, st>>>(g->d_pointCell, n, g->d_cursor, g->d_cellIndices);
        uint32_t* bigCount = g->d_bounds + 6;
        orderCellsKernel<<<divUp(cellCount, 256), 256, 0, st>>>(g->d_cellEnds, cellCount, g->d_cellIndices, g->d_bigCells, bigCount);
        bigCellsKernel<<<ctx->smCount * 2, 256, 0, st>>>(g->d_cellEnds, g->d_bigCells, bigCount, g->d_cellIndices);
        gatherKernel<<<divUp(n, 256), 256, 0, st>>>(d_pos, g->d_cellIndices, n, g->d_cellPoints);
    }
    SELAS_CUDA_CHECK(ctx, cudaGetLastError());
    return SELAS_OK;
}

}  // namespace

// for vcm.cu: a by-value view of a built grid, and the 3-kernel inclusive/exclusive scan (sums: n / 2048 + 1 words of scratch)
HashGridView hashGridView(const SelasHashGrid* g) {
    HashGridView v;
    v.bounds = g->d_bounds;
    v.cellEnds = g->d_cellEnds;
    v.cellPoints = g->d_cellPoints;
    v.inverseCellSize = g->inverseCellSize;
    v.radiusSq = g->radiusSq;
    v.cellCount = g->cellCount;
    v.n = (uint32_t)g->n;
    return v;
}
// a grid whose buffers hold up to n points / cellCount cells, so that per-iteration rebuilds never allocate
int hashGridReserve(SelasContext* ctx, SelasHashGrid** inout, size_t n, uint32_t cellCount) {
    SelasHashGrid* g = *inout;
    if (!g) {
        g = new (std::nothrow) SelasHashGrid();
        if (!g) return SELAS_ERR_MEMORY;
        g->ctx = ctx;
        *inout = g;
    }
    return reserve(g, n, cellCount);
}
uint32_t scanScratchWords(uint32_t n) { return divUp(n, kScanTile) + 1; }
int launchScanU32(SelasContext* ctx, const uint32_t* d_in, uint32_t n, uint32_t* d_sums, uint32_t* d_inclusive, uint32_t* d_exclusive,
                  cudaStream_t st) {
    if (!n) return SELAS_OK;
    uint32_t blocks = divUp(n, kScanTile);
    scanTileSums<<<blocks, kScanThreads, 0, st>>>(d_in, n, d_sums);
    scanSums<<<1, 1024, 0, st>>>(d_sums, blocks);
    scanTiles<<<blocks, kScanThreads, 0, st>>>(d_in, n, d_sums, d_inclusive, d_exclusive);
    SELAS_CUDA_CHECK(ctx, cudaGetLastError());
    return SELAS_OK;
}
}  // namespace selas

using namespace selas;

extern "C" {

int selas_cuda_hashgrid_build_device(SelasContext* ctx, size_t n, const float* d_positions3, float radius, uint32_t cellCount,
                                     SelasHashGrid** inout, void* stream) {
    if (!ctx || !inout) {
        setError(ctx, "selas_cuda_hashgrid_build_device: null argument");
        return SELAS_ERR_ARGUMENT;
    }
    SELAS_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    if (cellCount == 0 || cellCount >= 0x80000000u || !(radius > 0.0f) || n > 0xFFFFFFF0ull || (n && !d_positions3)) {
        setError(ctx, "selas_cuda_hashgrid_build: need 0 < cellCount < 2^31, radius > 0, n < 2^32");
        return SELAS_ERR_ARGUMENT;
    }
    SelasHashGrid* g = *inout;
    if (!g) {
        g = new (std::nothrow) SelasHashGrid();
        if (!g) return SELAS_ERR_MEMORY;
        g->ctx = ctx;
    }
    int rc = buildOnDevice(g, n, d_positions3, radius, cellCount, stream ? (cudaStream_t)stream : ctx->stream);
    if (rc) {
        if (!*inout) {
            releaseBuffers(g);
            delete g;
        }
        return rc;
    }
    *inout = g;
    return SELAS_OK;
}

int selas_cuda_hashgrid_build(SelasContext* ctx, size_t n, const float* positions3, float radius, uint32_t cellCount, SelasHashGrid** out) {
    if (!ctx || !out) {
        setError(ctx, "selas_cuda_hashgrid_build: null argument");
        return SELAS_ERR_ARGUMENT;
    }
    *out = nullptr;
    SELAS_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    if (n && !positions3) {
        setError(ctx, "selas_cuda_hashgrid_build: positions is null");
        return SELAS_ERR_ARGUMENT;
    }
    float* d_pos = nullptr;
    SELAS_CUDA_CHECK(ctx, cudaMalloc((void**)&d_pos, sizeof(float) * 3 * (n ? n : 1)));
    cudaError_t e = cudaMemcpyAsync(d_pos, positions3, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, ctx->stream);
    int rc = SELAS_OK;
    if (e != cudaSuccess) {
        setError(ctx, "upload of %zu positions failed: %s", n, cudaGetErrorString(e));
        rc = SELAS_ERR_CUDA;
    }
    if (!rc) rc = selas_cuda_hashgrid_build_device(ctx, n, d_pos, radius, cellCount, out, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_pos);
    return rc;
}

int selas_cuda_hashgrid_destroy(SelasHashGrid* g) {
    if (!g) return SELAS_OK;
    cudaSetDevice(g->ctx->device);
    cudaStreamSynchronize(g->ctx->stream);
    releaseBuffers(g);
    delete g;
    return SELAS_OK;
}

int selas_cuda_hashgrid_read(const SelasHashGrid* g, uint32_t* cellRangeEnds, uint32_t* cellIndices, float* aabbMinMax6) {
    if (!g) return SELAS_ERR_ARGUMENT;
    SelasContext* ctx = g->ctx;
    SELAS_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    SELAS_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    if (cellRangeEnds) SELAS_CUDA_CHECK(ctx, cudaMemcpy(cellRangeEnds, g->d_cellEnds, sizeof(uint32_t) * (size_t)g->cellCount, cudaMemcpyDeviceToHost));
    if (cellIndices && g->n) SELAS_CUDA_CHECK(ctx, cudaMemcpy(cellIndices, g->d_cellIndices, sizeof(uint32_t) * g->n, cudaMemcpyDeviceToHost));
    if (aabbMinMax6) {
        uint32_t b[6];
        SELAS_CUDA_CHECK(ctx, cudaMemcpy(b, g->d_bounds, sizeof(b), cudaMemcpyDeviceToHost));
        for (int k = 0; k < 6; ++k) aabbMinMax6[k] = floatFromOrdered(b[k]);
    }
    return SELAS_OK;
}

int selas_cuda_hashgrid_query_device(const SelasHashGrid* g, size_t m, const float* d_queries3, uint32_t* d_counts, uint32_t* d_indices,
                                     uint32_t maxPerQuery, void* stream) {
    if (!g) return SELAS_ERR_ARGUMENT;
    SelasContext* ctx = g->ctx;
    SELAS_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    if (m == 0) return SELAS_OK;
    if (!d_queries3 || !d_counts) {
        setError(ctx, "selas_cuda_hashgrid_query: null argument");
        return SELAS_ERR_ARGUMENT;
    }
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    static const bool perThread = getenv("SELAS_HASHGRID_PER_THREAD_QUERY") != nullptr;  // tuning knob: the first version of the kernel
    if (perThread)
        queryKernel<<<divUp(m, 128), 128, 0, st>>>(g->d_bounds, g->inverseCellSize, g->radiusSq, g->cellCount, g->d_cellEnds, g->d_cellPoints, m,
                                                   d_queries3, d_counts, d_indices, maxPerQuery);
    else {
        uint32_t grid = (uint32_t)std::min<size_t>((m + kQueryWarps * 32 - 1) / (kQueryWarps * 32), (size_t)ctx->smCount * 16);
        queryWarpKernel<<<grid, kQueryWarps * 32, 0, st>>>(g->d_bounds, g->inverseCellSize, g->radiusSq, g->cellCount, g->d_cellEnds,
                                                           g->d_cellPoints, m, d_queries3, d_counts, d_indices, maxPerQuery);
    }
    SELAS_CUDA_CHECK(ctx, cudaGetLastError());
    return SELAS_OK;
}

int selas_cuda_hashgrid_query(const SelasHashGrid* g, size_t m, const float* queries3, uint32_t* counts, uint32_t* indices,
                              uint32_t maxPerQuery) {
    if (!g) return SELAS_ERR_ARGUMENT;
    SelasContext* ctx = g->ctx;
    SELAS_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    if (m == 0) return SELAS_OK;
    if (!queries3 || !counts) {
        setError(ctx, "selas_cuda_hashgrid_query: null argument");
        return SELAS_ERR_ARGUMENT;
    }
    float* d_q = nullptr;
    uint32_t *d_c = nullptr, *d_i = nullptr;
    int rc = SELAS_OK;
    auto fail = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && !rc) {
            setError(ctx, "%s failed: %s", what, cudaGetErrorString(e));
            rc = SELAS_ERR_CUDA;
        }
    };
    fail(cudaMalloc((void**)&d_q, sizeof(float) * 3 * m), "cudaMalloc(queries)");
    fail(cudaMalloc((void**)&d_c, sizeof(uint32_t) * m), "cudaMalloc(counts)");
    if (indices && maxPerQuery) fail(cudaMalloc((void**)&d_i, sizeof(uint32_t) * m * maxPerQuery), "cudaMalloc(indices)");
    if (!rc) fail(cudaMemcpyAsync(d_q, queries3, sizeof(float) * 3 * m, cudaMemcpyHostToDevice, ctx->stream), "upload");
    if (!rc && d_i) fail(cudaMemsetAsync(d_i, 0xff, sizeof(uint32_t) * m * maxPerQuery, ctx->stream), "memset");
    if (!rc) rc = selas_cuda_hashgrid_query_device(g, m, d_q, d_c, d_i, maxPerQuery, ctx->stream);
    if (!rc) fail(cudaMemcpyAsync(counts, d_c, sizeof(uint32_t) * m, cudaMemcpyDeviceToHost, ctx->stream), "download counts");
    if (!rc && d_i) fail(cudaMemcpyAsync(indices, d_i, sizeof(uint32_t) * m * maxPerQuery, cudaMemcpyDeviceToHost, ctx->stream), "download indices");
    fail(cudaStreamSynchronize(ctx->stream), "synchronize");
    cudaFree(d_q);
    cudaFree(d_c);
    cudaFree(d_i);
    return rc;
}

}  // extern "C"
