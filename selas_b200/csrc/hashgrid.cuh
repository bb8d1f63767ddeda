// hashgrid.cuh -- device-side pieces of the VCM hash grid shared by hashgrid.cu (build / batch query) and vcm.cu (the merge
// kernel walks the grid itself): the cell hash, the order-preserving float encoding of the AABB, and a by-value view of a
// built grid.  Reference: Source/Applications/Selas/Source/VCMHashGrid.cpp:15-34, :107-155.
#pragma once
#include <cstring>

#include "vec.cuh"

namespace selas {

SELAS_D uint32_t orderedFromFloat(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
SELAS_HD float floatFromOrdered(uint32_t u) {
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

// GetCellIndex, VCMHashGrid.cpp:15-22.  The reference's `uint` is 64 bits wide (SystemLib/BasicTypes.h:47): the int32
// coordinates are sign-extended and the products, xor and modulo are 64-bit.
SELAS_D uint32_t cellHash(int32_t x, int32_t y, int32_t z, uint32_t cellCount) {
    uint64_t ux = (uint64_t)(int64_t)x, uy = (uint64_t)(int64_t)y, uz = (uint64_t)(int64_t)z;
    uint64_t h = (ux * 73856093ull) ^ (uy * 19349663ull) ^ (uz * 83492791ull);
    return (uint32_t)(h % (uint64_t)cellCount);
}

// What a kernel needs to walk a built grid (SearchHashGrid, VCMHashGrid.cpp:107-155).
struct HashGridView {
    const uint32_t* bounds;     // 6 order-preserving uints: min xyz, max xyz
    const uint32_t* cellEnds;   // cellRangeEnds[cellCount]
    const float4* cellPoints;   // n x {position, vertex index}, contiguous per cell, ascending index inside a cell
    float inverseCellSize, radiusSq;
    uint32_t cellCount;
    uint32_t n;
};

}  // namespace selas
