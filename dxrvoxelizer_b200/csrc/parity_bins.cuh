// parity_bins.cuh -- MODE_PARITY candidate lists by triangle-parallel 2-D binning (k_bin_columns, trace_parity.cu).
// (Kept apart from the kernel: the fused build once called it too -- binning the records for the grid of the previous
// voxelize while it still held them.  Measured: the build grew by the 12 us the voxelize lost; the binning is a chain of
// atomic round trips wherever it runs.  Dropped.)
//
// What is computed, per super-tile of SY x SZ voxel columns: the triangles whose (y,z) box meets the tile's rectangle of
// column centres.  One thread per sorted triangle turns its box into a conservative range of tiles, tests it exactly
// against the tiles' tabulated rectangles, and appends its slot to their lists with one atomic each.
#pragma once
#include "parity_common.cuh"

namespace dxrv
{
struct BinParams
{
    uint32_t N, z0, z1;        // grid size, slab (N == 0: no binning)
    uint32_t tilesY;           // super-tiles along y
    uint32_t candCap;
    float invNPow2;            // 1/N when N is a power of two, else 0
    uint32_t* candCount;       // [numTiles], zero on entry
    uint32_t* candList;        // [numTiles][candCap]
};

constexpr int kBinSY = 16, kBinSZ = 8;   // the super-tile of the fill kernel
#ifndef DXRV_BIN_EXACT_RECT
#define DXRV_BIN_EXACT_RECT 1
#endif

template <int SY, int SZ>
__device__ __forceinline__ uint32_t binTableFloats(const BinParams& bp) { return 2u * (bp.tilesY + (bp.z1 - bp.z0 + SZ - 1) / SZ); }

// yMin/yMax per tile column, zMin/zMax per tile row, from the exact column centres; block-wide, sync afterwards
template <int SY, int SZ>
__device__ __forceinline__ void binTablesSetup(const BinParams& bp, float* sRect)
{
    const uint32_t tilesY = bp.tilesY, tilesZ = (bp.z1 - bp.z0 + SZ - 1) / SZ;
    float* yMin = sRect; float* yMax = yMin + tilesY; float* zMin = yMax + tilesY; float* zMax = zMin + tilesZ;
    const float fN = (float)bp.N;
    for (uint32_t i = threadIdx.x; i < tilesY; i += blockDim.x)
    {
        const uint32_t sy0 = i * SY, yLast = min(sy0 + SY - 1, bp.N - 1);
        yMax[i] = -centreOf(sy0, fN, bp.invNPow2);   // scene Y decreases with y
        yMin[i] = -centreOf(yLast, fN, bp.invNPow2);
    }
    for (uint32_t i = threadIdx.x; i < tilesZ; i += blockDim.x)
    {
        const uint32_t sz0 = bp.z0 + i * SZ, zLast = min(sz0 + SZ - 1, bp.z1 - 1);
        zMin[i] = centreOf(sz0, fN, bp.invNPow2);
        zMax[i] = centreOf(zLast, fN, bp.invNPow2);
    }
}

// One triangle per lane (`has`: this lane holds one; slot j = its place in the sorted records); called by all 32 lanes.
template <int SY, int SZ>
__device__ __forceinline__ void binTriangleWarp(const BinParams& bp, const float* sRect, bool has, const float4& a, const float4& b,
                                                const float4& c, uint32_t j)
{
    const uint32_t tilesY = bp.tilesY, tilesZ = (bp.z1 - bp.z0 + SZ - 1) / SZ;
    const float* yMin = sRect; const float* yMax = yMin + tilesY; const float* zMin = yMax + tilesY; const float* zMax = zMin + tilesZ;
    const float fN = (float)bp.N, halfN = 0.5f * fN;
    const uint32_t lane = laneId();
    const int layers = (int)(bp.z1 - bp.z0);
    int ty0 = 0, ty1 = -1, tz0 = 0, tz1 = -1;
    float ylo = 0, yhi = 0, zlo = 0, zhi = 0;
    if (has)
    {
        ylo = fminf(fminf(a.y, b.y), c.y); yhi = fmaxf(fmaxf(a.y, b.y), c.y);
        zlo = fminf(fminf(a.z, b.z), c.z); zhi = fmaxf(fmaxf(a.z, b.z), c.z);
        // conservative index range (one voxel of slack; the exact compares below decide), then tiles
        const float yA = (1.0f - yhi) * halfN - 1.5f, yB = (1.0f - ylo) * halfN + 0.5f;
        const float zA = (zlo + 1.0f) * halfN - 1.5f - (float)bp.z0, zB = (zhi + 1.0f) * halfN + 0.5f - (float)bp.z0;
        if (yB >= 0.0f && yA <= fN - 1.0f && zB >= 0.0f && zA <= (float)(layers - 1))   // (false for NaN boxes)
        {
            ty0 = max((int)floorf(yA), 0) / SY; ty1 = min((int)ceilf(yB), (int)bp.N - 1) / SY;
            tz0 = max((int)floorf(zA), 0) / SZ; tz1 = min((int)ceilf(zB), layers - 1) / SZ;
        }
    }
#if DXRV_BIN_EXACT_RECT
    // The tiles that pass the exact compares form a sub-rectangle of the conservative one (the tabulated limits are
    // monotonic along each axis, and the y and z compares are independent): shrink the range here, once, instead of
    // testing every tile of the conservative range in the rounds below -- a triangle near a tile border no longer costs
    // its whole warp extra rounds (each batch of rounds is a round trip of atomics), and the rounds need no division.
    while (ty0 <= ty1 && !(ylo <= yMax[ty0] && yhi >= yMin[ty0])) ++ty0;
    while (ty1 >= ty0 && !(ylo <= yMax[ty1] && yhi >= yMin[ty1])) --ty1;
    while (tz0 <= tz1 && !(zlo <= zMax[tz0] && zhi >= zMin[tz0])) ++tz0;
    while (tz1 >= tz0 && !(zlo <= zMax[tz1] && zhi >= zMin[tz1])) --tz1;
#endif
    const uint32_t nu = (uint32_t)max(ty1 - ty0 + 1, 0), nv = (uint32_t)max(tz1 - tz0 + 1, 0), n = nu * nv;
    auto emitTile = [&](uint32_t slot, int ty, int tz, float bylo, float byhi, float bzlo, float bzhi) {
        if (bylo <= yMax[ty] && byhi >= yMin[ty] && bzlo <= zMax[tz] && bzhi >= zMin[tz])
        {
            const uint32_t tile = (uint32_t)tz * tilesY + (uint32_t)ty;
            const uint32_t at = atomicAdd(bp.candCount + tile, 1u);
            if (at < bp.candCap) bp.candList[(size_t)tile * bp.candCap + at] = slot;
        }
    };
    // Small rectangles, all lanes in step: neighbours in Morton order mostly hit the SAME tile, so the lanes of a warp
    // that do are served by one atomic (a crowded tile otherwise takes thousands of serialised same-address atomics).
    const uint32_t nSmall = n <= 16u ? n : 0u;
    const uint32_t rounds = __reduce_max_sync(0xffffffffu, nSmall);
    // four rounds per batch: their atomics are all issued before the first result is consumed (one L2 round trip
    // per batch instead of one per round -- the kernel is a latency chain, not a throughput problem)
#if DXRV_BIN_EXACT_RECT
    uint32_t cy = 0, rowTile = (uint32_t)tz0 * tilesY + (uint32_t)ty0;   // round q = tile (ty0 + cy, row of rowTile)
#endif
    for (uint32_t q0 = 0; q0 < rounds; q0 += 4u)
    {
        uint32_t tileK[4], peersK[4], atK[4];
#pragma unroll
        for (uint32_t k = 0; k < 4u; ++k)
        {
            const uint32_t q = q0 + k;
            bool hit = false;
            tileK[k] = 0; peersK[k] = 0; atK[k] = 0;
#if DXRV_BIN_EXACT_RECT
            if (q < nSmall)
            {
                hit = true;
                tileK[k] = rowTile + cy;
                if (++cy == nu) { cy = 0; rowTile += tilesY; }
            }
#else
            if (q < nSmall)
            {
                const int ty = ty0 + (int)(q % nu), tz = tz0 + (int)(q / nu);
                hit = ylo <= yMax[ty] && yhi >= yMin[ty] && zlo <= zMax[tz] && zhi >= zMin[tz];
                tileK[k] = (uint32_t)tz * tilesY + (uint32_t)ty;
            }
#endif
            const uint32_t act = __ballot_sync(0xffffffffu, hit);
            if (hit)
            {
                peersK[k] = __match_any_sync(act, tileK[k]);
                if ((int)lane == __ffs(peersK[k]) - 1) atK[k] = atomicAdd(bp.candCount + tileK[k], (uint32_t)__popc(peersK[k]));
            }
        }
#pragma unroll
        for (uint32_t k = 0; k < 4u; ++k)
        {
            __syncwarp();
            if (peersK[k] != 0u)
            {
                const uint32_t at = __shfl_sync(peersK[k], atK[k], __ffs(peersK[k]) - 1) + (uint32_t)__popc(peersK[k] & laneMaskLt());
                if (at < bp.candCap) bp.candList[(size_t)tileK[k] * bp.candCap + at] = j;
            }
        }
    }
    uint32_t big = __ballot_sync(0xffffffffu, n > 16u);
    while (big)
    {
        const int L = __ffs(big) - 1;
        big &= big - 1u;
        const uint32_t bn = __shfl_sync(0xffffffffu, n, L), bnu = __shfl_sync(0xffffffffu, nu, L);
        const int by0 = __shfl_sync(0xffffffffu, ty0, L), bz0 = __shfl_sync(0xffffffffu, tz0, L);
        const uint32_t bj = __shfl_sync(0xffffffffu, j, L);
        const float b0 = __shfl_sync(0xffffffffu, ylo, L), b1 = __shfl_sync(0xffffffffu, yhi, L);
        const float b2 = __shfl_sync(0xffffffffu, zlo, L), b3 = __shfl_sync(0xffffffffu, zhi, L);
        for (uint32_t q = lane; q < bn; q += 32u) emitTile(bj, by0 + (int)(q % bnu), bz0 + (int)(q / bnu), b0, b1, b2, b3);
    }
}
}  // namespace dxrv
