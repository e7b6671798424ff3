// scatter_parity.cu -- MODE_PARITY for FINE meshes on COARSE grids: triangle-parallel scatter (sm_100a).
//
// The tile kernels of trace_parity.cu parallelise over super-tiles of 16 x 8 columns.  The reference's own
// regime -- GRID_SIZE 64 (Content/Voxelizer.cpp:8), and every configuration where the triangles are smaller
// than the voxels -- has few tiles and thousands of candidate triangles in each: 100 k triangles on a 64^3
// grid are 32 tiles.  There the parallelism is in the TRIANGLES:
//
//   memset              the slab is cleared (it is small in this regime)
//   k_scatter_crossings one warp per chunk of 32 consecutive sorted triangles; the same two-level flattening as
//                       the tile kernel (triangles -> row units with a conservative y interval -> pairs), the
//                       same exact crossing test (Spec H, parity_common.cuh); a crossing toggles ONE bit of the
//                       grid with a global atomicXor -- neighbouring triangles are neighbours in Morton order,
//                       so the atomics of a warp land in a handful of L2 lines.  Most triangles cross no column
//                       centre at all and cost one record.
//   k_scatter_huge      the few triangles that cover thousands of columns (a ground plane under a detailed object)
//                       would keep one warp busy for milliseconds: k_scatter_crossings only lists them, and this
//                       kernel deals their ROWS to all the warps of the machine (it exits at once when the list is empty)
//   k_prefix_rows       occupancy = prefix-XOR of the toggles along x, in place: 128 bits per lane, ballot carry
//
// XOR commutes, so the result is bit-identical to the tile path's whatever the order of the atomics.
// The LBVH is not needed here (its sorted triangle records are); it is still built, the metric counts it.
#include <algorithm>
#include "kernels.h"
#include "parity_common.cuh"

namespace dxrv
{
namespace
{
constexpr int kScatterWarps = 4;
constexpr uint32_t kHugePairs = 8192;   // rows x bounding columns from which a triangle goes to k_scatter_huge
constexpr uint32_t kHugeCap = 4096;     // triangles that list holds (more are handled in line)

struct ScatterParams
{
    const Tri48* tris;
    uint32_t numTris;
    uint32_t N, P;
    uint32_t z0, z1;
    float invNPow2;          // 1/N when N is a power of two, else 0
    uint32_t* grid;
    unsigned long long* crossings;
    uint32_t* hugeCount;     // one word, zero on entry; reset by the prefix kernel
    uint32_t* hugeList;      // [hugeCap] sorted-triangle slots
    uint32_t hugeCap;
};

__global__ void __launch_bounds__(32 * kScatterWarps)
k_scatter_crossings(const ScatterParams prm)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t N = prm.N, P = prm.P, layers = prm.z1 - prm.z0;
    const float fN = (float)N, invNPow2 = prm.invNPow2, halfN = 0.5f * fN;
    float4* stage = reinterpret_cast<float4*>(smem);                       // [warps][32 records x 3 float4 + 32 row units]
    float* colY = reinterpret_cast<float*>(stage + kScatterWarps * 112);   // [N]      scene Y of the columns (decreasing)
    float* rowZ = colY + N;                                                // [layers] scene Z of the slab's rows
    for (uint32_t i = tid; i < N; i += blockDim.x) colY[i] = -centreOf(i, fN, invNPow2);
    for (uint32_t i = tid; i < layers; i += blockDim.x) rowZ[i] = centreOf(prm.z0 + i, fN, invNPow2);
    __syncthreads();

    float4* tab = stage + warp * 112u;
    uint2* units = reinterpret_cast<uint2*>(tab + 96);
    const uint32_t lt = laneMaskLt(), le = lt | (1u << lane);
    uint32_t myCrossings = 0;
    const uint32_t numChunks = (prm.numTris + 31u) / 32u;
    for (uint32_t chunk = blockIdx.x * kScatterWarps + warp; chunk < numChunks; chunk += gridDim.x * kScatterWarps)
    {
        const uint32_t slot = chunk * 32u + lane;
        float4 a = make_float4(0, 0, 0, 0), b = a, c = a;
        uint32_t zA = 0, h = 0;
        if (slot < prm.numTris)
        {
            const float4* t = reinterpret_cast<const float4*>(prm.tris + slot);
            a = __ldg(t); b = __ldg(t + 1); c = __ldg(t + 2);
            // rows of the slab whose centre may lie in [zlo, zhi]
            const float zlo = fminf(fminf(a.z, b.z), c.z), zhi = fmaxf(fmaxf(a.z, b.z), c.z);
            const int r0 = min(max((int)ceilf((zlo + 1.0f) * halfN - 0.5f - kIdxSlack) - (int)prm.z0, 0), (int)layers);
            const int r1 = min(max((int)floorf((zhi + 1.0f) * halfN - 0.5f + kIdxSlack) - (int)prm.z0, -1), (int)layers - 1);
            zA = (uint32_t)r0; h = (uint32_t)max(r1 - r0 + 1, 0);
            // a triangle over thousands of columns is listed for k_scatter_huge instead (when the list has room)
            const float ylo = fminf(fminf(a.y, b.y), c.y), yhi = fmaxf(fmaxf(a.y, b.y), c.y);
            const float wEst = fminf((yhi - ylo) * halfN + 2.0f, fN);
            if (h != 0u && (float)h * wEst > (float)kHugePairs)
            {
                const uint32_t at = atomicAdd(prm.hugeCount, 1u);
                if (at < prm.hugeCap) { prm.hugeList[at] = slot; h = 0; }
            }
        }
        // records {a.xyz, zA | b.xyz, first row unit | c.xyz, -} of the triangles that have rows; owner lookup
        // by start masks exactly as in k_trace_fill_columns
        uint32_t incl = h;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (uint32_t)o) incl += v;
        }
        const uint32_t totalRows = __shfl_sync(0xffffffffu, incl, 31);
        if (totalRows == 0u) continue;
        const uint32_t nzA = __ballot_sync(0xffffffffu, h != 0u);
        if (h != 0u)
        {
            float4* r = tab + __popc(nzA & lt) * 3u;
            a.w = __uint_as_float(zA); b.w = __uint_as_float(incl - h);
            r[0] = a; r[1] = b; r[2] = c;
        }
        __syncwarp();
        const uint32_t myStartA = lane < (uint32_t)__popc(nzA) ? __float_as_uint(tab[lane * 3u + 1u].w) : 0xffffffffu;
        uint32_t beforeA = 0;
#pragma unroll 1
        for (uint32_t r0 = 0; r0 < totalRows; r0 += 32u)
        {
            const uint32_t relA = myStartA - r0;
            const uint32_t startsA = __reduce_or_sync(0xffffffffu, relA < 32u ? 1u << relA : 0u);
            const uint32_t ownerA = beforeA + __popc(startsA & le) - 1u;
            beforeA += __popc(startsA);
            uint32_t cnt = 0, unit = 0;
            if (r0 + lane < totalRows)
            {
                const float4* r = tab + ownerA * 3u;
                const float4 ta = r[0], tb = r[1], tc = r[2];
                const uint32_t zl = __float_as_uint(ta.w) + (r0 + lane - __float_as_uint(tb.w));
                float lo, hi;
                rowIntervalY(ta, tb, tc, rowZ[zl], lo, hi);
                // columns whose centre may lie in [lo, hi] (colY decreases with y)
                const int ya = min(max((int)ceilf((1.0f - hi) * halfN - 0.5f - kIdxSlack), 0), (int)N);
                const int yb = min(max((int)floorf((1.0f - lo) * halfN - 0.5f + kIdxSlack), -1), (int)N - 1);
                cnt = (uint32_t)max(yb - ya + 1, 0);
                unit = ownerA | (zl << 5) | ((uint32_t)ya << 18);   // 5 + 13 + 13 bits: N <= 8192
            }
            uint32_t inclB = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const uint32_t v = __shfl_up_sync(0xffffffffu, inclB, o);
                if (lane >= (uint32_t)o) inclB += v;
            }
            const uint32_t totalPairs = __shfl_sync(0xffffffffu, inclB, 31);
            if (totalPairs == 0u) continue;
            const uint32_t nzB = __ballot_sync(0xffffffffu, cnt != 0u);
            __syncwarp();   // the previous window's pairs have read the row units
            if (cnt != 0u) units[__popc(nzB & lt)] = make_uint2(unit, inclB - cnt);
            __syncwarp();
            const uint32_t myStartB = lane < (uint32_t)__popc(nzB) ? units[lane].y : 0xffffffffu;
            uint32_t beforeB = 0;
#pragma unroll 1
            for (uint32_t p0 = 0; p0 < totalPairs; p0 += 32u)
            {
                const uint32_t relB = myStartB - p0;
                const uint32_t startsB = __reduce_or_sync(0xffffffffu, relB < 32u ? 1u << relB : 0u);
                const uint32_t ownerB = beforeB + __popc(startsB & le) - 1u;
                beforeB += __popc(startsB);
                const uint32_t p = p0 + lane;
                if (p < totalPairs)
                {
                    const uint2 u = units[ownerB];
                    const float4* r = tab + (u.x & 31u) * 3u;
                    const float4 ta = r[0], tb = r[1], tc = r[2];
                    const uint32_t y = (u.x >> 18) + (p - u.y), zl = (u.x >> 5) & 0x1fffu;
                    uint32_t ix;
                    if (columnCrossing(ta, tb, tc, colY[y], rowZ[zl], N, fN, invNPow2, ix))
                    {
                        ++myCrossings;
                        if (ix < N) atomicXor(prm.grid + ((size_t)zl * N + y) * P + (ix >> 5), 1u << (ix & 31u));
                    }
                }
            }
        }
        __syncwarp();   // the table is rewritten by the next chunk
    }
    for (int o = 16; o > 0; o >>= 1) myCrossings += __shfl_xor_sync(0xffffffffu, myCrossings, o);
    if (lane == 0 && myCrossings) atomicAdd(prm.crossings, (unsigned long long)myCrossings);
}

// The listed huge triangles: every warp of the grid takes rows of every one of them at a stride, its lanes the
// columns of the row's conservative interval.
__global__ void __launch_bounds__(128)
k_scatter_huge(const ScatterParams prm)
{
    const uint32_t count = min(__ldcg(prm.hugeCount), prm.hugeCap);
    if (count == 0u) return;
    const uint32_t lane = laneId();
    const uint32_t gwarp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), numWarps = gridDim.x * (blockDim.x >> 5);
    const uint32_t N = prm.N, P = prm.P, layers = prm.z1 - prm.z0;
    const float fN = (float)N, invNPow2 = prm.invNPow2, halfN = 0.5f * fN;
    uint32_t myCrossings = 0;
    for (uint32_t b = 0; b < count; ++b)
    {
        const float4* t = reinterpret_cast<const float4*>(prm.tris + __ldcg(prm.hugeList + b));
        const float4 ta = __ldg(t), tb = __ldg(t + 1), tc = __ldg(t + 2);
        const float zlo = fminf(fminf(ta.z, tb.z), tc.z), zhi = fmaxf(fmaxf(ta.z, tb.z), tc.z);
        const int r0 = min(max((int)ceilf((zlo + 1.0f) * halfN - 0.5f - kIdxSlack) - (int)prm.z0, 0), (int)layers);
        const int r1 = min(max((int)floorf((zhi + 1.0f) * halfN - 0.5f + kIdxSlack) - (int)prm.z0, -1), (int)layers - 1);
        for (int zl = r0 + (int)gwarp; zl <= r1; zl += (int)numWarps)
        {
            const float Zc = centreOf(prm.z0 + (uint32_t)zl, fN, invNPow2);
            float lo, hi;
            rowIntervalY(ta, tb, tc, Zc, lo, hi);
            const int ya = min(max((int)ceilf((1.0f - hi) * halfN - 0.5f - kIdxSlack), 0), (int)N);
            const int yb = min(max((int)floorf((1.0f - lo) * halfN - 0.5f + kIdxSlack), -1), (int)N - 1);
            for (int y = ya + (int)lane; y <= yb; y += 32)
            {
                uint32_t ix;
                if (columnCrossing(ta, tb, tc, -centreOf((uint32_t)y, fN, invNPow2), Zc, N, fN, invNPow2, ix))
                {
                    ++myCrossings;
                    if (ix < N) atomicXor(prm.grid + ((size_t)zl * N + (uint32_t)y) * P + (ix >> 5), 1u << (ix & 31u));
                }
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) myCrossings += __shfl_xor_sync(0xffffffffu, myCrossings, o);
    if (lane == 0 && myCrossings) atomicAdd(prm.crossings, (unsigned long long)myCrossings);
}

// toggles -> occupancy, in place.  kGroups = 128-bit groups per row (a power of two <= 32): a warp holds
// 32 / kGroups whole rows, one group per lane, and the carry into a lane is the parity of the lower lanes of its row.
template <int kGroups>
__global__ void __launch_bounds__(256)
k_prefix_rows_vec(uint4* __restrict__ grid, size_t numGroups, uint32_t* __restrict__ hugeCount)
{
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g == 0) *hugeCount = 0;   // leave the list empty for the next launch
    const uint32_t lane = laneId();
    uint4 t = g < numGroups ? grid[g] : make_uint4(0, 0, 0, 0);
    uint32_t par;
    t.x = prefixXor32(t.x); par = t.x >> 31;
    t.y = prefixXor32(t.y) ^ (0u - par); par = t.y >> 31;
    t.z = prefixXor32(t.z) ^ (0u - par); par = t.z >> 31;
    t.w = prefixXor32(t.w) ^ (0u - par); par = t.w >> 31;
    const uint32_t bal = __ballot_sync(0xffffffffu, par);
    const uint32_t segLo = lane & ~(uint32_t)(kGroups - 1);   // first lane of this row
    const uint32_t flip = 0u - (__popc(bal & laneMaskLt() & ~((1u << segLo) - 1u)) & 1u);
    t.x ^= flip; t.y ^= flip; t.z ^= flip; t.w ^= flip;
    if (g < numGroups) grid[g] = t;
}

// any row length: one thread per row
__global__ void __launch_bounds__(256)
k_prefix_rows_any(uint32_t* __restrict__ grid, size_t numRows, uint32_t P, uint32_t tailMask, uint32_t* __restrict__ hugeCount)
{
    const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row == 0) *hugeCount = 0;
    if (row >= numRows) return;
    uint32_t* w = grid + row * P;
    uint32_t carry = 0;
    for (uint32_t k = 0; k < P; ++k)
    {
        uint32_t v = prefixXor32(w[k]) ^ (0u - carry);
        carry = v >> 31;
        if (k == P - 1u) v &= tailMask;
        w[k] = v;
    }
}
}  // namespace

// Triangles at least as dense as a quarter of the columns: the average triangle then covers a few columns at most
// and the grid is small next to the mesh.  (A coarse mesh on the same grid stays on the tile path, which
// parallelises over area.)  N <= 2048 keeps the column tables in shared memory.
bool useScatterParity(uint32_t numTris, uint32_t N)
{
    return N <= 2048u && numTris >= 2u && (uint64_t)numTris * 4u >= (uint64_t)N * N;
}

int launchScatterFillColumns(cudaStream_t s, const BvhView& bvh, uint32_t N, uint32_t z0, uint32_t z1, uint32_t* grid,
                             uint32_t* walkBuf, unsigned long long* dCrossings, cudaEvent_t* ev)
{
    ScatterParams prm;
    prm.tris = bvh.tris; prm.numTris = bvh.numTris;
    prm.N = N; prm.P = (N + 31) / 32;
    prm.z0 = z0; prm.z1 = z1;
    prm.invNPow2 = ((N & (N - 1)) == 0) ? 1.0f / (float)N : 0.0f;
    prm.grid = grid; prm.crossings = dCrossings;
    // scratch shared with the tile path (same buffer, never used by both in one call): the list counter is a spare
    // word of its zero-initialised, self-cleaning head; the list itself lies in the candidate-list area
    uint32_t numTiles, candCap;
    parityTileCounts(N, z0, z1, numTiles, candCap);
    prm.hugeCount = walkBuf + 6;
    prm.hugeList = walkBuf + (parityScratchWords(N, z0, z1) - (size_t)numTiles * candCap);
    prm.hugeCap = (uint32_t)std::min<size_t>(kHugeCap, (size_t)numTiles * candCap);
    const size_t numRows = (size_t)(z1 - z0) * N, words = numRows * prm.P;
    cudaMemsetAsync(dCrossings, 0, sizeof(unsigned long long), s);
    if (ev) cudaEventRecord(ev[0], s);
    cudaMemsetAsync(grid, 0, words * sizeof(uint32_t), s);
    const size_t smemBytes = sizeof(float4) * kScatterWarps * 112 + sizeof(float) * ((size_t)N + (z1 - z0));
    const uint32_t numChunks = (bvh.numTris + 31u) / 32u;
    uint32_t blocks = (numChunks + kScatterWarps - 1) / kScatterWarps;
    if (blocks > 148u * 16u) blocks = 148u * 16u;
    k_scatter_crossings<<<blocks, 32 * kScatterWarps, smemBytes, s>>>(prm);
    k_scatter_huge<<<148 * 4, 128, 0, s>>>(prm);
    if (ev) cudaEventRecord(ev[1], s);
    const uint32_t groups = prm.P / 4u;
    const bool vec = (prm.P & 3u) == 0u && (N & 31u) == 0u && groups <= 32u && (groups & (groups - 1u)) == 0u;
    if (vec)
    {
        const size_t numGroups = words / 4u;
        const unsigned pb = (unsigned)((numGroups + 255) / 256);
        uint4* g4 = reinterpret_cast<uint4*>(grid);
        switch (groups)
        {
        case 1: k_prefix_rows_vec<1><<<pb, 256, 0, s>>>(g4, numGroups, prm.hugeCount); break;
        case 2: k_prefix_rows_vec<2><<<pb, 256, 0, s>>>(g4, numGroups, prm.hugeCount); break;
        case 4: k_prefix_rows_vec<4><<<pb, 256, 0, s>>>(g4, numGroups, prm.hugeCount); break;
        case 8: k_prefix_rows_vec<8><<<pb, 256, 0, s>>>(g4, numGroups, prm.hugeCount); break;
        case 16: k_prefix_rows_vec<16><<<pb, 256, 0, s>>>(g4, numGroups, prm.hugeCount); break;
        default: k_prefix_rows_vec<32><<<pb, 256, 0, s>>>(g4, numGroups, prm.hugeCount); break;
        }
    }
    else
    {
        const uint32_t tailMask = (N & 31u) ? ((1u << (N & 31u)) - 1u) : 0xffffffffu;
        k_prefix_rows_any<<<(unsigned)((numRows + 255) / 256), 256, 0, s>>>(grid, numRows, prm.P, tailMask, prm.hugeCount);
    }
    if (ev) cudaEventRecord(ev[2], s);
    return 3;
}
}  // namespace dxrv
