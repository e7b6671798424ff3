// trace_parity.cu -- MODE_PARITY: +x column rays, watertight crossings, parity fill (sm_100a).
//
// Replaces DispatchRays + raygenMain/TraceRay (reference Content/Voxelizer.cpp:351-369,
// Content/Shaders/DXRVoxelizer.hlsl:58-85) for the column-parity formulation of the solid test.
// The per-(column, triangle) arithmetic is Spec H / MODE_PARITY of oracle/dxrv_oracle.h.
//
// The grid is cut into super-tiles of SY x SZ voxel columns (y,z); a super-tile owns ALL voxels of its
// columns along x.  Two kernels, each with the occupancy its phase needs:
//
//   k_walk_columns       latency bound, almost no shared memory, 64 warps / SM.  One WARP per
//                        super-tile walks the LBVH cooperatively: every lane pops a different node
//                        from a shared-memory stack, tests both child boxes against the super-tile's
//                        (y,z) rectangle (exact compares: a crossing implies the column lies inside
//                        every ancestor box), pushes inner children and appends leaf children to the
//                        super-tile's candidate list in HBM.  All walks of a 1024^2-column grid are in
//                        flight at once, so the kernel takes about one root-to-leaf latency chain.
//   k_trace_fill_columns writes every word of the slab exactly once, no clear pass, no scatter to HBM.
//                        * Empty super-tiles (most of a real grid) are streamed by one "writer" CTA per SM
//                          with TMA bulk stores out of a zeroed shared buffer, concurrently with
//                        * the busy super-tiles, which the other CTAs take in order of decreasing work
//                          (crowded tiles are split into parts that merge through a scratch buffer).
//                          Tracing: candidate triangles -> row units (triangle x z row, with a
//                          conservative y interval) -> (row unit, column) pairs, both flattened over the
//                          32 lanes of a warp; every pair gets the exact crossing test, and a crossing
//                          XORs a suffix mask into the column's shared-memory bit row -- XOR makes the
//                          order of crossings irrelevant, so no per-column hit list or sort is needed.
//                          Write-out: bit 31 of every word (the parity of its crossings) tells which of
//                          the following whole words to flip; coalesced 128-bit stores.
//                        A super-tile whose candidate list overflowed (huge meshes) walks the tree itself,
//                        CTA-cooperatively, instead of reading a list.
#include <algorithm>
#include <cstdlib>
#include "kernels.h"
#include "parity_common.cuh"
#include "parity_bins.cuh"
#include "timeline_debug.cuh"

namespace dxrv
{
namespace
{
constexpr int kStackGuard = 64;   // >= LBVH depth bound (62): head-room kept for depth-first popping
constexpr int kWalkStack = 512;   // per-warp stack of k_walk_columns
constexpr int kWalkWarps = 8;     // warps (= super-tiles) per CTA of k_walk_columns (<= 32)
// in-kernel fallback walk of the fill kernel, per thread of the CTA: stack entries and leaf-ring entries (>= 3)
#ifndef DXRV_CHUNK_MAX
#define DXRV_CHUNK_MAX 32
#endif
#ifndef DXRV_FILL_CTAS
#define DXRV_FILL_CTAS 9
#endif
constexpr int kChunkMax = DXRV_CHUNK_MAX;                 // candidates a warp stages per chunk (listed-candidates path)
constexpr int kStagePerWarp = 3 * kChunkMax + 16;         // float4: kChunkMax records of 48 bytes + 32 row units of 8 bytes
constexpr int kCandPerThread = kChunkMax >= 32 ? 6 : 3;
constexpr int kStackPerThread = kStagePerWarp / 8 - kCandPerThread;   // the walk's buffers share the staging area

struct ParityParams
{
    const BvhNode* nodes;
    const Tri48* tris;
    uint32_t numTris;
    uint32_t N, P, Ps;       // grid size, words per global row, words per shared row
    uint32_t gprShift;       // log2(Ps / 4) when Ps <= 128 (Ps is then a power of two)
    uint32_t z0, z1;
    uint32_t tilesY;         // super-tiles along y
    uint32_t numTiles;
    uint32_t tilesPad;       // numTiles rounded up to a multiple of 32
    uint32_t numWriters;     // leading CTAs of the fill kernel that only write the empty super-tiles
#ifdef DXRV_EXPERIMENT
    uint32_t noTrace;
#endif
    uint32_t bulkStores;     // 0: the writers use ordinary stores (DXRV_NO_BULK_STORE=1; compute-sanitizer's
                             // initcheck does not see what cp.async.bulk writes)
    float invNPow2;          // 1/N when N is a power of two, else 0
    uint32_t* grid;
    uint32_t* bucketCount;   // [0] heavy entries, [2] empty tiles, [3] heavy slots, [4] extra parts, [5] cursor of the empty-tile writers, [6] (scatter path) huge-triangle count, [8..11] light tiles per class, [32 + smid] writer claim of an SM
    uint32_t* lightTiles;    // [kLightClasses][tilesPad]  light tiles, classed by candidate count: tile | count << 24
    uint32_t* emptyTiles;    // [numTiles]
    uint2* heavyEntries;     // [numTiles + kExtraParts]  {tile, part | parts << 8 | slot << 16}
    uint32_t* heavyArrive;   // [kHeavySlots] parts of a split tile that have merged (self-resetting)
    uint32_t* heavyScratch;  // [kHeavySlots][128 * Ps] merged toggle rows of split tiles (self-cleaning)
    uint32_t* candCount;     // [numTiles]  leaves found by k_walk_columns; > candCap = overflow
    uint32_t* candList;      // [numTiles][candCap]
    uint32_t candCap;
    unsigned long long* crossings;
    uint32_t* err;
};

// (y,z) rectangle of a super-tile, from the exact column centres
template <int SY, int SZ>
__device__ __forceinline__ void tileRect(const ParityParams& prm, uint32_t sy0, uint32_t sz0, float& rYmin, float& rYmax,
                                         float& rZmin, float& rZmax)
{
    const float fN = (float)prm.N;
    const uint32_t yLast = min(sy0 + SY - 1, prm.N - 1), zLast = min(sz0 + SZ - 1, prm.z1 - 1);
    rYmax = -centreOf(sy0, fN, prm.invNPow2);   // scene Y decreases with y
    rYmin = -centreOf(yLast, fN, prm.invNPow2);
    rZmin = centreOf(sz0, fN, prm.invNPow2);
    rZmax = centreOf(zLast, fN, prm.invNPow2);
}

// both children of `ni` against the rectangle
__device__ __forceinline__ void testNode(const BvhNode* __restrict__ nodes, uint32_t ni, float rYmin, float rYmax,
                                         float rZmin, float rZmax, bool& ov0, bool& ov1, uint32_t& c0, uint32_t& c1)
{
    const float4* q = reinterpret_cast<const float4*>(nodes + ni);
    const float4 b0 = __ldg(q), b1 = __ldg(q + 1);                       // (ylo, yhi, zlo, zhi) of child 0 / 1
    const uint4 ch = __ldg(reinterpret_cast<const uint4*>(q + 3));
    ov0 = b0.x <= rYmax && b0.y >= rYmin && b0.z <= rZmax && b0.w >= rZmin;
    ov1 = b1.x <= rYmax && b1.y >= rYmin && b1.z <= rZmax && b1.w >= rZmin;
    c0 = ch.x; c1 = ch.y;
}

// ---- kernel A: one warp per super-tile walks the tree and lists the leaves it may cross ---------
// The tile is then filed as "heavy", "light" or "empty".  The fill kernel takes its CTAs in that order,
// and a heavy tile is SPLIT: several CTAs each rasterise a share of its candidate list and merge their
// toggle rows through a scratch buffer with atomicXor (XOR commutes); the last one to arrive fills.
// Real meshes leave most of the (y,z) plane empty and put hundreds of triangles into a few tiles
// (surfaces seen edge-on): without the split those few CTAs are the kernel's critical path.
constexpr uint32_t kHeavyTile = 192;    // candidates from which a tile is scheduled first (<= 256: light entries pack the count into 8 bits)
constexpr int kLightClasses = 4;        // light tiles are scheduled by halving classes of candidate count:
                                        // [96,192) [48,96) [24,48) [1,24) -- longest work first, so that the
                                        // kernel's tail is made of its smallest work items
#ifndef DXRV_SPLIT_TILE
#define DXRV_SPLIT_TILE 512
#endif
#ifndef DXRV_PART_SIZE
#define DXRV_PART_SIZE 256
#endif
#ifndef DXRV_CHUNK_NUM
#define DXRV_CHUNK_NUM 1   // a warp's chunk: candidates / (NUM/DEN * warps), as a power of two in [4, 32].  1: at least one chunk per warp
#endif                   // (2 = two chunks per warp: 10 % more warp instructions, equal or up to 4 % slower on five mesh / grid pairs; 1/2: mixed)
#ifndef DXRV_CHUNK_DEN
#define DXRV_CHUNK_DEN 1
#endif
constexpr uint32_t kSplitTile = DXRV_SPLIT_TILE;    // candidates from which a tile is split ...
constexpr uint32_t kPartSize = DXRV_PART_SIZE;     // ... into parts of about this many candidates
constexpr uint32_t kMaxParts = 64;
constexpr uint32_t kHeavySlots = 1024;  // tiles that can be split per launch
constexpr uint32_t kExtraParts = 2048;  // extra CTAs (beyond one per tile) a launch provides

// File the (up to 32) tiles of a CTA: one atomicAdd per bucket per CTA instead of one per tile -- with
// thousands of tiles hammering three counters the serialised atomics were a third of the kernel.
// Called by one full warp; lane i files tile firstTile + i (count == 0xffffffff: no such tile).
__device__ __forceinline__ void fileTiles(const ParityParams& prm, uint32_t firstTile, uint32_t count)
{
    const uint32_t lane = laneId(), lt = laneMaskLt();
    const uint32_t tile = firstTile + lane;
    const bool active = count != 0xffffffffu;
    if (active) prm.candCount[tile] = count;
    const bool isEmpty = active && count == 0u, isLight = active && count > 0u && count < kHeavyTile;
    const bool isHeavy = active && count >= kHeavyTile;
    int cls = -1;
    if (isLight) cls = count >= kHeavyTile / 2 ? 0 : count >= kHeavyTile / 4 ? 1 : count >= kHeavyTile / 8 ? 2 : 3;
    const uint32_t mE = __ballot_sync(0xffffffffu, isEmpty);
    uint32_t mC[kLightClasses], baseC[kLightClasses], baseE = 0;
#pragma unroll
    for (int k = 0; k < kLightClasses; ++k) { mC[k] = __ballot_sync(0xffffffffu, cls == k); baseC[k] = 0; }
    if (lane == 0)
    {
        if (mE) baseE = atomicAdd(prm.bucketCount + 2, __popc(mE));
#pragma unroll
        for (int k = 0; k < kLightClasses; ++k)
            if (mC[k]) baseC[k] = atomicAdd(prm.bucketCount + 8 + k, __popc(mC[k]));
    }
    baseE = __shfl_sync(0xffffffffu, baseE, 0);
    if (isEmpty) prm.emptyTiles[baseE + __popc(mE & lt)] = tile;
#pragma unroll
    for (int k = 0; k < kLightClasses; ++k)
    {
        const uint32_t b = __shfl_sync(0xffffffffu, baseC[k], 0);
        if (cls == k) prm.lightTiles[(size_t)k * prm.tilesPad + b + __popc(mC[k] & lt)] = tile | (count << 24);   // count < kHeavyTile <= 256, tile < 2^24
    }

    uint32_t parts = isHeavy ? 1u : 0u, slot = 0xffffu;
    if (isHeavy && count <= prm.candCap && count >= kSplitTile)   // (an overflowed list is not split: that CTA walks itself)
    {
        parts = min((count + kPartSize - 1u) / kPartSize, kMaxParts);
        slot = atomicAdd(prm.bucketCount + 3, 1u);
        if (slot >= kHeavySlots || atomicAdd(prm.bucketCount + 4, parts - 1u) + parts - 1u > kExtraParts) { parts = 1u; slot = 0xffffu; }
    }
    uint32_t incl = parts;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += v;
    }
    const uint32_t totalParts = __shfl_sync(0xffffffffu, incl, 31);
    uint32_t baseH = 0;
    if (lane == 0 && totalParts) baseH = atomicAdd(prm.bucketCount, totalParts);
    baseH = __shfl_sync(0xffffffffu, baseH, 0) + incl - parts;
    for (uint32_t part = 0; part < parts; ++part) prm.heavyEntries[baseH + part] = make_uint2(tile, part | (parts << 8) | (slot << 16));
}

template <int SY, int SZ>
__global__ void __launch_bounds__(32 * kWalkWarps)
k_walk_columns(const ParityParams prm)
{
    DXRV_TL_SCOPE();
    __shared__ uint32_t sStack[kWalkWarps][kWalkStack];
    __shared__ uint32_t sCount[32];
    if (blockIdx.x == 0 && threadIdx.x == 0) *prm.crossings = 0ull;   // (accumulated by the fill kernel)
    const uint32_t lane = laneId(), warp = threadIdx.x >> 5;
    const uint32_t tile = blockIdx.x * kWalkWarps + warp;
    if (threadIdx.x < 32) sCount[threadIdx.x] = 0xffffffffu;
    __syncthreads();
    uint32_t count = 0xffffffffu;
    if (tile < prm.numTiles)
    {
        uint32_t* list = prm.candList + (size_t)tile * prm.candCap;
        if (prm.numTris <= 1)
        {
            count = prm.numTris;
            if (lane == 0 && prm.numTris) list[0] = 0;
        }
        else
        {
            const uint32_t sy0 = (tile % prm.tilesY) * SY;
            const uint32_t sz0 = prm.z0 + (tile / prm.tilesY) * SZ;
            float rYmin, rYmax, rZmin, rZmax;
            tileRect<SY, SZ>(prm, sy0, sz0, rYmin, rYmax, rZmin, rZmax);

            uint32_t* stack = sStack[warp];
            uint32_t sp = 1, guard = 0;  // warp-uniform
            count = 0;
            if (lane == 0) stack[0] = 0;
            __syncwarp();
            const uint32_t lt = laneMaskLt();
            while (sp > 0)
            {
                // pop up to 32 nodes from the top; keep kStackGuard entries of head-room so that the
                // depth-first tail (k == 1) can never overflow: pushes <= 2k, depth <= 62
                const int room = kWalkStack - kStackGuard - (int)sp;
                const uint32_t k = room >= 1 ? min(min(32u, sp), (uint32_t)room) : 1u;
                sp -= k;
                uint32_t c0 = 0, c1 = 0;
                bool ov0 = false, ov1 = false;
                if (lane < k) testNode(prm.nodes, stack[sp + lane], rYmin, rYmax, rZmin, rZmax, ov0, ov1, c0, c1);
                __syncwarp();
                const bool in0 = ov0 && !(c0 & kLeafFlag), in1 = ov1 && !(c1 & kLeafFlag);
                const bool lf0 = ov0 && (c0 & kLeafFlag), lf1 = ov1 && (c1 & kLeafFlag);
                const uint32_t mi0 = __ballot_sync(0xffffffffu, in0), mi1 = __ballot_sync(0xffffffffu, in1);
                const uint32_t ml0 = __ballot_sync(0xffffffffu, lf0), ml1 = __ballot_sync(0xffffffffu, lf1);
                const uint32_t nIn = __popc(mi0) + __popc(mi1), nLf = __popc(ml0) + __popc(ml1);
                if (sp + nIn > (uint32_t)kWalkStack || ++guard > 2u * prm.numTris + 64u)
                {
                    if (lane == 0) atomicMax(prm.err, (uint32_t)kErrStackOverflow);
                    break;
                }
                if (count + nLf > prm.candCap) { count = prm.candCap + 1u; break; }  // overflow: the fill kernel walks itself
                if (in0) stack[sp + __popc(mi0 & lt)] = c0;
                if (in1) stack[sp + __popc(mi0) + __popc(mi1 & lt)] = c1;
                sp += nIn;
                if (lf0) list[count + __popc(ml0 & lt)] = c0 & ~kLeafFlag;
                if (lf1) list[count + __popc(ml0) + __popc(ml1 & lt)] = c1 & ~kLeafFlag;
                count += nLf;
                __syncwarp();
            }
        }
    }
    if (lane == 0) sCount[warp] = count;
    __syncthreads();
    if (warp == 0) fileTiles(prm, blockIdx.x * kWalkWarps, sCount[lane]);
#ifdef DXRV_TIMELINE
    {
        uint32_t mx = 0;
        for (int i = 0; i < kWalkWarps; ++i) if (sCount[i] != 0xffffffffu && sCount[i] > mx) mx = sCount[i];
        DXRV_TL_ROLE(mx == 0 ? 1 : mx < 192 ? 2 : mx < 768 ? 3 : 4);
    }
#endif
}

// ---- kernel A', the default: the candidate lists WITHOUT a tree walk -----------------------------------------------
// What the walk computes is, per super-tile, the set of triangles whose (y,z) box meets the tile's rectangle of
// column centres.  That is a 2-D binning problem, and the parallelism is in the TRIANGLES: one thread per sorted
// triangle turns its box into a (conservative, then exactly tested) range of tiles and appends its slot to their
// lists with one atomic each -- a few microseconds for 100 k triangles, no dependent chain of node fetches (the walk
// is bound by the depth of the LBVH: 28 levels x one L2 round trip), and no hierarchy to build first.  The candidate
// SET of a tile is the same as the walk's (same boxes, same exact compares), only its order differs, and XOR makes
// the order irrelevant.  Rectangles of many tiles (scene-sized triangles) are walked by the whole warp.
// k_file_columns then files every tile as heavy / light / empty exactly as k_walk_columns does.
template <int SY, int SZ>
__global__ void __launch_bounds__(256)
k_bin_columns(const ParityParams prm)
{
    extern __shared__ float sRect[];   // exact rectangles of the tiles (parity_bins.cuh)
    if (blockIdx.x == 0 && threadIdx.x == 0) *prm.crossings = 0ull;   // (accumulated by the fill kernel)
    BinParams bp;
    bp.N = prm.N; bp.z0 = prm.z0; bp.z1 = prm.z1; bp.tilesY = prm.tilesY; bp.candCap = prm.candCap; bp.invNPow2 = prm.invNPow2;
    bp.candCount = prm.candCount; bp.candList = prm.candList;
    binTablesSetup<SY, SZ>(bp, sRect);
    __syncthreads();
    const uint32_t rounded = (prm.numTris + 31u) & ~31u;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < rounded; j += gridDim.x * blockDim.x)
    {
        float4 a = make_float4(0, 0, 0, 0), b = a, c = a;
        const bool has = j < prm.numTris;
        if (has)
        {
            const float4* t = reinterpret_cast<const float4*>(prm.tris + j);
            a = __ldg(t); b = __ldg(t + 1); c = __ldg(t + 2);
        }
        binTriangleWarp<SY, SZ>(bp, sRect, has, a, b, c, j);
    }
}

__global__ void __launch_bounds__(32 * kWalkWarps)
k_file_columns(const ParityParams prm)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) *prm.crossings = 0ull;   // (again: k_bin_columns does not run when the lists were ready)
    const uint32_t firstTile = (blockIdx.x * kWalkWarps + (threadIdx.x >> 5)) * 32u;
    if (firstTile >= prm.numTiles) return;
    const uint32_t tile = firstTile + laneId();
    uint32_t count = 0xffffffffu;
    if (tile < prm.numTiles) count = min(prm.candCount[tile], prm.candCap + 1u);   // > candCap: the fill kernel scans every triangle
    fileTiles(prm, firstTile, count);
}

// ---- empty super-tiles: nothing to trace, 16 KB of zeros to write.  A few dedicated "writer" CTAs
// (the first blocks of the fill kernel) stream them with fire-and-forget 128-bit stores while the
// other CTAs of the same SMs rasterise: the write stream of the ~80 % of a real grid that is empty
// overlaps the issue-bound tracing instead of following it.  The writers take as many empty tiles as
// they can stream while the busy tiles are traced (kWriterBytesPerWork per busy work item); what is
// left -- everything, for a mostly empty grid -- is written one tile per CTA by the launch's surplus
// CTAs (there is one CTA per tile, and empty tiles need none), i.e. by the whole machine.
constexpr uint32_t kCounterWords = 32 + 1024;   // bucketCount[32] + one writer claim per SM (%smid < 1024)
#ifndef DXRV_FILL_WAVES
#define DXRV_FILL_WAVES 4
#endif
constexpr uint32_t kFillWaves = DXRV_FILL_WAVES;   // (1 and 2 measured: 43.5 / 37.2 us for the dragon, 54.5 / 48.7 us for the bowl against 37.4 / 42.7; 8: no change)               // CTAs of the fill kernel per resident CTA slot of the machine
constexpr uint32_t kWriterBatch = 4;             // empty tiles a writer warp takes per grab

// zero layers zFirst, zFirst + zStep, ... of an empty tile (one warp)
template <int SY, int SZ>
__device__ __forceinline__ void zeroTileLayers(const ParityParams& prm, uint32_t tile, uint32_t zFirst, uint32_t zStep)
{
    const uint32_t lane = laneId();
    const uint32_t N = prm.N, P = prm.P;
    const size_t layerWords = (size_t)N * P;
    const uint4 zero4 = make_uint4(0, 0, 0, 0);
    const uint32_t sy0 = (tile % prm.tilesY) * SY, sz0 = prm.z0 + (tile / prm.tilesY) * SZ;
    const uint32_t runWords = min((uint32_t)SY, N - sy0) * P;   // the tile's rows of one layer are contiguous
    const uint32_t nz = min((uint32_t)SZ, prm.z1 - sz0);
    uint32_t* run = prm.grid + ((size_t)(sz0 - prm.z0 + zFirst) * N + sy0) * P;
    if ((P & 3u) == 0u)
    {
        // stores issued back to back from independent address registers (a store whose address register
        // is still being read by the previous one would stall the warp)
        const uint32_t run4 = runWords >> 2;
        for (uint32_t z = zFirst; z < nz; z += zStep, run += zStep * layerWords)
        {
            uint4* q = reinterpret_cast<uint4*>(run) + lane;
            uint32_t i = 0;
            for (; i + 128u <= run4; i += 128u) { q[i] = zero4; q[i + 32u] = zero4; q[i + 64u] = zero4; q[i + 96u] = zero4; }
            for (i += lane; i < run4; i += 32u) reinterpret_cast<uint4*>(run)[i] = zero4;
        }
    }
    else
    {
        for (uint32_t z = zFirst; z < nz; z += zStep, run += zStep * layerWords)
            for (uint32_t i = lane; i < runWords; i += 32u) run[i] = 0u;
    }
}

// dedicated writer warp `writerWarp` of `numWriterWarps`: entries writerWarp + k * numWriterWarps of the first
// nDedicated entries of the empty list; the tile numbers are fetched 32 at a time, one per lane, a batch ahead
// The stores are TMA bulk copies out of a zeroed shared-memory buffer (cp.async.bulk, one per layer of a
// tile, issued by one lane each): a warp's ordinary stores top out at a few GB/s -- a handful of warps per SM
// cannot keep the SM's share of the HBM write bandwidth busy -- while the bulk-copy engine has no such limit.
template <int SY, int SZ>
__device__ __forceinline__ void writeEmptyTiles(const ParityParams& prm, uint32_t nDedicated,
                                                uint32_t* zeros /* shared, >= SY * P words, 16-byte aligned */)
{
    const uint32_t lane = laneId();
    const uint32_t N = prm.N, P = prm.P;
    const bool bulk = (P & 3u) == 0u && prm.bulkStores != 0u;
    if (bulk)
    {
        for (uint32_t i = threadIdx.x; i < ((uint32_t)SY * P) >> 2; i += blockDim.x) reinterpret_cast<uint4*>(zeros)[i] = make_uint4(0, 0, 0, 0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy zeros -> visible to the bulk-copy engine
        __syncthreads();
    }
    const uint32_t zerosAddr = (uint32_t)__cvta_generic_to_shared(zeros);
    // kWriterBatch tiles per grab of the cursor, one grab ahead; lane j < kWriterBatch holds tile j of the batch
    auto grab = [&]() -> uint32_t {
        uint32_t first = 0;
        if (lane == 0) first = atomicAdd(prm.bucketCount + 5, kWriterBatch);
        first = __shfl_sync(0xffffffffu, first, 0);
        return (lane < kWriterBatch && first < nDedicated && first + lane < nDedicated) ? __ldg(prm.emptyTiles + first + lane) : 0xffffffffu;
    };
    uint32_t next = grab();
    for (;;)
    {
        const uint32_t mine = next;
        if (__shfl_sync(0xffffffffu, mine, 0) == 0xffffffffu) break;
        next = grab();
        for (uint32_t j = 0; j < kWriterBatch; ++j)
        {
            const uint32_t tile = __shfl_sync(0xffffffffu, mine, j);
            if (tile == 0xffffffffu) break;
            if (!bulk) { zeroTileLayers<SY, SZ>(prm, tile, 0u, 1u); continue; }
            const uint32_t sy0 = (tile % prm.tilesY) * SY, sz0 = prm.z0 + (tile / prm.tilesY) * SZ;
            const uint32_t runBytes = min((uint32_t)SY, N - sy0) * P * 4u;   // the tile's rows of one layer are contiguous
            const uint32_t nz = min((uint32_t)SZ, prm.z1 - sz0);
            if (lane < nz)
            {
                uint32_t* dst = prm.grid + ((size_t)(sz0 - prm.z0 + lane) * N + sy0) * P;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(zerosAddr), "r"(runBytes) : "memory");
            }
        }
        if (bulk)
        {
            // one group per batch of tiles; keep a few batches in flight
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
        }
    }
    if (bulk) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the zeros must outlive the copies
}

// ---- kernel B: W warps per CTA on a super-tile of SY x SZ columns.  W = 4 up to N = 1024; larger grids
// have longer bit rows, i.e. more shared memory per tile and fewer CTAs per SM, and get more warps per
// CTA instead (the tracing is dealt to warps in chunks, the write-out in rows: neither cares) --------
// One work item (a light tile, or one part of a heavy tile): trace its candidates, then write its rows out.
template <int W, int SY, int SZ>
__device__ __forceinline__ void traceFillItem(const ParityParams& prm, uint32_t bIdx, uint32_t nHeavy, const uint4& nLight, uint32_t* smem)
{
    constexpr int kThreads = 32 * W;
    constexpr int kCols = SY * SZ;
    constexpr int kRowsPerWarp = kCols / W;
    constexpr int kStackCap = kStackPerThread * kThreads, kCandCap = kCandPerThread * kThreads;
    constexpr int kIdStage = 2 * kThreads;   // candidate slots staged in shared memory per item
    static_assert(kCols % W == 0 && SY == 16, "rows are dealt to the warps in equal runs");
    __shared__ uint32_t sTop[2];   // fallback walk: stack height, double-buffered by iteration parity
    __shared__ uint32_t sCand;     // fallback walk: leaves queued so far
    __shared__ uint32_t sNext;     // listed candidates: next chunk to hand out

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t N = prm.N, P = prm.P, Ps = prm.Ps;
    const float fN = (float)N, invNPow2 = prm.invNPow2;

    // A crossing whose first inside voxel is ix flips every voxel >= ix of its column.  Inside the word of
    // ix that is one XOR with a suffix mask.  Every suffix mask holds bit 31, so bit 31 of a word of the row
    // is the parity of the crossings inside that word, and the write-out flips word w when an odd number of
    // the row's words before w have bit 31 set -- no second structure to keep.  (A single toggle bit per
    // crossing + a 128-bit prefix-XOR at write-out costs four times the ALU work.)
    uint32_t* rows = smem;                                     // [kCols][Ps] occupancy bits, column = zl*SY + yl
    float* tileY = reinterpret_cast<float*>(rows + (uint32_t)kCols * Ps);   // [SY] scene Y of the columns (decreasing)
    float* tileZ = tileY + SY;                                 // [SZ]
    uint32_t* stack = reinterpret_cast<uint32_t*>(tileZ + SZ); // [kStackCap]  (fallback walk only)
    uint32_t* cand = stack + kStackCap;                        // [kCandCap]   (fallback walk only)
    float4* stage = reinterpret_cast<float4*>(stack);          // [W][32 x 3]  triangle records of the listed-candidates path
    uint32_t* ids = cand + kCandCap;                           // [kIdStage]   first candidate slots of the item
    static_assert((kStackPerThread + kCandPerThread) * 8 >= kStagePerWarp && kCandPerThread >= 3 && kStackPerThread >= 4, "staging area must fit the fallback walk's buffers");

    // work items are numbered heavy parts first, then the light tiles class by class (see fileTiles)
    __shared__ uint32_t sIsLast;
    uint32_t tile, listed, part = 0, parts = 1, hslot = 0xffffu;
    if (bIdx < nHeavy)
    {
        const uint2 e = __ldg(prm.heavyEntries + bIdx);
        tile = e.x; part = e.y & 0xffu; parts = (e.y >> 8) & 0xffu; hslot = e.y >> 16;
        listed = __ldg(prm.candCount + tile);
    }
    else
    {
        bIdx -= nHeavy;
        uint32_t cls = 0;
        if (bIdx >= nLight.x) { bIdx -= nLight.x; cls = 1; if (bIdx >= nLight.y) { bIdx -= nLight.y; cls = 2; if (bIdx >= nLight.z) { bIdx -= nLight.z; cls = 3; } } }
        const uint32_t e = __ldg(prm.lightTiles + (size_t)cls * prm.tilesPad + bIdx);   // the count rides along: one dependent load less
        tile = e & 0xffffffu; listed = e >> 24;
    }
    const uint32_t sy0 = (tile % prm.tilesY) * SY;
    const uint32_t sz0 = prm.z0 + (tile / prm.tilesY) * SZ;
    const uint32_t yLast = min(sy0 + SY - 1, N - 1) - sy0, zLast = min(sz0 + SZ - 1, prm.z1 - 1) - sz0;
    DXRV_TL_STAMP(0);
    // this CTA's share of the tile's candidate list (a split tile has several parts)
    const uint32_t partBegin = (uint32_t)(((uint64_t)min(listed, prm.candCap) * part) / parts);
    const uint32_t mine = listed <= prm.candCap ? (uint32_t)(((uint64_t)listed * (part + 1u)) / parts) - partBegin : 0u;
    const uint32_t* list = prm.candList + (size_t)tile * prm.candCap + partBegin;
    // The first kIdStage candidate slots go to shared memory and their records are PREFETCHED into L1 while the rows
    // are being zeroed: the chunk loop below then neither waits for the list nor (mostly) for the records -- those two
    // dependent L2 round trips per chunk were the largest single stall of the kernel.
    uint32_t stagedId[kIdStage / kThreads];
#pragma unroll
    for (uint32_t k = 0; k < (uint32_t)(kIdStage / kThreads); ++k)
        stagedId[k] = (k * kThreads + tid < mine) ? __ldg(list + k * kThreads + tid) : 0xffffffffu;

    uint32_t myCrossings = 0;
    {
        for (uint32_t i = tid; i < ((uint32_t)kCols * Ps) >> 2; i += kThreads) reinterpret_cast<uint4*>(rows)[i] = make_uint4(0, 0, 0, 0);
        if (tid < SY) tileY[tid] = (sy0 + tid < N) ? -centreOf(sy0 + tid, fN, invNPow2) : INFINITY;
        if (tid >= 32 && tid < 32 + SZ) tileZ[tid - 32] = (sz0 + tid - 32 < prm.z1) ? centreOf(sz0 + tid - 32, fN, invNPow2) : INFINITY;
        if (tid == 0) { stack[0] = 0; sTop[0] = 1u; sTop[1] = 0; sCand = 0; sNext = 0; }
#pragma unroll
        for (uint32_t k = 0; k < (uint32_t)(kIdStage / kThreads); ++k)
        {
            ids[k * kThreads + tid] = stagedId[k];
            if (stagedId[k] != 0xffffffffu)
            {
                const char* rec = reinterpret_cast<const char*>(prm.tris + stagedId[k]);
                asm volatile("prefetch.global.L1 [%0];" ::"l"(rec));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(rec + 32));
            }
        }
        __syncthreads();
        DXRV_TL_STAMP(1);
        const float halfN = 0.5f * fN;

        // Up to 32 candidate triangles per call (one per lane; `has` marks real ones), processed by the
        // whole warp: every lane first finds the EXACT rectangle of columns whose centres lie in its
        // triangle's (y,z) box, a warp scan turns the rectangle sizes into one flat list of
        // (triangle, column) pairs, and the pairs are dealt round-robin to the 32 lanes -- a triangle
        // that spans many columns no longer serialises on one thread while the CTA waits at the barrier.
        // columnRange: rectangle of lane's triangle -> yA | zA << 8 | w << 16 and the pair count w * h.
        auto columnRange = [&](const float4& a, const float4& b, const float4& c, uint32_t& packed, uint32_t& n) {
            const float ylo = fminf(fminf(a.y, b.y), c.y), yhi = fmaxf(fmaxf(a.y, b.y), c.y);
            const float zlo = fminf(fminf(a.z, b.z), c.z), zhi = fmaxf(fmaxf(a.z, b.z), c.z);
            // float estimate of the index range, then fix-up against the tabulated exact centres
            // (tileY decreases with yl, tileZ increases with zl)
            int yA, zA, yB, zB;
            yA = min(max((int)floorf((1.0f - yhi) * halfN - 0.5f) - (int)sy0, 0), (int)yLast + 1);
            while (yA > 0 && tileY[yA - 1] <= yhi) --yA;
            while (yA <= (int)yLast && tileY[yA] > yhi) ++yA;
            yB = min(max((int)ceilf((1.0f - ylo) * halfN - 0.5f) - (int)sy0, -1), (int)yLast);
            while (yB < (int)yLast && tileY[yB + 1] >= ylo) ++yB;
            while (yB >= 0 && tileY[yB] < ylo) --yB;
            zA = min(max((int)floorf((zlo + 1.0f) * halfN - 0.5f) - (int)sz0, 0), (int)zLast + 1);
            while (zA > 0 && tileZ[zA - 1] >= zlo) --zA;
            while (zA <= (int)zLast && tileZ[zA] < zlo) ++zA;
            zB = min(max((int)ceilf((zhi + 1.0f) * halfN - 0.5f) - (int)sz0, -1), (int)zLast);
            while (zB < (int)zLast && tileZ[zB + 1] <= zhi) ++zB;
            while (zB >= 0 && tileZ[zB] > zhi) --zB;
            const int w = max(yB - yA + 1, 0), h = max(zB - zA + 1, 0);
            n = (uint32_t)(w * h);
            packed = (uint32_t)yA | ((uint32_t)zA << 8) | ((uint32_t)w << 16);
        };
        // one (triangle, column) pair: q-th column of the rectangle `packed`; inv = ceil(2^16 / w), so
        // that q / w == (q * inv) >> 16 exactly (q < 256, w <= 16)
        auto testPairAt = [&](const float4& ta, const float4& tb, const float4& tc, uint32_t yl, uint32_t zl) {
            uint32_t ix;
            if (columnCrossing(ta, tb, tc, tileY[yl], tileZ[zl], N, fN, invNPow2, ix))
            {
                ++myCrossings;
                if (ix < N)
                {
                    const uint32_t col = zl * SY + yl, w = ix >> 5;
                    atomicXor(&rows[col * Ps + w], 0xffffffffu << (ix & 31u));
                }
            }
        };
        auto testPair = [&](const float4& ta, const float4& tb, const float4& tc, uint32_t packed, uint32_t inv, uint32_t q) {
            const uint32_t ow = packed >> 16;
            const uint32_t qz = (q * inv) >> 16;
            testPairAt(ta, tb, tc, (packed & 0xffu) + (q - qz * ow), ((packed >> 8) & 0xffu) + qz);
        };

        // Staged variant (the normal path).  Work is flattened twice, so that all 32 lanes stay busy whatever
        // the triangles' shapes: triangles -> ROW UNITS (one z row of one triangle's columns) -> PAIRS
        // (row unit, column).  A row unit gets a conservative y interval of the triangle inside its row
        // (rowInterval), so only about one pair in ten is tested in vain -- the bounding rectangle of a
        // triangle holds three times more columns than the triangle crosses.  Triangles with rows are
        // compacted into a per-warp shared-memory table of 48-byte records {a.xyz, zA | h << 8 | b.xyz,
        // first row unit | c.xyz}, the row units of the current window into {record | zl << 8 | ya << 16,
        // first pair}.  The record that owns item p is found without a search: records start at increasing
        // item numbers, so with `starts` = bit mask of the records starting inside the current window of
        // 32 items (one warp-wide OR),  owner(p) = #records started before the window + popc(starts up to
        // p's lane) - 1.
        auto processWarpChunkStaged = [&](bool has, uint32_t slot) {
            float4* tab = stage + warp * (uint32_t)kStagePerWarp;          // kChunkMax records x 3 float4 ...
            uint2* units = reinterpret_cast<uint2*>(tab + 3 * kChunkMax);            // ... and 32 row units
            const uint32_t lt = laneMaskLt(), le = lt | (1u << lane);
            float4 a = make_float4(0, 0, 0, 0), b = a, c = a;
            uint32_t zA = 0, h = 0;
            if (has)
            {
                const float4* t = reinterpret_cast<const float4*>(prm.tris + slot);
                a = __ldg(t); b = __ldg(t + 1); c = __ldg(t + 2);
                // rows whose centre may lie in [zlo, zhi] (tileZ increases with zl)
                const float zlo = fminf(fminf(a.z, b.z), c.z), zhi = fmaxf(fmaxf(a.z, b.z), c.z);
                const int z0i = min(max((int)ceilf((zlo + 1.0f) * halfN - 0.5f - kIdxSlack) - (int)sz0, 0), (int)zLast + 1);
                const int z1i = min(max((int)floorf((zhi + 1.0f) * halfN - 0.5f + kIdxSlack) - (int)sz0, -1), (int)zLast);
                zA = (uint32_t)z0i; h = (uint32_t)max(z1i - z0i + 1, 0);
            }
            uint32_t incl = h;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= (uint32_t)o) incl += v;
            }
            const uint32_t totalRows = __shfl_sync(0xffffffffu, incl, 31);
            const uint32_t nzA = __ballot_sync(0xffffffffu, h != 0u);
            if (h != 0u)
            {
                float4* r = tab + __popc(nzA & lt) * 3u;
                a.w = __uint_as_float(zA | (h << 8)); b.w = __uint_as_float(incl - h);
                r[0] = a; r[1] = b; r[2] = c;
            }
            __syncwarp();
            const uint32_t myStartA = lane < (uint32_t)__popc(nzA) ? __float_as_uint(tab[lane * 3u + 1u].w) : 0xffffffffu;
            uint32_t beforeA = 0;   // records that start before the current window of row units
#pragma unroll 1
            for (uint32_t r0 = 0; r0 < totalRows; r0 += 32u)
            {
                const uint32_t relA = myStartA - r0;   // < 32 iff my record starts inside this window
                const uint32_t startsA = __reduce_or_sync(0xffffffffu, relA < 32u ? 1u << relA : 0u);
                const uint32_t ownerA = beforeA + __popc(startsA & le) - 1u;
                beforeA += __popc(startsA);
                uint32_t cnt = 0, unit = 0;
                if (r0 + lane < totalRows)
                {
                    const float4* r = tab + ownerA * 3u;
                    const float4 ta = r[0], tb = r[1], tc = r[2];
                    const uint32_t zl = (__float_as_uint(ta.w) & 0xffu) + (r0 + lane - __float_as_uint(tb.w));
                    float lo, hi;
                    rowIntervalY(ta, tb, tc, tileZ[zl], lo, hi);
                    // columns whose centre may lie in [lo, hi] (tileY decreases with yl)
                    const int ya = min(max((int)ceilf((1.0f - hi) * halfN - 0.5f - kIdxSlack) - (int)sy0, 0), (int)yLast + 1);
                    const int yb = min(max((int)floorf((1.0f - lo) * halfN - 0.5f + kIdxSlack) - (int)sy0, -1), (int)yLast);
                    cnt = (uint32_t)max(yb - ya + 1, 0);
                    unit = ownerA | (zl << 8) | ((uint32_t)ya << 16);
                }
                uint32_t inclB = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1)
                {
                    const uint32_t v = __shfl_up_sync(0xffffffffu, inclB, o);
                    if (lane >= (uint32_t)o) inclB += v;
                }
                const uint32_t totalPairs = __shfl_sync(0xffffffffu, inclB, 31);
                const uint32_t nzB = __ballot_sync(0xffffffffu, cnt != 0u);
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
                        const float4* r = tab + (u.x & 0xffu) * 3u;
                        const float4 ta = r[0], tb = r[1], tc = r[2];
                        testPairAt(ta, tb, tc, (u.x >> 16) + (p - u.y), (u.x >> 8) & 0xffu);
                    }
                }
                __syncwarp();   // the row units are rewritten by the next window
            }
            __syncwarp();       // the table is rewritten by the next chunk
        };

        // Direct variant (fallback walk only, where the staging area holds the walk's stack): records stay
        // in the registers of the lanes that loaded them and are fetched with shuffles.
        auto processWarpChunk = [&](bool has, uint32_t slot, uint32_t chunk) {
            uint32_t packed = 0, n = 0;
            if (has)
            {
                const float4* t = reinterpret_cast<const float4*>(prm.tris + slot);
                columnRange(__ldg(t), __ldg(t + 1), __ldg(t + 2), packed, n);
            }
            const uint32_t ow0 = packed >> 16;
            const uint32_t inv = ow0 > 0 ? (65535u + ow0) / ow0 : 0u;
            uint32_t incl = n;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= (uint32_t)o) incl += v;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            const uint32_t excl = incl - n;
#pragma unroll 1
            for (uint32_t p0 = 0; p0 < total; p0 += 32u)
            {
                const uint32_t p = p0 + lane;
                // owner = the last lane whose exclusive offset is <= p (empty triangles share their offset
                // with the next lane, so the last one is the one that really owns pair p)
                uint32_t owner = 0;
#pragma unroll 1
                for (uint32_t step = chunk >> 1; step > 0; step >>= 1)   // triangles live in lanes [0, chunk)
                {
                    const uint32_t e = __shfl_sync(0xffffffffu, excl, owner + step);
                    if (e <= p) owner += step;
                }
                const uint32_t q = p - __shfl_sync(0xffffffffu, excl, owner);
                const uint32_t oPacked = __shfl_sync(0xffffffffu, packed, owner);
                const uint32_t oInv = __shfl_sync(0xffffffffu, inv, owner);
                const uint32_t oSlot = __shfl_sync(0xffffffffu, slot, owner);
                if (p < total)
                {
                    // the owner's triangle was just loaded by the owner lane: these hit L1
                    const float4* t = reinterpret_cast<const float4*>(prm.tris + oSlot);
                    testPair(__ldg(t), __ldg(t + 1), __ldg(t + 2), oPacked, oInv, q);
                }
            }
        };

        if (listed <= prm.candCap)
        {
            // ---- candidates were listed by k_walk_columns: dealt to the warps in chunks of C (small
            // chunks when there are few candidates, so that all W warps get some) ----
            uint32_t C = (uint32_t)kChunkMax;
            while (C > 4u && mine * DXRV_CHUNK_DEN < C * (uint32_t)W * DXRV_CHUNK_NUM) C >>= 1;
            for (;;)   // the warps take chunks as they become free: pair counts per chunk vary a lot
            {
                uint32_t first = 0;
                if (lane == 0) first = atomicAdd(&sNext, C);
                first = __shfl_sync(0xffffffffu, first, 0);
                if (first >= mine) break;
                const bool has = lane < C && first + lane < mine;
                processWarpChunkStaged(has, has ? (first + lane < (uint32_t)kIdStage ? ids[first + lane] : __ldg(list + first + lane)) : 0u);
            }
        }
        else if (prm.nodes == nullptr)
        {
            // ---- overflowed list of a BINNED tile (no hierarchy was built): scan every triangle's box; the ones
            // that meet the tile go through the same ring of queued candidates as the walk below ----
            float rYmin, rYmax, rZmin, rZmax;
            tileRect<SY, SZ>(prm, sy0, sz0, rYmin, rYmax, rZmin, rZmax);
            uint32_t consumed = 0;
            const uint32_t lt = laneMaskLt();
            for (uint32_t base = 0; base < prm.numTris; base += (uint32_t)kThreads)
            {
                const uint32_t j = base + tid;
                bool ov = false;
                if (j < prm.numTris)
                {
                    const float4* t = reinterpret_cast<const float4*>(prm.tris + j);
                    const float4 a = __ldg(t), b = __ldg(t + 1), c = __ldg(t + 2);
                    ov = fminf(fminf(a.y, b.y), c.y) <= rYmax && fmaxf(fmaxf(a.y, b.y), c.y) >= rYmin &&
                         fminf(fminf(a.z, b.z), c.z) <= rZmax && fmaxf(fmaxf(a.z, b.z), c.z) >= rZmin;
                }
                const uint32_t m = __ballot_sync(0xffffffffu, ov);
                uint32_t off = 0;
                if (lane == 0 && m) off = atomicAdd(&sCand, (uint32_t)__popc(m));
                off = __shfl_sync(0xffffffffu, off, 0);
                if (ov) cand[(off + __popc(m & lt)) % kCandCap] = j;
                __syncthreads();
                const uint32_t produced = sCand;
                while (produced - consumed >= (uint32_t)kThreads)
                {
                    processWarpChunk(true, cand[(consumed + tid) % kCandCap], 32u);
                    consumed += kThreads;
                }
                __syncthreads();   // the drained ring entries may be overwritten by the next round
            }
            for (uint32_t produced = sCand, i0 = consumed + warp * 32u; i0 < produced; i0 += kThreads)
            {
                const bool has = i0 + lane < produced;
                processWarpChunk(has, has ? cand[(i0 + lane) % kCandCap] : 0u, 32u);
            }
        }
        else
        {
            // ---- fallback: CTA-cooperative walk (every thread pops a different node) ----
            float rYmin, rYmax, rZmin, rZmax;
            tileRect<SY, SZ>(prm, sy0, sz0, rYmin, rYmax, rZmin, rZmax);
            uint32_t parity = 0, guard = 0;
            uint32_t consumed = 0;  // ring entries already processed (uniform); sCand counts entries produced
            while (true)
            {
                __syncthreads();  // S1: pushes / counters of the previous iteration are visible
                const uint32_t sp = sTop[parity];
                const uint32_t produced = sCand;
                if (sp == 0) break;
                if (++guard > 2u * prm.numTris + 64u)
                {
                    if (tid == 0) atomicMax(prm.err, (uint32_t)kErrStackOverflow);
                    break;
                }
                // drain full batches of queued triangles first: at most kThreads-1 stay queued, so the
                // 2*kThreads pushes of this iteration always fit the ring (kCandCap >= 3*kThreads)
                while (produced - consumed >= (uint32_t)kThreads)
                {
                    processWarpChunk(true, cand[(consumed + tid) % kCandCap], 32u);
                    consumed += kThreads;
                }
                const int room = kStackCap - kStackGuard - (int)sp;
                const uint32_t k = room >= 1 ? min(min((uint32_t)kThreads, sp), (uint32_t)room) : 1u;
                const uint32_t base = sp - k;
                uint32_t c0 = 0, c1 = 0;
                bool ov0 = false, ov1 = false;
                if (tid < k) testNode(prm.nodes, stack[base + tid], rYmin, rYmax, rZmin, rZmax, ov0, ov1, c0, c1);
                if (tid == 0) sTop[parity ^ 1u] = base;   // nobody reads this word before S2
                __syncthreads();  // S2: every thread has read its stack / ring entries and both counters
                if (warp * 32u < k)
                {
                    const bool in0 = ov0 && !(c0 & kLeafFlag), in1 = ov1 && !(c1 & kLeafFlag);
                    const bool lf0 = ov0 && (c0 & kLeafFlag), lf1 = ov1 && (c1 & kLeafFlag);
                    const uint32_t mi0 = __ballot_sync(0xffffffffu, in0), mi1 = __ballot_sync(0xffffffffu, in1);
                    const uint32_t ml0 = __ballot_sync(0xffffffffu, lf0), ml1 = __ballot_sync(0xffffffffu, lf1);
                    const uint32_t lt = laneMaskLt();
                    const uint32_t nIn = __popc(mi0) + __popc(mi1), nLf = __popc(ml0) + __popc(ml1);
                    uint32_t offIn = 0, offLf = 0;
                    if (lane == 0)
                    {
                        if (nIn) offIn = atomicAdd(&sTop[parity ^ 1u], nIn);
                        if (nLf) offLf = atomicAdd(&sCand, nLf);
                    }
                    offIn = __shfl_sync(0xffffffffu, offIn, 0);
                    offLf = __shfl_sync(0xffffffffu, offLf, 0);
                    if (offIn + nIn > (uint32_t)kStackCap)
                    {
                        if (lane == 0) atomicMax(prm.err, (uint32_t)kErrStackOverflow);  // never expected; drop
                    }
                    else
                    {
                        if (in0) stack[offIn + __popc(mi0 & lt)] = c0;
                        if (in1) stack[offIn + __popc(mi0) + __popc(mi1 & lt)] = c1;
                    }
                    if (lf0) cand[(offLf + __popc(ml0 & lt)) % kCandCap] = c0 & ~kLeafFlag;
                    if (lf1) cand[(offLf + __popc(ml0) + __popc(ml1 & lt)) % kCandCap] = c1 & ~kLeafFlag;
                }
                parity ^= 1u;
            }
            __syncthreads();
            for (uint32_t produced = sCand, i0 = consumed + warp * 32u; i0 < produced; i0 += kThreads)
            {
                const bool has = i0 + lane < produced;
                processWarpChunk(has, has ? cand[(i0 + lane) % kCandCap] : 0u, 32u);
            }
        }
        __syncthreads();
        DXRV_TL_STAMP(2);

        if (parts > 1u)
        {
            // ---- split tile: merge this part's toggles into the tile's scratch rows; the last part to
            // arrive takes the merged rows back and carries on to the fill, the others are done ----
            uint32_t* scratch = prm.heavyScratch + (size_t)hslot * kCols * Ps;
            for (uint32_t i = tid; i < (uint32_t)kCols * Ps; i += kThreads)
            {
                const uint32_t v = rows[i];
                if (v) atomicXor(scratch + i, v);
            }
            __threadfence();
            __syncthreads();
            if (tid == 0) sIsLast = (atomicAdd(prm.heavyArrive + hslot, 1u) == parts - 1u) ? 1u : 0u;
            __syncthreads();
            if (!sIsLast)
            {
                for (int o = 16; o > 0; o >>= 1) myCrossings += __shfl_xor_sync(0xffffffffu, myCrossings, o);
                if (lane == 0 && myCrossings) atomicAdd(prm.crossings, (unsigned long long)myCrossings);
                return;
            }
            __threadfence();
            for (uint32_t i = tid; i < ((uint32_t)kCols * Ps) >> 2; i += kThreads)
            {
                uint4* src = reinterpret_cast<uint4*>(scratch) + i;
                reinterpret_cast<uint4*>(rows)[i] = __ldcg(src);
                *src = make_uint4(0, 0, 0, 0);          // leave the scratch clean for the next launch
            }
            if (tid == 0) prm.heavyArrive[hslot] = 0;
            __syncthreads();
        }
    }

    // ---- write-out, 4 words (128 bits) per lane: flip the words behind each crossing's word ----
    // warp w owns shared rows [32w, 32w+32): for every z of the super-tile SY consecutive y rows,
    // which are contiguous in the global grid.
    const uint32_t groupsPerRow = Ps >> 2;            // 128-bit groups per shared row
    const uint32_t tailMask = (N & 31u) ? ((1u << (N & 31u)) - 1u) : 0xffffffffu;
    const uint4* rows4 = reinterpret_cast<const uint4*>(rows) + (size_t)warp * kRowsPerWarp * groupsPerRow;
    const uint32_t warpGroups = (uint32_t)kRowsPerWarp * groupsPerRow;   // a multiple of 32 for every (W, Ps) launched

    // occupancy bits of the group this lane holds (t = four consecutive words of a row, as traced).  The groups of a
    // row sit in consecutive lanes (groupsPerRow <= 32: rows aligned to groupsPerRow lanes) or in consecutive
    // iterations (longer rows: `carry` = parity of the crossings in the row's earlier iterations, warp-uniform).
    const uint32_t lt = laneMaskLt();
    const uint32_t rowBelow = groupsPerRow <= 32u ? lt & ~((1u << (lane & ~(groupsPerRow - 1u))) - 1u) : lt;
    uint32_t carry = 0;
    auto occupancy = [&](uint4 t, uint32_t g0) -> uint4 {
        const uint32_t top = __ballot_sync(0xffffffffu, (int)t.x < 0) ^ __ballot_sync(0xffffffffu, (int)t.y < 0) ^
                             __ballot_sync(0xffffffffu, (int)t.z < 0) ^ __ballot_sync(0xffffffffu, (int)t.w < 0);
        uint32_t before = __popc(top & rowBelow);
        if (groupsPerRow > 32u)
        {
            if (g0 % groupsPerRow == 0u) carry = 0;
            before += carry;
            carry ^= __popc(top) & 1u;
        }
        const uint32_t mx = 0u - (before & 1u);
        const uint32_t my = mx ^ (uint32_t)((int)t.x >> 31), mz = my ^ (uint32_t)((int)t.y >> 31), mw = mz ^ (uint32_t)((int)t.z >> 31);
        t.x ^= mx; t.y ^= my; t.z ^= mz; t.w ^= mw;
        return t;
    };

    const bool interior = SY == 16 && (N & 127u) == 0u && groupsPerRow <= 32u && sy0 + SY <= N && sz0 + SZ <= prm.z1;
    if (interior)
    {
        // fast path (every super-tile of a grid with N % 128 == 0): the warp's rows are runs of up to 16
        // consecutive y rows of one z layer, each run contiguous in the grid.  The shared row pitch
        // (Ps, a power of two) may exceed the global one (P): the padding groups are skipped.
        uint4* base = reinterpret_cast<uint4*>(prm.grid + ((size_t)(sz0 - prm.z0) * N + sy0) * P);
        const uint32_t zStride = (uint32_t)(((size_t)N * P) >> 2);
        const uint32_t gprG = P >> 2;
        for (uint32_t g0 = 0; g0 < warpGroups; g0 += 32u)
        {
            const uint32_t g = g0 + lane;
            const uint32_t rr = g >> prm.gprShift, gi = g & (groupsPerRow - 1u);
            const uint32_t col = warp * (uint32_t)kRowsPerWarp + rr;
            const uint4 t = occupancy(rows4[g], g0);
            if (gi < gprG) base[(col >> 4) * zStride + (col & 15u) * gprG + gi] = t;
        }
    }
    else
    {
        for (uint32_t g0 = 0; g0 < warpGroups; g0 += 32u)
        {
            const uint32_t g = g0 + lane;
            uint32_t rowInWarp, gi;
            if (groupsPerRow <= 32u) { rowInWarp = g >> prm.gprShift; gi = g & (groupsPerRow - 1u); }
            else { rowInWarp = g0 / groupsPerRow; gi = g - rowInWarp * groupsPerRow; }
            uint4 t = occupancy(rows4[g], g0);

            const uint32_t col = warp * (uint32_t)kRowsPerWarp + rowInWarp;
            const uint32_t yl = col % SY, zl = col / SY;
            const uint32_t y = sy0 + yl, z = sz0 + zl;
            const uint32_t w0 = gi * 4u;
            if (y < N && z < prm.z1 && w0 < P)
            {
                uint32_t* dst = prm.grid + ((size_t)(z - prm.z0) * N + y) * P + w0;
                if (w0 + 4u <= P && (P & 3u) == 0u)
                {
                    if (w0 + 4u == P) t.w &= tailMask;
                    *reinterpret_cast<uint4*>(dst) = t;
                }
                else
                {
                    const uint32_t v[4] = {t.x, t.y, t.z, t.w};
                    for (uint32_t q = 0; q < 4u && w0 + q < P; ++q)
                        dst[q] = (w0 + q == P - 1u) ? (v[q] & tailMask) : v[q];
                }
            }
        }
    }

    // ---- statistics ----
    for (int o = 16; o > 0; o >>= 1) myCrossings += __shfl_xor_sync(0xffffffffu, myCrossings, o);
    if (lane == 0 && myCrossings) atomicAdd(prm.crossings, (unsigned long long)myCrossings);
}

// Order in which the launch's CTAs take the work items (numbered heavy parts first, then the light tiles by
// decreasing class, see fileTiles).  CTAs start in the order of their numbers, so this is the order in time.
#ifndef DXRV_ITEM_ORDER
#define DXRV_ITEM_ORDER 0
#endif
#ifndef DXRV_DYNAMIC_ITEMS
#define DXRV_DYNAMIC_ITEMS 0
#endif
struct ItemOrder
{
    uint32_t nWork, nHeavy, K, n;
    // smallest stride >= 0.618 n that is coprime to n: i -> i * K mod n visits the items evenly spread over the classes
    static __device__ __forceinline__ uint32_t goldenStride(uint32_t n)
    {
        if (n < 3u || n > 65535u) return 1u;                     // (i * K stays below 2^32; larger launches keep their order)
        uint32_t k = (n * 40503u) >> 16;                          // 0.618 n
        for (;; ++k)
        {
            uint32_t a = n, b = k % n;
            while (b) { const uint32_t t = a % b; a = b; b = t; }
            if (a == 1u) return k % n;
        }
    }
    __device__ __forceinline__ ItemOrder(uint32_t nWork_, uint32_t nHeavy_) : nWork(nWork_), nHeavy(nHeavy_), K(1u), n(nWork_)
    {
        if (DXRV_ITEM_ORDER == 2) K = goldenStride(n);
        if (DXRV_ITEM_ORDER == 4) { n = nWork - nHeavy; K = goldenStride(n); }
    }
    __device__ __forceinline__ uint32_t operator()(uint32_t i) const
    {
        if (DXRV_ITEM_ORDER == 1) return nWork - 1u - i;                                   // smallest first
        if (DXRV_ITEM_ORDER == 2) return (i * K) % n;                                // evenly mixed
        if (DXRV_ITEM_ORDER == 3) return i < nHeavy ? i : nWork - 1u - (i - nHeavy);       // heavy parts, then smallest first
        if (DXRV_ITEM_ORDER == 4) return i < nHeavy ? i : nHeavy + ((i - nHeavy) * K) % n;   // heavy parts, then mixed
        return i;                                                                          // largest first
    }
};

// The first CTAs of the launch write the empty tiles (one writer per SM); the others take the work items
// in order -- heavy parts first -- and, when there are more items than CTAs, further ones at a stride.
template <int W, int SY, int SZ>
__global__ void __launch_bounds__(32 * W, W == 4 ? DXRV_FILL_CTAS : W == 8 ? 4 : 2)
k_trace_fill_columns(const ParityParams prm)
{
    DXRV_TL_SCOPE();
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ uint32_t sDuplicate;
    const uint32_t nHeavy = __ldg(prm.bucketCount), nEmpty = __ldg(prm.bucketCount + 2);
    const uint4 nLight = __ldg(reinterpret_cast<const uint4*>(prm.bucketCount + 8));
    const uint32_t nWork = nHeavy + nLight.x + nLight.y + nLight.z + nLight.w;
    if (blockIdx.x < prm.numWriters)
    {
        // one writer per SM: an SM's write bandwidth is its share of the machine's, whoever issues the
        // stores, so a second writer CTA on the same SM would only take a CTA slot from the tracing.
        // (Twice as many writers as SMs are launched so that nearly every SM gets one.)
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        if (threadIdx.x == 0) sDuplicate = atomicExch(prm.bucketCount + 32 + (smid & 1023u), 1u);
        __syncthreads();
        if (sDuplicate != 0u) return;
        DXRV_TL_ROLE(1);
#ifdef DXRV_EXPERIMENT
        if (prm.bulkStores == 2u) return;   // (timing experiment: trace alone, the empty tiles stay unwritten)
#endif
        writeEmptyTiles<SY, SZ>(prm, nEmpty, smem);
        return;
    }
    const uint32_t stride = gridDim.x - prm.numWriters;
    uint32_t item = blockIdx.x - prm.numWriters;
    if (item < nWork) DXRV_TL_ROLE(item < nHeavy ? 2 : 3);
#ifdef DXRV_EXPERIMENT
    if (prm.noTrace) return;   // (timing experiment: the writers alone)
#endif
    const ItemOrder order(nWork, nHeavy);
#if DXRV_DYNAMIC_ITEMS
    // one CTA per resident slot; the first item is the CTA's own number, further ones come from a counter that is
    // read one item ahead (the atomic's round trip hides behind the item being traced)
    __shared__ uint32_t sItem;
    for (bool first = true; item < nWork; first = false)
    {
        if (!first) __syncthreads();   // the previous item's write-out has read the shared rows (and every thread sItem)
        if (threadIdx.x == 0) sItem = stride + atomicAdd(prm.bucketCount + 16, 1u);
        traceFillItem<W, SY, SZ>(prm, order(item), nHeavy, nLight, smem);
        __syncthreads();
        item = sItem;
    }
#else
    for (bool first = true; item < nWork; item += stride, first = false)
    {
        if (!first) __syncthreads();   // the previous item's write-out has read the shared rows
        traceFillItem<W, SY, SZ>(prm, order(item), nHeavy, nLight, smem);
    }
#endif
}

uint32_t sharedRowWords(uint32_t P)
{
    if (P <= 4) return 4;
    if (P <= 128)
    {
        uint32_t v = 4;
        while (v < P) v <<= 1;
        return v;
    }
    return (P + 127u) / 128u * 128u;
}

template <int W, int SY, int SZ>
void launchVariant(cudaStream_t s, ParityParams prm, cudaEvent_t* ev, bool binsReady)
{
    const size_t smemBytes = sizeof(uint32_t) * ((size_t)SY * SZ * prm.Ps + SY + SZ + (size_t)(kStackPerThread + kCandPerThread + 2) * 32 * W);
    static bool attrSet[64] = {};
    static int smCount[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attrSet[dev])
    {
        cudaFuncSetAttribute(k_trace_fill_columns<W, SY, SZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 64);
        cudaDeviceGetAttribute(&smCount[dev], cudaDevAttrMultiProcessorCount, dev);
        attrSet[dev] = true;
    }
    prm.numWriters = (dev >= 0 && dev < 64 && smCount[dev] > 0) ? 2u * (uint32_t)smCount[dev] : 296u;
    static const bool noBulk = [] { const char* e = std::getenv("DXRV_NO_BULK_STORE"); return e && e[0] && e[0] != '0'; }();
    prm.bulkStores = noBulk ? 0u : 1u;
#ifdef DXRV_EXPERIMENT
    if (const char* e = std::getenv("DXRV_EXP_NO_WRITERS")) if (e[0] == '1') prm.bulkStores = 2u;
    prm.noTrace = 0;
    if (const char* e = std::getenv("DXRV_EXP_NO_TRACE")) if (e[0] == '1') prm.noTrace = 1u;
#endif
    if (ev) cudaEventRecord(ev[0], s);
    if (prm.nodes)
        k_walk_columns<SY, SZ><<<(prm.numTiles + kWalkWarps - 1) / kWalkWarps, 32 * kWalkWarps, 0, s>>>(prm);
    else
    {
        // candidates by triangle-parallel binning (the default): no hierarchy needed
        const uint32_t tilesZ = (prm.z1 - prm.z0 + SZ - 1) / SZ;
        const uint32_t binBlocks = std::max(1u, std::min<uint32_t>((prm.numTris + 255u) / 256u, 148u * 8u));
        if (!binsReady) k_bin_columns<SY, SZ><<<binBlocks, 256, sizeof(float) * 2 * (prm.tilesY + tilesZ), s>>>(prm);
        const uint32_t groups = (prm.numTiles + 31u) / 32u;
        k_file_columns<<<(groups + kWalkWarps - 1) / kWalkWarps, 32 * kWalkWarps, 0, s>>>(prm);
    }
    if (ev) cudaEventRecord(ev[1], s);
#ifdef DXRV_TIMELINE
    {
        static const char* const kWalkRoles[5] = {"?", "empty CTA", "light CTA", "heavy CTA", "very heavy CTA"};
        if (std::getenv("DXRV_DBG_TIMELINE_WALK")) DXRV_TL_REPORT(s, (prm.numTiles + kWalkWarps - 1) / kWalkWarps, kWalkRoles, 5);
    }
#endif
    // enough CTAs for kFillWaves full waves of work items; more items than that are taken at a stride
    const uint32_t waves = DXRV_DYNAMIC_ITEMS ? 1u : kFillWaves;
    const uint32_t perSm = W == 4 ? (uint32_t)DXRV_FILL_CTAS : W == 8 ? 4u : 2u;
    const uint32_t workCtas = std::min<uint32_t>(prm.numTiles + kExtraParts, std::max(1u, waves * perSm * (prm.numWriters / 2u)));
    k_trace_fill_columns<W, SY, SZ><<<prm.numWriters + workCtas, 32 * W, smemBytes, s>>>(prm);
    if (ev) cudaEventRecord(ev[2], s);
    static const char* const kRoles[6] = {"?", "writer", "heavy part", "light tile", "merged part", "surplus"};
    (void)kRoles;
    DXRV_TL_REPORT(s, prm.numWriters + workCtas, kRoles, 6);
}
}  // namespace

// super-tile shape: 16 x 8 columns, 4 warps per CTA (9 CTAs / SM at N = 1024).  Smaller tiles than the
// obvious 16 x 16 spread the dense tiles of a real mesh over more SMs.
void parityTileCounts(uint32_t N, uint32_t z0, uint32_t z1, uint32_t& numTiles, uint32_t& candCap)
{
    const uint32_t SY = 16, SZ = 8;
    numTiles = ((N + SY - 1) / SY) * ((z1 - z0 + SZ - 1) / SZ);
    // candidate slots per tile: 2048, more when the tiles are few (a coarse grid puts thousands of triangles of a
    // surface seen edge-on into one tile; a list that overflows sends its tile to the slow in-kernel walk) --
    // up to 256 MiB of lists in all
    candCap = 2048;
    while (candCap < 32768u && (uint64_t)numTiles * candCap * 2u * sizeof(uint32_t) <= (256ull << 20)) candCap *= 2u;
}

// The per-tile candidate counters sit right behind the launch counters (ONE memset node clears both) and are sized for
// the whole grid, so that the zero-on-first-use head of the scratch depends on N only, not on the slab.
static size_t fullGridTilesPad(uint32_t N)
{
    uint32_t numTiles, candCap;
    parityTileCounts(N, 0, N, numTiles, candCap);
    return (numTiles + 31u) & ~31u;
}

// words of device scratch launchTraceFillColumns needs (see the layout below); the region up to
// parityScratchZeroWords() must be zero when first used (it is self-cleaning afterwards)
size_t parityScratchWords(uint32_t N, uint32_t z0, uint32_t z1)
{
    uint32_t numTiles, candCap;
    parityTileCounts(N, z0, z1, numTiles, candCap);
    const size_t tilesPad = (numTiles + 31u) & ~31u;
    const size_t Ps = sharedRowWords((N + 31) / 32);
    return kCounterWords + fullGridTilesPad(N) + kHeavySlots + (size_t)kHeavySlots * 128 * Ps + (1 + kLightClasses) * tilesPad + 2 * (tilesPad + kExtraParts) + (size_t)numTiles * candCap;
}

size_t parityScratchZeroWords(uint32_t N)
{
    const size_t Ps = sharedRowWords((N + 31) / 32);
    return kCounterWords + fullGridTilesPad(N) + kHeavySlots + (size_t)kHeavySlots * 128 * Ps;
}

int launchTraceFillColumns(cudaStream_t s, const BvhView& bvh, uint32_t N, uint32_t z0, uint32_t z1, uint32_t* grid,
                           uint32_t* walkBuf, unsigned long long* dCrossings, uint32_t* dErr, cudaEvent_t* ev, bool binsReady)
{
    ParityParams prm;
    prm.nodes = bvh.nodes; prm.tris = bvh.tris; prm.numTris = bvh.numTris;
    prm.N = N; prm.P = (N + 31) / 32; prm.Ps = sharedRowWords(prm.P);
    prm.gprShift = 0;
    while ((1u << prm.gprShift) < (prm.Ps >> 2)) ++prm.gprShift;
    prm.z0 = z0; prm.z1 = z1;
    prm.tilesY = (N + 15) / 16;
    parityTileCounts(N, z0, z1, prm.numTiles, prm.candCap);
    prm.invNPow2 = ((N & (N - 1)) == 0) ? 1.0f / (float)N : 0.0f;
    prm.grid = grid;
    const size_t tilesPad = (prm.numTiles + 31u) & ~31u;
    prm.tilesPad = (uint32_t)tilesPad;
    uint32_t* p = walkBuf;
    prm.bucketCount = p;  p += kCounterWords;
    prm.candCount = p;    p += fullGridTilesPad(N);
    prm.heavyArrive = p;  p += kHeavySlots;
    prm.heavyScratch = p; p += (size_t)kHeavySlots * 128 * prm.Ps;      // 16-byte aligned: all sizes are multiples of 4 words
    prm.lightTiles = p;   p += kLightClasses * tilesPad;
    prm.emptyTiles = p;   p += tilesPad;
    prm.heavyEntries = reinterpret_cast<uint2*>(p); p += 2 * (tilesPad + kExtraParts);
    prm.candList = p;
    prm.crossings = dCrossings; prm.err = dErr;
    // one memset node: the launch counters and, behind them, the tile counters the binning kernel counts up with atomics
    if (prm.nodes) binsReady = false;
    cudaMemsetAsync(prm.bucketCount, 0, (kCounterWords + ((prm.nodes || binsReady) ? 0 : tilesPad)) * sizeof(uint32_t), s);
    // warps per CTA by row length (see k_trace_fill_columns); every choice keeps rows-per-warp x groups-per-row
    // a multiple of 32
    if (prm.Ps <= 32) launchVariant<4, 16, 8>(s, prm, ev, binsReady);
    else if (prm.Ps <= 64) launchVariant<8, 16, 8>(s, prm, ev, binsReady);
    else launchVariant<16, 16, 8>(s, prm, ev, binsReady);
    return (prm.nodes || binsReady) ? 2 : 3;
}
}  // namespace dxrv
