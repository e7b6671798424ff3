// onesweep.cu -- stable LSD radix sort of (Morton key, triangle index) pairs for sm_100a.
//
// Single-pass-per-digit "onesweep" (Adinets & Merrill 2022): the digit histograms of all passes are
// computed up front (by the Morton kernel itself when the sort serves the LBVH build), and each
// 8-bit digit pass then reads every pair ONCE and writes it ONCE; the cross-tile prefix is resolved
// inside the pass by decoupled look-back over per-tile {flag, count} words.  Algorithmic HBM
// traffic: passes x 16 B per pair (+ 4 B per pair when the histogram needs its own read).
//
// Tile = 256 threads x kItems pairs (16 for bandwidth-bound sizes, 4 for small inputs where the
// per-tile latency chain matters more than the tile count).  Ranking inside a tile is
// warp-synchronous (match.any), which keeps the sort stable: order of equal digits =
// (warp, item, lane) = input order.
#include "kernels.h"

namespace dxrv
{
namespace
{
constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;   // 256
constexpr int kMaxPasses = 4;
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr uint32_t kSmallSort = 1u << 19;  // below this many pairs use the 1024-pair tile
#ifndef SORT_ITEMS
#define SORT_ITEMS 16
#endif
#ifndef SORT_MINBLOCKS
#define SORT_MINBLOCKS 4
#endif
#ifndef SORT_LB_WINDOW
#define SORT_LB_WINDOW 8
#endif
constexpr int kBigItems = SORT_ITEMS;      // items per thread of the bandwidth-bound variant

constexpr uint32_t kFlagAggregate = 1u << 30;
constexpr uint32_t kFlagInclusive = 2u << 30;
constexpr uint32_t kValueMask = (1u << 30) - 1u;

// ---- digit histograms for all passes (stand-alone sort only; the LBVH build fuses this into k_morton)
__global__ void __launch_bounds__(kSortThreads)
k_radix_histogram(const uint32_t* __restrict__ keys, uint32_t n, uint32_t* __restrict__ hist)
{
    __shared__ uint32_t sh[kMaxPasses][kRadix];
    for (int i = threadIdx.x; i < kMaxPasses * kRadix; i += blockDim.x) (&sh[0][0])[i] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const uint32_t k = __ldg(keys + i);
        atomicAdd(&sh[0][k & 255u], 1u);
        atomicAdd(&sh[1][(k >> 8) & 255u], 1u);
        atomicAdd(&sh[2][(k >> 16) & 255u], 1u);
        atomicAdd(&sh[3][k >> 24], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kMaxPasses * kRadix; i += blockDim.x)
    {
        const uint32_t c = (&sh[0][0])[i];
        if (c) atomicAdd(hist + i, c);
    }
}

// block-wide exclusive scan of one value per thread (256 threads); returns the exclusive prefix
__device__ __forceinline__ uint32_t blockExclusiveScan256(uint32_t v, uint32_t* warpSums /* [8] shared */)
{
    const uint32_t lane = laneId(), warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += t;
    }
    if (lane == 31) warpSums[warp] = inc;
    __syncthreads();
    uint32_t base = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w)
        if ((uint32_t)w < warp) base += warpSums[w];
    __syncthreads();
    return base + inc - v;
}

// ---- one digit pass ----------------------------------------------------------------------------------
template <int kItems>
__global__ void __launch_bounds__(kSortThreads, kItems >= 8 ? SORT_MINBLOCKS : 1)
k_onesweep_pass(const uint32_t* __restrict__ keysIn, const uint32_t* __restrict__ valsIn,
                uint32_t* __restrict__ keysOut, uint32_t* __restrict__ valsOut, uint32_t n, int shift,
                const uint32_t* __restrict__ digitCount, volatile uint32_t* __restrict__ lookback,
                uint32_t* __restrict__ tileCounter)
{
    constexpr int kTile = kSortThreads * kItems;
    // predecessor tiles fetched per round trip of the look-back: the small-tile variant is latency bound
    // (many tiles, few keys), the big-tile variant is register bound
    constexpr int kLookbackWindow = kItems >= 8 ? SORT_LB_WINDOW : 12;
    __shared__ uint32_t warpHist[kSortWarps][kRadix];  // 8 KB
    __shared__ uint32_t binStart[kRadix];
    __shared__ uint32_t globalBase[kRadix];
    __shared__ uint32_t sKeys[kTile];
    __shared__ uint32_t sVals[kTile];
    __shared__ uint32_t warpSums[kSortWarps];
    __shared__ uint32_t sTile;

    const uint32_t tid = threadIdx.x, lane = laneId(), warp = tid >> 5;
    // Tiles are numbered in the order blocks START, so every lower-numbered tile is already running
    // (or done) when this one looks back: the look-back can never wait on an unscheduled block.
    if (tid == 0) sTile = atomicAdd(tileCounter, 1u);
    for (int i = tid; i < kSortWarps * kRadix; i += kSortThreads) (&warpHist[0][0])[i] = 0;
    // exclusive digit bases of this pass from the global digit counts (256 values: one block scan)
    const uint32_t digitBase = blockExclusiveScan256(__ldg(digitCount + tid), warpSums);
    const uint32_t tile = sTile;
    const uint32_t base = tile * (uint32_t)kTile;
    const uint32_t valid = min((uint32_t)kTile, n - base);

    // ---- load (warp-striped: item i of lane l is element warp*32*kItems + i*32 + l) ----
    uint32_t key[kItems], val[kItems], rank[kItems];
    const uint32_t warpBase = warp * (32u * kItems);
#pragma unroll
    for (int i = 0; i < kItems; ++i)
    {
        const uint32_t local = warpBase + i * 32u + lane;
        const bool ok = local < valid;
        key[i] = ok ? __ldg(keysIn + base + local) : 0xffffffffu;
        val[i] = ok ? __ldg(valsIn + base + local) : 0u;
    }

    // ---- rank within the warp ----
#pragma unroll
    for (int i = 0; i < kItems; ++i)
    {
        const uint32_t d = (key[i] >> shift) & 255u;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if ((int)lane == leader)
        {
            old = warpHist[warp][d];
            warpHist[warp][d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[i] = old + __popc(peers & laneMaskLt());
        __syncwarp();
    }
    __syncthreads();

    // ---- per digit (thread d): prefix over warps, tile count ----
    uint32_t count = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w)
    {
        const uint32_t c = warpHist[w][tid];
        warpHist[w][tid] = count;
        count += c;
    }

    // The tile's digit counts are published at once; the look-back itself is delayed until the pairs sit in shared
    // memory in sorted order (that scatter needs only tile-local offsets): the predecessors, which started at about
    // the same time, have published by then, and the look-back rarely spins.
    volatile uint32_t* lb = lookback + (size_t)tile * kRadix;
    lb[tid] = (tile == 0 ? kFlagInclusive : kFlagAggregate) | count;

    const uint32_t start = blockExclusiveScan256(count, warpSums);
    binStart[tid] = start;
    __syncthreads();

    // ---- scatter into shared memory in sorted order ----
#pragma unroll
    for (int i = 0; i < kItems; ++i)
    {
        const uint32_t d = (key[i] >> shift) & 255u;
        const uint32_t pos = binStart[d] + warpHist[warp][d] + rank[i];
        sKeys[pos] = key[i];
        sVals[pos] = val[i];
    }

    // ---- decoupled look-back: exclusive count of digit `tid` over all previous tiles.  A window of
    // predecessors is fetched at once, so the latency chain is 1/window of the tile distance. ----
    uint32_t prev = 0;
    // (Tried and dropped for the 100 k-key regime: every tile summing ALL predecessor aggregates with independent
    // loads instead of the look-back chain -- 12.5 vs 10 us per pass: 98 tiles x 97 loads x 256 digits is more L2
    // traffic than the chain saves in latency.)
    if (tile != 0)
    {
        int j = (int)tile - 1;
        bool done = false;
        while (!done)
        {
            uint32_t v[kLookbackWindow];
#pragma unroll
            for (int q = 0; q < kLookbackWindow; ++q) v[q] = (j - q >= 0) ? lookback[(size_t)(j - q) * kRadix + tid] : 0u;
#pragma unroll
            for (int q = 0; q < kLookbackWindow; ++q)
            {
                if (done || j - q < 0) break;
                uint32_t x = v[q];
                while ((x & ~kValueMask) == 0u) x = lookback[(size_t)(j - q) * kRadix + tid];  // not published yet
                prev += x & kValueMask;
                if ((x & ~kValueMask) == kFlagInclusive) done = true;
            }
            j -= kLookbackWindow;
        }
        lb[tid] = kFlagInclusive | (prev + count);
    }
    globalBase[tid] = digitBase + prev - start;
    __syncthreads();

    // ---- coalesced write-out (padding keys sort to the end of the tile: positions >= valid) ----
    for (uint32_t p = tid; p < valid; p += kSortThreads)
    {
        const uint32_t k = sKeys[p];
        const uint32_t dst = globalBase[(k >> shift) & 255u] + p;
        keysOut[dst] = k;
        valsOut[dst] = sVals[p];
    }
}

// ---- one digit pass, bandwidth-bound sizes: 512 threads x 16 pairs = 8192 pairs per tile ---------------------------
// Same algorithm as k_onesweep_pass.  The decoupled look-back was a third of all instructions of the 4096-pair tile at
// 16.8 M pairs (all resident tiles start together, so the nearest INCLUSIVE prefix is about a wave of tiles back and
// every digit thread walks that distance); twice the tile halves the tiles in flight and the look-backs per pair.
// Measured: 103.6 -> 101.8 us per pass -- the pass stays latency bound.  Shared memory is dynamic (82 KB).
constexpr int kBigThreads = 512;
constexpr int kBigWarps = kBigThreads / 32;
constexpr int kBigTile = kBigThreads * kBigItems;
constexpr size_t kBigSmemBytes = sizeof(uint32_t) * ((size_t)kBigWarps * kRadix + 2 * kRadix + 2 * (size_t)kBigTile + kBigWarps + 4);

__global__ void __launch_bounds__(kBigThreads, 2)
k_onesweep_pass_big(const uint32_t* __restrict__ keysIn, const uint32_t* __restrict__ valsIn,
                    uint32_t* __restrict__ keysOut, uint32_t* __restrict__ valsOut, uint32_t n, int shift,
                    const uint32_t* __restrict__ digitCount, volatile uint32_t* __restrict__ lookback,
                    uint32_t* __restrict__ tileCounter)
{
    constexpr int kItems = kBigItems;
    constexpr int kLookbackWindow = 8;
    extern __shared__ __align__(16) uint32_t smemBig[];
    uint32_t (*warpHist)[kRadix] = reinterpret_cast<uint32_t (*)[kRadix]>(smemBig);   // [kBigWarps][256]
    uint32_t* binStart = smemBig + kBigWarps * kRadix;
    uint32_t* globalBase = binStart + kRadix;
    uint32_t* sKeys = globalBase + kRadix;
    uint32_t* sVals = sKeys + kBigTile;
    uint32_t* warpSums = sVals + kBigTile;       // [kBigWarps]
    uint32_t* sTile = warpSums + kBigWarps;

    const uint32_t tid = threadIdx.x, lane = laneId(), warp = tid >> 5;
    const bool digitThread = tid < (uint32_t)kRadix;
    if (tid == 0) *sTile = atomicAdd(tileCounter, 1u);
    for (int i = tid; i < kBigWarps * kRadix; i += kBigThreads) (&warpHist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = *sTile;
    const uint32_t base = tile * (uint32_t)kBigTile;
    const uint32_t valid = min((uint32_t)kBigTile, n - base);

    // block-wide exclusive scan of one value per DIGIT thread (threads >= 256 pass 0 and ignore the result)
    auto scanDigits = [&](uint32_t v) -> uint32_t {
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc += t;
        }
        if (lane == 31) warpSums[warp] = inc;
        __syncthreads();
        uint32_t b = 0;
#pragma unroll
        for (int w = 0; w < kRadix / 32; ++w)
            if ((uint32_t)w < warp) b += warpSums[w];
        __syncthreads();
        return b + inc - v;
    };
    const uint32_t digitBase = scanDigits(digitThread ? __ldg(digitCount + tid) : 0u);

    // ---- load (warp-striped) ----
    uint32_t key[kItems], val[kItems], rank[kItems];
    const uint32_t warpBase = warp * (32u * kItems);
#pragma unroll
    for (int i = 0; i < kItems; ++i)
    {
        const uint32_t local = warpBase + i * 32u + lane;
        const bool ok = local < valid;
        key[i] = ok ? __ldg(keysIn + base + local) : 0xffffffffu;
        val[i] = ok ? __ldg(valsIn + base + local) : 0u;
    }

    // ---- rank within the warp (stable) ----
#pragma unroll
    for (int i = 0; i < kItems; ++i)
    {
        const uint32_t d = (key[i] >> shift) & 255u;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if ((int)lane == leader)
        {
            old = warpHist[warp][d];
            warpHist[warp][d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[i] = old + __popc(peers & laneMaskLt());
        __syncwarp();
    }
    __syncthreads();

    // ---- per digit: prefix over warps, tile count; published at once ----
    uint32_t count = 0;
    if (digitThread)
    {
#pragma unroll
        for (int w = 0; w < kBigWarps; ++w)
        {
            const uint32_t c = warpHist[w][tid];
            warpHist[w][tid] = count;
            count += c;
        }
        lookback[(size_t)tile * kRadix + tid] = (tile == 0 ? kFlagInclusive : kFlagAggregate) | count;
    }
    const uint32_t start = scanDigits(count);
    if (digitThread) binStart[tid] = start;
    __syncthreads();

    // ---- scatter into shared memory in sorted order ----
#pragma unroll
    for (int i = 0; i < kItems; ++i)
    {
        const uint32_t d = (key[i] >> shift) & 255u;
        const uint32_t pos = binStart[d] + warpHist[warp][d] + rank[i];
        sKeys[pos] = key[i];
        sVals[pos] = val[i];
    }

    // ---- decoupled look-back (digit threads), delayed behind the scatter ----
    // (Tried and dropped, measured at 16.8 M pairs: a two-level look-back -- blocks of 32 tiles publishing block
    // aggregates -- was SLOWER, 127 vs 102 us per pass: the block words add a second spin chain.  The pass is bound by
    // the latency of this walk)
    if (digitThread)
    {
        uint32_t prev = 0;
        if (tile != 0)
        {
            int j = (int)tile - 1;
            bool done = false;
            while (!done)
            {
                uint32_t v[kLookbackWindow];
#pragma unroll
                for (int q = 0; q < kLookbackWindow; ++q) v[q] = (j - q >= 0) ? lookback[(size_t)(j - q) * kRadix + tid] : 0u;
#pragma unroll
                for (int q = 0; q < kLookbackWindow; ++q)
                {
                    if (done || j - q < 0) break;
                    uint32_t x = v[q];
                    while ((x & ~kValueMask) == 0u) x = lookback[(size_t)(j - q) * kRadix + tid];  // not published yet
                    prev += x & kValueMask;
                    if ((x & ~kValueMask) == kFlagInclusive) done = true;
                }
                j -= kLookbackWindow;
            }
            lookback[(size_t)tile * kRadix + tid] = kFlagInclusive | (prev + count);
        }
        globalBase[tid] = digitBase + prev - start;
    }
    __syncthreads();

    // ---- coalesced write-out ----
    for (uint32_t p = tid; p < valid; p += kBigThreads)
    {
        const uint32_t k = sKeys[p];
        const uint32_t dst = globalBase[(k >> shift) & 255u] + p;
        keysOut[dst] = k;
        valsOut[dst] = sVals[p];
    }
}

uint32_t tileSizeFor(uint32_t n) { return n < kSmallSort ? kSortThreads * 4u : (uint32_t)kBigTile; }
}  // namespace

static size_t passWords(size_t tiles) { return tiles * kRadix; }

uint32_t SortTemp::tilesFor(uint32_t n) { const uint32_t t = tileSizeFor(n); return (n + t - 1) / t; }

size_t SortTemp::bytesFor(uint32_t n)
{
    const size_t tiles = tilesFor(n) ? tilesFor(n) : 1;
    // [hist 4*256][tileCounter 4 (padded to 64)][per pass: look-back words tiles*256]
    return sizeof(uint32_t) * (kMaxPasses * kRadix + 64 + (size_t)kMaxPasses * passWords(tiles));
}

uint32_t* sortClearTemp(cudaStream_t s, void* tempBase, uint32_t n)
{
    cudaMemsetAsync(tempBase, 0, SortTemp::bytesFor(n), s);
    return static_cast<uint32_t*>(tempBase);
}

int radixSortPairs(cudaStream_t s, void* tempBase, uint32_t* keysA, uint32_t* valsA, uint32_t* keysB,
                   uint32_t* valsB, uint32_t n, int numPasses, bool histReady, bool* resultInB)
{
    if (resultInB) *resultInB = false;
    if (n < 2 || numPasses < 1) return 0;
    if (numPasses > kMaxPasses) numPasses = kMaxPasses;
    const uint32_t tiles = SortTemp::tilesFor(n);
    uint32_t* hist = static_cast<uint32_t*>(tempBase);
    uint32_t* tileCounter = hist + kMaxPasses * kRadix;
    uint32_t* lookback = tileCounter + 64;

    int launches = 0;
    if (!histReady)
    {
        sortClearTemp(s, tempBase, n);
        uint32_t histBlocks = (n + kSortThreads * 16 - 1) / (kSortThreads * 16);
        if (histBlocks > 148 * 4) histBlocks = 148 * 4;
        k_radix_histogram<<<histBlocks, kSortThreads, 0, s>>>(keysA, n, hist);
        ++launches;
    }
    {
        // two 82 KB CTAs per SM: dynamic shared memory above 48 KB is opt-in
        static bool carveSet[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 0 && dev < 64 && !carveSet[dev])
        {
            cudaFuncSetAttribute(k_onesweep_pass_big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBigSmemBytes);
            cudaFuncSetAttribute(k_onesweep_pass_big, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            carveSet[dev] = true;
        }
    }
    uint32_t *kin = keysA, *vin = valsA, *kout = keysB, *vout = valsB;
    for (int p = 0; p < numPasses; ++p)
    {
        if (n < kSmallSort)
            k_onesweep_pass<4><<<tiles, kSortThreads, 0, s>>>(kin, vin, kout, vout, n, p * kRadixBits, hist + p * kRadix,
                                                              lookback + (size_t)p * passWords(tiles), tileCounter + p);
        else
            k_onesweep_pass_big<<<tiles, kBigThreads, kBigSmemBytes, s>>>(kin, vin, kout, vout, n, p * kRadixBits, hist + p * kRadix,
                                                                         lookback + (size_t)p * passWords(tiles), tileCounter + p);
        ++launches;
        uint32_t* t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
    }
    if (resultInB) *resultInB = (numPasses & 1) != 0;
    return launches;
}
}  // namespace dxrv
