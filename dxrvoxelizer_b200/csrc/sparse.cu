// sparse.cu -- DXRV_FORMAT_SPARSE_BRICKS: a lossless compact form of the bit grid for the read-back.
//
// A solid voxelization is almost all empty space and solid interior: the dragon's 1024^3 grid is 94 % empty, and what
// is not empty is mostly full.  The dense bit grid is already 32x smaller than the reference's R10G10B10A2 texture
// (Content/Voxelizer.cpp:62-67), but its 128 MiB still take 2.4 ms over PCIe -- 95 % of an end-to-end step.  This
// format sends only the bricks the surface passes through.
//
//   brick   = 32 (x) x 4 (y) x 4 (z) voxels = 16 words of the dense grid: word column bx, rows y = 4 by + j, z = 4 bz + k
//             (rows beyond N / beyond the slab do not exist and count as empty).  Bricks are numbered
//             b = (bz * BY + by) * P + bx.
//   state   = 2 bits per brick, 16 bricks per uint32 (brick b: bits 2 (b & 15) of word b >> 4):
//             0 empty, 1 full (every existing voxel of the brick set), 2 mixed
//   payload = the 16 words {k = 0..3 {j = 0..3}} of every mixed brick, in brick order
//   blob    = header (16 uint32) | states | payload        (dxrv.h: dxrv_sparse_header)
//
//   k_brick_classify   one thread per brick: 16 loads (coalesced over bx), state, mixed bricks counted per block
//   k_brick_scan       exclusive scan of the block counts (one block), totals into the header
//   k_brick_pack       mixed bricks copy their 16 words to payload[rank]
// Decoding (dxrv_sparse_decode, host) is the inverse; tests require decode(encode(grid)) == grid for every case.
#include "kernels.h"

namespace dxrv
{
namespace
{
constexpr int kBrickThreads = (int)kSparseBlockBricks;

struct BrickGeom
{
    const uint32_t* grid;
    uint32_t N, P, layers, BY, BZ;
    uint32_t numBricks;
    uint32_t tailMask;     // valid bits of the last word of a row (N % 32), else all ones
};

__device__ __forceinline__ void loadBrick(const BrickGeom& g, uint32_t b, uint32_t (&w)[16], uint32_t& exists)
{
    const uint32_t bx = b % g.P, t = b / g.P, by = t % g.BY, bz = t / g.BY;
    exists = 0;
#pragma unroll
    for (uint32_t k = 0; k < 4u; ++k)
#pragma unroll
        for (uint32_t j = 0; j < 4u; ++j)
        {
            const uint32_t y = 4u * by + j, z = 4u * bz + k;
            const bool ok = y < g.N && z < g.layers;
            w[4u * k + j] = ok ? __ldg(g.grid + ((size_t)z * g.N + y) * g.P + bx) : 0u;
            exists |= ok ? 1u << (4u * k + j) : 0u;
        }
}

__global__ void __launch_bounds__(kBrickThreads)
k_brick_classify(const BrickGeom g, uint32_t* __restrict__ states, uint32_t* __restrict__ blockCounts)
{
    __shared__ uint32_t sCount;
    if (threadIdx.x == 0) sCount = 0;
    __syncthreads();
    const uint32_t b = blockIdx.x * kBrickThreads + threadIdx.x;
    uint32_t state = 0;
    if (b < g.numBricks)
    {
        uint32_t w[16], exists;
        loadBrick(g, b, w, exists);
        const uint32_t fullWord = (b % g.P == g.P - 1u) ? g.tailMask : 0xffffffffu;
        uint32_t any = 0;
        bool full = true;
#pragma unroll
        for (int i = 0; i < 16; ++i)
        {
            any |= w[i];
            if ((exists >> i) & 1u) full = full && w[i] == fullWord;
        }
        state = any == 0u ? 0u : (full ? 1u : 2u);
    }
    // 16 bricks per state word: OR over each half-warp
    uint32_t word = state << (2u * (threadIdx.x & 15u));
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) word |= __shfl_xor_sync(0xffffffffu, word, o);
    if ((threadIdx.x & 15u) == 0u && b < g.numBricks) states[b >> 4] = word;
    const uint32_t mixed = __ballot_sync(0xffffffffu, state == 2u);
    if ((threadIdx.x & 31u) == 0u && mixed) atomicAdd(&sCount, (uint32_t)__popc(mixed));
    __syncthreads();
    if (threadIdx.x == 0) blockCounts[blockIdx.x] = sCount;
}

__global__ void __launch_bounds__(1024)
k_brick_scan(uint32_t* __restrict__ blockCounts, uint32_t numBlocks, uint32_t* __restrict__ header, uint32_t padBegin, uint32_t padEnd)
{
    // the alignment gap between the states and the payload is part of the blob: it must not carry stale device memory
    for (uint32_t i = padBegin + threadIdx.x; i < padEnd; i += blockDim.x) header[i] = 0u;
    __shared__ uint32_t warpSums[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < numBlocks; base += 1024u)
    {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < numBlocks ? blockCounts[i] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (laneId() >= (uint32_t)o) inc += t;
        }
        const uint32_t warp = threadIdx.x >> 5;
        if (laneId() == 31) warpSums[warp] = inc;
        __syncthreads();
        uint32_t run = carry + inc - v;
        for (uint32_t w = 0; w < warp; ++w) run += warpSums[w];
        if (i < numBlocks) blockCounts[i] = run;
        __syncthreads();
        if (threadIdx.x == 1023) carry = run + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) header[9] = carry;   // numMixed
}

__global__ void __launch_bounds__(kBrickThreads)
k_brick_pack(const BrickGeom g, const uint32_t* __restrict__ states, const uint32_t* __restrict__ blockBase, uint4* __restrict__ payload)
{
    __shared__ uint32_t warpCount[kBrickThreads / 32];
    const uint32_t b = blockIdx.x * kBrickThreads + threadIdx.x;
    const bool mixed = b < g.numBricks && ((__ldg(states + (b >> 4)) >> (2u * (b & 15u))) & 3u) == 2u;
    const uint32_t m = __ballot_sync(0xffffffffu, mixed);
    const uint32_t warp = threadIdx.x >> 5;
    if (laneId() == 0) warpCount[warp] = (uint32_t)__popc(m);
    __syncthreads();
    if (!mixed) return;
    uint32_t rank = __ldg(blockBase + blockIdx.x) + (uint32_t)__popc(m & laneMaskLt());
    for (uint32_t w = 0; w < warp; ++w) rank += warpCount[w];
    uint32_t w[16], exists;
    loadBrick(g, b, w, exists);
    uint4* dst = payload + 4 * (size_t)rank;
    dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
    dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
    dst[2] = make_uint4(w[8], w[9], w[10], w[11]);
    dst[3] = make_uint4(w[12], w[13], w[14], w[15]);
}
}  // namespace

SparseLayout sparseLayout(uint32_t N, uint32_t layers)
{
    SparseLayout L;
    L.P = (N + 31) / 32; L.BY = (N + 3) / 4; L.BZ = (layers + 3) / 4;
    L.numBricks = L.P * L.BY * L.BZ;
    L.stateWords = (L.numBricks + 15) / 16;
    L.numBlocks = (L.numBricks + kBrickThreads - 1) / kBrickThreads;
    L.offStates = 64;
    L.offPayload = (L.offStates + (size_t)L.stateWords * 4 + 63) & ~(size_t)63;
    L.maxBytes = L.offPayload + (size_t)L.numBricks * 64;
    return L;
}

int launchSparseEncode(cudaStream_t s, const uint32_t* grid, uint32_t N, uint32_t z0, uint32_t z1, uint8_t* blob, uint32_t* blockCounts)
{
    const SparseLayout L = sparseLayout(N, z1 - z0);
    BrickGeom g;
    g.grid = grid; g.N = N; g.P = L.P; g.layers = z1 - z0; g.BY = L.BY; g.BZ = L.BZ; g.numBricks = L.numBricks;
    g.tailMask = (N & 31u) ? ((1u << (N & 31u)) - 1u) : 0xffffffffu;
    uint32_t* header = reinterpret_cast<uint32_t*>(blob);
    const uint32_t h[16] = {0x42525844u /* "DXRB" */, 1u, N, z0, z1, L.P, L.BY, L.BZ, L.numBricks, 0u /* numMixed */,
                            (uint32_t)L.offStates, (uint32_t)L.offPayload, 32u, 4u, 4u, 0u};
    cudaMemcpyAsync(header, h, sizeof(h), cudaMemcpyHostToDevice, s);
    uint32_t* states = reinterpret_cast<uint32_t*>(blob + L.offStates);
    k_brick_classify<<<L.numBlocks, kBrickThreads, 0, s>>>(g, states, blockCounts);
    k_brick_scan<<<1, 1024, 0, s>>>(blockCounts, L.numBlocks, header, (uint32_t)(L.offStates / 4) + L.stateWords, (uint32_t)(L.offPayload / 4));
    k_brick_pack<<<L.numBlocks, kBrickThreads, 0, s>>>(g, states, blockCounts, reinterpret_cast<uint4*>(blob + L.offPayload));
    return 3;
}
}  // namespace dxrv
