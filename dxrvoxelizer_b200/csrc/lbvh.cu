// lbvh.cu -- LBVH construction kernels for sm_100a.
//
// Replaces the driver's BLAS/TLAS build requested by Voxelizer::buildAccelerationStructures
// (reference Content/Voxelizer.cpp:264-326).  Pipeline (all HBM/latency bound, no tensor work):
//   k_bounds      min/max of all vertex positions -> {c, w}         (Voxelizer.cpp:52-57)
//   k_morton      per triangle: scene-space box centre -> 30-bit Morton key
//   (onesweep.cu) stable radix sort of (key, triangle)
//   k_hierarchy   Karras 2012 radix tree over the sorted keys (index tie-break for duplicates)
//   k_refit       leaves: scene-space triangle + box; bottom-up union with one atomic per node
#include "kernels.h"

namespace dxrv
{
namespace
{
__device__ __forceinline__ float3 loadPos(const uint8_t* verts, uint32_t stride, uint32_t i)
{
    const float* p = reinterpret_cast<const float*>(verts + (size_t)stride * i);
    return make_float3(__ldg(p), __ldg(p + 1), __ldg(p + 2));
}

__device__ __forceinline__ float3 scenePos(const uint8_t* verts, uint32_t stride, uint32_t i, float4 bound)
{
    const float3 p = loadPos(verts, stride, i);
    return make_float3(toScene(p.x, bound.x, bound.w), toScene(p.y, bound.y, bound.w),
                       toScene(p.z, bound.z, bound.w));
}

// ---- bounds ---------------------------------------------------------------------------------------
constexpr int kBoundsThreads = 256;

__device__ __forceinline__ void warpMinMax(float (&mn)[3], float (&mx)[3])
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int a = 0; a < 3; ++a)
        {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
}

__global__ void __launch_bounds__(kBoundsThreads)
k_bounds(const uint8_t* __restrict__ verts, uint32_t numVerts, uint32_t stride, float* __restrict__ bound,
         float* __restrict__ partials, uint32_t* __restrict__ counter)
{
    __shared__ float sm[kBoundsThreads / 32][6];
    __shared__ bool isLast;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < numVerts; i += gridDim.x * blockDim.x)
    {
        const float3 p = loadPos(verts, stride, i);
        mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
        mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
        mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
    }
    warpMinMax(mn, mx);
    const int warp = threadIdx.x >> 5;
    if (laneId() == 0)
        for (int a = 0; a < 3; ++a) { sm[warp][a] = mn[a]; sm[warp][3 + a] = mx[a]; }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        for (int w = 1; w < kBoundsThreads / 32; ++w)
            for (int a = 0; a < 3; ++a)
            {
                mn[a] = fminf(mn[a], sm[w][a]);
                mx[a] = fmaxf(mx[a], sm[w][3 + a]);
            }
        for (int a = 0; a < 3; ++a)
        {
            partials[6 * blockIdx.x + a] = mn[a];
            partials[6 * blockIdx.x + 3 + a] = mx[a];
        }
        __threadfence();
        isLast = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!isLast) return;

    // the last block folds the per-block partials and derives {c, w}
    __threadfence();
    for (int a = 0; a < 3; ++a) { mn[a] = INFINITY; mx[a] = -INFINITY; }
    for (uint32_t b = threadIdx.x; b < gridDim.x; b += blockDim.x)
        for (int a = 0; a < 3; ++a)
        {
            mn[a] = fminf(mn[a], __ldcg(&partials[6 * b + a]));
            mx[a] = fmaxf(mx[a], __ldcg(&partials[6 * b + 3 + a]));
        }
    warpMinMax(mn, mx);
    __syncthreads();
    if (laneId() == 0)
        for (int a = 0; a < 3; ++a) { sm[warp][a] = mn[a]; sm[warp][3 + a] = mx[a]; }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        for (int w = 1; w < kBoundsThreads / 32; ++w)
            for (int a = 0; a < 3; ++a)
            {
                mn[a] = fminf(mn[a], sm[w][a]);
                mx[a] = fmaxf(mx[a], sm[w][3 + a]);
            }
        // Voxelizer.cpp:52-57: centre = (max + min) / 2, w = max(ext) / 2
        const float ex = __fsub_rn(mx[0], mn[0]), ey = __fsub_rn(mx[1], mn[1]), ez = __fsub_rn(mx[2], mn[2]);
        bound[0] = __fdiv_rn(__fadd_rn(mx[0], mn[0]), 2.0f);
        bound[1] = __fdiv_rn(__fadd_rn(mx[1], mn[1]), 2.0f);
        bound[2] = __fdiv_rn(__fadd_rn(mx[2], mn[2]), 2.0f);
        float m = ey > ez ? ey : ez;
        m = ex > m ? ex : m;
        bound[3] = __fdiv_rn(m, 2.0f);
        *counter = 0;  // self-reset for the next build
    }
}

__global__ void k_set_bound(float cx, float cy, float cz, float w, float* bound)
{
    bound[0] = cx; bound[1] = cy; bound[2] = cz; bound[3] = w;
}

// ---- Morton keys -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t expandBits10(uint32_t v)
{
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__device__ __forceinline__ uint32_t quantize10(float c)
{
    // scene space is [-1,1]^3 by construction of {c, w}
    const float q = fminf(fmaxf((c + 1.0f) * 512.0f, 0.0f), 1023.0f);
    return (uint32_t)q;
}

__global__ void __launch_bounds__(256)
k_morton(MeshView m, const float* __restrict__ boundPtr, uint32_t* __restrict__ keys,
         uint32_t* __restrict__ vals, uint32_t* __restrict__ err)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m.numTris) return;
    const float4 bound = make_float4(__ldg(boundPtr), __ldg(boundPtr + 1), __ldg(boundPtr + 2), __ldg(boundPtr + 3));
    uint32_t i0 = __ldg(m.indices + 3 * (size_t)k), i1 = __ldg(m.indices + 3 * (size_t)k + 1),
             i2 = __ldg(m.indices + 3 * (size_t)k + 2);
    if (i0 >= m.numVerts || i1 >= m.numVerts || i2 >= m.numVerts)
    {
        atomicMax(err, (uint32_t)kErrBadIndex);
        i0 = i1 = i2 = 0;
    }
    const float3 a = scenePos(m.verts, m.stride, i0, bound);
    const float3 b = scenePos(m.verts, m.stride, i1, bound);
    const float3 c = scenePos(m.verts, m.stride, i2, bound);
    const float cx = 0.5f * (fminf(fminf(a.x, b.x), c.x) + fmaxf(fmaxf(a.x, b.x), c.x));
    const float cy = 0.5f * (fminf(fminf(a.y, b.y), c.y) + fmaxf(fmaxf(a.y, b.y), c.y));
    const float cz = 0.5f * (fminf(fminf(a.z, b.z), c.z) + fmaxf(fmaxf(a.z, b.z), c.z));
    keys[k] = (expandBits10(quantize10(cx)) << 2) | (expandBits10(quantize10(cy)) << 1) | expandBits10(quantize10(cz));
    vals[k] = k;
}

// ---- Karras hierarchy ------------------------------------------------------------------------------
// delta(i,j): length of the common prefix of the 64-bit augmented keys (key << 32 | index), -1 when
// j is out of range.
__device__ __forceinline__ int delta(const uint32_t* __restrict__ keys, int numLeaves, uint32_t ki, int i, int j)
{
    if (j < 0 || j >= numLeaves) return -1;
    const uint32_t kj = __ldg(keys + j);
    const uint32_t x = ki ^ kj;
    return x ? __clz(x) : 32 + __clz((uint32_t)i ^ (uint32_t)j);
}

__global__ void __launch_bounds__(256)
k_hierarchy(const uint32_t* __restrict__ keys, int numLeaves, BvhNode* __restrict__ nodes,
            uint32_t* __restrict__ nodeParent, uint32_t* __restrict__ leafParent)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numLeaves - 1) return;
    const uint32_t ki = __ldg(keys + i);
    const int d = (delta(keys, numLeaves, ki, i, i + 1) - delta(keys, numLeaves, ki, i, i - 1)) < 0 ? -1 : 1;
    const int dMin = delta(keys, numLeaves, ki, i, i - d);
    long long lMax = 2;
    while (true)
    {
        const long long j = (long long)i + lMax * d;
        if (j < 0 || j >= numLeaves || delta(keys, numLeaves, ki, i, (int)j) <= dMin) break;
        lMax <<= 1;
    }
    long long l = 0;
    for (long long t = lMax >> 1; t >= 1; t >>= 1)
    {
        const long long j = (long long)i + (l + t) * d;
        if (j >= 0 && j < numLeaves && delta(keys, numLeaves, ki, i, (int)j) > dMin) l += t;
    }
    const int j = i + (int)l * d;
    const int dNode = delta(keys, numLeaves, ki, i, j);
    long long s = 0;
    long long t = l;
    do
    {
        t = (t + 1) >> 1;
        const long long q = (long long)i + (s + t) * d;
        if (q >= 0 && q < numLeaves && delta(keys, numLeaves, ki, i, (int)q) > dNode) s += t;
    } while (t > 1);
    const int gamma = i + (int)s * d + min(d, 0);

    const int lo = min(i, j), hi = max(i, j);
    const bool leftLeaf = (lo == gamma), rightLeaf = (hi == gamma + 1);
    nodes[i].c0 = leftLeaf ? (kLeafFlag | (uint32_t)gamma) : (uint32_t)gamma;
    nodes[i].c1 = rightLeaf ? (kLeafFlag | (uint32_t)(gamma + 1)) : (uint32_t)(gamma + 1);
    // parent reference: node index | (1u << 31 when the child is the right one)
    if (leftLeaf) leafParent[gamma] = (uint32_t)i; else nodeParent[gamma] = (uint32_t)i;
    if (rightLeaf) leafParent[gamma + 1] = (uint32_t)i | 0x80000000u; else nodeParent[gamma + 1] = (uint32_t)i | 0x80000000u;
    if (i == 0) nodeParent[0] = 0xffffffffu;
}

// ---- refit -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_refit(MeshView m, const float* __restrict__ boundPtr, const uint32_t* __restrict__ sortedPrims,
        BvhNode* __restrict__ nodes, const uint32_t* __restrict__ nodeParent,
        const uint32_t* __restrict__ leafParent, uint32_t* __restrict__ flags, Tri48* __restrict__ tris,
        float* __restrict__ rootBox, uint32_t* __restrict__ err)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m.numTris) return;
    const float4 bound = make_float4(__ldg(boundPtr), __ldg(boundPtr + 1), __ldg(boundPtr + 2), __ldg(boundPtr + 3));
    const uint32_t k = __ldg(sortedPrims + j);
    uint32_t i0 = __ldg(m.indices + 3 * (size_t)k), i1 = __ldg(m.indices + 3 * (size_t)k + 1),
             i2 = __ldg(m.indices + 3 * (size_t)k + 2);
    if (i0 >= m.numVerts || i1 >= m.numVerts || i2 >= m.numVerts)
    {
        atomicMax(err, (uint32_t)kErrBadIndex);
        i0 = i1 = i2 = 0;
    }
    const float3 a = scenePos(m.verts, m.stride, i0, bound);
    const float3 b = scenePos(m.verts, m.stride, i1, bound);
    const float3 c = scenePos(m.verts, m.stride, i2, bound);
    Tri48 t;
    t.a = make_float4(a.x, a.y, a.z, __uint_as_float(k));
    t.b = make_float4(b.x, b.y, b.z, 0.0f);
    t.c = make_float4(c.x, c.y, c.z, 0.0f);
    tris[j] = t;

    float lo[3] = {fminf(fminf(a.x, b.x), c.x), fminf(fminf(a.y, b.y), c.y), fminf(fminf(a.z, b.z), c.z)};
    float hi[3] = {fmaxf(fmaxf(a.x, b.x), c.x), fmaxf(fmaxf(a.y, b.y), c.y), fmaxf(fmaxf(a.z, b.z), c.z)};

    if (m.numTris == 1)
    {
        for (int q = 0; q < 3; ++q) { rootBox[q] = lo[q]; rootBox[3 + q] = hi[q]; }
        return;
    }

    uint32_t p = __ldg(leafParent + j);
    for (int level = 0; level < 128; ++level)  // depth <= 62; the bound only guards against a corrupt tree
    {
        const uint32_t pi = p & 0x7fffffffu, slot = p >> 31;
        const uint32_t up = (pi != 0u) ? __ldg(nodeParent + pi) : 0u;  // issued early: overlaps the atomic
        BvhNode* n = nodes + pi;
        float4* yz = slot ? &n->yz1 : &n->yz0;
        float2* xx = reinterpret_cast<float2*>(&n->x01) + slot;
        *yz = make_float4(lo[1], hi[1], lo[2], hi[2]);
        *xx = make_float2(lo[0], hi[0]);
        __threadfence();  // release: the box must be visible before the arrival counter moves
        const uint32_t old = atomicAdd(flags + pi, 1u);
        if ((old & 1u) == 0u) return;  // first child to arrive: the sibling will carry on
        // second arrival: the sibling's box was released before its increment; read it past L1
        const float4 syz = __ldcg(slot ? &n->yz0 : &n->yz1);
        const float2 sxx = __ldcg(reinterpret_cast<const float2*>(&n->x01) + (1u - slot));
        lo[0] = fminf(lo[0], sxx.x); hi[0] = fmaxf(hi[0], sxx.y);
        lo[1] = fminf(lo[1], syz.x); hi[1] = fmaxf(hi[1], syz.y);
        lo[2] = fminf(lo[2], syz.z); hi[2] = fmaxf(hi[2], syz.w);
        if (pi == 0)
        {
            for (int q = 0; q < 3; ++q) { rootBox[q] = lo[q]; rootBox[3 + q] = hi[q]; }
            return;
        }
        p = up;
    }
    atomicMax(err, (uint32_t)kErrStackOverflow);
}

// ---- small utilities -------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_popcount(const uint32_t* __restrict__ words, size_t numWords, unsigned long long* __restrict__ total)
{
    unsigned long long c = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < numWords; i += (size_t)gridDim.x * blockDim.x)
        c += __popc(__ldg(words + i));
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (laneId() == 0 && c) atomicAdd(total, c);
}

__global__ void __launch_bounds__(256)
k_bits_to_u8(const uint32_t* __restrict__ words, uint32_t N, uint32_t P, size_t numVoxels, uint8_t* __restrict__ out)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < numVoxels; i += (size_t)gridDim.x * blockDim.x)
    {
        const size_t row = i / N;
        const uint32_t x = (uint32_t)(i - row * N);
        out[i] = (uint8_t)((__ldg(words + row * P + (x >> 5)) >> (x & 31)) & 1u);
    }
}
}  // namespace

void launchBounds(cudaStream_t s, const MeshView& m, float* dBound, float* dPartials, uint32_t* dCounter)
{
    int blocks = (int)((m.numVerts + kBoundsThreads * 4 - 1) / (kBoundsThreads * 4));
    blocks = blocks < 1 ? 1 : (blocks > kBoundsMaxBlocks ? kBoundsMaxBlocks : blocks);
    k_bounds<<<blocks, kBoundsThreads, 0, s>>>(m.verts, m.numVerts, m.stride, dBound, dPartials, dCounter);
}

void launchSetBound(cudaStream_t s, float cx, float cy, float cz, float w, float* dBound)
{
    k_set_bound<<<1, 1, 0, s>>>(cx, cy, cz, w, dBound);
}

void launchMorton(cudaStream_t s, const MeshView& m, const float* dBound, uint32_t* keys, uint32_t* vals, uint32_t* dErr)
{
    if (!m.numTris) return;
    k_morton<<<(m.numTris + 255) / 256, 256, 0, s>>>(m, dBound, keys, vals, dErr);
}

void launchHierarchy(cudaStream_t s, const uint32_t* sortedKeys, uint32_t numTris, BvhNode* nodes,
                     uint32_t* nodeParent, uint32_t* leafParent)
{
    if (numTris < 2) return;
    k_hierarchy<<<(numTris - 1 + 255) / 256, 256, 0, s>>>(sortedKeys, (int)numTris, nodes, nodeParent, leafParent);
}

void launchRefit(cudaStream_t s, const MeshView& m, const float* dBound, const uint32_t* sortedPrims, BvhNode* nodes,
                 const uint32_t* nodeParent, const uint32_t* leafParent, uint32_t* flags, Tri48* tris, float* rootBox,
                 uint32_t* dErr)
{
    if (!m.numTris) return;
    k_refit<<<(m.numTris + 255) / 256, 256, 0, s>>>(m, dBound, sortedPrims, nodes, nodeParent, leafParent, flags, tris,
                                                   rootBox, dErr);
}

void launchPopcount(cudaStream_t s, const uint32_t* words, size_t numWords, unsigned long long* dCount)
{
    cudaMemsetAsync(dCount, 0, sizeof(unsigned long long), s);
    if (!numWords) return;
    size_t blocks = (numWords + 256 * 8 - 1) / (256 * 8);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_popcount<<<(unsigned)blocks, 256, 0, s>>>(words, numWords, dCount);
}

void launchBitsToU8(cudaStream_t s, const uint32_t* words, uint32_t N, uint32_t layers, uint8_t* out)
{
    const size_t numVoxels = (size_t)layers * N * N;
    if (!numVoxels) return;
    size_t blocks = (numVoxels + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_bits_to_u8<<<(unsigned)blocks, 256, 0, s>>>(words, N, (N + 31) / 32, numVoxels, out);
}
}  // namespace dxrv
