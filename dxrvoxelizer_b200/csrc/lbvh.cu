// lbvh.cu -- LBVH construction kernels for sm_100a.
//
// Replaces the driver's BLAS/TLAS build requested by Voxelizer::buildAccelerationStructures
// (reference Content/Voxelizer.cpp:264-326).  Pipeline (all HBM/latency bound, no tensor work):
//   k_bounds          min/max of all vertex positions -> {c, w}     (Voxelizer.cpp:52-57)
//   k_morton          per triangle: scene-space box centre -> Morton key, fused digit histograms
//   (onesweep.cu)     stable radix sort of (key, triangle)
//   k_leaf_setup      per sorted leaf: scene-space triangle, its box, 16- and 256-leaf summary boxes, and the
//                     prefix / suffix unions inside every aligned group of 16 (the box pyramid)
//   k_box_level       coarser summary levels (16:1) for big meshes
//   k_hierarchy_topology  Karras 2012 radix tree over the sorted keys (index tie-break for duplicates);
//                     needs only the keys, so it runs BESIDE the two kernels above (side stream / graph fork)
//   k_node_boxes      each node's child boxes are the unions of CONTIGUOUS leaf ranges, answered from the
//                     pyramid with two independent loads per level -- no parent pointers, no atomics, no
//                     bottom-up latency chain (the classic refit walks leaf-to-root with a fence + atomic
//                     per level; k_refit_atomic does that for meshes of millions of triangles, where
//                     throughput counts and not latency)
#include <algorithm>
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

// ---- Morton keys ----------------------------------------------------------------------------------
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

// keys[k] = 30-bit Morton code >> keyShift (keyShift = 6 keeps 8 bits per axis: three radix passes are
// enough for small meshes).  The digit histograms of all `numPasses` radix passes are accumulated here
// (hist must be zero on entry), so the sort never re-reads the keys for counting.
__global__ void __launch_bounds__(256)
k_morton(MeshView m, const float* __restrict__ boundPtr, uint32_t* __restrict__ keys,
         uint32_t* __restrict__ vals, uint32_t keyShift, int numPasses, uint32_t* __restrict__ hist,
         uint32_t* __restrict__ err)
{
    __shared__ uint32_t sh[4][256];
    for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) (&sh[0][0])[i] = 0;
    __syncthreads();
    // grid-stride: a bounded number of CTAs, each flushing its histograms once (one CTA per 256 triangles meant 65 k
    // global atomics on each of the 1024 counters at 16.8 M triangles -- two thirds of the kernel's time)
    const float4 bound = make_float4(__ldg(boundPtr), __ldg(boundPtr + 1), __ldg(boundPtr + 2), __ldg(boundPtr + 3));
    // The key only ORDERS the triangles (any order gives the same grids): one reciprocal and three multiply-adds
    // instead of the nine IEEE divisions of the exact scene transform, which k_leaf_setup applies to the records.
    const float invW = 1.0f / bound.w;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < m.numTris; k += gridDim.x * blockDim.x)
    {
        uint32_t i0 = __ldg(m.indices + 3 * (size_t)k), i1 = __ldg(m.indices + 3 * (size_t)k + 1),
                 i2 = __ldg(m.indices + 3 * (size_t)k + 2);
        if (i0 >= m.numVerts || i1 >= m.numVerts || i2 >= m.numVerts)
        {
            atomicMax(err, (uint32_t)kErrBadIndex);
            i0 = i1 = i2 = 0;
        }
        const float3 a = loadPos(m.verts, m.stride, i0), b = loadPos(m.verts, m.stride, i1), c = loadPos(m.verts, m.stride, i2);
        const float cx = (0.5f * (fminf(fminf(a.x, b.x), c.x) + fmaxf(fmaxf(a.x, b.x), c.x)) - bound.x) * invW;
        const float cy = (0.5f * (fminf(fminf(a.y, b.y), c.y) + fmaxf(fmaxf(a.y, b.y), c.y)) - bound.y) * invW;
        const float cz = (0.5f * (fminf(fminf(a.z, b.z), c.z) + fmaxf(fmaxf(a.z, b.z), c.z)) - bound.z) * invW;
        const uint32_t key = ((expandBits10(quantize10(cx)) << 2) | (expandBits10(quantize10(cy)) << 1) |
                              expandBits10(quantize10(cz))) >> keyShift;
        keys[k] = key;
        vals[k] = k;
        // neighbouring triangles share their leading digits: one shared-memory atomic per distinct digit of the warp
        const uint32_t active = __activemask();
        const uint32_t first = (uint32_t)__ffs(active) - 1u;
        for (int p = 0; p < numPasses; ++p)
        {
            const uint32_t dgt = (key >> (8 * p)) & 255u;
            // the upper digits of 32 neighbouring triangles usually agree: one vote instead of a match
            if (__all_sync(active, dgt == __shfl_sync(active, dgt, (int)first)))
            {
                if (laneId() == first) atomicAdd(&sh[p][dgt], (uint32_t)__popc(active));
                continue;
            }
            const uint32_t peers = __match_any_sync(active, dgt);
            if ((uint32_t)(__ffs(peers) - 1) == laneId()) atomicAdd(&sh[p][dgt], (uint32_t)__popc(peers));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < numPasses * 256; i += blockDim.x)
    {
        const uint32_t c = (&sh[0][0])[i];
        if (c) atomicAdd(hist + i, c);
    }
}

// ---- leaves and the box pyramid ----------------------------------------------------------------------
// Box of a range of sorted leaves, stored as two float4: (ylo, yhi, zlo, zhi), (xlo, xhi, -, -).
// Level 0 = the leaves, level l = unions of 16^l consecutive leaves.
struct RangeBox
{
    float ylo, yhi, zlo, zhi, xlo, xhi;
    __device__ __forceinline__ void clear()
    {
        ylo = zlo = xlo = INFINITY;
        yhi = zhi = xhi = -INFINITY;
    }
    __device__ __forceinline__ void add(const float4& yz, const float4& x)
    {
        ylo = fminf(ylo, yz.x); yhi = fmaxf(yhi, yz.y);
        zlo = fminf(zlo, yz.z); zhi = fmaxf(zhi, yz.w);
        xlo = fminf(xlo, x.x);  xhi = fmaxf(xhi, x.y);
    }
    __device__ __forceinline__ void unite(const RangeBox& o)
    {
        ylo = fminf(ylo, o.ylo); yhi = fmaxf(yhi, o.yhi);
        zlo = fminf(zlo, o.zlo); zhi = fmaxf(zhi, o.zhi);
        xlo = fminf(xlo, o.xlo); xhi = fmaxf(xhi, o.xhi);
    }
    // this box of the lane `delta` lanes below (up = true) / above within its group of 16 lanes
    __device__ __forceinline__ RangeBox shifted16(bool up, unsigned delta) const
    {
        RangeBox r;
        if (up)
        {
            r.ylo = __shfl_up_sync(0xffffffffu, ylo, delta, 16); r.yhi = __shfl_up_sync(0xffffffffu, yhi, delta, 16);
            r.zlo = __shfl_up_sync(0xffffffffu, zlo, delta, 16); r.zhi = __shfl_up_sync(0xffffffffu, zhi, delta, 16);
            r.xlo = __shfl_up_sync(0xffffffffu, xlo, delta, 16); r.xhi = __shfl_up_sync(0xffffffffu, xhi, delta, 16);
        }
        else
        {
            r.ylo = __shfl_down_sync(0xffffffffu, ylo, delta, 16); r.yhi = __shfl_down_sync(0xffffffffu, yhi, delta, 16);
            r.zlo = __shfl_down_sync(0xffffffffu, zlo, delta, 16); r.zhi = __shfl_down_sync(0xffffffffu, zhi, delta, 16);
            r.xlo = __shfl_down_sync(0xffffffffu, xlo, delta, 16); r.xhi = __shfl_down_sync(0xffffffffu, xhi, delta, 16);
        }
        return r;
    }
    __device__ __forceinline__ void store(float4* dst) const
    {
        dst[0] = make_float4(ylo, yhi, zlo, zhi);
        dst[1] = make_float4(xlo, xhi, 0.0f, 0.0f);
    }
};

// Entry i of a level: its box at [stride*i], and (stride == 6) the union of the entries from the start of
// its aligned group of 16 up to it ("prefix", at +2) and from it to the end of the group ("suffix", at +4).
// With those, the part of a contiguous range that falls into one group costs ONE entry to read whenever
// the range enters or leaves the group across its boundary.  The top level (<= 16 entries) has boxes only.
struct Pyramid
{
    float4* level[kMaxBoxLevels];
    uint32_t count[kMaxBoxLevels];  // entries per level
    int numLevels;
    uint32_t stride;                // float4 per entry: 6, or 2 (boxes only: large meshes, k_refit_atomic)
};

// inclusive prefix / suffix unions over each group of 16 consecutive lanes
__device__ __forceinline__ void groupScans16(const RangeBox& box, RangeBox& pre, RangeBox& suf)
{
    const uint32_t hl = laneId() & 15u;
    pre = box; suf = box;
#pragma unroll
    for (unsigned o = 1; o < 16u; o <<= 1)
    {
        const RangeBox a = pre.shifted16(true, o), b = suf.shifted16(false, o);
        if (hl >= o) pre.unite(a);
        if (hl + o < 16u) suf.unite(b);
    }
}

__global__ void __launch_bounds__(256)
k_leaf_setup(MeshView m, const float* __restrict__ boundPtr, const uint32_t* __restrict__ sortedPrims,
             Tri48* __restrict__ tris, Pyramid pyr, float* __restrict__ rootBox, uint32_t* __restrict__ err)
{
    __shared__ float sGroup[16][6];
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t S = pyr.stride;
    const bool scans = S == 6u;
    RangeBox box;
    box.clear();
    if (j < m.numTris)
    {
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
        box.ylo = fminf(fminf(a.y, b.y), c.y); box.yhi = fmaxf(fmaxf(a.y, b.y), c.y);
        box.zlo = fminf(fminf(a.z, b.z), c.z); box.zhi = fmaxf(fmaxf(a.z, b.z), c.z);
        box.xlo = fminf(fminf(a.x, b.x), c.x); box.xhi = fmaxf(fmaxf(a.x, b.x), c.x);
        if (S == 6u) box.store(pyr.level[0] + (size_t)S * j);   // (large meshes: k_refit_atomic derives the leaf box from the record)
        if (m.numTris == 1)
        {
            rootBox[0] = box.xlo; rootBox[1] = box.ylo; rootBox[2] = box.zlo;
            rootBox[3] = box.xhi; rootBox[4] = box.yhi; rootBox[5] = box.zhi;
        }
    }
    if (pyr.numLevels < 2) return;   // (uniform) a single level is read directly
    // level 0 scans and level 1: 16 consecutive leaves = one half-warp (leaves past the end are neutral boxes)
    RangeBox pre, suf;
    groupScans16(box, pre, suf);
    if (scans && j < m.numTris)
    {
        pre.store(pyr.level[0] + (size_t)S * j + 2);
        suf.store(pyr.level[0] + (size_t)S * j + 4);
    }
    const uint32_t group = threadIdx.x >> 4;   // 16 groups = 16 level-1 entries per block
    if ((threadIdx.x & 15u) == 15u)
    {
        // lane 15's prefix is the union of its group
        if (16u * blockIdx.x + group < pyr.count[1]) pre.store(pyr.level[1] + (size_t)S * (16u * blockIdx.x + group));
        sGroup[group][0] = pre.ylo; sGroup[group][1] = pre.yhi; sGroup[group][2] = pre.zlo;
        sGroup[group][3] = pre.zhi; sGroup[group][4] = pre.xlo; sGroup[group][5] = pre.xhi;
    }
    if (pyr.numLevels < 3) return;   // (uniform) level 1 is the top: boxes only
    __syncthreads();
    // level 1 scans and level 2: the block's 16 level-1 entries, by its first half-warp
    if (threadIdx.x < 32)
    {
        const uint32_t t = threadIdx.x & 15u;
        RangeBox g;
        g.ylo = sGroup[t][0]; g.yhi = sGroup[t][1]; g.zlo = sGroup[t][2]; g.zhi = sGroup[t][3]; g.xlo = sGroup[t][4]; g.xhi = sGroup[t][5];
        groupScans16(g, pre, suf);
        const uint32_t e = 16u * blockIdx.x + t;
        if (scans && threadIdx.x < 16 && e < pyr.count[1])
        {
            pre.store(pyr.level[1] + (size_t)S * e + 2);
            suf.store(pyr.level[1] + (size_t)S * e + 4);
        }
        if (threadIdx.x == 15) pre.store(pyr.level[2] + (size_t)S * blockIdx.x);
    }
}

// ---- the whole build of a small mesh in ONE kernel -------------------------------------------------------------------
// A 100 k-triangle build is six dependent launches (bounds, keys, two or three radix passes, leaves), each at its
// latency floor of 7-10 us although the data (a few MB) would stream through the machine in about one.  Here every
// triangle lives in the registers of one thread of a grid that is resident as a whole (cooperative launch, one CTA of
// 1024 threads per SM, up to kFusedRounds triangles per thread), and the phases are separated by grid barriers (~1.5 us)
// instead of kernel boundaries:
//   vertex min/max -> BARRIER -> {c, w}, keys, per-(round, warp) digit counts -> BARRIER -> digit bases from all CTAs'
//   counts, stable ranks, scatter -> BARRIER -> next radix pass ... -> the last pass writes the sorted (key, triangle)
//   pairs AND, when no hierarchy follows, the scene-space triangle records straight into their sorted slots.
// The sort is the same stable LSD radix sort over the same keys as k_morton + onesweep (identical output): CTA c owns
// the items [c * rounds * 1024, ...) of the current order, item = (round, warp, lane); a (round, warp) counts its
// digits with match_any, a 256-thread scan turns the counts into exclusive offsets inside the CTA, and the CTAs'
// totals go through global memory [pass][digit][cta].
constexpr int kFusedThreads = 1024;
constexpr int kFusedRounds = 2;
constexpr int kFusedMaxCtas = 160;                // CTAs of the grid (one per SM; a multiple of 32)
constexpr uint32_t kFusedSpinLimit = 1u << 21;   // polls of a barrier word (~1 us each) before the kernel gives up

struct FusedBuildParams
{
    MeshView m;
    float* bound;        // out: {c, w}
    float* partials;     // [gridDim.x][6]
    uint32_t* hist;      // [numPasses][256][gridDim.x]
    uint32_t* sync;      // [0] barrier arrivals, [1] CTAs that have finished; both zero between launches
    uint32_t *keysA, *valsA, *keysB, *valsB;   // the last pass lands in A
    Tri48* tris;         // null: sorted pairs only (k_leaf_setup follows)
    uint32_t* err;
    uint32_t keyShift, numPasses, rounds, haveBound;
    float bnd[4];
};

__device__ __forceinline__ uint32_t loadAcquire(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// all CTAs of the (co-resident) grid; false = gave up waiting (error raised, the caller returns)
__device__ __forceinline__ bool gridBarrier(uint32_t* sync, uint32_t& target, uint32_t* err, uint32_t* sOk)
{
    __syncthreads();
    if (threadIdx.x == 0)
    {
        target += gridDim.x;
        __threadfence();
        atomicAdd(sync, 1u);
        uint32_t spins = 0, ok = 1u;
        while (loadAcquire(sync) < target)
            if (++spins > kFusedSpinLimit) { ok = 0u; atomicMax(err, (uint32_t)kErrBarrierTimeout); break; }
        *sOk = ok;
        __threadfence();
    }
    __syncthreads();
    return *sOk != 0u;
}

__global__ void __launch_bounds__(kFusedThreads, 1)
k_build_fused(const FusedBuildParams p)
{
    __shared__ uint16_t sTab[kFusedRounds * 32][256];   // digit counts per (round, warp) -> exclusive offsets inside the CTA
    __shared__ uint32_t sBase[256];
    __shared__ uint32_t sSeg[4][256];
    __shared__ float sRed[32][6];
    __shared__ float sBound[4];
    __shared__ uint32_t sWarpTot[8];
    __shared__ uint32_t sOk;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, G = gridDim.x, cta = blockIdx.x;
    const uint32_t T = p.m.numTris, R = p.rounds;
    const uint32_t itemBase = cta * R * (uint32_t)kFusedThreads;
    const uint32_t lt = laneMaskLt();
    uint32_t target = 0;

    // ---- this thread's triangles: object-space box centres (issued before anything waits) ----
    float cx[kFusedRounds], cy[kFusedRounds], cz[kFusedRounds];
#pragma unroll
    for (int r = 0; r < kFusedRounds; ++r)
    {
        cx[r] = cy[r] = cz[r] = 0.0f;
        const uint32_t k = itemBase + (uint32_t)r * kFusedThreads + tid;
        if ((uint32_t)r < R && k < T)
        {
            uint32_t i0 = __ldg(p.m.indices + 3 * (size_t)k), i1 = __ldg(p.m.indices + 3 * (size_t)k + 1), i2 = __ldg(p.m.indices + 3 * (size_t)k + 2);
            if (i0 >= p.m.numVerts || i1 >= p.m.numVerts || i2 >= p.m.numVerts)
            {
                atomicMax(p.err, (uint32_t)kErrBadIndex);
                i0 = i1 = i2 = 0;
            }
            const float3 a = loadPos(p.m.verts, p.m.stride, i0), b = loadPos(p.m.verts, p.m.stride, i1), c = loadPos(p.m.verts, p.m.stride, i2);
            cx[r] = 0.5f * (fminf(fminf(a.x, b.x), c.x) + fmaxf(fmaxf(a.x, b.x), c.x));
            cy[r] = 0.5f * (fminf(fminf(a.y, b.y), c.y) + fmaxf(fmaxf(a.y, b.y), c.y));
            cz[r] = 0.5f * (fminf(fminf(a.z, b.z), c.z) + fmaxf(fmaxf(a.z, b.z), c.z));
        }
    }

    // ---- {c, w} (Voxelizer.cpp:52-57), over ALL vertices as the reference's computeAABB ----
    if (!p.haveBound)
    {
        float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (uint32_t i = cta * kFusedThreads + tid; i < p.m.numVerts; i += G * kFusedThreads)
        {
            const float3 q = loadPos(p.m.verts, p.m.stride, i);
            mn[0] = fminf(mn[0], q.x); mx[0] = fmaxf(mx[0], q.x);
            mn[1] = fminf(mn[1], q.y); mx[1] = fmaxf(mx[1], q.y);
            mn[2] = fminf(mn[2], q.z); mx[2] = fmaxf(mx[2], q.z);
        }
        warpMinMax(mn, mx);
        if (lane == 0)
            for (int a = 0; a < 3; ++a) { sRed[warp][a] = mn[a]; sRed[warp][3 + a] = mx[a]; }
        __syncthreads();
        if (warp == 0)
        {
            for (int a = 0; a < 3; ++a) { mn[a] = sRed[lane][a]; mx[a] = sRed[lane][3 + a]; }
            warpMinMax(mn, mx);
            if (lane == 0)
                for (int a = 0; a < 3; ++a) { p.partials[6 * cta + a] = mn[a]; p.partials[6 * cta + 3 + a] = mx[a]; }
        }
        if (!gridBarrier(p.sync, target, p.err, &sOk)) return;
        if (warp == 0)
        {
            for (int a = 0; a < 3; ++a) { mn[a] = INFINITY; mx[a] = -INFINITY; }
            for (uint32_t b = lane; b < G; b += 32u)
                for (int a = 0; a < 3; ++a)
                {
                    mn[a] = fminf(mn[a], __ldcg(p.partials + 6 * b + a));
                    mx[a] = fmaxf(mx[a], __ldcg(p.partials + 6 * b + 3 + a));
                }
            warpMinMax(mn, mx);
            if (lane == 0)
            {
                const float ex = __fsub_rn(mx[0], mn[0]), ey = __fsub_rn(mx[1], mn[1]), ez = __fsub_rn(mx[2], mn[2]);
                sBound[0] = __fdiv_rn(__fadd_rn(mx[0], mn[0]), 2.0f);
                sBound[1] = __fdiv_rn(__fadd_rn(mx[1], mn[1]), 2.0f);
                sBound[2] = __fdiv_rn(__fadd_rn(mx[2], mn[2]), 2.0f);
                float m = ey > ez ? ey : ez;
                m = ex > m ? ex : m;
                sBound[3] = __fdiv_rn(m, 2.0f);
            }
        }
    }
    else if (tid < 4) sBound[tid] = p.bnd[tid];
    __syncthreads();
    const float4 bound = make_float4(sBound[0], sBound[1], sBound[2], sBound[3]);
    if (cta == 0 && tid < 4) p.bound[tid] = sBound[tid];

    // ---- keys: exactly k_morton's ----
    uint32_t key[kFusedRounds], val[kFusedRounds], rank[kFusedRounds];
    {
        const float invW = 1.0f / bound.w;
#pragma unroll
        for (int r = 0; r < kFusedRounds; ++r)
        {
            const float x = (cx[r] - bound.x) * invW, y = (cy[r] - bound.y) * invW, z = (cz[r] - bound.z) * invW;
            key[r] = ((expandBits10(quantize10(x)) << 2) | (expandBits10(quantize10(y)) << 1) | expandBits10(quantize10(z))) >> p.keyShift;
            val[r] = itemBase + (uint32_t)r * kFusedThreads + tid;
            rank[r] = 0;
        }
    }

    // ---- stable LSD radix passes ----
    for (uint32_t pass = 0; pass < p.numPasses; ++pass)
    {
        const bool last = pass + 1u == p.numPasses;
        const bool outA = ((p.numPasses - 1u - pass) & 1u) == 0u;
        const uint32_t* keysIn = outA ? p.keysB : p.keysA;
        const uint32_t* valsIn = outA ? p.valsB : p.valsA;
        uint32_t* keysOut = outA ? p.keysA : p.keysB;
        uint32_t* valsOut = outA ? p.valsA : p.valsB;
        const uint32_t shift = 8u * pass;
        for (uint32_t i = tid; i < (uint32_t)(sizeof(sTab) / sizeof(uint4)); i += kFusedThreads) reinterpret_cast<uint4*>(&sTab[0][0])[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kFusedRounds; ++r)
        {
            const uint32_t i = itemBase + (uint32_t)r * kFusedThreads + tid;
            const bool valid = (uint32_t)r < R && i < T;
            if (pass > 0u && valid) { key[r] = __ldcg(keysIn + i); val[r] = __ldcg(valsIn + i); }
            const uint32_t vm = __ballot_sync(0xffffffffu, valid);
            if (valid)
            {
                const uint32_t d = (key[r] >> shift) & 255u;
                const uint32_t peers = __match_any_sync(vm, d);
                rank[r] = (uint32_t)__popc(peers & lt);
                if (rank[r] == 0u) sTab[r * 32 + (int)warp][d] = (uint16_t)__popc(peers);
            }
        }
        __syncthreads();
        {
            // counts -> exclusive offsets inside the CTA, per digit over the (round, warp) rows in order: four threads per
            // digit scan a quarter of the rows each, then add the quarters before theirs
            const uint32_t d = tid & 255u, q = tid >> 8, rows = R * 8u;
            uint32_t run = 0;
            for (uint32_t vw = q * rows; vw < (q + 1u) * rows; ++vw)
            {
                const uint32_t c = sTab[vw][d];
                sTab[vw][d] = (uint16_t)run;
                run += c;
            }
            sSeg[q][d] = run;
            __syncthreads();
            uint32_t add = 0, all = 0;
#pragma unroll
            for (uint32_t k = 0; k < 4u; ++k) { const uint32_t v = sSeg[k][d]; add += k < q ? v : 0u; all += v; }
            if (q > 0u)
                for (uint32_t vw = q * rows; vw < (q + 1u) * rows; ++vw) sTab[vw][d] = (uint16_t)(sTab[vw][d] + add);
            if (q == 0u) p.hist[((size_t)pass * 256u + d) * G + cta] = all;
        }
        if (!gridBarrier(p.sync, target, p.err, &sOk)) return;
        {
            // digit bases: warp w takes the digits w, w + 32, ...; its lanes read the counts of all CTAs for a digit with
            // independent coalesced loads ([pass][digit][cta]) and fold them with warp reductions
            uint32_t v[8][kFusedMaxCtas / 32];
#pragma unroll
            for (uint32_t k = 0; k < 8u; ++k)
#pragma unroll
                for (uint32_t j = 0; j < (uint32_t)(kFusedMaxCtas / 32); ++j)
                {
                    const uint32_t c = j * 32u + lane;
                    v[k][j] = c < G ? __ldcg(p.hist + ((size_t)pass * 256u + (k * 32u + warp)) * G + c) : 0u;
                }
#pragma unroll
            for (uint32_t k = 0; k < 8u; ++k)
            {
                uint32_t total = 0, before = 0;
#pragma unroll
                for (uint32_t j = 0; j < (uint32_t)(kFusedMaxCtas / 32); ++j) { total += v[k][j]; before += (j * 32u + lane < cta) ? v[k][j] : 0u; }
                total = __reduce_add_sync(0xffffffffu, total);
                before = __reduce_add_sync(0xffffffffu, before);
                if (lane == 0) { sSeg[0][k * 32u + warp] = total; sSeg[1][k * 32u + warp] = before; }
            }
        }
        __syncthreads();
        if (tid < 256u)
        {
            const uint32_t total = sSeg[0][tid], before = sSeg[1][tid];
            uint32_t incl = total;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= (uint32_t)o) incl += v;
            }
            if (lane == 31u) sWarpTot[warp] = incl;
            sBase[tid] = incl - total + before;   // (the earlier warps' totals are added below)
        }
        __syncthreads();
        if (tid < 256u)
        {
            uint32_t add = 0;
            for (uint32_t w = 0; w < warp; ++w) add += sWarpTot[w];
            sBase[tid] += add;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kFusedRounds; ++r)
        {
            const uint32_t i = itemBase + (uint32_t)r * kFusedThreads + tid;
            if ((uint32_t)r < R && i < T)
            {
                const uint32_t d = (key[r] >> shift) & 255u;
                const uint32_t dst = sBase[d] + sTab[r * 32 + (int)warp][d] + rank[r];
                keysOut[dst] = key[r];
                valsOut[dst] = val[r];
                if (last && p.tris)
                {
                    // the sorted leaf record (k_leaf_setup's, without the box pyramid nobody reads in this mode)
                    const uint32_t k = val[r];
                    uint32_t i0 = __ldg(p.m.indices + 3 * (size_t)k), i1 = __ldg(p.m.indices + 3 * (size_t)k + 1), i2 = __ldg(p.m.indices + 3 * (size_t)k + 2);
                    if (i0 >= p.m.numVerts || i1 >= p.m.numVerts || i2 >= p.m.numVerts) i0 = i1 = i2 = 0;   // (error raised above)
                    const float3 a = scenePos(p.m.verts, p.m.stride, i0, bound);
                    const float3 b = scenePos(p.m.verts, p.m.stride, i1, bound);
                    const float3 c = scenePos(p.m.verts, p.m.stride, i2, bound);
                    Tri48 t;
                    t.a = make_float4(a.x, a.y, a.z, __uint_as_float(k));
                    t.b = make_float4(b.x, b.y, b.z, 0.0f);
                    t.c = make_float4(c.x, c.y, c.z, 0.0f);
                    p.tris[dst] = t;
                }
            }
        }
        if (!last && !gridBarrier(p.sync, target, p.err, &sOk)) return;
    }

    // ---- leave the barrier words zero for the next launch: the last CTA to get here knows everybody is past them ----
    __syncthreads();
    if (tid == 0)
    {
        __threadfence();
        if (atomicAdd(p.sync + 1, 1u) == G - 1u) { p.sync[0] = 0u; p.sync[1] = 0u; }
    }
}

// level l (>= 3) from level l-1: one thread per entry, 16 children each; also the children's scans
__global__ void __launch_bounds__(128)
k_box_level(float4* __restrict__ src, uint32_t srcCount, float4* __restrict__ dst, uint32_t dstCount, uint32_t S)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dstCount) return;
    // all 32 loads are issued before the first min/max: one memory round trip, not sixteen
    float4 yz[16], xx[16];
#pragma unroll
    for (uint32_t q = 0; q < 16u; ++q)
    {
        const uint32_t c = min(16u * i + q, srcCount - 1u);   // clamped duplicates do not change a union
        yz[q] = __ldg(src + (size_t)S * c);
        xx[q] = __ldg(src + (size_t)S * c + 1);
    }
    RangeBox run;
    run.clear();
#pragma unroll
    for (uint32_t q = 0; q < 16u; ++q)
    {
        run.add(yz[q], xx[q]);
        if (S == 6u && 16u * i + q < srcCount) run.store(src + (size_t)S * (16u * i + q) + 2);
    }
    run.store(dst + (size_t)S * i);
    if (S != 6u) return;
    run.clear();
#pragma unroll
    for (int q = 15; q >= 0; --q)
    {
        run.add(yz[q], xx[q]);
        if (16u * i + (uint32_t)q < srcCount) run.store(src + (size_t)S * (16u * i + (uint32_t)q) + 4);
    }
}

// union of `n` (<= 16) consecutive entries starting at `first`; loads are issued in batches of 8 entries
// (16 independent 128-bit loads in flight) instead of one dependent round trip per entry
__device__ __forceinline__ void addRun(RangeBox& box, const float4* __restrict__ lev, uint32_t S, uint32_t first, uint32_t n)
{
    if (n <= 2u)   // most nodes are tiny: do not pay for a padded batch
    {
        if (n >= 1u) box.add(__ldg(lev + (size_t)S * first), __ldg(lev + (size_t)S * first + 1));
        if (n == 2u) box.add(__ldg(lev + (size_t)S * (first + 1u)), __ldg(lev + (size_t)S * (first + 1u) + 1));
        return;
    }
#pragma unroll 1
    for (uint32_t base = 0; base < n; base += 8u)
    {
        float4 yz[8], xx[8];
#pragma unroll
        for (uint32_t q = 0; q < 8u; ++q)
        {
            const uint32_t c = first + min(base + q, n - 1u);   // clamped duplicates do not change a union
            yz[q] = __ldg(lev + (size_t)S * c);
            xx[q] = __ldg(lev + (size_t)S * c + 1);
        }
#pragma unroll
        for (uint32_t q = 0; q < 8u; ++q) box.add(yz[q], xx[q]);
    }
}

// Union of the leaf boxes [first, last].  Per level the range is cut at the multiples of 16: the piece before
// the first cut is the SUFFIX of entry lo, the piece after the last cut the PREFIX of entry hi-1, the whole
// groups between are entries of the next level.  Two loads per level, none depending on another (the adds of
// a level are delayed until the next level's loads are issued), and at most 16 direct entries at the end.
__device__ __forceinline__ RangeBox rangeQuery(const Pyramid& pyr, uint32_t first, uint32_t last)
{
    RangeBox box;
    box.clear();
    const uint32_t S = pyr.stride;
    const float4 none0 = make_float4(INFINITY, -INFINITY, INFINITY, -INFINITY), none1 = make_float4(INFINITY, -INFINITY, 0.0f, 0.0f);
    float4 pYz0 = none0, pXx0 = none1, pYz1 = none0, pXx1 = none1;   // loads in flight
    uint32_t lo = first, hi = last + 1u;  // half open, in units of level-l entries
    for (int l = 0; l < pyr.numLevels && lo < hi; ++l)
    {
        const float4* lev = pyr.level[l];
        const bool top = l == pyr.numLevels - 1;
        if (top || hi - lo <= 15u)
        {
            if (!top && S == 6u && (lo >> 4) != ((hi - 1u) >> 4))
            {
                // two neighbouring groups: suffix of lo + prefix of hi-1
                const float4 a0 = __ldg(lev + (size_t)S * lo + 4), a1 = __ldg(lev + (size_t)S * lo + 5);
                const float4 b0 = __ldg(lev + (size_t)S * (hi - 1u) + 2), b1 = __ldg(lev + (size_t)S * (hi - 1u) + 3);
                box.add(a0, a1); box.add(b0, b1);
            }
            else addRun(box, lev, S, lo, hi - lo);
            break;
        }
        const uint32_t head = (16u - (lo & 15u)) & 15u;   // entries up to the next multiple of 16
        const uint32_t tail = hi & 15u;                    // entries after the last multiple of 16
        float4 nYz0 = none0, nXx0 = none1, nYz1 = none0, nXx1 = none1;
        if (S == 6u)
        {
            if (head) { nYz0 = __ldg(lev + (size_t)S * lo + 4); nXx0 = __ldg(lev + (size_t)S * lo + 5); }
            if (tail) { nYz1 = __ldg(lev + (size_t)S * (hi - 1u) + 2); nXx1 = __ldg(lev + (size_t)S * (hi - 1u) + 3); }
        }
        else
        {
            addRun(box, lev, S, lo, head);
            addRun(box, lev, S, hi - tail, tail);
        }
        box.add(pYz0, pXx0); box.add(pYz1, pXx1);
        pYz0 = nYz0; pXx0 = nXx0; pYz1 = nYz1; pXx1 = nXx1;
        lo = (lo + head) >> 4; hi = (hi - tail) >> 4;
    }
    box.add(pYz0, pXx0); box.add(pYz1, pXx1);
    return box;
}

// ---- Karras hierarchy + child boxes ----------------------------------------------------------------
// delta(i,j): length of the common prefix of the 64-bit augmented keys (key << 32 | index), -1 when
// j is out of range.
__device__ __forceinline__ int delta(const uint32_t* __restrict__ keys, int numLeaves, uint32_t ki, int i, int j)
{
    if (j < 0 || j >= numLeaves) return -1;
    const uint32_t kj = __ldg(keys + j);
    const uint32_t x = ki ^ kj;
    return x ? __clz(x) : 32 + __clz((uint32_t)i ^ (uint32_t)j);
}

// Topology only (child references + leaf range of every node): it needs nothing but the sorted keys, so it
// runs concurrently with k_leaf_setup / k_box_level; the child boxes follow in k_node_boxes or k_refit_atomic.
// kLatency = true : small meshes, where everything is latency bound (100 k triangles): multi-probe searches
// kLatency = false: millions of nodes, mostly tiny ranges, throughput counts: classic one-probe-per-step searches,
//                   and parent references for k_refit_atomic (O(T) traffic instead of O(T log T) pyramid reads)
template <bool kLatency>
__global__ void __launch_bounds__(128, 8)
k_hierarchy_topology(const uint32_t* __restrict__ keys, int numLeaves, BvhNode* __restrict__ nodes,
                     uint32_t* __restrict__ nodeParent, uint32_t* __restrict__ leafParent)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numLeaves - 1) return;
    const uint32_t ki = __ldg(keys + i);
    const int d = (delta(keys, numLeaves, ki, i, i + 1) - delta(keys, numLeaves, ki, i, i - 1)) < 0 ? -1 : 1;
    const int dMin = delta(keys, numLeaves, ki, i, i - d);
    // The three searches below are chains of dependent key loads (every probe is an L2 round trip
    // for the nodes near the root), so each round issues several independent probes at once.
    // delta(i, i + l*d) > bound is monotone in l: true up to the end of the node's range, false beyond.
    auto inRange = [&](long long l, int bound) {
        const long long q = (long long)i + l * d;
        return q >= 0 && q < numLeaves && delta(keys, numLeaves, ki, i, (int)q) > bound;
    };
    long long l = 0, s = 0;
    int j, dNode;
    if (kLatency)
    {
        // latency regime (small meshes): several independent probes per round
        long long lMax = 2;
        while (true)
        {
            const bool p0 = inRange(lMax, dMin), p1 = inRange(lMax * 2, dMin), p2 = inRange(lMax * 4, dMin), p3 = inRange(lMax * 8, dMin);
            if (!p0) break;
            if (!p1) { lMax *= 2; break; }
            if (!p2) { lMax *= 4; break; }
            if (!p3) { lMax *= 8; break; }
            lMax *= 16;
        }
        // largest l < lMax with inRange(l): two bits per round
        l = 0;
        long long t = lMax >> 1;
        while (t >= 1)
        {
            const long long h = t >> 1;
            const bool a = inRange(l + t, dMin);
            const bool b0 = h >= 1 && inRange(l + h, dMin), b1 = h >= 1 && inRange(l + t + h, dMin);
            if (a) { l += t; if (b1) l += h; }
            else if (b0) l += h;
            t = h >> 1;
            if (h < 1) break;
        }
        j = i + (int)l * d;
        dNode = delta(keys, numLeaves, ki, i, j);
        // split: largest s in [0, l) with delta(i, i + s*d) > dNode, by the same two-bits-per-round search
        // over the power-of-two ladder ceil(l/2), ceil(l/4), ... , 1
        s = 0;
        t = l;
        do
        {
            t = (t + 1) >> 1;
            const long long t2 = (t > 1) ? ((t + 1) >> 1) : 0;
            const bool a = inRange(s + t, dNode);
            const bool b0 = t2 >= 1 && inRange(s + t2, dNode), b1 = t2 >= 1 && inRange(s + t + t2, dNode);
            if (a) { s += t; if (t2 >= 1 && b1) s += t2; }
            else if (t2 >= 1 && b0) s += t2;
            if (t2 >= 1) t = t2;
        } while (t > 1);
    }
    else
    {
        // throughput regime (millions of nodes, mostly tiny ranges): the classic one-probe-per-step searches
        long long lMax = 2;
        while (inRange(lMax, dMin)) lMax <<= 1;
        for (long long t = lMax >> 1; t >= 1; t >>= 1)
            if (inRange(l + t, dMin)) l += t;
        j = i + (int)l * d;
        dNode = delta(keys, numLeaves, ki, i, j);
        long long t = l;
        do
        {
            t = (t + 1) >> 1;
            if (inRange(s + t, dNode)) s += t;
        } while (t > 1);
    }
    const int gamma = i + (int)s * d + min(d, 0);

    // node i covers the sorted leaves [first, last]; its children cover [first, gamma] and [gamma+1, last]
    const int first = min(i, j), last = max(i, j);
    const bool leftLeaf = (first == gamma), rightLeaf = (last == gamma + 1);
    const uint32_t c0 = leftLeaf ? (kLeafFlag | (uint32_t)gamma) : (uint32_t)gamma;
    const uint32_t c1 = rightLeaf ? (kLeafFlag | (uint32_t)(gamma + 1)) : (uint32_t)(gamma + 1);
    float4* dst = reinterpret_cast<float4*>(nodes + i);
    // child references + the node's leaf range (debug / tests)
    dst[3] = make_float4(__uint_as_float(c0), __uint_as_float(c1), __uint_as_float((uint32_t)first), __uint_as_float((uint32_t)last));
    if (!kLatency)
    {
        // parent reference: node index | (1u << 31 when the child is the right one)
        if (leftLeaf) leafParent[gamma] = (uint32_t)i; else nodeParent[gamma] = (uint32_t)i;
        if (rightLeaf) leafParent[gamma + 1] = (uint32_t)i | 0x80000000u; else nodeParent[gamma + 1] = (uint32_t)i | 0x80000000u;
    }
}

// ---- child boxes by range union from the pyramid (small meshes: no dependency chain between nodes) ------
__device__ __forceinline__ void storeChildBoxes(BvhNode* node, const RangeBox& b0, const RangeBox& b1, bool isRoot, float* rootBox)
{
    float4* dst = reinterpret_cast<float4*>(node);
    dst[0] = make_float4(b0.ylo, b0.yhi, b0.zlo, b0.zhi);
    dst[1] = make_float4(b1.ylo, b1.yhi, b1.zlo, b1.zhi);
    dst[2] = make_float4(b0.xlo, b0.xhi, b1.xlo, b1.xhi);
    if (isRoot)
    {
        rootBox[0] = fminf(b0.xlo, b1.xlo); rootBox[1] = fminf(b0.ylo, b1.ylo); rootBox[2] = fminf(b0.zlo, b1.zlo);
        rootBox[3] = fmaxf(b0.xhi, b1.xhi); rootBox[4] = fmaxf(b0.yhi, b1.yhi); rootBox[5] = fmaxf(b0.zhi, b1.zhi);
    }
}

__global__ void __launch_bounds__(128, 4)
k_node_boxes(int numLeaves, BvhNode* __restrict__ nodes, const __grid_constant__ Pyramid pyr, float* __restrict__ rootBox)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numLeaves - 1) return;
    // {c0, c1, first, last} as written by k_hierarchy_topology; c0 = gamma (| leaf flag)
    const uint4 t = *reinterpret_cast<const uint4*>(reinterpret_cast<const float4*>(nodes + i) + 3);
    const uint32_t gamma = t.x & ~kLeafFlag, first = t.z, last = t.w;
    const uint32_t n = last - first + 1u;
    if (n <= 16u)
    {
        // 19 nodes in 20: few leaves -- read them once, in one or two batches, and deal them to the two children
        const float4* lev = pyr.level[0];
        const uint32_t S = pyr.stride, n0 = gamma - first + 1u;   // leaves of child 0
        RangeBox b0, b1;
        b0.clear(); b1.clear();
#pragma unroll 1
        for (uint32_t base = 0; base < n; base += 8u)
        {
            float4 yz[8], xx[8];
#pragma unroll
            for (uint32_t q = 0; q < 8u; ++q)
            {
                const uint32_t c = first + min(base + q, n - 1u);   // clamped duplicates do not change a union
                yz[q] = __ldg(lev + (size_t)S * c);
                xx[q] = __ldg(lev + (size_t)S * c + 1);
            }
#pragma unroll
            for (uint32_t q = 0; q < 8u; ++q)
            {
                if (min(base + q, n - 1u) < n0) b0.add(yz[q], xx[q]);
                else b1.add(yz[q], xx[q]);
            }
        }
        storeChildBoxes(nodes + i, b0, b1, i == 0, rootBox);
        return;
    }
    storeChildBoxes(nodes + i, rangeQuery(pyr, first, gamma), rangeQuery(pyr, gamma + 1u, last), i == 0, rootBox);
}

// ---- bottom-up refit with one atomic per node (large meshes) -----------------------------------------
// Every leaf starts with its box (pyramid level 0), stores it into its parent's slot, and the SECOND child
// to arrive at a node (atomic arrival counter, zeroed per build) unions both boxes and carries on upwards.
__global__ void __launch_bounds__(256)
k_refit_atomic(int numLeaves, const Tri48* __restrict__ tris, BvhNode* __restrict__ nodes,
               const uint32_t* __restrict__ nodeParent, const uint32_t* __restrict__ leafParent, uint32_t* __restrict__ flags,
               float* __restrict__ rootBox, uint32_t* __restrict__ err)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= numLeaves) return;
    const float4* t = reinterpret_cast<const float4*>(tris + j);
    const float4 a = __ldg(t), b = __ldg(t + 1), c = __ldg(t + 2);
    float4 yz = make_float4(fminf(fminf(a.y, b.y), c.y), fmaxf(fmaxf(a.y, b.y), c.y), fminf(fminf(a.z, b.z), c.z), fmaxf(fmaxf(a.z, b.z), c.z));
    float4 xx = make_float4(fminf(fminf(a.x, b.x), c.x), fmaxf(fmaxf(a.x, b.x), c.x), 0.0f, 0.0f);
    uint32_t p = __ldg(leafParent + j);
    for (int level = 0; level < 128; ++level)  // depth <= 62; the bound only guards against a corrupt tree
    {
        const uint32_t pi = p & 0x7fffffffu, slot = p >> 31;
        const uint32_t up = (pi != 0u) ? __ldg(nodeParent + pi) : 0u;  // issued early: overlaps the atomic
        BvhNode* n = nodes + pi;
        *(slot ? &n->yz1 : &n->yz0) = yz;
        *(reinterpret_cast<float2*>(&n->x01) + slot) = make_float2(xx.x, xx.y);
        __threadfence();  // release: the box must be visible before the arrival counter moves
        if (atomicAdd(flags + pi, 1u) == 0u) return;  // first child to arrive: the sibling will carry on
        __threadfence();  // acquire: order the loads below after the arrival counter (PTX memory model)
        // second arrival: the sibling's box was released before its increment; read it past L1
        const float4 syz = __ldcg(slot ? &n->yz0 : &n->yz1);
        const float2 sxx = __ldcg(reinterpret_cast<const float2*>(&n->x01) + (1u - slot));
        yz = make_float4(fminf(yz.x, syz.x), fmaxf(yz.y, syz.y), fminf(yz.z, syz.z), fmaxf(yz.w, syz.w));
        xx.x = fminf(xx.x, sxx.x); xx.y = fmaxf(xx.y, sxx.y);
        if (pi == 0)
        {
            rootBox[0] = xx.x; rootBox[1] = yz.x; rootBox[2] = yz.z;
            rootBox[3] = xx.y; rootBox[4] = yz.y; rootBox[5] = yz.w;
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

// One level of the occupancy pyramid: a voxel of the coarser level is set when any of its 2x2x2
// children is (conservative: safe for empty-space skipping).  One thread per destination word:
// 64 source voxels along x (two words) of 2 rows and 2 layers are OR-ed and pair-compacted.
__device__ __forceinline__ uint32_t compactPairs(uint32_t x)
{
    x = (x | (x >> 1)) & 0x55555555u;          // OR of each bit pair, kept in the even bit
    x = (x | (x >> 1)) & 0x33333333u;
    x = (x | (x >> 2)) & 0x0f0f0f0fu;
    x = (x | (x >> 4)) & 0x00ff00ffu;
    x = (x | (x >> 8)) & 0x0000ffffu;
    return x;
}

__global__ void __launch_bounds__(256)
k_mip_reduce(const uint32_t* __restrict__ src, uint32_t Ns, uint32_t Ps, uint32_t* __restrict__ dst, uint32_t Nd,
             uint32_t Pd, uint32_t layersDst)
{
    const size_t total = (size_t)layersDst * Nd * Pd;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    {
        const uint32_t w = (uint32_t)(i % Pd);
        const size_t row = i / Pd;
        const uint32_t y = (uint32_t)(row % Nd), z = (uint32_t)(row / Nd);
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (uint32_t dz = 0; dz < 2; ++dz)
#pragma unroll
            for (uint32_t dy = 0; dy < 2; ++dy)
            {
                const uint32_t* r = src + ((size_t)(2 * z + dz) * Ns + (2 * y + dy)) * Ps;
                if (2 * w < Ps) lo |= __ldg(r + 2 * w);
                if (2 * w + 1 < Ps) hi |= __ldg(r + 2 * w + 1);
            }
        dst[i] = compactPairs(lo) | (compactPairs(hi) << 16);
    }
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

void launchMorton(cudaStream_t s, const MeshView& m, const float* dBound, uint32_t* keys, uint32_t* vals,
                  uint32_t keyShift, int numPasses, uint32_t* hist, uint32_t* dErr)
{
    if (!m.numTris) return;
    const uint32_t blocks = std::min<uint32_t>((m.numTris + 255) / 256, 148u * 8u);
    k_morton<<<blocks, 256, 0, s>>>(m, dBound, keys, vals, keyShift, numPasses, hist, dErr);
}

size_t fusedBuildScratchBytes(int smCount) { return sizeof(uint32_t) * (16 + (size_t)kMaxFusedPasses * (size_t)smCount * 256u); }

bool fusedBuildPlan(uint32_t numTris, int smCount, uint32_t& ctas, uint32_t& rounds)
{
    if (numTris < 2 || smCount < 1) return false;
    ctas = std::min<uint32_t>(std::min<uint32_t>((uint32_t)smCount, kFusedMaxCtas), (numTris + kFusedThreads - 1) / kFusedThreads);
    rounds = (numTris + ctas * kFusedThreads - 1) / (ctas * kFusedThreads);
    return rounds <= (uint32_t)kFusedRounds;
}

bool fusedBuildSupported(int device)
{
    int coop = 0, perSm = 0;
    if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device) != cudaSuccess || !coop) { cudaGetLastError(); return false; }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_build_fused, kFusedThreads, 0) != cudaSuccess || perSm < 1) { cudaGetLastError(); return false; }
    return true;
}

bool launchFusedBuild(cudaStream_t s, const MeshView& m, int smCount, const float* bnd, float* dBound, float* dPartials, void* scratch,
                      uint32_t* keysA, uint32_t* valsA, uint32_t* keysB, uint32_t* valsB, Tri48* tris, uint32_t keyShift, int numPasses, uint32_t* dErr)
{
    FusedBuildParams p;
    uint32_t ctas = 0, rounds = 0;
    if (!fusedBuildPlan(m.numTris, smCount, ctas, rounds) || numPasses < 1 || numPasses > kMaxFusedPasses) return false;
    p.m = m; p.bound = dBound; p.partials = dPartials;
    p.sync = static_cast<uint32_t*>(scratch); p.hist = p.sync + 16;
    p.keysA = keysA; p.valsA = valsA; p.keysB = keysB; p.valsB = valsB; p.tris = tris; p.err = dErr;
    p.keyShift = keyShift; p.numPasses = (uint32_t)numPasses; p.rounds = rounds; p.haveBound = bnd ? 1u : 0u;
    for (int i = 0; i < 4; ++i) p.bnd[i] = bnd ? bnd[i] : 0.0f;
    void* args[] = {&p};
    const cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(k_build_fused), dim3(ctas), dim3(kFusedThreads), args, 0, s);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return true;
}

size_t boxPyramidFloat4s(uint32_t numTris)
{
    size_t total = 0;
    uint32_t c = numTris ? numTris : 1;
    for (int l = 0; l < kMaxBoxLevels; ++l)
    {
        total += 6 * (size_t)c + 6;
        if (c <= 16) break;
        c = (c + 15) / 16;
    }
    return total;
}

int launchLeavesAndHierarchy(cudaStream_t s, const SideStream* side, const MeshView& m, const float* dBound, const uint32_t* sortedKeys,
                             const uint32_t* sortedPrims, BvhNode* nodes, Tri48* tris, float4* pyramidMem,
                             uint32_t* refitScratch, float* rootBox, uint32_t* dErr, int parts)
{
    if (!m.numTris) return 0;
    const bool doLeaves = (parts & kBuildLeaves) != 0, doTree = (parts & kBuildTree) != 0;
    const bool atomicRefit = useAtomicRefit(m.numTris);
    Pyramid pyr;
    uint32_t c = m.numTris;
    float4* p = pyramidMem;
    pyr.numLevels = 0;
    pyr.stride = atomicRefit ? 2u : 6u;
    for (int l = 0; l < kMaxBoxLevels; ++l)
    {
        pyr.level[l] = p; pyr.count[l] = c; pyr.numLevels = l + 1;
        p += (size_t)pyr.stride * c + pyr.stride;
        if (c <= 16 || atomicRefit) break;   // the bottom-up refit only needs the leaf boxes (level 0)
        c = (c + 15) / 16;
    }
    for (int l = pyr.numLevels; l < kMaxBoxLevels; ++l) { pyr.level[l] = nullptr; pyr.count[l] = 0; }
    int launches = 0;
    // Two independent chains once the keys are sorted: {leaves, triangle records, box pyramid} and {topology}.
    // With a side stream they run concurrently (also under stream capture: the events become graph edges).
    const bool forked = side && side->stream && m.numTris > 1 && doTree && (doLeaves || pyr.numLevels > 3);
    cudaStream_t sb = forked ? side->stream : s;
    if (forked)
    {
        cudaEventRecord(side->fork, s);
        cudaStreamWaitEvent(sb, side->fork, 0);
    }
    if (doLeaves)
    {
        k_leaf_setup<<<(m.numTris + 255) / 256, 256, 0, sb>>>(m, dBound, sortedPrims, tris, pyr, rootBox, dErr);
        ++launches;
    }
    if (doTree)   // the coarser pyramid levels serve k_node_boxes only
        for (int l = 3; l < pyr.numLevels; ++l)
        {
            k_box_level<<<(pyr.count[l] + 127) / 128, 128, 0, sb>>>(pyr.level[l - 1], pyr.count[l - 1], pyr.level[l], pyr.count[l], pyr.stride);
            ++launches;
        }
    if (forked) cudaEventRecord(side->join, sb);
    if (m.numTris > 1 && doTree)
    {
        const uint32_t blocks = (m.numTris - 1 + 127) / 128;
        // refitScratch = [nodeParent T][leafParent T][flags T]
        uint32_t* nodeParent = refitScratch;
        uint32_t* leafParent = refitScratch + m.numTris;
        uint32_t* flags = refitScratch + 2 * (size_t)m.numTris;
        if (!atomicRefit) k_hierarchy_topology<true><<<blocks, 128, 0, s>>>(sortedKeys, (int)m.numTris, nodes, nullptr, nullptr);
        else
        {
            cudaMemsetAsync(flags, 0, sizeof(uint32_t) * m.numTris, s);
            k_hierarchy_topology<false><<<blocks, 128, 0, s>>>(sortedKeys, (int)m.numTris, nodes, nodeParent, leafParent);
        }
        if (forked) cudaStreamWaitEvent(s, side->join, 0);
        if (!atomicRefit) k_node_boxes<<<blocks, 128, 0, s>>>((int)m.numTris, nodes, pyr, rootBox);
        else
            k_refit_atomic<<<(m.numTris + 255) / 256, 256, 0, s>>>((int)m.numTris, tris, nodes, nodeParent, leafParent, flags,
                                                                   rootBox, dErr);
        launches += 2;
    }
    return launches;
}

bool useAtomicRefit(uint32_t numTris) { return numTris > (1u << 19); }

void launchPopcount(cudaStream_t s, const uint32_t* words, size_t numWords, unsigned long long* dCount)
{
    cudaMemsetAsync(dCount, 0, sizeof(unsigned long long), s);
    if (!numWords) return;
    size_t blocks = (numWords + 256 * 8 - 1) / (256 * 8);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_popcount<<<(unsigned)blocks, 256, 0, s>>>(words, numWords, dCount);
}

void launchMipReduce(cudaStream_t s, const uint32_t* src, uint32_t Ns, uint32_t layersSrc, uint32_t* dst)
{
    const uint32_t Nd = Ns / 2, layersDst = layersSrc / 2;
    const size_t total = (size_t)layersDst * Nd * ((Nd + 31) / 32);
    if (!total) return;
    size_t blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_mip_reduce<<<(unsigned)blocks, 256, 0, s>>>(src, Ns, (Ns + 31) / 32, dst, Nd, (Nd + 31) / 32, layersDst);
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
