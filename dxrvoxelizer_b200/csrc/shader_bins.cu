// shader_bins.cu -- MODE_SHADER on sm_100a: direction bins, an EXACT accelerator for the reference's ray family.
//
// The reference casts one ray per voxel with origin = voxel centre and direction = normalize(origin)
// (Content/Shaders/DXRVoxelizer.hlsl:44-53): every ray lies on a line through the grid centre.  B200 has no RT
// cores, and 2^30 incoherent-origin rays through a software BVH cost ~50 dependent node fetches each.  But for THIS
// family the set of triangles a ray can hit depends on its direction only, and the order along the ray on the
// distance from the centre only.  So (hard part 3):
//
//   build (once per acceleration structure, lazily at the first MODE_SHADER voxelize):
//     k_bins_scatter<count>   per triangle: central projection onto the 6 faces of a cube map of R x R cells per
//                             face (R ~ sqrt(T/1.5)); conservative (u,v) rectangle of its clipped projection,
//                             dilated by the margin below -> one count per covered cell
//     k_bins_scan_*           exclusive scan of the cell counts
//     k_bins_scatter<fill>    same walk, entries {rmin, rmax, triangle slot} into the cells' lists
//     k_bins_finish           per cell: sort by rmin, list header {first, count, max rmax}
//   trace (k_trace_shader_bins, one lane per voxel, one warp per 32-voxel word of the bit grid):
//     voxel -> cube-map cell of its direction -> list.  With rho = |origin|: the whole cell is skipped when
//     rho > max rmax (the voxel lies outside every surface layer of that direction: most of the grid);
//     an entry is skipped when rmax < rho (behind the origin) and the sorted list is left when
//     rmin > rho + best tc (cannot beat the closest hit so far).  Everything not culled goes through the same
//     per-pair arithmetic as the LBVH kernel (shader_common.cuh: Spec H of oracle/dxrv_oracle.h).
//
// Exactness.  A pair (ray, triangle) is a hit candidate only if Spec H's slab test AND watertight test pass.
//  (1) direction cull: the watertight test evaluates exact edge functions of vertices carrying rounding errors of
//      a few ulp of coordinates <= 4 in magnitude, i.e. it is an exact test against a triangle whose corners
//      moved by < 1e-6; the bins use 1e-5 (kVertexSlack).  A corner at distance >= r from the centre moved by
//      delta changes its cube-map coordinate by <= 2*sqrt(3)*delta/r, hence the margin kVertexSlack*4/r (+1e-5 for
//      the rounding of the voxel's own (u,v) and of the direction).  Triangles closer than kNearRadius to the
//      centre (where that margin explodes) go on a NEAR LIST every ray tests.
//  (2) radial culls: points of the triangle's box B on the ray have parameter t = |p| - rho (the ray is radial,
//      up to 1e-7), so the exact slab interval lies inside [rminB - rho, rmaxB - rho], rminB/rmaxB = min/max
//      distance of B from the centre.  Spec H computes tin/tout with relative error <= 2e-7 (one subtraction, one
//      product with a rounded reciprocal), absolute <= 1e-6 for |t| <= 4 (larger |t| keep their sign and
//      magnitude class).  tc is clamped into [tin, tout], so  rmaxB + kRadialSlack < rho  =>  tout < 0  =>  no hit,
//      and  rminB - kRadialSlack > rho + best  =>  tc > best  =>  cannot win (ties included: strict inequalities
//      with 2e-5 of slack never cull an equal tc).
//  The closest hit is the lexicographic minimum of (tc, primitive), so the order of evaluation is irrelevant.
//  Triangles with a non-finite coordinate are not binned (Spec H leaves pairs with NaN operands to the fallthrough
//  of its comparisons; tests/test_gpu_voxelize.py treats the neighbourhood of such triangles as unspecified).
//  tests/test_gpu_shader_bins.py checks bins == LBVH walk == oracle on meshes, soups and the near-list case.
//
// Budget: the entry lists are capped (clamp(48 T, 2^20, 2^27) entries); if a mesh of huge triangles needs more, a
// device flag is raised by the scan, every bins kernel returns at once and k_trace_shader (LBVH walk) runs
// instead -- decided on the device, no host round trip, the same CUDA graph.
#include "kernels.h"
#include "shader_common.cuh"

#include <cmath>
#include <cstdlib>
#include <cstring>

namespace dxrv
{
namespace
{
constexpr float kVertexSlack = 1e-5f;
constexpr float kRadialSlack = 2e-5f;
constexpr float kNearRadius = 1e-3f;
constexpr uint32_t kBinsUnsorted = 0x80000000u;
constexpr uint32_t kScanTile = 2048;      // cells per block of the local scan
constexpr int kSortCap = 512;             // longest list sorted in shared memory by one warp
constexpr int kFinishThreads = 128;

struct TriGeo
{
    float ax, ay, az, bx, by, bz, cx, cy, cz;
    float rminS, rmaxS;   // radial extent of the box, slack included
    float rlo;            // lower bound of the triangle's distance from the centre
};

__device__ __forceinline__ float len3(float x, float y, float z) { return sqrtf(x * x + y * y + z * z); }

// returns false for triangles that are not binned (non-finite coordinates)
__device__ __forceinline__ bool triGeometry(const Tri48* __restrict__ tris, uint32_t slot, TriGeo& g)
{
    const float4* t = reinterpret_cast<const float4*>(tris + slot);
    const float4 a = __ldg(t), b = __ldg(t + 1), c = __ldg(t + 2);
    g.ax = a.x; g.ay = a.y; g.az = a.z; g.bx = b.x; g.by = b.y; g.bz = b.z; g.cx = c.x; g.cy = c.y; g.cz = c.z;
    const float s = fabsf(a.x) + fabsf(a.y) + fabsf(a.z) + fabsf(b.x) + fabsf(b.y) + fabsf(b.z) + fabsf(c.x) + fabsf(c.y) + fabsf(c.z);
    if (!(s < 1e30f)) return false;   // NaN or infinite somewhere
    const float lox = fminf(fminf(a.x, b.x), c.x), hix = fmaxf(fmaxf(a.x, b.x), c.x);
    const float loy = fminf(fminf(a.y, b.y), c.y), hiy = fmaxf(fmaxf(a.y, b.y), c.y);
    const float loz = fminf(fminf(a.z, b.z), c.z), hiz = fmaxf(fmaxf(a.z, b.z), c.z);
    // min / max distance of the box from the centre
    const float nx = fmaxf(fmaxf(lox, -hix), 0.0f), ny = fmaxf(fmaxf(loy, -hiy), 0.0f), nz = fmaxf(fmaxf(loz, -hiz), 0.0f);
    const float fx = fmaxf(fabsf(lox), fabsf(hix)), fy = fmaxf(fabsf(loy), fabsf(hiy)), fz = fmaxf(fabsf(loz), fabsf(hiz));
    const float rminB = len3(nx, ny, nz), rmaxB = len3(fx, fy, fz);
    g.rminS = fmaxf(rminB * 0.99999f - kRadialSlack, 0.0f);
    g.rmaxS = rmaxB * 1.00001f + kRadialSlack;
    // lower bounds of the triangle's own distance: the box; a corner minus the longest edge; the plane (when the
    // normal is well conditioned)
    const float la = len3(a.x, a.y, a.z), lb = len3(b.x, b.y, b.z), lc = len3(c.x, c.y, c.z);
    const float e1x = b.x - a.x, e1y = b.y - a.y, e1z = b.z - a.z, e2x = c.x - a.x, e2y = c.y - a.y, e2z = c.z - a.z;
    const float e1 = len3(e1x, e1y, e1z), e2 = len3(e2x, e2y, e2z), e3 = len3(c.x - b.x, c.y - b.y, c.z - b.z);
    const float lmax = fmaxf(fmaxf(la, lb), lc);
    float r = fmaxf(rminB, lmax - fmaxf(fmaxf(e1, e2), e3));
    const float px = e1y * e2z - e1z * e2y, py = e1z * e2x - e1x * e2z, pz = e1x * e2y - e1y * e2x;
    const float pl = len3(px, py, pz);
    if (pl > 1e-3f * e1 * e2) r = fmaxf(r, fabsf(px * a.x + py * a.y + pz * a.z) / pl - 2e-3f * lmax);
    g.rlo = r * 0.9999f - 1e-5f;
    return true;
}

// Conservative cell rectangle of the triangle's central projection on one cube-map face.
// face = 2 * major axis + (negative side); (u, v) = the two other coordinates (cyclic order) over |major|.
__device__ __forceinline__ bool faceRect(const TriGeo& g, int face, uint32_t R, int& iu0, int& iu1, int& iv0, int& iv1)
{
    const int m = face >> 1;
    const float sgn = (face & 1) ? -1.0f : 1.0f;
    float z[3], x[3], y[3];
    z[0] = sgn * pick(g.ax, g.ay, g.az, m); z[1] = sgn * pick(g.bx, g.by, g.bz, m); z[2] = sgn * pick(g.cx, g.cy, g.cz, m);
    const int ia = (m + 1) % 3, ib = (m + 2) % 3;
    x[0] = pick(g.ax, g.ay, g.az, ia); x[1] = pick(g.bx, g.by, g.bz, ia); x[2] = pick(g.cx, g.cy, g.cz, ia);
    y[0] = pick(g.ax, g.ay, g.az, ib); y[1] = pick(g.bx, g.by, g.bz, ib); y[2] = pick(g.cx, g.cy, g.cz, ib);
    // points of the triangle inside this face's pyramid have z >= |p| / sqrt(3) >= 0.577 rlo: clip at z >= zc = rlo / 2
    const float zc = 0.5f * g.rlo;
    float umin = INFINITY, umax = -INFINITY, vmin = INFINITY, vmax = -INFINITY;
#pragma unroll
    for (int i = 0; i < 3; ++i)
    {
        const int j = (i + 1) % 3;
        const bool in_i = z[i] >= zc, in_j = z[j] >= zc;
        if (in_i)
        {
            const float inv = 1.0f / z[i];
            const float u = x[i] * inv, v = y[i] * inv;
            umin = fminf(umin, u); umax = fmaxf(umax, u); vmin = fminf(vmin, v); vmax = fmaxf(vmax, v);
        }
        if (in_i != in_j)
        {
            const float t = (zc - z[i]) / (z[j] - z[i]);
            const float inv = 1.0f / zc;
            const float u = (x[i] + t * (x[j] - x[i])) * inv, v = (y[i] + t * (y[j] - y[i])) * inv;
            umin = fminf(umin, u); umax = fmaxf(umax, u); vmin = fminf(vmin, v); vmax = fmaxf(vmax, v);
        }
    }
    if (!(umin <= umax)) return false;   // nothing in front of the clip plane
    const float margin = 4.0f * kVertexSlack / g.rlo + 1e-5f;
    umin -= margin; umax += margin; vmin -= margin; vmax += margin;
    if (umin > 1.0f || umax < -1.0f || vmin > 1.0f || vmax < -1.0f) return false;
    const float h = 0.5f * (float)R;
    const int last = (int)R - 1;
    iu0 = max(0, min(last, (int)floorf((fmaxf(umin, -1.0f) + 1.0f) * h)));
    iu1 = max(0, min(last, (int)floorf((fminf(umax, 1.0f) + 1.0f) * h)));
    iv0 = max(0, min(last, (int)floorf((fmaxf(vmin, -1.0f) + 1.0f) * h)));
    iv1 = max(0, min(last, (int)floorf((fminf(vmax, 1.0f) + 1.0f) * h)));
    return true;
}

template <bool kFill>
__device__ __forceinline__ void emit(const ShaderBinsView& bins, uint32_t cell, uint32_t slot, float rminS, float rmaxS)
{
    if (!kFill) atomicAdd(bins.cursors + cell, 1u);
    else
    {
        const uint32_t idx = __ldg(bins.blockSums + (cell / kScanTile)) + atomicAdd(bins.cursors + cell, 1u);
        bins.entries[idx] = make_uint4(__float_as_uint(rminS), __float_as_uint(rmaxS), slot, 0u);
    }
}

// One warp = 32 consecutive sorted triangles (neighbours in Morton order: their rectangles overlap heavily).
// Per cube-map face the warp rasterises the UNION of its rectangles with one lane per CELL: the lane counts the
// warp's triangles covering its cell and issues ONE atomic for all of them (14x fewer atomics than one per
// (triangle, cell), and 32 of them in flight at once instead of a dependent chain per triangle).  Warps whose
// union is large (triangles far apart, scene-sized triangles) fall back to one lane per triangle, rectangles of
// more than 32 cells walked by the whole warp.
template <bool kFill>
__global__ void __launch_bounds__(256)
k_bins_scatter(const Tri48* __restrict__ tris, uint32_t numTris, const ShaderBinsView bins)
{
    if (kFill && bins.state[1] != 0u) return;   // over budget: the LBVH walk takes over
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = laneId();
    const uint32_t warpBase = j - lane;
    TriGeo g;
    g.rminS = g.rmaxS = 0.0f;
    bool ok = j < numTris && triGeometry(tris, j, g);
    if (ok && !(g.rlo >= kNearRadius))
    {
        if (!kFill)
        {
            const uint32_t idx = atomicAdd(bins.state + 2, 1u);
            if (idx < bins.nearCap) bins.nearList[idx] = j;
            else bins.state[1] = 1u;
        }
        ok = false;
    }
    const uint32_t RR = bins.R * bins.R;
    for (int f = 0; f < 6; ++f)
    {
        int iu0 = 0, iu1 = -1, iv0 = 0, iv1 = -1;
        const bool v = ok && faceRect(g, f, bins.R, iu0, iu1, iv0, iv1);
        if (!__any_sync(0xffffffffu, v)) continue;
        const int U0 = __reduce_min_sync(0xffffffffu, v ? iu0 : 0x7fffffff), U1 = __reduce_max_sync(0xffffffffu, v ? iu1 : -1);
        const int V0 = __reduce_min_sync(0xffffffffu, v ? iv0 : 0x7fffffff), V1 = __reduce_max_sync(0xffffffffu, v ? iv1 : -1);
        const uint32_t W = (uint32_t)(U1 - U0 + 1), H = (uint32_t)(V1 - V0 + 1), area = W * H;
        if (W <= 255u && H <= 255u && area <= 512u)
        {
            // own rectangle relative to the union's origin, 8 bits per bound; an invalid lane covers nothing (lo > hi)
            const uint32_t packed = v ? ((uint32_t)(iu0 - U0) | ((uint32_t)(iu1 - U0) << 8) | ((uint32_t)(iv0 - V0) << 16) | ((uint32_t)(iv1 - V0) << 24))
                                      : 0x000000ffu;
            for (uint32_t c0 = 0; c0 < area; c0 += 32u)
            {
                const uint32_t c = c0 + lane;
                const uint32_t cu = c % W, cv = c / W;
                const bool inb = c < area;
                uint32_t k = 0;
#pragma unroll 8
                for (int t = 0; t < 32; ++t)
                {
                    const uint32_t p = __shfl_sync(0xffffffffu, packed, t);
                    k += (inb && cu >= (p & 255u) && cu <= ((p >> 8) & 255u) && cv >= ((p >> 16) & 255u) && cv <= (p >> 24)) ? 1u : 0u;
                }
                const uint32_t cell = (uint32_t)f * RR + (uint32_t)(V0 + (int)cv) * bins.R + (uint32_t)(U0 + (int)cu);
                if (!kFill)
                {
                    if (k) atomicAdd(bins.cursors + cell, k);
                }
                else
                {
                    uint32_t idx = 0;
                    if (k) idx = __ldg(bins.blockSums + (cell / kScanTile)) + atomicAdd(bins.cursors + cell, k);
                    if (__any_sync(0xffffffffu, k != 0u))
                    {
#pragma unroll 4
                        for (int t = 0; t < 32; ++t)
                        {
                            const uint32_t p = __shfl_sync(0xffffffffu, packed, t);
                            const float rm = __shfl_sync(0xffffffffu, g.rminS, t), rx = __shfl_sync(0xffffffffu, g.rmaxS, t);
                            if (inb && cu >= (p & 255u) && cu <= ((p >> 8) & 255u) && cv >= ((p >> 16) & 255u) && cv <= (p >> 24))
                                bins.entries[idx++] = make_uint4(__float_as_uint(rm), __float_as_uint(rx), warpBase + (uint32_t)t, 0u);
                        }
                    }
                }
            }
            continue;
        }
        const uint32_t nu = v ? (uint32_t)(iu1 - iu0 + 1) : 0u, nv = v ? (uint32_t)(iv1 - iv0 + 1) : 0u;
        const uint32_t n = nu * nv;
        if (n > 0u && n <= 64u)
            for (uint32_t c = 0; c < n; ++c)
                emit<kFill>(bins, (uint32_t)f * RR + (uint32_t)(iv0 + (int)(c / nu)) * bins.R + (uint32_t)(iu0 + (int)(c % nu)), j, g.rminS, g.rmaxS);
        if (n > 64u && !kFill)
        {
            // large rectangle (a triangle near the grid centre, or a scene-sized one): listed once by the counting pass
            // and rasterised by the whole grid in k_bins_big, in both passes
            const uint32_t idx = atomicAdd(bins.state + 3, 1u);
            if (idx < bins.bigCap)
            {
                bins.bigRects[2 * idx] = make_uint4(j, (uint32_t)f, (uint32_t)iu0, nu);
                bins.bigRects[2 * idx + 1] = make_uint4((uint32_t)iv0, n, __float_as_uint(g.rminS), __float_as_uint(g.rmaxS));
            }
            else bins.state[1] = 1u;
        }
    }
}

// The listed large rectangles: one CTA per rectangle (grid-stride), one thread per cell.
template <bool kFill>
__global__ void __launch_bounds__(256)
k_bins_big(const ShaderBinsView bins)
{
    if (bins.state[1] != 0u) return;
    const uint32_t count = min(bins.state[3], bins.bigCap);
    const uint32_t RR = bins.R * bins.R;
    for (uint32_t item = blockIdx.x; item < count; item += gridDim.x)
    {
        const uint4 a = __ldg(bins.bigRects + 2 * item), b = __ldg(bins.bigRects + 2 * item + 1);
        for (uint32_t c = threadIdx.x; c < b.y; c += blockDim.x)
            emit<kFill>(bins, a.y * RR + (b.x + c / a.w) * bins.R + (a.z + c % a.w), a.x, __uint_as_float(b.z), __uint_as_float(b.w));
    }
}

// ---- exclusive scan of the cell counts (in place), two kernels; the block bases stay in blockSums -------------
__global__ void __launch_bounds__(256)
k_bins_scan_local(uint32_t* __restrict__ cursors, uint32_t numCells, uint32_t* __restrict__ blockSums)
{
    __shared__ uint32_t warpSums[8];
    const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * 8u;
    uint32_t v[8];
    uint32_t sum = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q)
    {
        v[q] = (base + q < numCells) ? cursors[base + q] : 0u;
        sum += v[q];
    }
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (laneId() >= (uint32_t)o) inc += t;
    }
    const uint32_t warp = threadIdx.x >> 5;
    if (laneId() == 31) warpSums[warp] = inc;
    __syncthreads();
    uint32_t run = inc - sum;
    for (uint32_t w = 0; w < warp; ++w) run += warpSums[w];
#pragma unroll
    for (int q = 0; q < 8; ++q)
    {
        if (base + q < numCells) cursors[base + q] = run;
        run += v[q];
    }
    if (threadIdx.x == 255) blockSums[blockIdx.x] = run;
}

__global__ void __launch_bounds__(1024)
k_bins_scan_sums(uint32_t* __restrict__ blockSums, uint32_t numBlocks, uint32_t* __restrict__ state, uint32_t cap)
{
    __shared__ uint32_t warpSums[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < numBlocks; base += 1024u)
    {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < numBlocks ? blockSums[i] : 0u;
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
        if (i < numBlocks) blockSums[i] = run;
        __syncthreads();
        if (threadIdx.x == 1023) carry = run + v;
        __syncthreads();
    }
    if (threadIdx.x == 0)
    {
        state[0] = carry;
        if (carry > cap) state[1] = 1u;
    }
}

// ---- per cell: sort by (rmax, slot), suffix minima of rmin, header -----------------------------------------------
// Order by rmax: the entries behind a ray's origin (rmax < rho) are a PREFIX of the list, found by bisection; the
// walk may stop at entry i when the smallest rmin of entries i.. (stored in .w) exceeds rho + best tc.
__device__ __forceinline__ bool entryLess(const uint4& a, const uint4& b) { return a.y < b.y || (a.y == b.y && a.z < b.z); }

// Insertion sort of one list by its own thread, then the suffix minima; returns the largest rmax.
__device__ __forceinline__ float sortOwnList(uint4* e, uint32_t n)
{
    for (uint32_t i = 1; i < n; ++i)
    {
        const uint4 x = e[i];
        uint32_t k = i;
        while (k > 0u)
        {
            const uint4 p = e[k - 1];
            if (!entryLess(x, p)) break;
            e[k] = p;
            --k;
        }
        e[k] = x;
    }
    uint32_t run = 0x7f800000u;   // +inf; non-negative floats order like their bit patterns
    for (uint32_t i = n; i-- > 0u;)
    {
        run = min(run, e[i].x);
        e[i].w = run;
    }
    return n ? __uint_as_float(e[n - 1].y) : 0.0f;
}

__device__ __forceinline__ uint4 cellHeader(uint32_t first, uint32_t n, uint32_t flags, float rmaxAll)
{
    // {first entry, count | flags, max rmax, (max rmax)^2 rounded up: lets the tracer reject a voxel on rho^2}
    return make_uint4(first, n | flags, __float_as_uint(rmaxAll), __float_as_uint(__fmul_ru(rmaxAll, rmaxAll) * 1.000001f));
}

constexpr uint32_t kOwnSortMax = 48;      // longest list sorted by its own thread

// One thread per cell, one warp per 32 consecutive cells -- whose lists are one contiguous piece of the entry
// array: it is staged in shared memory (coalesced), every lane sorts its own list there, and it goes back coalesced.
// Longer lists are listed for k_bins_finish_big (one warp each, spread over the whole grid).
__global__ void __launch_bounds__(kFinishThreads)
k_bins_finish(const ShaderBinsView bins, uint32_t numCells)
{
    __shared__ uint4 sh[kFinishThreads / 32][kSortCap];
    if (bins.state[1] != 0u) return;
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = laneId(), warp = threadIdx.x >> 5;
    uint32_t first = 0, n = 0;
    if (c < numCells)
    {
        const uint32_t base = __ldg(bins.blockSums + (c / kScanTile));
        first = base + ((c % kScanTile) ? bins.cursors[c - 1] : 0u);
        n = base + bins.cursors[c] - first;
    }
    float rmaxAll = 0.0f;
    const bool own = n <= kOwnSortMax;
    // The warp's lists are one contiguous piece of the entry array.  It is staged through shared memory in batches of
    // consecutive cells (as many as fit kSortCap entries): coalesced in, every lane sorts its own list, coalesced out.
    const uint32_t endMine = first + n;
    uint32_t startLane = 0;
    while (startLane < 32u)
    {
        const uint32_t bFirst = __shfl_sync(0xffffffffu, first, startLane);
        const uint32_t fits = __ballot_sync(0xffffffffu, lane >= startLane && endMine - bFirst <= (uint32_t)kSortCap);
        // first + n is monotone over the lanes, so the fitting lanes are a run starting at startLane
        const uint32_t run = (uint32_t)__popc(fits);
        if (run == 0u) { ++startLane; continue; }        // a single list longer than the staging buffer: not sorted here
        const uint32_t bEnd = __shfl_sync(0xffffffffu, endMine, startLane + run - 1u);
        const uint32_t total = bEnd - bFirst;
        const bool mine = lane >= startLane && lane < startLane + run;
        if (total > 0u)
        {
            for (uint32_t i = lane; i < total; i += 32u) sh[warp][i] = bins.entries[bFirst + i];
            __syncwarp();
            if (mine && own) rmaxAll = sortOwnList(&sh[warp][first - bFirst], n);
            __syncwarp();
            for (uint32_t i = lane; i < total; i += 32u) bins.entries[bFirst + i] = sh[warp][i];
            __syncwarp();
        }
        startLane += run;
    }
    if (c < numCells)
    {
        if (own) bins.cells[c] = cellHeader(first, n, 0u, rmaxAll);
        else
        {
            const uint32_t idx = atomicAdd(bins.state + 4, 1u);
            if (idx < bins.bigCap) bins.bigCells[idx] = c;
            // (more long lists than the table holds: the rest stay unsorted, which is slower but still exact)
            uint32_t mx = 0u;
            if (idx >= bins.bigCap)
                for (uint32_t i = 0; i < n; ++i) mx = max(mx, bins.entries[first + i].y);
            bins.cells[c] = cellHeader(first, n, kBinsUnsorted, __uint_as_float(mx));
        }
    }
}

// One warp per listed cell: bitonic sort in shared memory (lists up to kSortCap entries), suffix minima, header.
__global__ void __launch_bounds__(kFinishThreads)
k_bins_finish_big(const ShaderBinsView bins)
{
    __shared__ uint4 sh[kFinishThreads / 32][kSortCap];
    if (bins.state[1] != 0u) return;
    const uint32_t lane = laneId(), warp = threadIdx.x >> 5;
    const uint32_t count = min(bins.state[4], bins.bigCap);
    for (uint32_t item = blockIdx.x * (kFinishThreads / 32) + warp; item < count; item += gridDim.x * (kFinishThreads / 32))
    {
        const uint32_t c = __ldg(bins.bigCells + item);
        const uint4 hdr = bins.cells[c];
        const uint32_t n = hdr.y & ~kBinsUnsorted;
        uint4* e = bins.entries + hdr.x;
        uint32_t mx = 0u;
        if (n <= (uint32_t)kSortCap)
        {
            uint32_t M = 32;
            while (M < n) M <<= 1;
            for (uint32_t i = lane; i < M; i += 32u) sh[warp][i] = i < n ? e[i] : make_uint4(0u, 0xffffffffu, 0xffffffffu, 0u);
            __syncwarp();
            for (uint32_t k = 2; k <= M; k <<= 1)
                for (uint32_t jj = k >> 1; jj > 0u; jj >>= 1)
                {
                    for (uint32_t i = lane; i < M; i += 32u)
                    {
                        const uint32_t l = i ^ jj;
                        if (l > i)
                        {
                            const uint4 a = sh[warp][i], b = sh[warp][l];
                            const bool up = (i & k) == 0u;
                            if (up ? entryLess(b, a) : entryLess(a, b)) { sh[warp][i] = b; sh[warp][l] = a; }
                        }
                    }
                    __syncwarp();
                }
            // suffix minima of rmin: every lane scans its own contiguous chunk backwards, then the chunks are chained
            const uint32_t chunk = (n + 31u) / 32u;
            const uint32_t lo = min(lane * chunk, n), hi = min(lo + chunk, n);
            uint32_t run = 0x7f800000u;
            for (uint32_t i = hi; i-- > lo;) run = min(run, sh[warp][i].x);
            // exclusive suffix over lanes: min of the chunk minima of all higher lanes
            uint32_t suf = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const uint32_t t = __shfl_down_sync(0xffffffffu, suf, o);
                if (lane + (uint32_t)o < 32u) suf = min(suf, t);
            }
            uint32_t after = __shfl_down_sync(0xffffffffu, suf, 1);
            if (lane == 31u) after = 0x7f800000u;
            run = after;
            for (uint32_t i = hi; i-- > lo;)
            {
                run = min(run, sh[warp][i].x);
                sh[warp][i].w = run;
            }
            __syncwarp();
            for (uint32_t i = lane; i < n; i += 32u) e[i] = sh[warp][i];
            mx = n ? sh[warp][n - 1].y : 0u;
            __syncwarp();
            if (lane == 0u) bins.cells[c] = cellHeader(hdr.x, n, 0u, __uint_as_float(mx));
        }
        else
        {
            for (uint32_t i = lane; i < n; i += 32u) mx = max(mx, e[i].y);
            mx = __reduce_max_sync(0xffffffffu, mx);
            if (lane == 0u) bins.cells[c] = cellHeader(hdr.x, n, kBinsUnsorted, __uint_as_float(mx));
        }
    }
}

// ---- trace --------------------------------------------------------------------------------------------------------
// (Tried and dropped, measured on dragon / bowl / bunny at 512^3: a RAY POOL -- classify 128 voxels cheaply, queue the
// ones with candidates in shared memory, run the queue on 32 lanes that refill individually or in batches.  It was
// 10-40 % SLOWER than this kernel at every refill threshold: the compaction destroys the coherence of the sub-brick
// (neighbouring rays read the same entries and triangles) and the queue management costs more than the idle lanes.)
// One warp = a sub-brick of 8 (x) x 2 (y) x 2 (z) voxels: neighbouring rays share cube-map cells and radii, so the
// lanes' list walks have similar lengths and read the same entries / triangles (a 32 x 1 x 1 row spans up to 15 cells).
// The warp's 32 result bits are four bytes of four different grid words; each is stored as one byte (every byte of
// the slab is written exactly once, by exactly one warp -- no barrier, no atomics, no clear).
constexpr int kTraceThreads = 128;
#ifndef TRACE_MINBLOCKS
#define TRACE_MINBLOCKS 9   // 56 registers, 36 warps / SM: measured best of 7..10 (8-10 % over 72 registers)
#endif

struct RayK
{
    float Ox, Oy, Oz, Dx, Dy, Dz;
    float S1, S2, Sz;       // D[K1]/D[KZ], D[K2]/D[KZ], 1/D[KZ] with K1 = (KZ+1)%3, K2 = (KZ+2)%3
};

template <int K> __device__ __forceinline__ float comp(float x, float y, float z) { return K == 0 ? x : (K == 1 ? y : z); }

// Spec H steps 1-4 for one pair, specialised for the ray's dominant axis KZ (no per-pair selects).  Two identities
// keep this bit-identical to shader_common.cuh::testTriangle: (i) a pair is a hit iff slab test AND watertight test
// pass, so the (cheaper to fail) watertight test runs first; (ii) Spec H swaps kx and ky when D[kz] < 0, which
// negates U, V, W and det exactly (fl(a - b) = -fl(b - a)) and leaves the sign test, t = T/det and the
// barycentrics unchanged -- so the natural order (K1, K2) serves both signs.
template <int KZ>
__device__ __forceinline__ void testTriangleK(const RayK& r, const Tri48* __restrict__ tris, uint32_t slot, BestHit& best)
{
    constexpr int K1 = (KZ + 1) % 3, K2 = (KZ + 2) % 3;
    const float4* t = reinterpret_cast<const float4*>(tris + slot);
    const float4 a = __ldg(t), b = __ldg(t + 1), c = __ldg(t + 2);
    const float Ax3 = __fsub_rn(a.x, r.Ox), Ay3 = __fsub_rn(a.y, r.Oy), Az3 = __fsub_rn(a.z, r.Oz);
    const float Bx3 = __fsub_rn(b.x, r.Ox), By3 = __fsub_rn(b.y, r.Oy), Bz3 = __fsub_rn(b.z, r.Oz);
    const float Cx3 = __fsub_rn(c.x, r.Ox), Cy3 = __fsub_rn(c.y, r.Oy), Cz3 = __fsub_rn(c.z, r.Oz);
    const float Akz = comp<KZ>(Ax3, Ay3, Az3), Bkz = comp<KZ>(Bx3, By3, Bz3), Ckz = comp<KZ>(Cx3, Cy3, Cz3);
    const float Ax = __fsub_rn(comp<K1>(Ax3, Ay3, Az3), __fmul_rn(r.S1, Akz));
    const float Ay = __fsub_rn(comp<K2>(Ax3, Ay3, Az3), __fmul_rn(r.S2, Akz));
    const float Bx = __fsub_rn(comp<K1>(Bx3, By3, Bz3), __fmul_rn(r.S1, Bkz));
    const float By = __fsub_rn(comp<K2>(Bx3, By3, Bz3), __fmul_rn(r.S2, Bkz));
    const float Cx = __fsub_rn(comp<K1>(Cx3, Cy3, Cz3), __fmul_rn(r.S1, Ckz));
    const float Cy = __fsub_rn(comp<K2>(Cx3, Cy3, Cz3), __fmul_rn(r.S2, Ckz));
    float U, V, W;
    edgeValues(Ax, Ay, Bx, By, Cx, Cy, U, V, W);
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return;
    const float det = __fadd_rn(__fadd_rn(U, V), W);
    if (det == 0.0f) return;
    // the pair passed the watertight test (rare): only now the slab interval, with the reciprocals Spec H prescribes
    float ix = __fdiv_rn(1.0f, r.Dx), iy = __fdiv_rn(1.0f, r.Dy), iz = __fdiv_rn(1.0f, r.Dz);
    if (ix > kFltMax) ix = kFltMax; if (ix < -kFltMax) ix = -kFltMax;
    if (iy > kFltMax) iy = kFltMax; if (iy < -kFltMax) iy = -kFltMax;
    if (iz > kFltMax) iz = kFltMax; if (iz < -kFltMax) iz = -kFltMax;
    float tin = 0.0f, tout = kTMax;
    {
        float t0 = __fmul_rn(__fsub_rn(fminsel(fminsel(a.x, b.x), c.x), r.Ox), ix), t1 = __fmul_rn(__fsub_rn(fmaxsel(fmaxsel(a.x, b.x), c.x), r.Ox), ix);
        tin = fmaxsel(fminsel(t0, t1), tin); tout = fminsel(fmaxsel(t0, t1), tout);
        t0 = __fmul_rn(__fsub_rn(fminsel(fminsel(a.y, b.y), c.y), r.Oy), iy); t1 = __fmul_rn(__fsub_rn(fmaxsel(fmaxsel(a.y, b.y), c.y), r.Oy), iy);
        tin = fmaxsel(fminsel(t0, t1), tin); tout = fminsel(fmaxsel(t0, t1), tout);
        t0 = __fmul_rn(__fsub_rn(fminsel(fminsel(a.z, b.z), c.z), r.Oz), iz); t1 = __fmul_rn(__fsub_rn(fmaxsel(fmaxsel(a.z, b.z), c.z), r.Oz), iz);
        tin = fmaxsel(fminsel(t0, t1), tin); tout = fminsel(fmaxsel(t0, t1), tout);
    }
    if (!(tin <= tout)) return;
    const float tt = __fdiv_rn(weighted3(U, __fmul_rn(r.Sz, Akz), V, __fmul_rn(r.Sz, Bkz), W, __fmul_rn(r.Sz, Ckz)), det);
    const float tc = fminsel(fmaxsel(tt, tin), tout);
    if (!(tc > 0.0f && tc < kTMax)) return;
    const uint32_t prim = __float_as_uint(a.w);
    if (tc < best.tc || (tc == best.tc && prim < best.prim))
    {
        best.tc = tc; best.prim = prim;
        best.bx = __fdiv_rn(V, det); best.by = __fdiv_rn(W, det);
    }
}

struct TraceGeom
{
    uint32_t unitsX, unitsY, unitsZ;   // sub-bricks per axis: 4 P, ceil(N / 2), ceil(layers / 2)
    uint32_t strideX, strideY, strideZ;  // total warps of the grid, decomposed in the same mixed radix
    uint32_t layers;
};

__global__ void __launch_bounds__(kTraceThreads, TRACE_MINBLOCKS)
k_trace_shader_bins(const ShaderParams prm, const ShaderBinsView bins, const float* __restrict__ centres, const TraceGeom g)
{
    if (__ldg(bins.state + 1) != 0u) return;   // over budget: k_trace_shader (LBVH walk) produces the grid
    const uint32_t lane = laneId();
    const uint32_t N = prm.N, P = prm.P, R = bins.R;
    const float halfR = 0.5f * (float)R;
    const uint32_t nearCount = min(__ldg(bins.state + 2), bins.nearCap);
    const uint32_t lx = lane & 7u, ly = (lane >> 3) & 1u, lz = lane >> 4;
    uint8_t* gridBytes = reinterpret_cast<uint8_t*>(prm.grid);
    // first sub-brick of this warp, then mixed-radix increments by the total number of warps
    uint32_t ux, uy, uz;
    {
        const uint64_t w0 = (uint64_t)blockIdx.x * (kTraceThreads / 32) + (threadIdx.x >> 5);
        ux = (uint32_t)(w0 % g.unitsX);
        const uint64_t t = w0 / g.unitsX;
        uy = (uint32_t)(t % g.unitsY);
        uz = (uint32_t)(t / g.unitsY);
    }
    while (uz < g.unitsZ)
    {
        const uint32_t x = ux * 8u + lx, y = uy * 2u + ly, zl = uz * 2u + lz;
        const bool valid = x < N && y < N && zl < g.layers;
        bool inside = false, active = false, sorted = true;
        uint32_t texel = 0, n = 0;
        int kz = 0;
        float rho = 0.0f;
        const uint4* e = bins.entries;
        RayK r;
        r.Ox = r.Oy = r.Oz = r.Dx = r.Dy = r.Dz = r.S1 = r.S2 = r.Sz = 0.0f;
        if (valid)
        {
            r.Ox = __ldg(centres + x);
            r.Oy = -__ldg(centres + y);
            r.Oz = __ldg(centres + prm.z0 + zl);
            if (!(r.Ox == 0.0f && r.Oy == 0.0f && r.Oz == 0.0f) && prm.numTris > 0)
            {
                // cube-map cell of the direction (first largest |component| wins ties; any consistent rule will do:
                // the rectangles are dilated and clamped onto the closed face)
                const float ax = fabsf(r.Ox), ay = fabsf(r.Oy), az = fabsf(r.Oz);
                int m = 0; float pm = ax;
                if (ay > pm) { m = 1; pm = ay; }
                if (az > pm) { m = 2; pm = az; }
                const float inv = __fdividef(1.0f, pm);
                const float u = pick(r.Oy, r.Oz, r.Ox, m) * inv, v = pick(r.Oz, r.Ox, r.Oy, m) * inv;
                const int last = (int)R - 1;
                const int iu = max(0, min(last, (int)((u + 1.0f) * halfR))), iv = max(0, min(last, (int)((v + 1.0f) * halfR)));
                const uint32_t face = 2u * (uint32_t)m + (pick(r.Ox, r.Oy, r.Oz, m) < 0.0f ? 1u : 0u);
                const uint4 hdr = __ldg(bins.cells + ((size_t)face * R + (uint32_t)iv) * R + (uint32_t)iu);
                // rho^2 against the cell's (max rmax)^2: most voxels lie outside every surface layer of their direction
                const float rho2 = __fadd_rn(__fadd_rn(__fmul_rn(r.Ox, r.Ox), __fmul_rn(r.Oy, r.Oy)), __fmul_rn(r.Oz, r.Oz));
                n = hdr.y & ~kBinsUnsorted;
                if (n > 0u && !(__uint_as_float(hdr.w) < rho2)) { e = bins.entries + hdr.x; sorted = (hdr.y & kBinsUnsorted) == 0u; }
                else n = 0u;
                if (n > 0u || nearCount)
                {
                    rho = __fsqrt_rn(rho2);   // = rayLength(): Spec H's |O|
                    r.Dx = __fdiv_rn(r.Ox, rho); r.Dy = __fdiv_rn(r.Oy, rho); r.Dz = __fdiv_rn(r.Oz, rho);
                    // Spec H: kz = axis of the largest |D| (first wins ties); chosen on D itself (two different |O|
                    // components may round to the same |D|)
                    float md = fabsf(r.Dx);
                    if (fabsf(r.Dy) > md) { kz = 1; md = fabsf(r.Dy); }
                    if (fabsf(r.Dz) > md) { kz = 2; }
                    const float dk = pick(r.Dx, r.Dy, r.Dz, kz);
                    r.S1 = __fdiv_rn(pick(r.Dy, r.Dz, r.Dx, kz), dk);
                    r.S2 = __fdiv_rn(pick(r.Dz, r.Dx, r.Dy, kz), dk);
                    r.Sz = __fdiv_rn(1.0f, dk);
                    active = true;
                }
            }
        }
        BestHit best;
        best.tc = INFINITY; best.prim = 0xffffffffu; best.bx = 0.0f; best.by = 0.0f;
        for (uint32_t k = 0; k < nearCount; ++k)
            if (active)
            {
                const uint32_t slot = __ldg(bins.nearList + k);
                if (kz == 0) testTriangleK<0>(r, prm.tris, slot, best);
                else if (kz == 1) testTriangleK<1>(r, prm.tris, slot, best);
                else testTriangleK<2>(r, prm.tris, slot, best);
            }
        // "while-while": every lane first runs ahead to its next candidate (a cheap scan of 16-byte entries), then the
        // lanes that found one test together -- the exact tests are the expensive part and must not run at a few
        // lanes per instruction.
        bool walking = active && n > 0u;
        uint32_t i = 0;
        if (walking && sorted && n > 8u)
        {
            // entries behind the origin (rmax < rho) are a prefix of the list (sorted by rmax): bisect over it
            uint32_t lo = 0, hi = n;
            while (lo < hi)
            {
                const uint32_t mid = (lo + hi) >> 1;
                if (__uint_as_float(__ldg(&e[mid].y)) < rho) lo = mid + 1u; else hi = mid;
            }
            i = lo;
        }
        while (__any_sync(0xffffffffu, walking))
        {
            uint32_t slot = 0xffffffffu;
            if (walking)
            {
                while (i < n)
                {
                    const uint4 en = __ldg(e + i);
                    ++i;
                    if (__uint_as_float(en.y) < rho) continue;                 // behind the origin
                    const float reach = __fadd_rn(rho, best.tc);
                    if (sorted && __uint_as_float(en.w) > reach) { i = n; break; }   // nothing from here on can beat the best hit
                    if (__uint_as_float(en.x) > reach) continue;               // this one cannot
                    slot = en.z;
                    break;
                }
                walking = slot != 0xffffffffu;
            }
            __syncwarp();
            if (walking)
            {
                if (kz == 0) testTriangleK<0>(r, prm.tris, slot, best);
                else if (kz == 1) testTriangleK<1>(r, prm.tris, slot, best);
                else testTriangleK<2>(r, prm.tris, slot, best);
            }
        }
        if (best.prim != 0xffffffffu)
        {
            RaySetup rs;
            rs.Dx = r.Dx; rs.Dy = r.Dy; rs.Dz = r.Dz;
            inside = shadeHit(prm, rs, best, texel);
        }
        const uint32_t bits = __ballot_sync(0xffffffffu, inside);
        if (lane < 4u)
        {
            // row `lane` of the sub-brick = (ly, lz) = (lane & 1, lane >> 1): its 8 voxels are lanes 8*lane .. 8*lane+7
            const uint32_t yy = uy * 2u + (lane & 1u), zz = uz * 2u + (lane >> 1);
            if (yy < N && zz < g.layers && ux * 8u < P * 32u)
                gridBytes[(((size_t)zz * N + yy) * P) * 4u + ux] = (uint8_t)((bits >> (8u * lane)) & 0xffu);
        }
        if (prm.texels && valid) prm.texels[((size_t)zl * N + y) * N + x] = texel;
        ux += g.strideX; if (ux >= g.unitsX) { ux -= g.unitsX; ++uy; }
        uy += g.strideY; if (uy >= g.unitsY) { uy -= g.unitsY; ++uz; }
        uz += g.strideZ;
    }
}

__global__ void k_voxel_centres(float* __restrict__ centres, uint32_t N)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) centres[i] = voxelCentre(i, (float)N);
}
}  // namespace

// ---- host side ------------------------------------------------------------------------------------------------------
uint32_t shaderBinsResolution(uint32_t numTris)
{
    if (const char* e = std::getenv("DXRV_SHADER_BINS_R"))
    {
        const int r = std::atoi(e);
        if (r >= 8 && r <= 4096 && (r & (r - 1)) == 0) return (uint32_t)r;
    }
    const double want = std::sqrt((double)(numTris ? numTris : 1) / 1.5);
    int lg = (int)std::lround(std::log2(want > 1.0 ? want : 1.0));
    lg = lg < 3 ? 3 : (lg > 11 ? 11 : lg);
    return 1u << lg;
}

ShaderBinsSizes shaderBinsSizes(uint32_t numTris)
{
    ShaderBinsSizes s;
    s.R = shaderBinsResolution(numTris);
    const size_t cells = 6 * (size_t)s.R * s.R;
    size_t cap = 48 * (size_t)numTris;
    if (cap < (1u << 20)) cap = 1u << 20;
    if (cap > (1u << 27)) cap = 1u << 27;
    s.cap = (uint32_t)cap;
    s.nearCap = 4096;
    s.bigCap = 1u << 16;
    s.numBlocks = (uint32_t)((cells + kScanTile - 1) / kScanTile);
    auto align = [](size_t b) { return (b + 255) & ~(size_t)255; };
    s.offCells = 0;
    s.offEntries = s.offCells + align(cells * sizeof(uint4));
    s.offCursors = s.offEntries + align(cap * sizeof(uint4));
    s.offBlockSums = s.offCursors + align((cells + 8) * sizeof(uint32_t));
    s.offNear = s.offBlockSums + align((s.numBlocks + 1) * sizeof(uint32_t));
    s.offBigRects = s.offNear + align(s.nearCap * sizeof(uint32_t));
    s.offBigCells = s.offBigRects + align(2 * (size_t)s.bigCap * sizeof(uint4));
    s.offState = s.offBigCells + align((size_t)s.bigCap * sizeof(uint32_t));
    s.bytes = s.offState + 256;
    return s;
}

ShaderBinsView shaderBinsView(void* base, const ShaderBinsSizes& s)
{
    uint8_t* p = static_cast<uint8_t*>(base);
    ShaderBinsView v;
    v.cells = reinterpret_cast<uint4*>(p + s.offCells);
    v.entries = reinterpret_cast<uint4*>(p + s.offEntries);
    v.cursors = reinterpret_cast<uint32_t*>(p + s.offCursors);
    v.blockSums = reinterpret_cast<uint32_t*>(p + s.offBlockSums);
    v.nearList = reinterpret_cast<uint32_t*>(p + s.offNear);
    v.state = reinterpret_cast<uint32_t*>(p + s.offState);
    v.bigRects = reinterpret_cast<uint4*>(p + s.offBigRects);
    v.bigCells = reinterpret_cast<uint32_t*>(p + s.offBigCells);
    v.R = s.R; v.cap = s.cap; v.nearCap = s.nearCap; v.bigCap = s.bigCap;
    return v;
}

int launchShaderBinsBuild(cudaStream_t s, const BvhView& bvh, void* base, const ShaderBinsSizes& sz, bool forceOverflow)
{
    const ShaderBinsView v = shaderBinsView(base, sz);
    const uint32_t cells = 6u * sz.R * sz.R;
    cudaMemsetAsync(v.state, 0, 32, s);
    if (forceOverflow)
    {
        // DXRV_SHADER_PATH=bvh: raise the flag, build nothing
        cudaMemsetAsync(v.state + 1, 1, 1, s);   // low byte = 1
        return 0;
    }
    cudaMemsetAsync(v.cursors, 0, sizeof(uint32_t) * ((size_t)cells + 8), s);
    const uint32_t T = bvh.numTris;
    const uint32_t tb = (T + 255) / 256;
    if (T)
    {
        k_bins_scatter<false><<<tb, 256, 0, s>>>(bvh.tris, T, v);
        k_bins_big<false><<<148 * 4, 256, 0, s>>>(v);
    }
    k_bins_scan_local<<<sz.numBlocks, 256, 0, s>>>(v.cursors, cells, v.blockSums);
    k_bins_scan_sums<<<1, 1024, 0, s>>>(v.blockSums, sz.numBlocks, v.state, sz.cap);
    if (T)
    {
        k_bins_scatter<true><<<tb, 256, 0, s>>>(bvh.tris, T, v);
        k_bins_big<true><<<148 * 4, 256, 0, s>>>(v);
    }
    k_bins_finish<<<(cells + kFinishThreads - 1) / kFinishThreads, kFinishThreads, 0, s>>>(v, cells);
    k_bins_finish_big<<<148 * 4, kFinishThreads, 0, s>>>(v);
    return T ? 8 : 4;
}

void launchTraceShaderBins(cudaStream_t s, const BvhView& bvh, const MeshView& m, uint32_t N, uint32_t z0, uint32_t z1,
                           uint32_t* grid, uint32_t* texels, uint32_t* dErr, void* base, const ShaderBinsSizes& sz, float* centres)
{
    const ShaderBinsView v = shaderBinsView(base, sz);
    ShaderParams prm;
    prm.nodes = bvh.nodes; prm.tris = bvh.tris; prm.numTris = bvh.numTris;
    prm.verts = m.verts; prm.stride = m.stride; prm.indices = m.indices;
    prm.N = N; prm.P = (N + 31) / 32; prm.z0 = z0;
    prm.numWords = (uint64_t)(z1 - z0) * N * prm.P;
    prm.grid = grid; prm.texels = texels; prm.err = dErr; prm.binsState = v.state;
    k_voxel_centres<<<(N + 255) / 256, 256, 0, s>>>(centres, N);
    TraceGeom g;
    g.layers = z1 - z0;
    g.unitsX = prm.P * 4u; g.unitsY = (N + 1u) / 2u; g.unitsZ = (g.layers + 1u) / 2u;
    const uint64_t units = (uint64_t)g.unitsX * g.unitsY * g.unitsZ;
    const uint64_t wantBlocks = (units + (kTraceThreads / 32) - 1) / (kTraceThreads / 32);
    const uint64_t capBlocks = 148ull * 16ull * 8ull;
    const uint32_t blocks = (uint32_t)(wantBlocks < capBlocks ? wantBlocks : capBlocks);
    uint64_t stride = (uint64_t)blocks * (kTraceThreads / 32);
    g.strideX = (uint32_t)(stride % g.unitsX); stride /= g.unitsX;
    g.strideY = (uint32_t)(stride % g.unitsY); stride /= g.unitsY;
    g.strideZ = (uint32_t)stride;
    k_trace_shader_bins<<<blocks, kTraceThreads, 0, s>>>(prm, v, centres, g);
}
}  // namespace dxrv
