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

// One thread per triangle (sorted slot); rectangles of more than 32 cells are walked by the whole warp.
template <bool kFill>
__global__ void __launch_bounds__(256)
k_bins_scatter(const Tri48* __restrict__ tris, uint32_t numTris, const ShaderBinsView bins)
{
    if (kFill && bins.state[1] != 0u) return;   // over budget: the LBVH walk takes over
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = laneId();
    TriGeo g;
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
        const uint32_t nu = v ? (uint32_t)(iu1 - iu0 + 1) : 0u, nv = v ? (uint32_t)(iv1 - iv0 + 1) : 0u;
        const uint32_t n = nu * nv;
        if (n > 0u && n <= 32u)
            for (uint32_t c = 0; c < n; ++c)
                emit<kFill>(bins, (uint32_t)f * RR + (uint32_t)(iv0 + (int)(c / nu)) * bins.R + (uint32_t)(iu0 + (int)(c % nu)), j, g.rminS, g.rmaxS);
        uint32_t big = __ballot_sync(0xffffffffu, n > 32u);
        while (big)
        {
            const int L = __ffs(big) - 1;
            big &= big - 1u;
            const uint32_t bn = __shfl_sync(0xffffffffu, n, L), bnu = __shfl_sync(0xffffffffu, nu, L);
            const int bu0 = __shfl_sync(0xffffffffu, iu0, L), bv0 = __shfl_sync(0xffffffffu, iv0, L);
            const uint32_t bj = __shfl_sync(0xffffffffu, j, L);
            const float bmin = __shfl_sync(0xffffffffu, g.rminS, L), bmax = __shfl_sync(0xffffffffu, g.rmaxS, L);
            for (uint32_t c = lane; c < bn; c += 32u)
                emit<kFill>(bins, (uint32_t)f * RR + (uint32_t)(bv0 + (int)(c / bnu)) * bins.R + (uint32_t)(bu0 + (int)(c % bnu)), bj, bmin, bmax);
        }
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

// ---- per cell: sort by (rmin, slot), header ----------------------------------------------------------------------
__device__ __forceinline__ bool entryLess(const uint4& a, const uint4& b) { return a.x < b.x || (a.x == b.x && a.z < b.z); }

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
    uint32_t flags = 0;
    float rmaxAll = 0.0f;
    if (n > 0u && n <= 16u)
    {
        // insertion sort in place (the list is this thread's own)
        uint4* e = bins.entries + first;
        for (uint32_t i = 0; i < n; ++i)
        {
            const uint4 x = e[i];
            rmaxAll = fmaxf(rmaxAll, __uint_as_float(x.y));
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
    }
    uint32_t big = __ballot_sync(0xffffffffu, n > 16u);
    while (big)
    {
        const int L = __ffs(big) - 1;
        big &= big - 1u;
        const uint32_t bf = __shfl_sync(0xffffffffu, first, L), bn = __shfl_sync(0xffffffffu, n, L);
        uint4* e = bins.entries + bf;
        float mx = 0.0f;
        if (bn <= (uint32_t)kSortCap)
        {
            uint32_t M = 32;
            while (M < bn) M <<= 1;
            for (uint32_t i = lane; i < M; i += 32u)
            {
                const uint4 x = i < bn ? e[i] : make_uint4(0xffffffffu, 0u, 0xffffffffu, 0u);
                if (i < bn) mx = fmaxf(mx, __uint_as_float(x.y));
                sh[warp][i] = x;
            }
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
            for (uint32_t i = lane; i < bn; i += 32u) e[i] = sh[warp][i];
            __syncwarp();
        }
        else
        {
            for (uint32_t i = lane; i < bn; i += 32u) mx = fmaxf(mx, __uint_as_float(e[i].y));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if ((int)lane == L)
        {
            rmaxAll = mx;
            if (bn > (uint32_t)kSortCap) flags = kBinsUnsorted;
        }
    }
    if (c < numCells) bins.cells[c] = make_uint4(first, n | flags, __float_as_uint(rmaxAll), 0u);
}

// ---- trace --------------------------------------------------------------------------------------------------------
constexpr int kTraceThreads = 128;

__global__ void __launch_bounds__(kTraceThreads)
k_trace_shader_bins(const ShaderParams prm, const ShaderBinsView bins)
{
    if (__ldg(bins.state + 1) != 0u) return;   // over budget: k_trace_shader (LBVH walk) produces the grid
    const uint32_t lane = laneId();
    const uint32_t N = prm.N, P = prm.P, R = bins.R;
    const float fN = (float)N, halfR = 0.5f * (float)R;
    const uint32_t nearCount = min(__ldg(bins.state + 2), bins.nearCap);
    for (uint64_t word = (uint64_t)blockIdx.x * (kTraceThreads / 32) + (threadIdx.x >> 5); word < prm.numWords;
         word += (uint64_t)gridDim.x * (kTraceThreads / 32))
    {
        const uint64_t row = word / P;
        const uint32_t x = (uint32_t)(word - row * P) * 32u + lane;
        const uint32_t y = (uint32_t)(row % N), z = prm.z0 + (uint32_t)(row / N);

        bool inside = false;
        uint32_t texel = 0;
        RaySetup r;
        r.Ox = voxelCentre(x, fN);
        r.Oy = -voxelCentre(y, fN);
        r.Oz = voxelCentre(z, fN);
        const bool live = x < N && !(r.Ox == 0.0f && r.Oy == 0.0f && r.Oz == 0.0f) && prm.numTris > 0;
        if (live)
        {
            // cube-map cell of the direction (first largest |component| wins ties, any consistent rule will do:
            // the rectangles are dilated and clamped onto the closed face)
            const float ax = fabsf(r.Ox), ay = fabsf(r.Oy), az = fabsf(r.Oz);
            int m = 0; float pm = ax;
            if (ay > pm) { m = 1; pm = ay; }
            if (az > pm) { m = 2; pm = az; }
            const float inv = 1.0f / pm;
            const float u = pick(r.Oy, r.Oz, r.Ox, m) * inv, v = pick(r.Oz, r.Ox, r.Oy, m) * inv;
            const int last = (int)R - 1;
            const int iu = max(0, min(last, (int)((u + 1.0f) * halfR))), iv = max(0, min(last, (int)((v + 1.0f) * halfR)));
            const uint32_t face = 2u * (uint32_t)m + (pick(r.Ox, r.Oy, r.Oz, m) < 0.0f ? 1u : 0u);
            const uint4 hdr = __ldg(bins.cells + ((size_t)face * R + (uint32_t)iv) * R + (uint32_t)iu);
            const uint32_t n = hdr.y & ~kBinsUnsorted;
            const float rho = rayLength(r.Ox, r.Oy, r.Oz);
            const bool any = n > 0u && !(__uint_as_float(hdr.z) < rho);
            if (any || nearCount)
            {
                raySetup(r, rho);
                BestHit best;
                best.tc = INFINITY; best.prim = 0xffffffffu; best.bx = 0.0f; best.by = 0.0f;
                for (uint32_t k = 0; k < nearCount; ++k) testTriangle(r, prm.tris, __ldg(bins.nearList + k), best);
                if (any)
                {
                    const bool sorted = (hdr.y & kBinsUnsorted) == 0u;
                    const uint4* e = bins.entries + hdr.x;
                    for (uint32_t i = 0; i < n; ++i)
                    {
                        const uint4 en = __ldg(e + i);
                        if (__uint_as_float(en.y) < rho) continue;                       // behind the origin
                        if (__uint_as_float(en.x) > __fadd_rn(rho, best.tc))             // cannot beat the best hit
                        {
                            if (sorted) break;
                            continue;
                        }
                        testTriangle(r, prm.tris, en.z, best);
                    }
                }
                if (best.prim != 0xffffffffu) inside = shadeHit(prm, r, best, texel);
            }
        }
        const uint32_t bits = __ballot_sync(0xffffffffu, inside);
        if (lane == 0) prm.grid[word] = bits;
        if (prm.texels && x < N) prm.texels[row * N + x] = texel;
    }
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
    s.numBlocks = (uint32_t)((cells + kScanTile - 1) / kScanTile);
    auto align = [](size_t b) { return (b + 255) & ~(size_t)255; };
    s.offCells = 0;
    s.offEntries = s.offCells + align(cells * sizeof(uint4));
    s.offCursors = s.offEntries + align(cap * sizeof(uint4));
    s.offBlockSums = s.offCursors + align((cells + 8) * sizeof(uint32_t));
    s.offNear = s.offBlockSums + align((s.numBlocks + 1) * sizeof(uint32_t));
    s.offState = s.offNear + align(s.nearCap * sizeof(uint32_t));
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
    v.R = s.R; v.cap = s.cap; v.nearCap = s.nearCap;
    return v;
}

int launchShaderBinsBuild(cudaStream_t s, const BvhView& bvh, void* base, const ShaderBinsSizes& sz, bool forceOverflow)
{
    const ShaderBinsView v = shaderBinsView(base, sz);
    const uint32_t cells = 6u * sz.R * sz.R;
    cudaMemsetAsync(v.state, 0, 16, s);
    if (forceOverflow)
    {
        // DXRV_SHADER_PATH=bvh: raise the flag, build nothing
        cudaMemsetAsync(v.state + 1, 1, 1, s);   // low byte = 1
        return 0;
    }
    cudaMemsetAsync(v.cursors, 0, sizeof(uint32_t) * ((size_t)cells + 8), s);
    const uint32_t T = bvh.numTris;
    const uint32_t tb = (T + 255) / 256;
    if (T) k_bins_scatter<false><<<tb, 256, 0, s>>>(bvh.tris, T, v);
    k_bins_scan_local<<<sz.numBlocks, 256, 0, s>>>(v.cursors, cells, v.blockSums);
    k_bins_scan_sums<<<1, 1024, 0, s>>>(v.blockSums, sz.numBlocks, v.state, sz.cap);
    if (T) k_bins_scatter<true><<<tb, 256, 0, s>>>(bvh.tris, T, v);
    k_bins_finish<<<(cells + kFinishThreads - 1) / kFinishThreads, kFinishThreads, 0, s>>>(v, cells);
    return T ? 5 : 3;
}

void launchTraceShaderBins(cudaStream_t s, const BvhView& bvh, const MeshView& m, uint32_t N, uint32_t z0, uint32_t z1,
                           uint32_t* grid, uint32_t* texels, uint32_t* dErr, void* base, const ShaderBinsSizes& sz)
{
    const ShaderBinsView v = shaderBinsView(base, sz);
    ShaderParams prm;
    prm.nodes = bvh.nodes; prm.tris = bvh.tris; prm.numTris = bvh.numTris;
    prm.verts = m.verts; prm.stride = m.stride; prm.indices = m.indices;
    prm.N = N; prm.P = (N + 31) / 32; prm.z0 = z0;
    prm.numWords = (uint64_t)(z1 - z0) * N * prm.P;
    prm.grid = grid; prm.texels = texels; prm.err = dErr; prm.binsState = v.state;
    const uint64_t want = (prm.numWords + (kTraceThreads / 32) - 1) / (kTraceThreads / 32);
    const uint64_t cap = 148ull * 16ull * 8ull;
    k_trace_shader_bins<<<(unsigned)(want < cap ? want : cap), kTraceThreads, 0, s>>>(prm, v);
}
}  // namespace dxrv
