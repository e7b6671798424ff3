// trace_parity.cu -- MODE_PARITY: +x column rays, watertight crossings, parity fill (sm_100a).
//
// Replaces DispatchRays + raygenMain/TraceRay (reference Content/Voxelizer.cpp:351-369,
// Content/Shaders/DXRVoxelizer.hlsl:58-85) for the column-parity formulation of the solid test.
// The per-(column, triangle) arithmetic is Spec H / MODE_PARITY of oracle/dxrv_oracle.h.
//
// One warp owns a tile of TY x TZ = 32 voxel columns (y,z) and ALL their voxels along x:
//   1. warp-cooperative BVH walk: every lane pops a different node from a shared-memory stack,
//      tests both child boxes against the tile's (y,z) rectangle (exact compares: a crossing implies
//      the column lies inside every ancestor box), pushes inner children, queues leaf children;
//   2. queued triangles are processed 32 at a time, one triangle per lane: each covered column gets
//      the exact crossing test and toggles ONE bit (first voxel whose centre is beyond the crossing)
//      in the tile's shared-memory bit rows -- XOR makes the order of crossings irrelevant, so no
//      per-column hit list or sort is needed;
//   3. the toggles become occupancy by an inclusive prefix-XOR along x (in-register per 128-bit
//      group, ballot carry across lanes) and every word of the slab is written exactly once with
//      coalesced 128-bit stores -- no clear pass, no scatter to HBM.
// HBM traffic is the grid (N^3/8 bytes, written once) plus node/triangle reads that mostly hit L2.
#include "kernels.h"

namespace dxrv
{
namespace
{
constexpr int kParityWarps = 4;
constexpr int kParityThreads = kParityWarps * 32;
constexpr int kStackCap = 512;   // entries per warp
constexpr int kStackSlack = 96;  // switch to depth-first popping above kStackCap - kStackSlack
constexpr int kCandCap = 128;    // queued leaf references per warp

__device__ __forceinline__ uint32_t prefixXor32(uint32_t v)
{
    v ^= v << 1; v ^= v << 2; v ^= v << 4; v ^= v << 8; v ^= v << 16;
    return v;
}

// exact sign of edge(P,Q) = P.p*Q.q - P.q*Q.p including the (+e, +e^2) tie rule, given the exactly
// evaluated double value
__device__ __forceinline__ int edgeSignExact(double e, float Pp, float Pq, float Qp, float Qq)
{
    if (e > 0.0) return 1;
    if (e < 0.0) return -1;
    if (Pq != Qq) return (Pq > Qq) ? 1 : -1;
    return (Qp > Pp) - (Qp < Pp);
}

// Spec H, MODE_PARITY: does the line {(s, Y, Z)} cross triangle (a,b,c)?  On a crossing returns the
// first toggled voxel ix in [0, N].
__device__ __forceinline__ bool columnCrossing(const float4& a, const float4& b, const float4& c, float Y, float Z,
                                               uint32_t N, float fN, uint32_t& ixOut)
{
    const float Ap = __fsub_rn(a.y, Y), Aq = __fsub_rn(a.z, Z);
    const float Bp = __fsub_rn(b.y, Y), Bq = __fsub_rn(b.z, Z);
    const float Cp = __fsub_rn(c.y, Y), Cq = __fsub_rn(c.z, Z);
    float U = diffOfProducts(Cp, Bq, Cq, Bp);
    float V = diffOfProducts(Ap, Cq, Aq, Cp);
    float W = diffOfProducts(Bp, Aq, Bq, Ap);
    if (U != 0.0f && V != 0.0f && W != 0.0f)
    {
        // a non-zero float difference of two rounded products has the exact sign
        const bool pos = U > 0.0f;
        if ((V > 0.0f) != pos || (W > 0.0f) != pos) return false;
    }
    else
    {
        const double Ud = diffOfProductsD(Cp, Bq, Cq, Bp);
        const double Vd = diffOfProductsD(Ap, Cq, Aq, Cp);
        const double Wd = diffOfProductsD(Bp, Aq, Bq, Ap);
        const int sU = edgeSignExact(Ud, Cp, Cq, Bp, Bq);
        const int sV = edgeSignExact(Vd, Ap, Aq, Cp, Cq);
        const int sW = edgeSignExact(Wd, Bp, Bq, Ap, Aq);
        if (!(sU == sV && sV == sW && sU != 0)) return false;
        U = (float)Ud; V = (float)Vd; W = (float)Wd;
    }
    const float det = __fadd_rn(__fadd_rn(U, V), W);
    if (det == 0.0f) return false;
    const float d = __fdiv_rn(weighted3(U, a.x, V, b.x, W, c.x), det);

    // smallest x with centre(x) > d: estimate, then fix up with the exact predicate
    float g = floorf((d + 1.0f) * 0.5f * fN + 0.5f);
    if (!(g > 0.0f)) g = 0.0f;
    if (g > fN) g = fN;
    uint32_t ix = (uint32_t)g;
    while (ix > 0 && voxelCentre(ix - 1, fN) > d) --ix;
    while (ix < N && !(voxelCentre(ix, fN) > d)) ++ix;
    ixOut = ix;
    return true;
}

struct ParityParams
{
    const BvhNode* nodes;
    const Tri48* tris;
    uint32_t numTris;
    uint32_t N, P, Ps;      // grid size, words per global row, words per shared row
    uint32_t z0, z1;
    uint32_t tilesY, numTiles;
    uint32_t* grid;
    unsigned long long* crossings;
    uint32_t* err;
};

template <int TY>
__global__ void __launch_bounds__(kParityThreads)
k_trace_fill_columns(const ParityParams prm)
{
    constexpr int TZ = 32 / TY;
    extern __shared__ __align__(16) uint32_t smem[];
    const uint32_t lane = laneId(), warp = threadIdx.x >> 5;
    const uint32_t tile = blockIdx.x * kParityWarps + warp;
    if (tile >= prm.numTiles) return;

    const uint32_t N = prm.N, P = prm.P, Ps = prm.Ps;
    const float fN = (float)N;
    const uint32_t perWarp = 32u * Ps + kStackCap + kCandCap + 64u;
    uint32_t* rows = smem + warp * perWarp;          // [32][Ps] toggle / occupancy bits
    uint32_t* stack = rows + 32u * Ps;
    uint32_t* cand = stack + kStackCap;
    float* tileY = reinterpret_cast<float*>(cand + kCandCap);  // [TY] scene Y of the tile's columns
    float* tileZ = tileY + 32;                                  // [TZ]

    const uint32_t ty0 = (tile % prm.tilesY) * TY;
    const uint32_t tz0 = prm.z0 + (tile / prm.tilesY) * TZ;

    // ---- tile setup ----
    for (uint32_t i = lane; i < 8u * Ps; i += 32u) reinterpret_cast<uint4*>(rows)[i] = make_uint4(0, 0, 0, 0);
    if (lane < TY) tileY[lane] = (ty0 + lane < N) ? -voxelCentre(ty0 + lane, fN) : INFINITY;
    if (lane < TZ) tileZ[lane] = (tz0 + lane < prm.z1) ? voxelCentre(tz0 + lane, fN) : INFINITY;
    __syncwarp();
    // scene Y decreases with y; the last VALID column bounds the rectangle
    const uint32_t yLast = min(ty0 + TY - 1, N - 1) - ty0, zLast = min(tz0 + TZ - 1, prm.z1 - 1) - tz0;
    const float rYmax = tileY[0], rYmin = tileY[yLast];
    const float rZmin = tileZ[0], rZmax = tileZ[zLast];

    uint32_t myCrossings = 0;

    // one queued triangle per lane
    auto processTriangle = [&](uint32_t slot) {
        const float4* t = reinterpret_cast<const float4*>(prm.tris + slot);
        const float4 a = __ldg(t), b = __ldg(t + 1), c = __ldg(t + 2);
        const float ylo = fminf(fminf(a.y, b.y), c.y), yhi = fmaxf(fmaxf(a.y, b.y), c.y);
        const float zlo = fminf(fminf(a.z, b.z), c.z), zhi = fmaxf(fmaxf(a.z, b.z), c.z);
#pragma unroll 1
        for (int zl = 0; zl < TZ; ++zl)
        {
            const float Z = tileZ[zl];
            if (Z < zlo || Z > zhi) continue;
#pragma unroll 1
            for (int yl = 0; yl < TY; ++yl)
            {
                const float Y = tileY[yl];
                if (Y < ylo || Y > yhi) continue;
                uint32_t ix;
                if (columnCrossing(a, b, c, Y, Z, N, fN, ix))
                {
                    ++myCrossings;
                    if (ix < N) atomicXor(&rows[(uint32_t)(zl * TY + yl) * Ps + (ix >> 5)], 1u << (ix & 31u));
                }
            }
        }
    };

    // ---- 1+2: cooperative walk ----
    if (prm.numTris == 1)
    {
        if (lane == 0) processTriangle(0);
    }
    else if (prm.numTris > 1)
    {
        uint32_t sp = 1, nc = 0;  // warp-uniform
        if (lane == 0) stack[0] = 0;
        __syncwarp();
        uint32_t guard = 0;
        while (sp > 0)
        {
            const uint32_t k = (sp <= (uint32_t)(kStackCap - kStackSlack)) ? min(32u, sp) : 1u;
            sp -= k;
            const bool has = lane < k;
            uint32_t c0 = 0, c1 = 0;
            bool ov0 = false, ov1 = false;
            if (has)
            {
                const uint32_t ni = stack[sp + lane];
                const float4* q = reinterpret_cast<const float4*>(prm.nodes + ni);
                const float4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
                const uint4 q3 = __ldg(reinterpret_cast<const uint4*>(q + 3));
                // child0: lo=(q0.x,q0.y,q0.z) hi=(q0.w,q1.x,q1.y); child1: lo=(q1.z,q1.w,q2.x) hi=(q2.y,q2.z,q2.w)
                ov0 = q0.y <= rYmax && q1.x >= rYmin && q0.z <= rZmax && q1.y >= rZmin;
                ov1 = q1.w <= rYmax && q2.z >= rYmin && q2.x <= rZmax && q2.w >= rZmin;
                c0 = q3.x; c1 = q3.y;
            }
            __syncwarp();
            const bool in0 = ov0 && !(c0 & kLeafFlag), in1 = ov1 && !(c1 & kLeafFlag);
            const bool lf0 = ov0 && (c0 & kLeafFlag), lf1 = ov1 && (c1 & kLeafFlag);
            const uint32_t mi0 = __ballot_sync(0xffffffffu, in0), mi1 = __ballot_sync(0xffffffffu, in1);
            const uint32_t ml0 = __ballot_sync(0xffffffffu, lf0), ml1 = __ballot_sync(0xffffffffu, lf1);
            const uint32_t lt = laneMaskLt();
            const uint32_t pushes = __popc(mi0) + __popc(mi1);
            if (sp + pushes > (uint32_t)kStackCap || ++guard > 4u * prm.numTris + 64u)
            {
                if (lane == 0) atomicMax(prm.err, (uint32_t)kErrStackOverflow);
                break;
            }
            if (in0) stack[sp + __popc(mi0 & lt)] = c0;
            if (in1) stack[sp + __popc(mi0) + __popc(mi1 & lt)] = c1;
            sp += pushes;
            if (lf0) cand[nc + __popc(ml0 & lt)] = c0 & ~kLeafFlag;
            if (lf1) cand[nc + __popc(ml0) + __popc(ml1 & lt)] = c1 & ~kLeafFlag;
            nc += __popc(ml0) + __popc(ml1);
            __syncwarp();
            while (nc >= 32u)
            {
                nc -= 32u;
                processTriangle(cand[nc + lane]);
                __syncwarp();
            }
        }
        if (lane < nc) processTriangle(cand[lane]);
    }
    __syncwarp();

    // ---- 3: prefix-XOR along x and write-out, 4 words (128 bits) per lane ----
    const uint32_t groupsPerRow = Ps >> 2;            // 128-bit groups per shared row (power of two or multiple of 32)
    const uint32_t totalGroups = 32u * groupsPerRow;  // multiple of 32
    const uint32_t tailMask = (N & 31u) ? ((1u << (N & 31u)) - 1u) : 0xffffffffu;
    uint32_t runCarry = 0;                            // carry along a row spanning several 32-group chunks
    for (uint32_t g0 = 0; g0 < totalGroups; g0 += 32u)
    {
        const uint32_t g = g0 + lane;
        const uint32_t row = g / groupsPerRow, gi = g - row * groupsPerRow;
        uint4 t = reinterpret_cast<const uint4*>(rows)[g];
        uint32_t par = 0;
        t.x = prefixXor32(t.x); par = t.x >> 31;
        t.y = prefixXor32(t.y) ^ (0u - par); par = t.y >> 31;
        t.z = prefixXor32(t.z) ^ (0u - par); par = t.z >> 31;
        t.w = prefixXor32(t.w) ^ (0u - par); par = t.w >> 31;
        const uint32_t bal = __ballot_sync(0xffffffffu, par);
        uint32_t carry;
        if (groupsPerRow >= 32u)
        {
            // the whole chunk is one row segment
            if ((g0 % groupsPerRow) == 0) runCarry = 0;
            carry = (__popc(bal & laneMaskLt()) & 1u) ^ runCarry;
            runCarry ^= (__popc(bal) & 1u);
        }
        else
        {
            // 32 / groupsPerRow rows per chunk: carry only from lower lanes of the same row.
            // NOTE `par` of a lane already includes its own lower words, but not lower lanes.
            const uint32_t segLo = lane - gi;  // first lane of this row
            const uint32_t segMask = laneMaskLt() & ~((1u << segLo) - 1u);
            carry = __popc(bal & segMask) & 1u;
        }
        // `par` bits were computed without the incoming carry: XOR of the lanes' own parities is
        // exactly the carry because prefix parity is linear.
        const uint32_t flip = 0u - carry;
        t.x ^= flip; t.y ^= flip; t.z ^= flip; t.w ^= flip;

        const uint32_t yl = row % TY, zl = row / TY;
        const uint32_t y = ty0 + yl, z = tz0 + zl;
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

    // ---- statistics ----
    for (int o = 16; o > 0; o >>= 1) myCrossings += __shfl_xor_sync(0xffffffffu, myCrossings, o);
    if (lane == 0 && myCrossings) atomicAdd(prm.crossings, (unsigned long long)myCrossings);
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
}  // namespace

void launchTraceFillColumns(cudaStream_t s, const BvhView& bvh, uint32_t N, uint32_t z0, uint32_t z1, uint32_t* grid,
                            unsigned long long* dCrossings, uint32_t* dErr, int smCount)
{
    (void)smCount;
    constexpr int TY = 8, TZ = 32 / TY;
    ParityParams prm;
    prm.nodes = bvh.nodes; prm.tris = bvh.tris; prm.numTris = bvh.numTris;
    prm.N = N; prm.P = (N + 31) / 32; prm.Ps = sharedRowWords(prm.P);
    prm.z0 = z0; prm.z1 = z1;
    prm.tilesY = (N + TY - 1) / TY;
    prm.numTiles = prm.tilesY * ((z1 - z0 + TZ - 1) / TZ);
    prm.grid = grid; prm.crossings = dCrossings; prm.err = dErr;
    const size_t smemBytes = sizeof(uint32_t) * kParityWarps * (32u * prm.Ps + kStackCap + kCandCap + 64u);
    static bool attrSet[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attrSet[dev])
    {
        cudaFuncSetAttribute(k_trace_fill_columns<TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        attrSet[dev] = true;
    }
    cudaMemsetAsync(dCrossings, 0, sizeof(unsigned long long), s);
    const uint32_t blocks = (prm.numTiles + kParityWarps - 1) / kParityWarps;
    k_trace_fill_columns<TY><<<blocks, kParityThreads, smemBytes, s>>>(prm);
}
}  // namespace dxrv
