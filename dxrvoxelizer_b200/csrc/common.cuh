// common.cuh -- device data layout and the fp32 arithmetic contract shared by the sm_100a kernels.
//
// The contract ("Spec H") is stated normatively in oracle/dxrv_oracle.h; the CPU oracle implements
// it in plain C with contraction disabled, the kernels implement it here with explicit
// round-to-nearest intrinsics (__fmul_rn / __fadd_rn / __fsub_rn / __fdiv_rn / __fsqrt_rn are never
// contracted into FMAs by nvcc).  The two implementations are independent; tests require them to
// agree bit for bit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dxrv
{
constexpr float kTMax = 10000.0f;      // ray.TMax, Content/Shaders/DXRVoxelizer.hlsl:77
constexpr float kThreshold = 0.12f;    // THRESHOLD, DXRVoxelizer.hlsl:5
constexpr uint32_t kLeafFlag = 0x80000000u;
constexpr float kFltMax = 3.402823466e+38f;

// ---- HBM layout ---------------------------------------------------------------------------------
// Internal node i (0 <= i < T-1; node 0 is the root).  64 bytes = two 32 B sectors, fetched as
// 128-bit loads.  Each node stores the boxes of its two CHILDREN, so one fetch decides both.  The
// (y,z) extents come first: the column tracer (rays along x) needs only yz0, yz1 and the child
// references (three loads), the closest-hit tracer reads all four.
//   yz0 = child0 (ylo, yhi, zlo, zhi)      yz1 = child1 (ylo, yhi, zlo, zhi)
//   x01 = (child0 xlo, child0 xhi, child1 xlo, child1 xhi)
//   c0, c1 = child references: kLeafFlag | sortedTriangleSlot, or internal node index
struct __align__(16) BvhNode
{
    float4 yz0, yz1, x01;
    uint32_t c0, c1;
    uint32_t pad0, pad1;
};
static_assert(sizeof(BvhNode) == 64, "BvhNode must be 64 bytes");

// Scene-space triangle in Morton-sorted order: 48 bytes, three 128-bit loads.
//   a.w carries the ORIGINAL primitive index (PrimitiveIndex() of the reference, hlsl:93).
struct __align__(16) Tri48
{
    float4 a, b, c;
};
static_assert(sizeof(Tri48) == 48, "Tri48 must be 48 bytes");

// Device-side status word: set by kernels, reported lazily by the host API.
enum DeviceError : uint32_t
{
    kErrNone = 0,
    kErrBadIndex = 1,       // an index >= numVerts
    kErrStackOverflow = 2,  // traversal stack exhausted (never expected)
    kErrBarrierTimeout = 3, // the fused build's grid barrier gave up (its CTAs were not co-resident)
};

// ---- arithmetic contract -----------------------------------------------------------------------
__device__ __forceinline__ float fminsel(float a, float b) { return (a < b) ? a : b; }
__device__ __forceinline__ float fmaxsel(float a, float b) { return (a > b) ? a : b; }

// centre(i) = ((float)i + 0.5f) / (float)N * 2.0f - 1.0f            (hlsl:46)
__device__ __forceinline__ float voxelCentre(uint32_t i, float fN)
{
    return __fsub_rn(__fmul_rn(__fdiv_rn(__fadd_rn((float)i, 0.5f), fN), 2.0f), 1.0f);
}

// p' = (p - c) / w per component                                    (Voxelizer.cpp:304-310)
__device__ __forceinline__ float toScene(float p, float c, float w) { return __fdiv_rn(__fsub_rn(p, c), w); }

// a*b - c*d with each product and the difference rounded once
__device__ __forceinline__ float diffOfProducts(float a, float b, float c, float d)
{
    return __fsub_rn(__fmul_rn(a, b), __fmul_rn(c, d));
}
__device__ __forceinline__ double diffOfProductsD(float a, float b, float c, float d)
{
    return __dsub_rn(__dmul_rn((double)a, (double)b), __dmul_rn((double)c, (double)d));
}

// U = edge(C,B), V = edge(A,C), W = edge(B,A) with edge(P,Q) = P.p*Q.q - P.q*Q.p; float first, all
// three recomputed in double (and rounded to float) when any is exactly zero.
__device__ __forceinline__ void edgeValues(float Ap, float Aq, float Bp, float Bq, float Cp, float Cq,
                                           float& U, float& V, float& W)
{
    U = diffOfProducts(Cp, Bq, Cq, Bp);
    V = diffOfProducts(Ap, Cq, Aq, Cp);
    W = diffOfProducts(Bp, Aq, Bq, Ap);
    if (U == 0.0f || V == 0.0f || W == 0.0f)
    {
        U = (float)diffOfProductsD(Cp, Bq, Cq, Bp);
        V = (float)diffOfProductsD(Ap, Cq, Aq, Cp);
        W = (float)diffOfProductsD(Bp, Aq, Bq, Ap);
    }
}

// (U*x + V*y) + W*z, every operation rounded once
__device__ __forceinline__ float weighted3(float U, float x, float V, float y, float W, float z)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(U, x), __fmul_rn(V, y)), __fmul_rn(W, z));
}

__device__ __forceinline__ uint32_t laneId() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t laneMaskLt()
{
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
}  // namespace dxrv
