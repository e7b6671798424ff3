// parity_common.cuh -- MODE_PARITY arithmetic shared by the tile kernels (trace_parity.cu) and the
// triangle-parallel kernels (scatter_parity.cu): Spec H crossing test (oracle/dxrv_oracle.h) and the conservative
// culling helpers.  The culling may only ever add (triangle, column) pairs, never drop one that crosses.
#pragma once
#include "common.cuh"

namespace dxrv
{
__device__ __forceinline__ uint32_t prefixXor32(uint32_t v)
{
    v ^= v << 1; v ^= v << 2; v ^= v << 4; v ^= v << 8; v ^= v << 16;
    return v;
}

// centre(i) of Spec H.  For a power-of-two grid dividing by N is exact scaling, so multiplying by
// 1/N is bit-identical to the IEEE division and several times cheaper.
__device__ __forceinline__ float centreOf(uint32_t i, float fN, float invNPow2)
{
    const float h = __fadd_rn((float)i, 0.5f);
    const float q = (invNPow2 != 0.0f) ? __fmul_rn(h, invNPow2) : __fdiv_rn(h, fN);
    return __fsub_rn(__fmul_rn(q, 2.0f), 1.0f);
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
                                               uint32_t N, float fN, float invNPow2, uint32_t& ixOut)
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

    // smallest x with centre(x) > d: estimate, then fix up with the exact predicate.  The estimate's
    // error is far below 1/64 of a voxel for every supported N, so the (exact, but costlier) fix-up is
    // only needed when d sits that close to a voxel centre.
    const float gf = (d + 1.0f) * 0.5f * fN + 0.5f;
    float g = floorf(gf);
    const float fr = gf - g;
    const bool nearCentre = !(fr > 0.015625f && fr < 0.984375f);   // also true for NaN / inf
    if (!(g > 0.0f)) g = 0.0f;
    if (g > fN) g = fN;
    uint32_t ix = (uint32_t)g;
    if (nearCentre)
    {
        while (ix > 0 && centreOf(ix - 1, fN, invNPow2) > d) --ix;
        while (ix < N && !(centreOf(ix, fN, invNPow2) > d)) ++ix;
    }
    ixOut = ix;
    return true;
}

// index units; covers the rounding of the float index estimates (floor/ceil of (1 +- v) * N/2 - 0.5) for any N
constexpr float kIdxSlack = 0.02f;

// Conservative y extent of triangle (a,b,c) within the row z = Zc.  A crossing of column (Y, Zc)
// means (0,0) lies in the triangle of the ROUNDED differences (a.y - Y, a.z - Zc)..., whose
// vertices are within 2^-23 of the exact ones: so some point of the exact triangle lies within
// 2^-23 of (Y, Zc) in y and in z, i.e. Y is within 2^-23 of the y extent of the triangle inside the
// slab |z - Zc| <= m (m = 1e-6 > 2^-23).  That extent is spanned by the parts of the edges inside
// the slab: a steep edge (|dz| >= 64 m) stays within |dy| / 64 of its point at Zc there, a shallow
// one is taken whole.  1e-5 on top covers the float evaluation.
__device__ __forceinline__ void rowIntervalEdge(const float4& P, const float4& Q, float Zc, float& lo, float& hi)
{
    const float m = 1e-6f;
    if (fmaxf(P.z, Q.z) < Zc - m || fminf(P.z, Q.z) > Zc + m) return;
    const float dz = Q.z - P.z, dy = Q.y - P.y;
    float l = fminf(P.y, Q.y), u = fmaxf(P.y, Q.y);
    if (fabsf(dz) >= 64.0f * m)
    {
        const float t = fminf(fmaxf(__fdividef(Zc - P.z, dz), 0.0f), 1.0f);
        const float yc = P.y + t * dy, e = fabsf(dy) * (1.0f / 64.0f);
        l = yc - e; u = yc + e;
    }
    lo = fminf(lo, l); hi = fmaxf(hi, u);
}
__device__ __forceinline__ void rowIntervalY(const float4& a, const float4& b, const float4& c, float Zc, float& lo, float& hi)
{
    lo = INFINITY; hi = -INFINITY;
    rowIntervalEdge(a, b, Zc, lo, hi); rowIntervalEdge(b, c, Zc, lo, hi); rowIntervalEdge(c, a, Zc, lo, hi);
    lo -= 1e-5f; hi += 1e-5f;
}
}  // namespace dxrv
