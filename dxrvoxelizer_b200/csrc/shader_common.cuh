// shader_common.cuh -- MODE_SHADER per-ray / per-pair arithmetic (Spec H of oracle/dxrv_oracle.h) shared by the two
// closest-hit kernels (shader_bins.cu: direction bins, the default; trace_shader.cu: LBVH walk, the overflow path).
//
//   generateRay     Content/Shaders/DXRVoxelizer.hlsl:44-53   pos=(idx+.5)/N*2-1, pos.y=-pos.y, D=normalize(pos)
//   TraceRay        DXRVoxelizer.hlsl:80                      closest hit, no culling, 0 < t < 10000
//   closestHitMain  DXRVoxelizer.hlsl:90-119,132-140          normal lerp, dot(normalize(N), D) > 0.12
// Both kernels evaluate exactly these functions for every (ray, triangle) pair they do not cull, and cull only
// pairs that provably fail them, so their grids are identical bit for bit.
#pragma once
#include "common.cuh"

namespace dxrv
{
__device__ __forceinline__ float pick(float x, float y, float z, int k) { return k == 0 ? x : (k == 1 ? y : z); }

struct RaySetup
{
    float Ox, Oy, Oz, Dx, Dy, Dz, ix, iy, iz;
    float Sx, Sy, Sz;
    int kx, ky, kz;
};

struct BestHit
{
    float tc;
    uint32_t prim;
    float bx, by;
};

// O must be set; len = sqrt((Ox*Ox + Oy*Oy) + Oz*Oz) (Spec H normalize), not zero
__device__ __forceinline__ void raySetup(RaySetup& r, float len)
{
    r.Dx = __fdiv_rn(r.Ox, len); r.Dy = __fdiv_rn(r.Oy, len); r.Dz = __fdiv_rn(r.Oz, len);
    r.ix = __fdiv_rn(1.0f, r.Dx); r.iy = __fdiv_rn(1.0f, r.Dy); r.iz = __fdiv_rn(1.0f, r.Dz);
    if (r.ix > kFltMax) r.ix = kFltMax; if (r.ix < -kFltMax) r.ix = -kFltMax;
    if (r.iy > kFltMax) r.iy = kFltMax; if (r.iy < -kFltMax) r.iy = -kFltMax;
    if (r.iz > kFltMax) r.iz = kFltMax; if (r.iz < -kFltMax) r.iz = -kFltMax;
    int kz = 0; float m = fabsf(r.Dx);
    if (fabsf(r.Dy) > m) { kz = 1; m = fabsf(r.Dy); }
    if (fabsf(r.Dz) > m) { kz = 2; }
    int kx = (kz + 1) % 3, ky = (kx + 1) % 3;
    const float dz = pick(r.Dx, r.Dy, r.Dz, kz);
    if (dz < 0.0f) { const int t = kx; kx = ky; ky = t; }
    r.kx = kx; r.ky = ky; r.kz = kz;
    r.Sx = __fdiv_rn(pick(r.Dx, r.Dy, r.Dz, kx), dz);
    r.Sy = __fdiv_rn(pick(r.Dx, r.Dy, r.Dz, ky), dz);
    r.Sz = __fdiv_rn(1.0f, dz);
}

__device__ __forceinline__ float rayLength(float Ox, float Oy, float Oz)
{
    return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(Ox, Ox), __fmul_rn(Oy, Oy)), __fmul_rn(Oz, Oz)));
}

__device__ __forceinline__ bool slabTest(const RaySetup& r, float lox, float loy, float loz, float hix, float hiy,
                                         float hiz, float& tin, float& tout)
{
    tin = 0.0f; tout = kTMax;
    float t0 = __fmul_rn(__fsub_rn(lox, r.Ox), r.ix), t1 = __fmul_rn(__fsub_rn(hix, r.Ox), r.ix);
    tin = fmaxsel(fminsel(t0, t1), tin); tout = fminsel(fmaxsel(t0, t1), tout);
    t0 = __fmul_rn(__fsub_rn(loy, r.Oy), r.iy); t1 = __fmul_rn(__fsub_rn(hiy, r.Oy), r.iy);
    tin = fmaxsel(fminsel(t0, t1), tin); tout = fminsel(fmaxsel(t0, t1), tout);
    t0 = __fmul_rn(__fsub_rn(loz, r.Oz), r.iz); t1 = __fmul_rn(__fsub_rn(hiz, r.Oz), r.iz);
    tin = fmaxsel(fminsel(t0, t1), tin); tout = fminsel(fmaxsel(t0, t1), tout);
    return tin <= tout;
}

// Spec H steps 1-4 for one (ray, triangle) pair
__device__ __forceinline__ void testTriangle(const RaySetup& r, const Tri48* __restrict__ tris, uint32_t slot, BestHit& best)
{
    const float4* t = reinterpret_cast<const float4*>(tris + slot);
    const float4 a = __ldg(t), b = __ldg(t + 1), c = __ldg(t + 2);
    const uint32_t prim = __float_as_uint(a.w);
    float tin, tout;
    if (!slabTest(r, fminsel(fminsel(a.x, b.x), c.x), fminsel(fminsel(a.y, b.y), c.y), fminsel(fminsel(a.z, b.z), c.z),
                  fmaxsel(fmaxsel(a.x, b.x), c.x), fmaxsel(fmaxsel(a.y, b.y), c.y), fmaxsel(fmaxsel(a.z, b.z), c.z), tin, tout))
        return;
    const float Ax3 = __fsub_rn(a.x, r.Ox), Ay3 = __fsub_rn(a.y, r.Oy), Az3 = __fsub_rn(a.z, r.Oz);
    const float Bx3 = __fsub_rn(b.x, r.Ox), By3 = __fsub_rn(b.y, r.Oy), Bz3 = __fsub_rn(b.z, r.Oz);
    const float Cx3 = __fsub_rn(c.x, r.Ox), Cy3 = __fsub_rn(c.y, r.Oy), Cz3 = __fsub_rn(c.z, r.Oz);
    const float Akz = pick(Ax3, Ay3, Az3, r.kz), Bkz = pick(Bx3, By3, Bz3, r.kz), Ckz = pick(Cx3, Cy3, Cz3, r.kz);
    const float Ax = __fsub_rn(pick(Ax3, Ay3, Az3, r.kx), __fmul_rn(r.Sx, Akz));
    const float Ay = __fsub_rn(pick(Ax3, Ay3, Az3, r.ky), __fmul_rn(r.Sy, Akz));
    const float Bx = __fsub_rn(pick(Bx3, By3, Bz3, r.kx), __fmul_rn(r.Sx, Bkz));
    const float By = __fsub_rn(pick(Bx3, By3, Bz3, r.ky), __fmul_rn(r.Sy, Bkz));
    const float Cx = __fsub_rn(pick(Cx3, Cy3, Cz3, r.kx), __fmul_rn(r.Sx, Ckz));
    const float Cy = __fsub_rn(pick(Cx3, Cy3, Cz3, r.ky), __fmul_rn(r.Sy, Ckz));
    float U, V, W;
    edgeValues(Ax, Ay, Bx, By, Cx, Cy, U, V, W);
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return;
    const float det = __fadd_rn(__fadd_rn(U, V), W);
    if (det == 0.0f) return;
    const float tt = __fdiv_rn(weighted3(U, __fmul_rn(r.Sz, Akz), V, __fmul_rn(r.Sz, Bkz), W, __fmul_rn(r.Sz, Ckz)), det);
    const float tc = fminsel(fmaxsel(tt, tin), tout);
    if (!(tc > 0.0f && tc < kTMax)) return;
    if (tc < best.tc || (tc == best.tc && prim < best.prim))
    {
        best.tc = tc; best.prim = prim;
        best.bx = __fdiv_rn(V, det); best.by = __fdiv_rn(W, det);
    }
}

__device__ __forceinline__ uint32_t unorm10(float v)
{
    if (!(v > 0.0f)) return 0u;
    if (v > 1.0f) v = 1.0f;
    return (uint32_t)__fadd_rn(__fmul_rn(v, 1023.0f), 0.5f);
}

struct ShaderParams
{
    const BvhNode* nodes;
    const Tri48* tris;
    uint32_t numTris;
    const uint8_t* verts;
    uint32_t stride;
    const uint32_t* indices;
    uint32_t N, P, z0;
    uint64_t numWords;
    uint32_t* grid;
    uint32_t* texels;
    uint32_t* err;
    const uint32_t* binsState;   // [1] != 0: the direction bins overflowed their budget -> the LBVH walk runs instead
};

// closestHitMain: interpolate the (object-space) vertex normals with the barycentrics of vertices 1 and 2,
// normalise, compare against the ray direction.  Returns inside; texel = the UAV value when inside.
__device__ __forceinline__ bool shadeHit(const ShaderParams& prm, const RaySetup& r, const BestHit& best, uint32_t& texel)
{
    const uint32_t i0 = __ldg(prm.indices + 3 * (size_t)best.prim), i1 = __ldg(prm.indices + 3 * (size_t)best.prim + 1),
                   i2 = __ldg(prm.indices + 3 * (size_t)best.prim + 2);
    const float* n0 = reinterpret_cast<const float*>(prm.verts + (size_t)prm.stride * i0 + 12);
    const float* n1 = reinterpret_cast<const float*>(prm.verts + (size_t)prm.stride * i1 + 12);
    const float* n2 = reinterpret_cast<const float*>(prm.verts + (size_t)prm.stride * i2 + 12);
    float nrm[3];
#pragma unroll
    for (int q = 0; q < 3; ++q)
    {
        const float v0 = __ldg(n0 + q), v1 = __ldg(n1 + q), v2 = __ldg(n2 + q);
        nrm[q] = __fadd_rn(__fadd_rn(v0, __fmul_rn(best.bx, __fsub_rn(v1, v0))), __fmul_rn(best.by, __fsub_rn(v2, v0)));
    }
    const float nl = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(nrm[0], nrm[0]), __fmul_rn(nrm[1], nrm[1])), __fmul_rn(nrm[2], nrm[2])));
    const float nx = __fdiv_rn(nrm[0], nl), ny = __fdiv_rn(nrm[1], nl), nz = __fdiv_rn(nrm[2], nl);
    const float d = __fadd_rn(__fadd_rn(__fmul_rn(nx, r.Dx), __fmul_rn(ny, r.Dy)), __fmul_rn(nz, r.Dz));
    const bool inside = d > kThreshold;
    texel = inside ? (unorm10(nx) | (unorm10(ny) << 10) | (unorm10(nz) << 20) | (3u << 30)) : 0u;
    return inside;
}
}  // namespace dxrv
