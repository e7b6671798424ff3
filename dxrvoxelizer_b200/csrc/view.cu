// view.cu -- headless port of the reference's viewer pass (the step right after the voxelization path).
//
// Restates Content/Shaders/PSRayCast.hlsl:61-187 (ScreenToLocal, ComputeStartPoint, GetSample, main) and
// VSScreenQuad.hlsl: a full-screen pass that ray-marches the voxel grid (128 steps, 32 light steps per
// lit sample, trilinear LINEAR_CLAMP sampling of the grid's alpha = occupancy, PSRayCast.hlsl:108) and
// writes an R8G8B8A8 image.  min16float is evaluated in fp32 (the reference allows either precision).
// Not on the hot path; SURVEY.md section 8f item 3.  Checked against oracle/dxrv_oracle.c (tolerance on
// the 8-bit image, since HLSL's normalize/sqrt and the sampler's weights are implementation defined).
#include "kernels.h"

namespace dxrv
{
namespace
{
constexpr int kNumSamples = 128;       // NUM_SAMPLES
constexpr int kNumLightSamples = 32;   // NUM_LIGHT_SAMPLES
constexpr float kAbsorption = 1.0f;
constexpr float kZeroThreshold = 0.01f;

struct ViewParams
{
    const uint32_t* grid;
    uint32_t N, P;
    uint32_t width, height;
    float m[16];        // screenToLocal, row-vector convention: p' = (x, y, z, 1) * M
    float eye[3], light[3];
    uint32_t* image;    // RGBA8
};

__device__ __forceinline__ float occupancyAt(const ViewParams& v, int x, int y, int z)
{
    const int n = (int)v.N - 1;
    x = min(max(x, 0), n); y = min(max(y, 0), n); z = min(max(z, 0), n);   // CLAMP addressing
    return (float)((__ldg(v.grid + ((size_t)z * v.N + y) * v.P + (x >> 5)) >> (x & 31)) & 1u);
}

// GetSample: trilinear fetch of the alpha channel (1 inside, 0 outside), then min(density * 8, 16)
__device__ __forceinline__ float getSample(const ViewParams& v, float tx, float ty, float tz)
{
    const float fN = (float)v.N;
    const float ux = tx * fN - 0.5f, uy = ty * fN - 0.5f, uz = tz * fN - 0.5f;
    const float x0 = floorf(ux), y0 = floorf(uy), z0 = floorf(uz);
    const float fx = ux - x0, fy = uy - y0, fz = uz - z0;
    const int ix = (int)x0, iy = (int)y0, iz = (int)z0;
    const float c000 = occupancyAt(v, ix, iy, iz), c100 = occupancyAt(v, ix + 1, iy, iz);
    const float c010 = occupancyAt(v, ix, iy + 1, iz), c110 = occupancyAt(v, ix + 1, iy + 1, iz);
    const float c001 = occupancyAt(v, ix, iy, iz + 1), c101 = occupancyAt(v, ix + 1, iy, iz + 1);
    const float c011 = occupancyAt(v, ix, iy + 1, iz + 1), c111 = occupancyAt(v, ix + 1, iy + 1, iz + 1);
    const float c00 = c000 + fx * (c100 - c000), c10 = c010 + fx * (c110 - c010);
    const float c01 = c001 + fx * (c101 - c001), c11 = c011 + fx * (c111 - c011);
    const float c0 = c00 + fy * (c10 - c00), c1 = c01 + fy * (c11 - c01);
    const float density = c0 + fz * (c1 - c0);
    return fminf(density * 8.0f, 16.0f);
}

__device__ __forceinline__ float signf(float a) { return (a > 0.0f) ? 1.0f : ((a < 0.0f) ? -1.0f : 0.0f); }
__device__ __forceinline__ float saturate(float a) { return fminf(fmaxf(a, 0.0f), 1.0f); }
__device__ __forceinline__ uint32_t unorm8(float a) { return (uint32_t)(saturate(a) * 255.0f + 0.5f); }

__global__ void __launch_bounds__(256)
k_raycast_view(const ViewParams v)
{
    const uint32_t px = blockIdx.x * 16 + (threadIdx.x & 15), py = blockIdx.y * 16 + (threadIdx.x >> 4);
    if (px >= v.width || py >= v.height) return;
    const float clearColor[3] = {0.0f, 0.2f, 0.4f};   // CLEAR_COLOR, Content/SharedConst.h:8

    // ScreenToLocal(float3(sspos.xy, 0)): the point on the near plane
    const float sx = (float)px + 0.5f, sy = (float)py + 0.5f;
    float h[4];
    for (int c = 0; c < 4; ++c) h[c] = sx * v.m[0 * 4 + c] + sy * v.m[1 * 4 + c] + v.m[3 * 4 + c];   // z = 0
    float pos[3] = {h[0] / h[3], h[1] / h[3], h[2] / h[3]};
    float dir[3] = {pos[0] - v.eye[0], pos[1] - v.eye[1], pos[2] - v.eye[2]};
    const float dl = sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
    dir[0] /= dl; dir[1] /= dl; dir[2] /= dl;

    // ComputeStartPoint
    bool isHit = true;
    if (!(fabsf(pos[0]) <= 1.0f && fabsf(pos[1]) <= 1.0f && fabsf(pos[2]) <= 1.0f))
    {
        float U = 3.402823466e+38f;
        isHit = false;
        for (int i = 0; i < 3; ++i)
        {
            const float u = (-signf(dir[i]) - pos[i]) / dir[i];
            if (u < 0.0f) continue;
            const int j = (i + 1) % 3, k = (i + 2) % 3;
            if (fabsf(dir[j] * u + pos[j]) > 1.0f) continue;
            if (fabsf(dir[k] * u + pos[k]) > 1.0f) continue;
            if (u < U) { U = u; isHit = true; }
        }
        for (int i = 0; i < 3; ++i) pos[i] = fminf(fmaxf(dir[i] * U + pos[i], -1.0f), 1.0f);
    }
    if (!isHit)
    {
        v.image[(size_t)py * v.width + px] = unorm8(clearColor[0]) | (unorm8(clearColor[1]) << 8) | (unorm8(clearColor[2]) << 16);
        return;
    }

    const float maxDist = 2.0f * sqrtf(3.0f);
    const float stepScale = maxDist / kNumSamples, lightStepScale = maxDist / kNumLightSamples;
    const float step[3] = {dir[0] * stepScale, dir[1] * stepScale, dir[2] * stepScale};
    const float ll = sqrtf(v.light[0] * v.light[0] + v.light[1] * v.light[1] + v.light[2] * v.light[2]);
    const float lightStep[3] = {v.light[0] / ll * lightStepScale, v.light[1] / ll * lightStepScale, v.light[2] / ll * lightStepScale};

    float transmit = 1.0f, scatter = 0.0f;
    for (int i = 0; i < kNumSamples; ++i)
    {
        if (fabsf(pos[0]) > 1.0f || fabsf(pos[1]) > 1.0f || fabsf(pos[2]) > 1.0f) break;
        const float density = getSample(v, 0.5f * pos[0] + 0.5f, -0.5f * pos[1] + 0.5f, 0.5f * pos[2] + 0.5f);
        if (density > kZeroThreshold)
        {
            const float scaledDens = density * stepScale;
            transmit *= saturate(1.0f - scaledDens * kAbsorption);
            if (transmit < kZeroThreshold) break;
            float lightTrans = 1.0f;
            float lp[3] = {pos[0] + lightStep[0], pos[1] + lightStep[1], pos[2] + lightStep[2]};
            for (int j = 0; j < kNumLightSamples; ++j)
            {
                if (fabsf(lp[0]) > 1.0f || fabsf(lp[1]) > 1.0f || fabsf(lp[2]) > 1.0f) break;
                const float lightDens = getSample(v, 0.5f * lp[0] + 0.5f, -0.5f * lp[1] + 0.5f, 0.5f * lp[2] + 0.5f);
                lightTrans *= saturate(1.0f - kAbsorption * lightStepScale * lightDens);
                if (lightTrans < kZeroThreshold) break;
                lp[0] += lightStep[0]; lp[1] += lightStep[1]; lp[2] += lightStep[2];
            }
            scatter += lightTrans * transmit * scaledDens;
        }
        pos[0] += step[0]; pos[1] += step[1]; pos[2] += step[2];
    }
    uint32_t rgba = 0xff000000u;
    for (int c = 0; c < 3; ++c)
    {
        float r = scatter * 0.8f + 0.2f;
        const float cc = clearColor[c] * clearColor[c];
        r = r + transmit * (cc - r);            // lerp(result, clear^2, transmit)
        rgba |= unorm8(sqrtf(r)) << (8 * c);
    }
    v.image[(size_t)py * v.width + px] = rgba;
}
}  // namespace

void launchRaycastView(cudaStream_t s, const uint32_t* grid, uint32_t N, uint32_t width, uint32_t height,
                       const float screenToLocal[16], const float eye[3], const float light[3], uint32_t* image)
{
    ViewParams v;
    v.grid = grid; v.N = N; v.P = (N + 31) / 32; v.width = width; v.height = height; v.image = image;
    for (int i = 0; i < 16; ++i) v.m[i] = screenToLocal[i];
    for (int i = 0; i < 3; ++i) { v.eye[i] = eye[i]; v.light[i] = light[i]; }
    const dim3 blocks((width + 15) / 16, (height + 15) / 16);
    k_raycast_view<<<blocks, 256, 0, s>>>(v);
}
}  // namespace dxrv
