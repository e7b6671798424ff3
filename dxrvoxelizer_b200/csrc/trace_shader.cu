// trace_shader.cu -- MODE_SHADER: exact restatement of the reference's DXR shaders on sm_100a.
//
//   raygenMain      Content/Shaders/DXRVoxelizer.hlsl:58-85   index decode, radial ray, store
//   generateRay     DXRVoxelizer.hlsl:44-53                   pos=(idx+.5)/N*2-1, pos.y=-pos.y, D=normalize(pos)
//   TraceRay        DXRVoxelizer.hlsl:80                      closest hit, no culling, 0 < t < 10000
//   closestHitMain  DXRVoxelizer.hlsl:90-119,132-140          normal lerp, dot(normalize(N), D) > 0.12
//   missMain        DXRVoxelizer.hlsl:145-148                 nothing
// The driver-defined parts (BVH, ray/triangle arithmetic) follow Spec H of oracle/dxrv_oracle.h:
// watertight test, hit distance clamped into the triangle's own slab interval, closest hit = the
// lexicographic minimum of (tc, primitive index) -- so the result does not depend on the hierarchy.
//
// This file is the GENERAL closest-hit kernel: per-lane stack traversal of the LBVH with near-child-first
// ordering (one warp = one 32-voxel word, one __ballot_sync, one store).  Since round 2 the default MODE_SHADER
// path is shader_bins.cu (direction bins: an exact accelerator for this radial ray family); this kernel runs
// when the bins' entry budget overflows (device-side flag, no host round trip) or when DXRV_SHADER_PATH=bvh.
#include "kernels.h"
#include "shader_common.cuh"

namespace dxrv
{
namespace
{
constexpr int kShaderThreads = 128;
constexpr int kLocalStack = 64;  // LBVH depth <= 62 (30 key bits + 32 index bits)

__global__ void __launch_bounds__(kShaderThreads)
k_trace_shader(const ShaderParams prm)
{
    // runs only when the direction bins are not usable (shader_bins.cu): forced, or their entry budget overflowed
    if (prm.binsState && __ldg(prm.binsState + 1) == 0u) return;
    const uint32_t lane = laneId();
    const uint32_t N = prm.N, P = prm.P;
    const float fN = (float)N;
    for (uint64_t word = (uint64_t)blockIdx.x * (kShaderThreads / 32) + (threadIdx.x >> 5); word < prm.numWords;
         word += (uint64_t)gridDim.x * (kShaderThreads / 32))
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
            raySetup(r, rayLength(r.Ox, r.Oy, r.Oz));
            BestHit best;
            best.tc = INFINITY; best.prim = 0xffffffffu; best.bx = 0.0f; best.by = 0.0f;

            if (prm.numTris == 1) testTriangle(r, prm.tris, 0, best);
            else
            {
                uint32_t stack[kLocalStack];
                int sp = 0;
                uint32_t node = 0;
                uint32_t guard = 0;
                while (true)
                {
                    const float4* q = reinterpret_cast<const float4*>(prm.nodes + node);
                    const float4 yz0 = __ldg(q), yz1 = __ldg(q + 1), x01 = __ldg(q + 2);
                    const uint4 q3 = __ldg(reinterpret_cast<const uint4*>(q + 3));
                    float tin0, tout0, tin1, tout1;
                    bool h0 = slabTest(r, x01.x, yz0.x, yz0.z, x01.y, yz0.y, yz0.w, tin0, tout0) && !(tin0 > best.tc);
                    bool h1 = slabTest(r, x01.z, yz1.x, yz1.z, x01.w, yz1.y, yz1.w, tin1, tout1) && !(tin1 > best.tc);
                    if (h0 && (q3.x & kLeafFlag)) { testTriangle(r, prm.tris, q3.x & ~kLeafFlag, best); h0 = false; }
                    if (h1 && (q3.y & kLeafFlag))
                    {
                        if (!(tin1 > best.tc)) testTriangle(r, prm.tris, q3.y & ~kLeafFlag, best);
                        h1 = false;
                    }
                    if (h0 && h1)
                    {
                        const bool firstIs1 = tin1 < tin0;
                        if (sp >= kLocalStack || ++guard > 4u * prm.numTris + 64u) { atomicMax(prm.err, (uint32_t)kErrStackOverflow); break; }
                        stack[sp++] = firstIs1 ? q3.x : q3.y;
                        node = firstIs1 ? q3.y : q3.x;
                    }
                    else if (h0) node = q3.x;
                    else if (h1) node = q3.y;
                    else
                    {
                        if (sp == 0) break;
                        node = stack[--sp];
                    }
                    if (++guard > 8u * prm.numTris + 64u) { atomicMax(prm.err, (uint32_t)kErrStackOverflow); break; }
                }
            }
            if (best.prim != 0xffffffffu) inside = shadeHit(prm, r, best, texel);
        }

        const uint32_t bits = __ballot_sync(0xffffffffu, inside);
        if (lane == 0) prm.grid[word] = bits;
        if (prm.texels && x < N) prm.texels[row * N + x] = texel;
    }
}
}  // namespace

void launchTraceShaderBvh(cudaStream_t s, const BvhView& bvh, const MeshView& m, uint32_t N, uint32_t z0, uint32_t z1,
                          uint32_t* grid, uint32_t* texels, uint32_t* dErr, const uint32_t* binsState)
{
    ShaderParams prm;
    prm.nodes = bvh.nodes; prm.tris = bvh.tris; prm.numTris = bvh.numTris;
    prm.verts = m.verts; prm.stride = m.stride; prm.indices = m.indices;
    prm.N = N; prm.P = (N + 31) / 32; prm.z0 = z0;
    prm.numWords = (uint64_t)(z1 - z0) * N * prm.P;
    prm.grid = grid; prm.texels = texels; prm.err = dErr; prm.binsState = binsState;
    // grid-stride over the words: any slab of any N <= 8192 fits (a one-shot grid overflowed 2^31-1 blocks from N ~ 6500)
    const uint64_t want = (prm.numWords + (kShaderThreads / 32) - 1) / (kShaderThreads / 32);
    const uint64_t cap = 148ull * 16ull;
    k_trace_shader<<<(unsigned)(want < cap ? want : cap), kShaderThreads, 0, s>>>(prm);
}
}  // namespace dxrv
