// ctx.h -- the context behind the opaque dxrv_ctx handle and the small helpers shared by api.cu and comm.cu.
// Internal to libdxrv.so.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>   // header-only: ranges cost nothing unless a profiler injects itself

#include <cstdio>
#include <string>
#include <vector>

#include "../../include/dxrv.h"
#include "kernels.h"

namespace dxrv
{
std::string& globalError();  // obj_capi.cpp
}

using namespace dxrv;

struct dxrv_ctx
{
    int device = 0;
    int smCount = 148;
    cudaStream_t ownStream = nullptr;
    SideStream side{};               // fork/join partner of `stream` inside a build
    cudaStream_t stream = nullptr;
    std::string err;

    // mesh (either borrowed device pointers or owned staging copies)
    uint8_t* vertsOwned = nullptr; size_t vertsCap = 0;
    uint32_t* idxOwned = nullptr;  size_t idxCap = 0;
    MeshView mesh{};
    bool haveBvh = false;
    // The hierarchy (topology + child boxes) is built with the rest when the previous consumer traversed it, else on
    // demand by the first dxrv_voxelize that does (include/dxrv.h, dxrv_build_bvh).
    bool treeBuilt = false, treeWanted = false;
    // Small meshes are built by ONE cooperative kernel (k_build_fused); it writes no box pyramid, so a hierarchy asked for
    // later redoes the leaf pass (pyramidBuilt).  fusedBuild goes off for good if the device cannot run it.
    bool fusedBuild = true, pyramidBuilt = false;
    void* fusedScratch = nullptr;
    // MODE_PARITY candidate lists in walkBuf: valid for the current acceleration structure and exactly this grid / slab
    // (left by a tile-path voxelize; the next voxelize with the same parameters starts at k_file_columns)
    struct BinState { bool valid = false; uint32_t N = 0, z0 = 0, z1 = 0; } binsReady;

    // LBVH
    uint32_t *keysA = nullptr, *keysB = nullptr, *valsA = nullptr, *valsB = nullptr;
    BvhNode* nodes = nullptr;
    Tri48* tris = nullptr;
    float4* pyramid = nullptr;          // leaf boxes + 16:1 summary levels
    uint32_t* refitScratch = nullptr; size_t refitCap = 0;   // parents + arrival flags (large meshes only)
    void* sortTemp = nullptr; size_t sortTempCap = 0;
    size_t capTris = 0;

    // small device scalars: [0..3] bound, [4..9] root box, then counters
    float* dBound = nullptr;
    float* dRootBox = nullptr;
    float* dPartials = nullptr;
    float* dCentres = nullptr;          // voxel-centre table of the current N (MODE_SHADER)
    uint32_t* dCounter = nullptr;
    uint32_t* dErr = nullptr;
    unsigned long long* dCrossings = nullptr;
    unsigned long long* dCount = nullptr;
    void* dSmall = nullptr;

    // grid
    uint32_t* gridOwned = nullptr; size_t gridCap = 0;
    uint32_t* gridTarget = nullptr; size_t gridTargetBytes = 0;
    uint32_t* texels = nullptr; size_t texCap = 0;
    uint8_t* u8Temp = nullptr; size_t u8Cap = 0;
    uint8_t* sparseBuf = nullptr; size_t sparseCap = 0;   // DXRV_FORMAT_SPARSE_BRICKS blob + block counts
    uint8_t* hostBlob = nullptr; size_t hostBlobCap = 0;  // pinned staging of the blob (dxrv_voxelize_to_host, sparse transport)
    uint32_t* hostRanks = nullptr; size_t hostRanksCap = 0;  // pinned: the encoder's per-block ranks (handed to the host pass)
    bool hostFillBegun = false;                           // dxrv_voxelize_mesh_to_host has started the host pool's pass already
    uint32_t readBack = 0;                                // DXRV_READ_BACK_* (dxrv_set_read_back)
    uint64_t lastD2hBytes = 0;                            // DXRV_INFO_LAST_D2H_BYTES
    uint32_t* mips = nullptr; size_t mipCap = 0;      // occupancy pyramid levels 1.. (concatenated)
    uint32_t mipLevels = 0;                            // levels incl. level 0; 0 = not built for the current grid
    uint32_t* walkBuf = nullptr; size_t walkCap = 0, walkZeroed = 0;  // MODE_PARITY candidate lists + split-tile scratch
    uint8_t* binsBuf = nullptr; size_t binsCap = 0;                   // MODE_SHADER direction bins (shader_bins.cu)
    ShaderBinsSizes binsSizes{};
    bool binsValid = false;                                            // built for the current acceleration structure
    uint32_t N = 0, z0 = 0, z1 = 0, mode = 0;
    bool haveGrid = false, haveTexels = false;

    // CUDA graphs: the kernel sequence of a build / voxelize call is captured once per distinct
    // parameter set and replayed afterwards (the per-launch gaps matter at 100 k triangles)
    struct GraphEntry { std::vector<uint8_t> key; cudaGraphExec_t exec = nullptr; uint64_t launches = 0, lastUse = 0; };
    std::vector<GraphEntry> graphs;
    uint64_t graphClock = 0;
    bool useGraphs = true;

    cudaEvent_t copyDone = nullptr;
    cudaStream_t copyStream = nullptr;                   // dxrv_voxelize_to_host: D2H of chunk k beside the fill of chunk k+1
    cudaEvent_t chunkDone[2] = {nullptr, nullptr};
    cudaEvent_t prof[3] = {nullptr, nullptr, nullptr};  // MODE_PARITY kernel timing (dxrv_set_profiling)
    cudaEvent_t profBuild[4] = {nullptr, nullptr, nullptr, nullptr};  // build start, keys ready, sorted, done
    bool profiling = false, profValid = false, profBuildValid = false;
    uint64_t launches = 0;

    // multi-GPU (comm.cu): NCCL communicator (loaded with dlopen), gathered full grid, peer access
    void* comm = nullptr;             // ncclComm_t
    int commRank = 0, commWorld = 1;
    uint32_t* gridFull = nullptr; size_t gridFullCap = 0;   // the whole N^3 grid: gather destination / peer-shared target
    uint32_t* dSlabs = nullptr;       // [2 * world] slab ranges of all ranks (device) for dxrv_gather_grid
    dxrv_ctx* gridOwner = nullptr;    // set by dxrv_share_grid_target: whose gridFull our slab lands in
    bool meshReplicated = false;      // ctx->mesh came from dxrv_bcast_mesh (context-owned device buffers)
    bool haveFull = false;            // gridFull holds a gathered / shared full grid of fullN^3 voxels
    uint32_t fullN = 0;
};


namespace dxrv
{
inline int fail(dxrv_ctx* c, int code, const std::string& msg)
{
    if (c) c->err = msg; else globalError() = msg;
    return code;
}

inline int cudaFail(dxrv_ctx* c, cudaError_t e, const char* what)
{
    char buf[256];
    std::snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
    cudaGetLastError();  // clear the non-sticky error state
    return fail(c, e == cudaErrorMemoryAllocation ? DXRV_ERR_OOM : DXRV_ERR_CUDA, buf);
}

#define DXRV_CUDA(call)                                                   \
    do {                                                                  \
        cudaError_t e_ = (call);                                          \
        if (e_ != cudaSuccess) return cudaFail(ctx, e_, #call);           \
    } while (0)

// NVTX range per API phase (the reference ships WinPixEventRuntime markers it never calls)
struct NvtxRange
{
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

struct DeviceGuard
{
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

template <typename T>
cudaError_t ensure(T*& p, size_t& cap, size_t need)
{
    if (need <= cap && p) return cudaSuccess;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    const size_t grow = need + need / 8;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p), grow ? grow : 16);
    if (e == cudaSuccess) cap = grow ? grow : 16;
    return e;
}

inline size_t slabWords(uint32_t N, uint32_t z0, uint32_t z1) { return (size_t)(z1 - z0) * N * ((N + 31) / 32); }

// api.cu: LBVH build from ctx->mesh (device pointers)
int buildContextMesh(dxrv_ctx* ctx, const float bound[4]);
// comm.cu: destroy the communicator of a context (no-op without one)
void commRelease(dxrv_ctx* ctx);
}  // namespace dxrv
