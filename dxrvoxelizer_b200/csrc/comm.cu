// comm.cu -- multi-GPU entry points of the C ABI (include/dxrv.h, "multi-GPU" section).
//
// The reference is single-GPU (XUSGRayTracing.h:386 SetNodeMask is never called); north_star asks for z-slab sharding
// over the 8 B200s of one box: mesh replicated by NCCL broadcast over NVLink, every GPU builds the identical LBVH,
// slabs gathered only when a full grid is requested.  Two deployment shapes are served:
//   * one process per GPU (torchrun / MPI style): dxrv_comm_get_unique_id on one rank, dxrv_comm_init on all;
//   * one process driving several GPUs (the C++ host class, DXRVoxelizer::SetGpuCount): dxrv_comm_init_all, and
//     collective calls of the contexts bracketed by dxrv_group_begin / dxrv_group_end (ncclGroupStart/End).
// The gather itself has two forms: dxrv_gather_grid (NCCL send/recv or broadcasts into the full grid), and the fused
// form, where the fill kernel's 128-bit stores land directly in the owner's full grid over NVLink
// (dxrv_share_grid_target for contexts of one process, dxrv_ipc_* + dxrv_set_grid_target across processes).
//
// NCCL is loaded with dlopen("libnccl.so.2") at the first dxrv_comm_* call, so libdxrv.so itself has no NCCL
// dependency: a single-GPU user never needs the library, and inside a PyTorch process the copy torch already
// mapped is reused (one NCCL per process).
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>

#include "ctx.h"

namespace
{
struct NcclApi
{
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

NcclApi& nccl()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names)
            if ((api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL)) != nullptr) break;
        if (!api.handle) { api.error = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
        auto sym = [&](const char* name) { void* p = dlsym(api.handle, name); if (!p && api.error.empty()) api.error = std::string("libnccl lacks ") + name; return p; };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
        api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
        api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    });
    return api;
}

int ncclFail(dxrv_ctx* ctx, ncclResult_t r, const char* what)
{
    NcclApi& n = nccl();
    return fail(ctx, DXRV_ERR_CUDA, std::string(what) + ": " + (n.GetErrorString ? n.GetErrorString(r) : "NCCL error"));
}

#define DXRV_NCCL(call)                                                   \
    do {                                                                  \
        ncclResult_t r_ = (call);                                         \
        if (r_ != ncclSuccess) return ncclFail(ctx, r_, #call);           \
    } while (0)

int needNccl(dxrv_ctx* ctx)
{
    NcclApi& n = nccl();
    if (!n.error.empty() || !n.handle) return fail(ctx, DXRV_ERR_UNSUPPORTED, n.error.empty() ? "NCCL not available" : n.error);
    return DXRV_OK;
}

int needComm(dxrv_ctx* ctx, const char* who)
{
    if (!ctx) return DXRV_ERR_INVALID_ARG;
    if (!ctx->comm) return fail(ctx, DXRV_ERR_INVALID_ARG, std::string(who) + ": call dxrv_comm_init first");
    return needNccl(ctx);
}
}  // namespace

namespace dxrv
{
void commRelease(dxrv_ctx* ctx)
{
    if (ctx && ctx->comm && nccl().CommDestroy) nccl().CommDestroy(static_cast<ncclComm_t>(ctx->comm));
    if (ctx) { ctx->comm = nullptr; ctx->commRank = 0; ctx->commWorld = 1; }
}
}  // namespace dxrv

extern "C" {

int dxrv_comm_get_unique_id(void* id128)
{
    if (!id128) return DXRV_ERR_INVALID_ARG;
    int rc = needNccl(nullptr);
    if (rc) return rc;
    static_assert(sizeof(ncclUniqueId) == DXRV_COMM_ID_BYTES, "unique id size");
    ncclUniqueId id;
    ncclResult_t r = nccl().GetUniqueId(&id);
    if (r != ncclSuccess) return ncclFail(nullptr, r, "ncclGetUniqueId");
    std::memcpy(id128, &id, sizeof(id));
    return DXRV_OK;
}

int dxrv_comm_init(dxrv_ctx* ctx, const void* id128, int rank, int world)
{
    if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_comm_init: bad arguments");
    int rc = needNccl(ctx);
    if (rc) return rc;
    DeviceGuard g(ctx->device);
    commRelease(ctx);
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    DXRV_NCCL(nccl().CommInitRank(&comm, world, id, rank));
    ctx->comm = comm; ctx->commRank = rank; ctx->commWorld = world;
    return DXRV_OK;
}

int dxrv_comm_init_all(dxrv_ctx** ctxs, int count)
{
    if (!ctxs || count < 1) return DXRV_ERR_INVALID_ARG;
    dxrv_ctx* ctx = ctxs[0];
    int rc = needNccl(ctx);
    if (rc) return rc;
    ncclUniqueId id;
    DXRV_NCCL(nccl().GetUniqueId(&id));
    std::vector<ncclComm_t> comms((size_t)count, nullptr);
    DXRV_NCCL(nccl().GroupStart());
    for (int i = 0; i < count; ++i)
    {
        if (!ctxs[i]) { nccl().GroupEnd(); return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_comm_init_all: null context"); }
        commRelease(ctxs[i]);
        cudaSetDevice(ctxs[i]->device);
        ncclResult_t r = nccl().CommInitRank(&comms[(size_t)i], count, id, i);
        if (r != ncclSuccess) { nccl().GroupEnd(); return ncclFail(ctx, r, "ncclCommInitRank"); }
    }
    DXRV_NCCL(nccl().GroupEnd());
    for (int i = 0; i < count; ++i) { ctxs[i]->comm = comms[(size_t)i]; ctxs[i]->commRank = i; ctxs[i]->commWorld = count; }
    return DXRV_OK;
}

int dxrv_comm_destroy(dxrv_ctx* ctx)
{
    if (!ctx) return DXRV_ERR_INVALID_ARG;
    DeviceGuard g(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    commRelease(ctx);
    return DXRV_OK;
}

int dxrv_group_begin(void)
{
    int rc = needNccl(nullptr);
    if (rc) return rc;
    return nccl().GroupStart() == ncclSuccess ? DXRV_OK : DXRV_ERR_CUDA;
}

int dxrv_group_end(void)
{
    int rc = needNccl(nullptr);
    if (rc) return rc;
    return nccl().GroupEnd() == ncclSuccess ? DXRV_OK : DXRV_ERR_CUDA;
}

int dxrv_bcast_u32(dxrv_ctx* ctx, uint32_t* values, uint32_t count, int root)
{
    int rc = needComm(ctx, "dxrv_bcast_u32");
    if (rc) return rc;
    if (!values || count == 0 || count > 64 || root < 0 || root >= ctx->commWorld) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_bcast_u32: bad arguments");
    DeviceGuard g(ctx->device);
    if (!ctx->dSlabs) DXRV_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->dSlabs), sizeof(uint32_t) * 2 * 1024));
    uint32_t* d = ctx->dSlabs + 1024;   // upper half: small messages
    if (ctx->commRank == root) DXRV_CUDA(cudaMemcpyAsync(d, values, sizeof(uint32_t) * count, cudaMemcpyHostToDevice, ctx->stream));
    DXRV_NCCL(nccl().Broadcast(d, d, count, ncclUint32, root, static_cast<ncclComm_t>(ctx->comm), ctx->stream));
    DXRV_CUDA(cudaMemcpyAsync(values, d, sizeof(uint32_t) * count, cudaMemcpyDeviceToHost, ctx->stream));
    DXRV_CUDA(cudaStreamSynchronize(ctx->stream));
    return DXRV_OK;
}

int dxrv_bcast_mesh(dxrv_ctx* ctx, const void* vertices, uint32_t numVerts, uint32_t strideBytes, const uint32_t* indices,
                    uint32_t numIndices, int root)
{
    int rc = needComm(ctx, "dxrv_bcast_mesh");
    if (rc) return rc;
    if (root < 0 || root >= ctx->commWorld) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_bcast_mesh: bad root");
    if (numVerts == 0 || strideBytes < 12 || (strideBytes & 3u) || numIndices % 3u)
        return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_bcast_mesh: numVerts / strideBytes / numIndices must be valid on EVERY rank (dxrv_bcast_u32 carries them)");
    NvtxRange range("dxrv broadcast mesh (NCCL)");
    const bool isRoot = ctx->commRank == root;
    if (isRoot && (!vertices || (numIndices && !indices))) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_bcast_mesh: the root needs the host arrays");
    DeviceGuard g(ctx->device);
    const size_t vBytes = (size_t)numVerts * strideBytes, iBytes = (size_t)numIndices * sizeof(uint32_t);
    cudaError_t e = ensure(ctx->vertsOwned, ctx->vertsCap, vBytes);
    if (e != cudaSuccess) return cudaFail(ctx, e, "cudaMalloc(vertices)");
    e = ensure(ctx->idxOwned, ctx->idxCap, iBytes);
    if (e != cudaSuccess) return cudaFail(ctx, e, "cudaMalloc(indices)");
    if (isRoot)
    {
        // upload heaps of Voxelizer::createVB / createIB (Voxelizer.cpp:115-138), on the root only
        DXRV_CUDA(cudaMemcpyAsync(ctx->vertsOwned, vertices, vBytes, cudaMemcpyHostToDevice, ctx->stream));
        if (iBytes) DXRV_CUDA(cudaMemcpyAsync(ctx->idxOwned, indices, iBytes, cudaMemcpyHostToDevice, ctx->stream));
        DXRV_CUDA(cudaEventRecord(ctx->copyDone, ctx->stream));
    }
    ncclComm_t comm = static_cast<ncclComm_t>(ctx->comm);
    DXRV_NCCL(nccl().Broadcast(ctx->vertsOwned, ctx->vertsOwned, vBytes, ncclUint8, root, comm, ctx->stream));
    if (iBytes) DXRV_NCCL(nccl().Broadcast(ctx->idxOwned, ctx->idxOwned, iBytes, ncclUint8, root, comm, ctx->stream));
    ctx->mesh = MeshView{ctx->vertsOwned, numVerts, strideBytes, ctx->idxOwned, numIndices / 3u};
    ctx->haveBvh = false;
    ctx->meshReplicated = true;
    if (isRoot)
    {
        // the host arrays are borrowed for the duration of the call only
        cudaError_t ce = cudaEventSynchronize(ctx->copyDone);
        if (ce != cudaSuccess) return cudaFail(ctx, ce, "upload");
    }
    return DXRV_OK;
}

int dxrv_build_bvh_replicated(dxrv_ctx* ctx, const float bound[4])
{
    if (!ctx) return DXRV_ERR_INVALID_ARG;
    if (!ctx->meshReplicated) return fail(ctx, DXRV_ERR_NO_BVH, "dxrv_build_bvh_replicated: call dxrv_bcast_mesh first");
    if (bound && !(bound[3] > 0.0f)) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_build_bvh_replicated: bound[3] (half extent) must be > 0");
    DeviceGuard g(ctx->device);
    return buildContextMesh(ctx, bound);
}

static int ensureFullGrid(dxrv_ctx* ctx, size_t bytes)
{
    if (ctx->gridFull && ctx->gridFullCap >= bytes) return DXRV_OK;
    if (ctx->gridFull) { cudaFree(ctx->gridFull); ctx->gridFull = nullptr; ctx->gridFullCap = 0; }
    DXRV_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->gridFull), bytes));
    ctx->gridFullCap = bytes;
    return DXRV_OK;
}

static int gatherWithTable(dxrv_ctx* ctx, int root, const uint32_t* all /* {z0, z1} per rank */)
{
    NvtxRange range("dxrv gather slabs (NCCL)");
    const int world = ctx->commWorld, me = ctx->commRank;
    ncclComm_t comm = static_cast<ncclComm_t>(ctx->comm);
    const uint32_t N = ctx->N;
    if (all[2 * me] != ctx->z0 || all[2 * me + 1] != ctx->z1) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_gather_grid: the slab table disagrees with this rank's last dxrv_voxelize");
    std::vector<char> covered(N, 0);
    for (int r = 0; r < world; ++r)
    {
        if (all[2 * r] > all[2 * r + 1] || all[2 * r + 1] > N) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_gather_grid: bad slab table");
        for (uint32_t z = all[2 * r]; z < all[2 * r + 1]; ++z)
        {
            if (covered[z]) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_gather_grid: overlapping slabs");
            covered[z] = 1;
        }
    }
    const bool receive = root < 0 || me == root;
    const size_t layerWords = (size_t)N * ((N + 31) / 32);
    if (receive)
    {
        const int rc = ensureFullGrid(ctx, layerWords * N * sizeof(uint32_t));
        if (rc) return rc;
        // layers nobody computed stay zero (a partial job: some ranks idle)
        bool holes = false;
        for (uint32_t z = 0; z < N; ++z) holes |= !covered[z];
        if (holes) DXRV_CUDA(cudaMemsetAsync(ctx->gridFull, 0, layerWords * N * sizeof(uint32_t), ctx->stream));
    }
    const uint32_t* slab = ctx->gridTarget ? ctx->gridTarget : ctx->gridOwned;
    const bool inPlace = receive && slab == ctx->gridFull + layerWords * ctx->z0;   // fused target: already there
    DXRV_NCCL(nccl().GroupStart());
    ncclResult_t r = ncclSuccess;
    for (int p = 0; p < world && r == ncclSuccess; ++p)
    {
        const size_t off = layerWords * all[2 * p], cnt = layerWords * (all[2 * p + 1] - all[2 * p]);
        if (cnt == 0) continue;   // empty slab: every rank skips it
        if (root < 0) r = nccl().Broadcast(p == me ? slab : ctx->gridFull + off, ctx->gridFull + off, cnt, ncclUint32, p, comm, ctx->stream);
        else if (p == me) { if (me != root) r = nccl().Send(slab, cnt, ncclUint32, root, comm, ctx->stream); }
        else if (me == root) r = nccl().Recv(ctx->gridFull + off, cnt, ncclUint32, p, comm, ctx->stream);
    }
    ncclResult_t re = nccl().GroupEnd();
    if (r != ncclSuccess) return ncclFail(ctx, r, "gather");
    if (re != ncclSuccess) return ncclFail(ctx, re, "ncclGroupEnd");
    if (root >= 0 && me == root && !inPlace && ctx->z1 > ctx->z0)
        DXRV_CUDA(cudaMemcpyAsync(ctx->gridFull + layerWords * ctx->z0, slab, layerWords * (ctx->z1 - ctx->z0) * sizeof(uint32_t),
                                  cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->haveFull = receive;
    ctx->fullN = N;
    return DXRV_OK;
}

int dxrv_gather_grid(dxrv_ctx* ctx, int root)
{
    int rc = needComm(ctx, "dxrv_gather_grid");
    if (rc) return rc;
    if (!ctx->haveGrid) return fail(ctx, DXRV_ERR_NO_GRID, "dxrv_gather_grid: call dxrv_voxelize first");
    const int world = ctx->commWorld, me = ctx->commRank;
    if (root >= world) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_gather_grid: bad root");
    if (world > 340) return fail(ctx, DXRV_ERR_UNSUPPORTED, "dxrv_gather_grid: too many ranks");
    DeviceGuard g(ctx->device);
    ncclComm_t comm = static_cast<ncclComm_t>(ctx->comm);
    if (!ctx->dSlabs) DXRV_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->dSlabs), sizeof(uint32_t) * 2 * 1024));
    // every rank learns every slab range {z0, z1, N}
    uint32_t mine[3] = {ctx->z0, ctx->z1, ctx->N};
    std::vector<uint32_t> all((size_t)3 * world), table((size_t)2 * world);
    DXRV_CUDA(cudaMemcpyAsync(ctx->dSlabs + 3 * me, mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->stream));
    DXRV_NCCL(nccl().AllGather(ctx->dSlabs + 3 * me, ctx->dSlabs, 3, ncclUint32, comm, ctx->stream));
    DXRV_CUDA(cudaMemcpyAsync(all.data(), ctx->dSlabs, sizeof(uint32_t) * 3 * world, cudaMemcpyDeviceToHost, ctx->stream));
    DXRV_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int r = 0; r < world; ++r)
    {
        if (all[3 * r + 2] != ctx->N) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_gather_grid: the ranks voxelized different grids");
        table[2 * r] = all[3 * r]; table[2 * r + 1] = all[3 * r + 1];
    }
    return gatherWithTable(ctx, root, table.data());
}

int dxrv_gather_grid_slabs(dxrv_ctx* ctx, int root, const uint32_t* slabs)
{
    int rc = needComm(ctx, "dxrv_gather_grid_slabs");
    if (rc) return rc;
    if (!slabs) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_gather_grid_slabs: null slab table");
    if (!ctx->haveGrid) return fail(ctx, DXRV_ERR_NO_GRID, "dxrv_gather_grid_slabs: call dxrv_voxelize first");
    if (root >= ctx->commWorld) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_gather_grid_slabs: bad root");
    DeviceGuard g(ctx->device);
    return gatherWithTable(ctx, root, slabs);
}

int dxrv_full_grid_device(dxrv_ctx* ctx, void** d_ptr, size_t* bytes)
{
    if (!ctx || !d_ptr || !bytes) return DXRV_ERR_INVALID_ARG;
    if (!ctx->haveFull || !ctx->gridFull) return fail(ctx, DXRV_ERR_NO_GRID, "dxrv_full_grid_device: no gathered grid on this rank");
    *d_ptr = ctx->gridFull;
    *bytes = (size_t)ctx->fullN * ctx->fullN * ((ctx->fullN + 31) / 32) * sizeof(uint32_t);
    return DXRV_OK;
}

int dxrv_fetch_full_grid(dxrv_ctx* ctx, void* hostDst, size_t bytes)
{
    if (!ctx || !hostDst) return DXRV_ERR_INVALID_ARG;
    if (!ctx->haveFull || !ctx->gridFull) return fail(ctx, DXRV_ERR_NO_GRID, "dxrv_fetch_full_grid: no gathered grid on this rank");
    const size_t need = (size_t)ctx->fullN * ctx->fullN * ((ctx->fullN + 31) / 32) * sizeof(uint32_t);
    if (bytes != need) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_fetch_full_grid: bytes must be N * N * ceil(N / 32) * 4");
    DeviceGuard g(ctx->device);
    DXRV_CUDA(cudaMemcpyAsync(hostDst, ctx->gridFull, need, cudaMemcpyDeviceToHost, ctx->stream));
    DXRV_CUDA(cudaStreamSynchronize(ctx->stream));
    return DXRV_OK;
}

int dxrv_share_grid_target(dxrv_ctx* ctx, dxrv_ctx* owner, uint32_t N, uint32_t slabBegin, uint32_t slabEnd)
{
    if (!ctx || !owner) return DXRV_ERR_INVALID_ARG;
    if (N == 0 || N > 8192 || slabBegin >= slabEnd || slabEnd > N) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_share_grid_target: bad slab");
    const size_t layerWords = (size_t)N * ((N + 31) / 32);
    if ((layerWords * slabBegin * sizeof(uint32_t)) & 15u)
        return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_share_grid_target: the slab offset must be 16-byte aligned (N * ceil(N/32) * slabBegin divisible by 4)");
    {
        DeviceGuard g(owner->device);
        if (!owner->gridFull || owner->gridFullCap < layerWords * N * sizeof(uint32_t)) cudaStreamSynchronize(owner->stream);   // about to reallocate
        const int rc = ensureFullGrid(owner, layerWords * N * sizeof(uint32_t));
        if (rc) return fail(ctx, rc, std::string("dxrv_share_grid_target: ") + dxrv_last_error(owner));
        owner->haveFull = true;
        owner->fullN = N;
    }
    if (ctx->device != owner->device)
    {
        DeviceGuard g(ctx->device);
        int can = 0;
        DXRV_CUDA(cudaDeviceCanAccessPeer(&can, ctx->device, owner->device));
        if (!can) return fail(ctx, DXRV_ERR_UNSUPPORTED, "dxrv_share_grid_target: no peer access between the two GPUs");
        cudaError_t e = cudaDeviceEnablePeerAccess(owner->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cudaFail(ctx, e, "cudaDeviceEnablePeerAccess");
        cudaGetLastError();
    }
    ctx->gridOwner = owner;
    return dxrv_set_grid_target(ctx, owner->gridFull + layerWords * slabBegin, layerWords * (slabEnd - slabBegin) * sizeof(uint32_t));
}

}  // extern "C"
