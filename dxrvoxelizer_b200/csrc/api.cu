// api.cu -- the C ABI of libdxrv.so (include/dxrv.h): context, device memory, launch sequencing.
//
// Host-side sequencing only; every byte of the hot path is produced by the kernels in lbvh.cu,
// onesweep.cu, trace_parity.cu and trace_shader.cu.  There is deliberately no CPU fallback.
#include <cuda_runtime.h>

#include <cstdio>
#include <chrono>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <utility>
#include <vector>
#include <algorithm>
#include <cstdlib>

#include "ctx.h"
#include "host_pool.h"
#include "sparse_host.h"

namespace
{
int checkDeviceError(dxrv_ctx* ctx)
{
    uint32_t e = 0;
    DXRV_CUDA(cudaMemcpyAsync(&e, ctx->dErr, sizeof(e), cudaMemcpyDeviceToHost, ctx->stream));
    DXRV_CUDA(cudaStreamSynchronize(ctx->stream));
    if (e == kErrNone) return DXRV_OK;
    cudaMemsetAsync(ctx->dErr, 0, sizeof(uint32_t), ctx->stream);
    ctx->walkZeroed = 0;  // a kernel that bailed out may have left the split-tile scratch dirty
    ctx->binsReady.valid = false;
    if (e == kErrBarrierTimeout)
    {
        // the fused build's CTAs were not co-resident after all: clean its barrier words, use the multi-kernel build from now on
        if (ctx->fusedScratch) cudaMemsetAsync(ctx->fusedScratch, 0, 64, ctx->stream);
        ctx->fusedBuild = false;
        ctx->haveBvh = false;
        return fail(ctx, DXRV_ERR_CUDA, "fused build: grid barrier timed out (fused build disabled for this context; build again)");
    }
    if (e == kErrBadIndex) return fail(ctx, DXRV_ERR_INVALID_ARG, "index buffer references a vertex >= numVerts");
    return fail(ctx, DXRV_ERR_CUDA, "traversal stack overflow / corrupt hierarchy");
}

// Run `enqueue` (which only launches kernels / memsets on ctx->stream) through a cached CUDA graph.
template <class F>
int runCaptured(dxrv_ctx* ctx, const std::vector<uint8_t>& key, F&& enqueue)
{
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (!ctx->useGraphs || ctx->profiling || cudaStreamIsCapturing(ctx->stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone)
    {
        cudaGetLastError();
        enqueue();
        return DXRV_OK;
    }
    ++ctx->graphClock;
    for (auto& g : ctx->graphs)
        if (g.key == key)
        {
            g.lastUse = ctx->graphClock;
            DXRV_CUDA(cudaGraphLaunch(g.exec, ctx->stream));
            ctx->launches += g.launches;
            return DXRV_OK;
        }
    const uint64_t before = ctx->launches;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    bool ok = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok)
    {
        enqueue();
        ok = cudaStreamEndCapture(ctx->stream, &graph) == cudaSuccess && graph != nullptr;
    }
    // A graph of the same shape with other parameters (the next mesh of a batch: same kernels, other pointers)
    // is patched into an executable graph we already have -- instantiating costs far more than the kernels
    // of a small mesh take.  Candidates: same kind of call (first key word), same number of launches.
    if (ok)
    {
        const uint64_t launchesNow = ctx->launches - before;
        // ... but only once a dozen graphs of that kind exist: a call sequence that CYCLES through a few parameter sets
        // (the z sub-slabs of dxrv_voxelize_to_host, the slabs of a multi-GPU host) gets one executable graph each
        // and replays them, instead of patching one graph back and forth (capture + update cost ~0.1 ms of CPU time).
        dxrv_ctx::GraphEntry* best = nullptr;
        size_t sameKind = 0;
        for (auto& g : ctx->graphs)
            if (g.launches == launchesNow && g.key.size() >= 4 && key.size() >= 4 && std::equal(key.begin(), key.begin() + 4, g.key.begin()))
            {
                ++sameKind;
                if (!best || g.lastUse < best->lastUse) best = &g;   // least recently used of its kind
            }
        if (sameKind < 12) best = nullptr;
        if (best)
        {
            cudaGraphExecUpdateResultInfo info;
            if (cudaGraphExecUpdate(best->exec, graph, &info) == cudaSuccess)
            {
                cudaGraphDestroy(graph);
                best->key = key; best->lastUse = ctx->graphClock;
                DXRV_CUDA(cudaGraphLaunch(best->exec, ctx->stream));
                return DXRV_OK;
            }
            cudaGetLastError();
            // a failed update leaves that executable graph in an unspecified state: drop it
            cudaGraphExecDestroy(best->exec);
            ctx->graphs.erase(ctx->graphs.begin() + (best - ctx->graphs.data()));
        }
    }
    if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
    if (graph) cudaGraphDestroy(graph);
    if (!ok)
    {
        // capture is an optimisation only: fall back to plain stream launches
        cudaGetLastError();
        ctx->launches = before;
        ctx->useGraphs = false;
        enqueue();
        return DXRV_OK;
    }
    if (ctx->graphs.size() >= 32)
    {
        size_t victim = 0;
        for (size_t i = 1; i < ctx->graphs.size(); ++i) if (ctx->graphs[i].lastUse < ctx->graphs[victim].lastUse) victim = i;
        cudaGraphExecDestroy(ctx->graphs[victim].exec);
        ctx->graphs.erase(ctx->graphs.begin() + victim);
    }
    dxrv_ctx::GraphEntry e;
    e.key = key; e.exec = exec; e.launches = ctx->launches - before; e.lastUse = ctx->graphClock;
    ctx->graphs.push_back(e);
    DXRV_CUDA(cudaGraphLaunch(exec, ctx->stream));
    return DXRV_OK;
}

template <class T>
void keyPush(std::vector<uint8_t>& k, const T& v)
{
    const uint8_t* p = reinterpret_cast<const uint8_t*>(&v);
    k.insert(k.end(), p, p + sizeof(T));
}

int buildOnDevice(dxrv_ctx* ctx, const float bound[4])
{
    NvtxRange range("dxrv build (bounds, Morton, onesweep, leaves[, hierarchy])");
    const MeshView& m = ctx->mesh;
    const uint32_t T = m.numTris;
    ctx->haveBvh = false;
    ctx->haveGrid = false;
    ctx->binsValid = false;
    ctx->binsReady.valid = false;
    if (T > ctx->capTris || !ctx->nodes)
    {
        const size_t cap = T + T / 8 + 16;
        uint32_t** u32s[] = {&ctx->keysA, &ctx->keysB, &ctx->valsA, &ctx->valsB};
        for (auto pp : u32s) { if (*pp) cudaFree(*pp); *pp = nullptr; }
        if (ctx->nodes) cudaFree(ctx->nodes); ctx->nodes = nullptr;
        if (ctx->tris) cudaFree(ctx->tris); ctx->tris = nullptr;
        if (ctx->pyramid) cudaFree(ctx->pyramid); ctx->pyramid = nullptr;
        ctx->capTris = 0;
        for (auto pp : u32s) DXRV_CUDA(cudaMalloc(reinterpret_cast<void**>(pp), sizeof(uint32_t) * cap));
        DXRV_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->nodes), sizeof(BvhNode) * cap));
        DXRV_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->tris), sizeof(Tri48) * cap));
        DXRV_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->pyramid), sizeof(float4) * boxPyramidFloat4s((uint32_t)cap)));
        ctx->capTris = cap;
    }
    if (useAtomicRefit(T))
    {
        cudaError_t e = ensure(ctx->refitScratch, ctx->refitCap, sizeof(uint32_t) * 3 * (size_t)T);
        if (e != cudaSuccess) return cudaFail(ctx, e, "cudaMalloc(refit scratch)");
    }
    {
        size_t need = SortTemp::bytesFor(T);
        cudaError_t e = ensure(ctx->sortTemp, ctx->sortTempCap, need);
        if (e != cudaSuccess) return cudaFail(ctx, e, "cudaMalloc(sort temp)");
    }

    cudaStream_t s = ctx->stream;
    // Small meshes: the whole build in one cooperative kernel (lbvh.cu, k_build_fused); DXRV_FUSED_BUILD=0 keeps the
    // multi-kernel path (tests run both and require identical structures).
    // Measured inside a build + voxelize step (tools/build_time.py): 100 k triangles 39.5 -> 33 us, 200 k 41 -> 34 us; at
    // 20 k both take ~27 us (four grid barriers cost what five kernel boundaries do), so small meshes keep the kernels
    // that can overlap across streams.  DXRV_FUSED_BUILD=1 fuses whatever fits.
    const char* fusedEnv = std::getenv("DXRV_FUSED_BUILD");
    const bool fusedOff = fusedEnv && fusedEnv[0] == '0', fusedAll = fusedEnv && fusedEnv[0] == '1';
    uint32_t fusedCtas = 0, fusedRounds = 0;
    bool fused = ctx->fusedBuild && !fusedOff && (fusedAll || T >= kFusedBuildMinTris) && fusedBuildPlan(T, ctx->smCount, fusedCtas, fusedRounds);
    if (fused && !ctx->fusedScratch)
    {
        if (!fusedBuildSupported(ctx->device) || cudaMalloc(&ctx->fusedScratch, fusedBuildScratchBytes(ctx->smCount)) != cudaSuccess)
        {
            cudaGetLastError();
            ctx->fusedBuild = fused = false;
        }
        else DXRV_CUDA(cudaMemsetAsync(ctx->fusedScratch, 0, fusedBuildScratchBytes(ctx->smCount), s));
    }
    // small meshes: 8 bits per axis (24-bit keys, three radix passes); large: all 30 bits
    // Key width.  With the hierarchy (a consumer traverses it): 8 bits per axis for small meshes (three radix passes),
    // all 30 bits for large ones -- the tree's quality depends on it.  Without (MODE_PARITY: the sorted order only
    // gives neighbouring records neighbouring slots; triangles of one cell keep their mesh order, the sort is
    // stable): 16-bit keys, two passes, whatever the size (the top bits of the 30-bit code: 5-6 bits per axis).
    // Measured with RANDOMLY ORDERED input (tools/shuffle_test.py: 5.2 M triangles, 512^3 / 1024^3): the consumers
    // take 0.111 / 0.207 ms after two passes, 0.110 / 0.201 ms after four.  A hierarchy built later over such keys
    // is valid (ties are broken by index) but coarser; the next build widens the keys again.
    const bool withTree = ctx->treeWanted;
    const bool small = T <= (1u << 18);
    uint32_t keyShift = withTree ? (small ? 6u : 0u) : 14u;
    int numPasses = withTree ? (small ? 3 : 4) : 2;
    if (const char* e = std::getenv("DXRV_KEY_PASSES"))   // experiment switch: 2, 3 or 4 radix passes
    {
        const int p = std::atoi(e);
        if (p >= 2 && p <= 4) { numPasses = p; keyShift = 30u - 8u * (uint32_t)p + (p == 4 ? 2u : 0u); }
    }
    // with an odd number of passes start in the B buffers, so that the sorted result is always in A
    // (a single triangle is not sorted at all: it stays where the Morton kernel wrote it)
    const bool startInB = (numPasses & 1) && T >= 2;
    uint32_t* k0 = startInB ? ctx->keysB : ctx->keysA;
    uint32_t* v0 = startInB ? ctx->valsB : ctx->valsA;
    uint32_t* k1 = startInB ? ctx->keysA : ctx->keysB;
    uint32_t* v1 = startInB ? ctx->valsA : ctx->valsB;
    const float bnd[4] = {bound ? bound[0] : 0.0f, bound ? bound[1] : 0.0f, bound ? bound[2] : 0.0f, bound ? bound[3] : 0.0f};
    const bool haveBound = bound != nullptr;

    std::vector<uint8_t> key;
    keyPush(key, (uint32_t)0xB01Du);
    keyPush(key, m.verts); keyPush(key, m.numVerts); keyPush(key, m.stride); keyPush(key, m.indices); keyPush(key, m.numTris);
    keyPush(key, (uint32_t)haveBound); keyPush(key, bnd); keyPush(key, (uint32_t)withTree); keyPush(key, (uint32_t)fused);
    keyPush(key, ctx->keysA); keyPush(key, ctx->keysB); keyPush(key, ctx->valsA); keyPush(key, ctx->valsB);
    keyPush(key, ctx->nodes); keyPush(key, ctx->tris); keyPush(key, ctx->pyramid); keyPush(key, ctx->sortTemp); keyPush(key, ctx->refitScratch);
    const int rc = runCaptured(ctx, key, [&]() {
        const bool prof = ctx->profiling;   // (runCaptured launches directly, without a graph, while profiling)
        if (prof) cudaEventRecord(ctx->profBuild[0], s);
        if (fused && launchFusedBuild(s, m, ctx->smCount, haveBound ? bnd : nullptr, ctx->dBound, ctx->dPartials, ctx->fusedScratch, ctx->keysA, ctx->valsA,
                                      ctx->keysB, ctx->valsB, withTree ? nullptr : ctx->tris, keyShift, numPasses, ctx->dErr))
        {
            ctx->launches += 1;
            if (prof) { cudaEventRecord(ctx->profBuild[1], s); cudaEventRecord(ctx->profBuild[2], s); }   // (no separate sort phase)
            if (withTree)
                ctx->launches += (uint64_t)launchLeavesAndHierarchy(s, &ctx->side, m, ctx->dBound, ctx->keysA, ctx->valsA, ctx->nodes, ctx->tris,
                                                                    ctx->pyramid, ctx->refitScratch, ctx->dRootBox, ctx->dErr, kBuildLeaves | kBuildTree);
            if (prof) { cudaEventRecord(ctx->profBuild[3], s); ctx->profBuildValid = true; }
            return;
        }
        if (fused) ctx->fusedBuild = false;   // the cooperative launch was refused: multi-kernel builds from now on
        if (haveBound) { launchSetBound(s, bnd[0], bnd[1], bnd[2], bnd[3], ctx->dBound); }
        else { launchBounds(s, m, ctx->dBound, ctx->dPartials, ctx->dCounter); }
        ctx->launches += 1;
        if (T > 0)
        {
            uint32_t* hist = sortClearTemp(s, ctx->sortTemp, T);
            launchMorton(s, m, ctx->dBound, k0, v0, keyShift, numPasses, hist, ctx->dErr);
            ctx->launches += 1;
            if (prof) cudaEventRecord(ctx->profBuild[1], s);
            ctx->launches += (uint64_t)radixSortPairs(s, ctx->sortTemp, k0, v0, k1, v1, T, numPasses, true, nullptr);
            if (prof) cudaEventRecord(ctx->profBuild[2], s);
            ctx->launches += (uint64_t)launchLeavesAndHierarchy(s, &ctx->side, m, ctx->dBound, ctx->keysA, ctx->valsA, ctx->nodes, ctx->tris,
                                                                ctx->pyramid, ctx->refitScratch, ctx->dRootBox, ctx->dErr,
                                                                withTree ? (kBuildLeaves | kBuildTree) : kBuildLeaves);
        }
        if (prof) { cudaEventRecord(ctx->profBuild[3], s); ctx->profBuildValid = T > 0; }
    });
    if (rc) return rc;
    DXRV_CUDA(cudaGetLastError());
    ctx->haveBvh = true;
    ctx->treeBuilt = withTree || T < 2;
    ctx->pyramidBuilt = withTree || !fused;   // the fused kernel writes the sorted records only; a later hierarchy redoes the leaves
    return DXRV_OK;
}

// Enqueue the hierarchy of the current build if no consumer has needed it yet (called inside a voxelize's launch
// sequence, so it is captured into the same CUDA graph).
void enqueueTree(dxrv_ctx* ctx)
{
    if (ctx->treeBuilt) return;
    ctx->launches += (uint64_t)launchLeavesAndHierarchy(ctx->stream, &ctx->side, ctx->mesh, ctx->dBound, ctx->keysA, ctx->valsA, ctx->nodes, ctx->tris,
                                                        ctx->pyramid, ctx->refitScratch, ctx->dRootBox, ctx->dErr,
                                                        ctx->pyramidBuilt ? kBuildTree : (kBuildLeaves | kBuildTree));
}

int validateMeshArgs(dxrv_ctx* ctx, const void* v, uint32_t numVerts, uint32_t stride, const uint32_t* idx, uint32_t numIndices,
                     const float bound[4])
{
    if (!ctx) return DXRV_ERR_INVALID_ARG;
    if (!v || numVerts == 0) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_build_bvh: no vertices");
    if (stride < 12 || (stride & 3u)) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_build_bvh: strideBytes must be >= 12 and a multiple of 4");
    if (numIndices % 3u) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_build_bvh: numIndices must be a multiple of 3");
    if (numIndices && !idx) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_build_bvh: null index buffer");
    if (bound && !(bound[3] > 0.0f)) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_build_bvh: bound[3] (half extent) must be > 0");
    return DXRV_OK;
}
// dxrv_voxelize_to_host through the compact transport: voxelize the slab, encode it on the device (sparse.cu), bring the
// blob back into pinned staging memory and expand it into hostDst -- whose zeroing by the host pool starts before the
// GPU does.  DXRV_ERR_UNSUPPORTED: the grid did not compress (nothing copied; the slab is resident in ctx->gridOwned).
bool hostFillTwoPass()
{
    static const bool twoPass = [] { const char* e = std::getenv("DXRV_HOST_FILL"); return e && !std::strcmp(e, "twopass"); }();
    return twoPass;
}

// transport of dxrv_voxelize_to_host for a slab of `bytes`: 1 dense copy, 2 compact blob + host pass
uint32_t toHostTransport(const dxrv_ctx* ctx, size_t bytes)
{
    uint32_t transport = ctx->readBack;
    if (const char* e = std::getenv("DXRV_TO_HOST")) transport = !std::strcmp(e, "dense") ? 1u : (!std::strcmp(e, "sparse") ? 2u : transport);
    // (a host thread writes ~12 GB/s, the link copies ~55 GB/s when one GPU has it to itself and 16-28 GB/s when the
    // eight GPUs of a box read back together.  Measured with the ranks of an 8-GPU box sharing 32 cores, 4 threads each:
    // one-pass host grid 0.94 ms against 1.32-1.48 ms for the dense copy -- the old two-pass expansion was 1.41 ms, hence
    // the threshold of eight threads it had; below four threads nothing has been measured and the copy stays)
    if (transport == 0u) transport = (bytes >= (8u << 20) && hostPoolThreads() >= 4u) ? 2u : 1u;
    return transport;
}

int toHostArgsCheck(dxrv_ctx* ctx, uint32_t N, uint32_t mode, uint32_t slabBegin, uint32_t slabEnd, const void* hostDst, size_t bytes)
{
    if (!ctx || !hostDst) return DXRV_ERR_INVALID_ARG;
    if (slabBegin >= slabEnd || slabEnd > N || N == 0 || N > 8192) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_voxelize_to_host: need 0 <= slabBegin < slabEnd <= N <= 8192");
    if (mode & DXRV_EMIT_TEXELS) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_voxelize_to_host: bit grid only");
    const size_t layerBytes = (size_t)N * ((N + 31) / 32) * sizeof(uint32_t);
    if (bytes != layerBytes * (slabEnd - slabBegin)) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_voxelize_to_host: bytes does not match the slab size");
    if (ctx->gridTarget) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_voxelize_to_host: not with an external grid target");
    return DXRV_OK;
}

int voxelizeToHostSparse(dxrv_ctx* ctx, uint32_t N, uint32_t mode, uint32_t slabBegin, uint32_t slabEnd, void* hostDst, size_t bytes)
{
    NvtxRange range("dxrv voxelize to host (sparse transport + host expansion)");
    const SparseLayout L = sparseLayout(N, slabEnd - slabBegin);
    const size_t countsOff = (L.maxBytes + 255) & ~(size_t)255;
    cudaError_t e = ensure(ctx->sparseBuf, ctx->sparseCap, countsOff + (size_t)L.numBlocks * sizeof(uint32_t));
    if (e != cudaSuccess) return cudaFail(ctx, e, "cudaMalloc(sparse bricks)");
    auto ensureStaging = [&](size_t need, size_t keep) -> cudaError_t {   // keep: leading bytes that survive a re-allocation
        if (need <= ctx->hostBlobCap && ctx->hostBlob) return cudaSuccess;
        uint8_t* fresh = nullptr;
        const size_t cap = need + need / 2;
        const cudaError_t err = cudaHostAlloc(reinterpret_cast<void**>(&fresh), cap, cudaHostAllocDefault);
        if (err != cudaSuccess) return err;
        if (ctx->hostBlob) { std::memcpy(fresh, ctx->hostBlob, std::min(keep, ctx->hostBlobCap)); cudaFreeHost(ctx->hostBlob); }
        ctx->hostBlob = fresh; ctx->hostBlobCap = cap;
        return cudaSuccess;
    };
    if ((e = ensureStaging(L.offPayload + (1u << 20), 0)) != cudaSuccess) return cudaFail(ctx, e, "cudaHostAlloc(blob staging)");
    if (ctx->hostRanksCap < L.numBlocks)
    {
        if (ctx->hostRanks) cudaFreeHost(ctx->hostRanks);
        ctx->hostRanks = nullptr; ctx->hostRanksCap = 0;
        if (cudaHostAlloc(reinterpret_cast<void**>(&ctx->hostRanks), (size_t)L.numBlocks * sizeof(uint32_t), cudaHostAllocDefault) == cudaSuccess) ctx->hostRanksCap = L.numBlocks;
        else { ctx->hostRanks = nullptr; cudaGetLastError(); }   // (not fatal: the host pass counts the states itself)
    }
    bool haveRanks = false;
    // DXRV_DBG_E2E=1: the phases of this call on stderr (development aid)
    static const bool dbgPhases = [] { const char* e = std::getenv("DXRV_DBG_E2E"); return e && e[0] == '1'; }();
    auto nowUs = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double tBegin = dbgPhases ? nowUs() : 0.0;
    // The host pool starts on the caller's grid at once (zeroing it from the outside of the slab inwards) and is handed
    // the blob as soon as it is here: the brick layers still to do are then written with their final contents in the
    // same pass (sparse_host.cpp).  DXRV_HOST_FILL=twopass: zero everything, then expand (the earlier scheme).
    const bool twoPass = hostFillTwoPass();
    if (twoPass) hostZeroBegin(hostDst, bytes);   // runs beside everything up to hostZeroWait()
    else if (ctx->hostFillBegun) ctx->hostFillBegun = false;   // (dxrv_voxelize_mesh_to_host started the pass before the upload)
    else hostFillBegin(hostDst, N, slabBegin, slabEnd - slabBegin);
    int rc = dxrv_voxelize(ctx, N, mode, slabBegin, slabEnd);
    if (rc == DXRV_OK)
    {
        ctx->launches += (uint64_t)launchSparseEncode(ctx->stream, ctx->gridOwned, N, slabBegin, slabEnd, ctx->sparseBuf,
                                                      reinterpret_cast<uint32_t*>(ctx->sparseBuf + countsOff));
        // the encoder's per-block ranks (32 KB for 1024^3: the host pass then has nothing to count), the header (number of
        // mixed bricks) and the brick states in one go; then exactly the payload that exists
        haveRanks = ctx->hostRanks != nullptr &&
                    cudaMemcpyAsync(ctx->hostRanks, ctx->sparseBuf + countsOff, (size_t)L.numBlocks * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess;
        if (cudaMemcpyAsync(ctx->hostBlob, ctx->sparseBuf, L.offPayload, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess)
            rc = cudaFail(ctx, cudaGetLastError(), "sparse transport: header copy");
    }
    size_t total = 0;
    if (rc == DXRV_OK)
    {
        const uint32_t numMixed = reinterpret_cast<const uint32_t*>(ctx->hostBlob)[9];
        total = L.offPayload + (size_t)numMixed * 64;
        if (total > bytes / 2) rc = DXRV_ERR_UNSUPPORTED;
        else if ((e = ensureStaging(total, L.offPayload)) != cudaSuccess) rc = cudaFail(ctx, e, "cudaHostAlloc(blob staging)");
        else if (numMixed && (cudaMemcpyAsync(ctx->hostBlob + L.offPayload, ctx->sparseBuf + L.offPayload, total - L.offPayload, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
                              cudaStreamSynchronize(ctx->stream) != cudaSuccess))
            rc = cudaFail(ctx, cudaGetLastError(), "sparse transport: payload copy");
    }
    const double tGpu = dbgPhases ? nowUs() : 0.0;
    SparseBlobView v;
    bool good = rc == DXRV_OK && sparseParse(ctx->hostBlob, total, v);
    double tZero = 0.0;
    if (twoPass)
    {
        hostZeroWait();
        tZero = dbgPhases ? nowUs() : 0.0;
        if (rc != DXRV_OK) return rc;
        good = good && sparseExpand(v, static_cast<uint32_t*>(hostDst), true);
    }
    else
    {
        good = good && hostFillPublish(v, haveRanks ? ctx->hostRanks : nullptr, kSparseBlockBricks);
        tZero = dbgPhases ? nowUs() : 0.0;
        const bool filled = hostFillWait();      // (always: the pool must be done with hostDst before this call returns)
        if (rc != DXRV_OK) return rc;
        good = good && filled;
    }
    if (!good) return fail(ctx, DXRV_ERR_CUDA, "sparse transport: inconsistent blob");
    if (dbgPhases)
        std::fprintf(stderr, "dxrv e2e phases (us): voxelize + encode + blob copy %.0f | %s %.0f | %s %.0f | call %.0f\n", tGpu - tBegin,
                     twoPass ? "rest of the zeroing" : "publish", tZero - tGpu, twoPass ? "expansion" : "rest of the pass", nowUs() - tZero, nowUs() - tBegin);
    ctx->lastD2hBytes = total;
    return checkDeviceError(ctx);
}
}  // namespace

namespace dxrv
{
int buildContextMesh(dxrv_ctx* ctx, const float bound[4]) { return buildOnDevice(ctx, bound); }
}  // namespace dxrv

extern "C" {

int dxrv_create(dxrv_ctx** out, int cuda_device)
{
    if (!out) return fail(nullptr, DXRV_ERR_INVALID_ARG, "dxrv_create: null out");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
    {
        cudaGetLastError();
        return fail(nullptr, DXRV_ERR_CUDA, std::string("dxrv_create: no CUDA device (") + cudaGetErrorString(e) +
                                                "); this library has no CPU fallback");
    }
    if (cuda_device < 0 || cuda_device >= count) return fail(nullptr, DXRV_ERR_INVALID_ARG, "dxrv_create: bad device ordinal");
    dxrv_ctx* ctx = new (std::nothrow) dxrv_ctx();
    if (!ctx) return fail(nullptr, DXRV_ERR_OOM, "dxrv_create: out of memory");
    ctx->device = cuda_device;
    DeviceGuard g(cuda_device);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, cuda_device)) != cudaSuccess) { delete ctx; return cudaFail(nullptr, e, "cudaGetDeviceProperties"); }
    if (prop.major < 10)
    {
        delete ctx;
        return fail(nullptr, DXRV_ERR_UNSUPPORTED, "dxrv_create: kernels are built for sm_100a (Blackwell) only");
    }
    ctx->smCount = prop.multiProcessorCount;
    if (const char* ng = std::getenv("DXRV_NO_GRAPHS")) ctx->useGraphs = !(ng[0] && ng[0] != '0');
    if ((e = cudaStreamCreateWithFlags(&ctx->ownStream, cudaStreamNonBlocking)) != cudaSuccess) { delete ctx; return cudaFail(nullptr, e, "cudaStreamCreate"); }
    ctx->stream = ctx->ownStream;
    if (cudaStreamCreateWithFlags(&ctx->side.stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->side.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->side.join, cudaEventDisableTiming) != cudaSuccess)
    {
        cudaGetLastError();
        ctx->side = SideStream{};   // no side stream: builds run serially
    }
    if ((e = cudaEventCreateWithFlags(&ctx->copyDone, cudaEventDisableTiming)) != cudaSuccess) { cudaStreamDestroy(ctx->ownStream); delete ctx; return cudaFail(nullptr, e, "cudaEventCreate"); }
    const size_t smallBytes = 64 * sizeof(float) + 6 * kBoundsMaxBlocks * sizeof(float) + 8192 * sizeof(float);
    if ((e = cudaMalloc(&ctx->dSmall, smallBytes)) != cudaSuccess) { cudaStreamDestroy(ctx->ownStream); delete ctx; return cudaFail(nullptr, e, "cudaMalloc"); }
    cudaMemset(ctx->dSmall, 0, smallBytes);
    float* f = static_cast<float*>(ctx->dSmall);
    ctx->dBound = f;                                                   // 4 floats
    ctx->dRootBox = f + 4;                                             // 6 floats
    ctx->dCounter = reinterpret_cast<uint32_t*>(f + 12);
    ctx->dErr = reinterpret_cast<uint32_t*>(f + 13);
    ctx->dCrossings = reinterpret_cast<unsigned long long*>(f + 16);   // 8-byte aligned
    ctx->dCount = reinterpret_cast<unsigned long long*>(f + 18);
    ctx->dPartials = f + 64;
    ctx->dCentres = f + 64 + 6 * kBoundsMaxBlocks;
    *out = ctx;
    return DXRV_OK;
}

void dxrv_destroy(dxrv_ctx* ctx)
{
    if (!ctx) return;
    DeviceGuard g(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    commRelease(ctx);
    void* ptrs[] = {ctx->gridFull, ctx->dSlabs, ctx->vertsOwned, ctx->idxOwned, ctx->keysA, ctx->keysB, ctx->valsA, ctx->valsB, ctx->nodes, ctx->tris,
                    ctx->pyramid, ctx->refitScratch, ctx->sortTemp, ctx->dSmall, ctx->gridOwned, ctx->texels, ctx->u8Temp, ctx->walkBuf, ctx->mips, ctx->sparseBuf,
                    ctx->binsBuf, ctx->fusedScratch};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (auto& g : ctx->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    if (ctx->hostBlob) cudaFreeHost(ctx->hostBlob);
    if (ctx->hostRanks) cudaFreeHost(ctx->hostRanks);
    if (ctx->copyDone) cudaEventDestroy(ctx->copyDone);
    for (cudaEvent_t e : ctx->chunkDone) if (e) cudaEventDestroy(e);
    if (ctx->copyStream) cudaStreamDestroy(ctx->copyStream);
    for (cudaEvent_t e : ctx->prof) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->profBuild) if (e) cudaEventDestroy(e);
    if (ctx->side.fork) cudaEventDestroy(ctx->side.fork);
    if (ctx->side.join) cudaEventDestroy(ctx->side.join);
    if (ctx->side.stream) cudaStreamDestroy(ctx->side.stream);
    if (ctx->ownStream) cudaStreamDestroy(ctx->ownStream);
    cudaGetLastError();
    delete ctx;
}

const char* dxrv_last_error(const dxrv_ctx* ctx) { return ctx ? ctx->err.c_str() : globalError().c_str(); }

int dxrv_set_stream(dxrv_ctx* ctx, void* cuda_stream)
{
    if (!ctx) return DXRV_ERR_INVALID_ARG;
    DeviceGuard g(ctx->device);
    DXRV_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->ownStream;
    return DXRV_OK;
}

int dxrv_set_profiling(dxrv_ctx* ctx, int enable)
{
    if (!ctx) return DXRV_ERR_INVALID_ARG;
    DeviceGuard g(ctx->device);
    if (enable)
    {
        for (cudaEvent_t& e : ctx->prof)
            if (!e) DXRV_CUDA(cudaEventCreate(&e));
        for (cudaEvent_t& e : ctx->profBuild)
            if (!e) DXRV_CUDA(cudaEventCreate(&e));
    }
    ctx->profiling = enable != 0;
    ctx->profValid = false;
    ctx->profBuildValid = false;
    return DXRV_OK;
}

int dxrv_synchronize(dxrv_ctx* ctx)
{
    if (!ctx) return DXRV_ERR_INVALID_ARG;
    DeviceGuard g(ctx->device);
    DXRV_CUDA(cudaStreamSynchronize(ctx->stream));
    return checkDeviceError(ctx);
}

int dxrv_build_bvh(dxrv_ctx* ctx, const void* vertices, uint32_t numVerts, uint32_t strideBytes, const uint32_t* indices,
                   uint32_t numIndices, const float bound[4])
{
    int rc = validateMeshArgs(ctx, vertices, numVerts, strideBytes, indices, numIndices, bound);
    if (rc) return rc;
    DeviceGuard g(ctx->device);
    const size_t vBytes = (size_t)numVerts * strideBytes, iBytes = (size_t)numIndices * sizeof(uint32_t);
    cudaError_t e = ensure(ctx->vertsOwned, ctx->vertsCap, vBytes);
    if (e != cudaSuccess) return cudaFail(ctx, e, "cudaMalloc(vertices)");
    e = ensure(ctx->idxOwned, ctx->idxCap, iBytes);
    if (e != cudaSuccess) return cudaFail(ctx, e, "cudaMalloc(indices)");
    // upload heaps of Voxelizer::createVB / createIB (Voxelizer.cpp:115-138)
    DXRV_CUDA(cudaMemcpyAsync(ctx->vertsOwned, vertices, vBytes, cudaMemcpyHostToDevice, ctx->stream));
    if (iBytes) DXRV_CUDA(cudaMemcpyAsync(ctx->idxOwned, indices, iBytes, cudaMemcpyHostToDevice, ctx->stream));
    DXRV_CUDA(cudaEventRecord(ctx->copyDone, ctx->stream));
    ctx->mesh = MeshView{ctx->vertsOwned, numVerts, strideBytes, ctx->idxOwned, numIndices / 3u};
    rc = buildOnDevice(ctx, bound);
    // the host arrays are only borrowed for the duration of this call: wait for the uploads (not the build)
    cudaError_t ce = cudaEventSynchronize(ctx->copyDone);
    if (ce != cudaSuccess) return cudaFail(ctx, ce, "upload");
    return rc;
}

int dxrv_build_bvh_device(dxrv_ctx* ctx, const void* d_vertices, uint32_t numVerts, uint32_t strideBytes,
                          const uint32_t* d_indices, uint32_t numIndices, const float bound[4])
{
    int rc = validateMeshArgs(ctx, d_vertices, numVerts, strideBytes, d_indices, numIndices, bound);
    if (rc) return rc;
    if ((reinterpret_cast<uintptr_t>(d_vertices) & 3u) || (reinterpret_cast<uintptr_t>(d_indices) & 3u))
        return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_build_bvh_device: buffers must be 4-byte aligned");
    DeviceGuard g(ctx->device);
    ctx->mesh = MeshView{static_cast<const uint8_t*>(d_vertices), numVerts, strideBytes, d_indices, numIndices / 3u};
    return buildOnDevice(ctx, bound);
}

int dxrv_get_bound(dxrv_ctx* ctx, float bound[4])
{
    if (!ctx || !bound) return DXRV_ERR_INVALID_ARG;
    if (!ctx->haveBvh) return fail(ctx, DXRV_ERR_NO_BVH, "dxrv_get_bound: no acceleration structure built");
    DeviceGuard g(ctx->device);
    DXRV_CUDA(cudaMemcpyAsync(bound, ctx->dBound, 4 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    DXRV_CUDA(cudaStreamSynchronize(ctx->stream));
    return DXRV_OK;
}

int dxrv_voxelize(dxrv_ctx* ctx, uint32_t N, uint32_t mode, uint32_t slabBegin, uint32_t slabEnd)
{
    if (!ctx) return DXRV_ERR_INVALID_ARG;
    if (!ctx->haveBvh) return fail(ctx, DXRV_ERR_NO_BVH, "dxrv_voxelize: call dxrv_build_bvh first");
    const uint32_t algo = mode & DXRV_MODE_MASK;
    NvtxRange range(algo == DXRV_MODE_PARITY ? "dxrv voxelize MODE_PARITY" : "dxrv voxelize MODE_SHADER");
    const bool wantTexels = (mode & DXRV_EMIT_TEXELS) != 0;
    if (algo != DXRV_MODE_SHADER && algo != DXRV_MODE_PARITY) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_voxelize: unknown mode");
    if (mode & ~(uint32_t)(DXRV_MODE_MASK | DXRV_EMIT_TEXELS)) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_voxelize: unknown mode flags");
    if (wantTexels && algo != DXRV_MODE_SHADER) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_voxelize: DXRV_EMIT_TEXELS needs DXRV_MODE_SHADER");
    if (N == 0 || N > 8192) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_voxelize: N must be in [1, 8192]");
    if (slabBegin > slabEnd || slabEnd > N) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_voxelize: need 0 <= slabBegin <= slabEnd <= N");
    if (slabBegin == slabEnd)
    {
        // an empty slab (more GPUs than layers): nothing to compute, but the context takes part in a later gather
        ctx->N = N; ctx->z0 = slabBegin; ctx->z1 = slabEnd; ctx->mode = mode;
        ctx->haveGrid = true; ctx->haveTexels = false; ctx->mipLevels = 0;
        return DXRV_OK;
    }
    if (algo == DXRV_MODE_SHADER && ctx->mesh.stride < 24)
        return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_voxelize: MODE_SHADER needs vertex normals (strideBytes >= 24)");
    DeviceGuard g(ctx->device);

    const size_t bytes = slabWords(N, slabBegin, slabEnd) * sizeof(uint32_t);
    uint32_t* grid = nullptr;
    if (ctx->gridTarget)
    {
        if (ctx->gridTargetBytes < bytes) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_voxelize: grid target too small for this slab");
        grid = ctx->gridTarget;
    }
    else
    {
        cudaError_t e = ensure(ctx->gridOwned, ctx->gridCap, bytes);
        if (e != cudaSuccess) return cudaFail(ctx, e, "cudaMalloc(grid)");
        grid = ctx->gridOwned;
    }
    uint32_t* texels = nullptr;
    if (wantTexels)
    {
        const size_t tb = (size_t)(slabEnd - slabBegin) * N * N * sizeof(uint32_t);
        cudaError_t e = ensure(ctx->texels, ctx->texCap, tb);
        if (e != cudaSuccess) return cudaFail(ctx, e, "cudaMalloc(texels)");
        texels = ctx->texels;
    }

    BvhView bvh{ctx->nodes, ctx->tris, ctx->dRootBox, ctx->mesh.numTris};   // (nodes dropped below when the consumer does not walk)
    if (algo == DXRV_MODE_PARITY)
    {
        const size_t walkBytes = sizeof(uint32_t) * parityScratchWords(N, slabBegin, slabEnd);
        const size_t zeroBytes = sizeof(uint32_t) * parityScratchZeroWords(N);
        // (a re-allocation may hand back the very same address, so compare capacities, not pointers)
        const bool reallocated = !ctx->walkBuf || walkBytes > ctx->walkCap;
        cudaError_t e = ensure(ctx->walkBuf, ctx->walkCap, walkBytes);
        if (e != cudaSuccess) return cudaFail(ctx, e, "cudaMalloc(walk lists)");
        if (reallocated) ctx->binsReady.valid = false;
        if (reallocated || zeroBytes != ctx->walkZeroed)
        {
            // the split-tile scratch is self-cleaning; it only needs zeroing when (re)allocated or resized
            DXRV_CUDA(cudaMemsetAsync(ctx->walkBuf, 0, zeroBytes, ctx->stream));
            ctx->walkZeroed = zeroBytes;
        }
    }
    bool buildBins = false, useBins = false;
    if (algo == DXRV_MODE_SHADER)
    {
        // Direction bins of the current acceleration structure: built by the first MODE_SHADER voxelize that is large
        // enough to pay for them (the build costs about as much as 100^3 rays through the LBVH), then reused by
        // every later voxelize until the next build.  DXRV_SHADER_PATH=bins|bvh forces one kernel.
        useBins = N >= 160u || ctx->binsValid;
        if (const char* f = std::getenv("DXRV_SHADER_PATH"))
        {
            if (!std::strcmp(f, "bvh")) useBins = false;
            else if (!std::strcmp(f, "bins")) useBins = true;
        }
        const ShaderBinsSizes sz = shaderBinsSizes(ctx->mesh.numTris);
        if (useBins && (!ctx->binsValid || sz.R != ctx->binsSizes.R || sz.cap != ctx->binsSizes.cap || !ctx->binsBuf))
        {
            cudaError_t e = ensure(ctx->binsBuf, ctx->binsCap, sz.bytes);
            if (e != cudaSuccess) return cudaFail(ctx, e, "cudaMalloc(direction bins)");
            ctx->binsSizes = sz;
            ctx->binsValid = false;
            buildBins = true;
        }
    }
    // fine mesh on a coarse grid: triangle-parallel scatter instead of the tile kernels (DXRV_PARITY_PATH=tiles|scatter forces one)
    bool scatter = algo == DXRV_MODE_PARITY && useScatterParity(ctx->mesh.numTris, N);
    if (const char* f = std::getenv("DXRV_PARITY_PATH"))
    {
        if (!std::strcmp(f, "tiles")) scatter = false;
        else if (!std::strcmp(f, "scatter") && algo == DXRV_MODE_PARITY && ctx->mesh.numTris >= 2 && N <= 2048) scatter = true;
    }
    std::vector<uint8_t> key;
    keyPush(key, (uint32_t)0x70C5u);
    keyPush(key, (uint32_t)scatter);
    keyPush(key, N); keyPush(key, mode); keyPush(key, slabBegin); keyPush(key, slabEnd);
    keyPush(key, grid); keyPush(key, texels); keyPush(key, ctx->walkBuf);
    // MODE_PARITY tile path: candidates by triangle-parallel binning (default) or by the LBVH walk (DXRV_PARITY_CANDIDATES=walk)
    bool walkTree = false;
    if (const char* f = std::getenv("DXRV_PARITY_CANDIDATES")) walkTree = !std::strcmp(f, "walk");
    const bool usesTree = algo == DXRV_MODE_SHADER || (!scatter && walkTree);
    const bool needTree = usesTree && !ctx->treeBuilt;
    if (algo == DXRV_MODE_PARITY && !(walkTree && !scatter)) bvh.nodes = nullptr;
    const bool tileBins = algo == DXRV_MODE_PARITY && !scatter && !walkTree;
    const bool binsReady = tileBins && ctx->binsReady.valid && ctx->binsReady.N == N && ctx->binsReady.z0 == slabBegin && ctx->binsReady.z1 == slabEnd;
    keyPush(key, (uint32_t)needTree); keyPush(key, (uint32_t)walkTree); keyPush(key, (uint32_t)binsReady);
    keyPush(key, ctx->binsBuf); keyPush(key, (uint32_t)buildBins); keyPush(key, (uint32_t)useBins); keyPush(key, ctx->binsSizes.R);
    keyPush(key, ctx->nodes); keyPush(key, ctx->tris);
    keyPush(key, ctx->mesh.verts); keyPush(key, ctx->mesh.numVerts); keyPush(key, ctx->mesh.stride); keyPush(key, ctx->mesh.indices); keyPush(key, ctx->mesh.numTris);
    const int rc = runCaptured(ctx, key, [&]() {
        if (needTree) enqueueTree(ctx);
        if (algo == DXRV_MODE_PARITY)
        {
            if (scatter)
                ctx->launches += (uint64_t)launchScatterFillColumns(ctx->stream, bvh, N, slabBegin, slabEnd, grid, ctx->walkBuf, ctx->dCrossings,
                                                                    ctx->profiling ? ctx->prof : nullptr);
            else
                ctx->launches += (uint64_t)launchTraceFillColumns(ctx->stream, bvh, N, slabBegin, slabEnd, grid, ctx->walkBuf,
                                                                  ctx->dCrossings, ctx->dErr, ctx->profiling ? ctx->prof : nullptr, binsReady);
            ctx->profValid = ctx->profiling;
        }
        else
        {
            if (useBins)
            {
                if (buildBins) ctx->launches += (uint64_t)launchShaderBinsBuild(ctx->stream, bvh, ctx->binsBuf, ctx->binsSizes, false);
                // exactly one of the two does the work (device-side overflow flag of the bins)
                launchTraceShaderBins(ctx->stream, bvh, ctx->mesh, N, slabBegin, slabEnd, grid, texels, ctx->dErr, ctx->binsBuf, ctx->binsSizes, ctx->dCentres);
                launchTraceShaderBvh(ctx->stream, bvh, ctx->mesh, N, slabBegin, slabEnd, grid, texels, ctx->dErr,
                                     shaderBinsView(ctx->binsBuf, ctx->binsSizes).state);
                ctx->launches += 2;
            }
            else launchTraceShaderBvh(ctx->stream, bvh, ctx->mesh, N, slabBegin, slabEnd, grid, texels, ctx->dErr, nullptr);
            ctx->launches += 1;
        }
    });
    if (rc) return rc;
    DXRV_CUDA(cudaGetLastError());
    ctx->N = N; ctx->z0 = slabBegin; ctx->z1 = slabEnd; ctx->mode = mode;
    ctx->haveGrid = true; ctx->haveTexels = wantTexels; ctx->mipLevels = 0;
    if (needTree) ctx->treeBuilt = ctx->pyramidBuilt = true;
    if (algo == DXRV_MODE_PARITY)
    {
        // the lists this call left (or used) stay valid for the same structure, grid and slab; the other paths overwrite them
        ctx->binsReady.valid = tileBins; ctx->binsReady.N = N; ctx->binsReady.z0 = slabBegin; ctx->binsReady.z1 = slabEnd;
    }
    ctx->treeWanted = usesTree;   // the next build includes the hierarchy iff this consumer traversed it
    if (algo == DXRV_MODE_SHADER && useBins) ctx->binsValid = true;
    return DXRV_OK;
}

int dxrv_fetch_grid(dxrv_ctx* ctx, void* hostDst, size_t bytes, uint32_t format)
{
    if (!ctx || !hostDst) return DXRV_ERR_INVALID_ARG;
    if (!ctx->haveGrid) return fail(ctx, DXRV_ERR_NO_GRID, "dxrv_fetch_grid: call dxrv_voxelize first");
    NvtxRange range("dxrv fetch grid (D2H)");
    DeviceGuard g(ctx->device);
    const uint32_t* grid = ctx->gridTarget ? ctx->gridTarget : ctx->gridOwned;
    const uint32_t layers = ctx->z1 - ctx->z0;
    const void* src = nullptr;
    size_t need = 0;
    if (format == DXRV_FORMAT_BITS) { src = grid; need = slabWords(ctx->N, ctx->z0, ctx->z1) * sizeof(uint32_t); }
    else if (format == DXRV_FORMAT_U8)
    {
        need = (size_t)layers * ctx->N * ctx->N;
        cudaError_t e = ensure(ctx->u8Temp, ctx->u8Cap, need);
        if (e != cudaSuccess) return cudaFail(ctx, e, "cudaMalloc(u8)");
        launchBitsToU8(ctx->stream, grid, ctx->N, layers, ctx->u8Temp);
        ctx->launches += 1;
        src = ctx->u8Temp;
    }
    else if (format == DXRV_FORMAT_R10G10B10A2)
    {
        if (!ctx->haveTexels) return fail(ctx, DXRV_ERR_NO_GRID, "dxrv_fetch_grid: texels need DXRV_MODE_SHADER | DXRV_EMIT_TEXELS");
        src = ctx->texels; need = (size_t)layers * ctx->N * ctx->N * sizeof(uint32_t);
    }
    else return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_fetch_grid: unknown format");
    if (bytes != need) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_fetch_grid: bytes does not match the slab size in this format");
    DXRV_CUDA(cudaMemcpyAsync(hostDst, src, need, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->lastD2hBytes = need;
    return checkDeviceError(ctx);
}

int dxrv_voxelize_to_host(dxrv_ctx* ctx, uint32_t N, uint32_t mode, uint32_t slabBegin, uint32_t slabEnd, void* hostDst, size_t bytes,
                          uint32_t chunks)
{
    if (const int bad = toHostArgsCheck(ctx, N, mode, slabBegin, slabEnd, hostDst, bytes)) return bad;
    const size_t layerBytes = (size_t)N * ((N + 31) / 32) * sizeof(uint32_t);
    const uint32_t layers = slabEnd - slabBegin;
    DeviceGuard g(ctx->device);
    // Transport (dxrv_set_read_back).  The dense grid over one PCIe link is the floor of the copying path (128 MiB: 2.4 ms);
    // a solid voxelization is mostly empty space and solid interior, so by default a large slab comes back as
    // DXRV_FORMAT_SPARSE_BRICKS (a few MB) and is written into hostDst by the host pool's threads in one pass (zeros from
    // the outside of the slab inwards while the GPU is still computing, final contents once the blob is here).  The host's
    // memory write bandwidth is then the floor (0.69 ms for 128 MiB on 16 cores; the call takes 0.63-0.75 ms).  A grid
    // that does not compress (more than half of the dense size) is copied densely after all.
    {
        if (toHostTransport(ctx, bytes) == 2u)
        {
            const int rcS = voxelizeToHostSparse(ctx, N, mode, slabBegin, slabEnd, hostDst, bytes);
            if (rcS != DXRV_ERR_UNSUPPORTED) return rcS;   // UNSUPPORTED: did not compress; the slab is resident, copy it densely
            DXRV_CUDA(cudaMemcpyAsync(hostDst, ctx->gridOwned, bytes, cudaMemcpyDeviceToHost, ctx->stream));
            ctx->lastD2hBytes = bytes + 64;
            return checkDeviceError(ctx);
        }
    }
    if (chunks < 1) chunks = 1;
    if (chunks > layers) chunks = layers;
    if (layerBytes % 16) chunks = 1;   // sub-slab offsets must keep the 128-bit store alignment
    if (!ctx->copyStream)
    {
        DXRV_CUDA(cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking));
        for (cudaEvent_t& e : ctx->chunkDone) DXRV_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    cudaError_t e = ensure(ctx->gridOwned, ctx->gridCap, bytes);
    if (e != cudaSuccess) return cudaFail(ctx, e, "cudaMalloc(grid)");
    uint32_t* base = ctx->gridOwned;
    int rc = DXRV_OK;
    // chunk k is computed on the context's stream into its place of the slab; its copy waits for it on the copy stream
    // while chunk k + 1 is computed
    for (uint32_t k = 0; k < chunks && rc == DXRV_OK; ++k)
    {
        const uint32_t a = slabBegin + (uint32_t)((uint64_t)layers * k / chunks), b = slabBegin + (uint32_t)((uint64_t)layers * (k + 1) / chunks);
        const size_t off = layerBytes * (a - slabBegin), len = layerBytes * (b - a);
        ctx->gridTarget = base + off / sizeof(uint32_t);
        ctx->gridTargetBytes = len;
        rc = dxrv_voxelize(ctx, N, mode, a, b);
        if (rc != DXRV_OK) break;
        cudaEvent_t ev = ctx->chunkDone[k % 2];
        DXRV_CUDA(cudaEventRecord(ev, ctx->stream));
        DXRV_CUDA(cudaStreamWaitEvent(ctx->copyStream, ev, 0));
        DXRV_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(hostDst) + off, reinterpret_cast<uint8_t*>(base) + off, len, cudaMemcpyDeviceToHost, ctx->copyStream));
    }
    ctx->gridTarget = nullptr;
    ctx->gridTargetBytes = 0;
    if (rc != DXRV_OK) { cudaStreamSynchronize(ctx->copyStream); return rc; }
    DXRV_CUDA(cudaStreamSynchronize(ctx->copyStream));
    // the context now describes the whole slab, resident in its own grid
    ctx->N = N; ctx->z0 = slabBegin; ctx->z1 = slabEnd; ctx->mode = mode;
    ctx->haveGrid = true; ctx->haveTexels = false; ctx->mipLevels = 0;
    ctx->lastD2hBytes = bytes;
    return checkDeviceError(ctx);
}

int dxrv_voxelize_mesh_to_host(dxrv_ctx* ctx, const void* vertices, uint32_t numVerts, uint32_t strideBytes, const uint32_t* indices,
                               uint32_t numIndices, const float bound[4], uint32_t N, uint32_t mode, uint32_t slabBegin, uint32_t slabEnd,
                               void* hostDst, size_t bytes, uint32_t chunks)
{
    // dxrv_build_bvh + dxrv_voxelize_to_host as ONE call: a frame of a deforming mesh (upload, rebuild, voxelize, read back).
    // What the single call buys: with the compact transport the host pool starts on hostDst BEFORE the upload and the build,
    // i.e. ~0.1 ms earlier, and the pass over the host grid is what bounds the call.
    int rc = validateMeshArgs(ctx, vertices, numVerts, strideBytes, indices, numIndices, bound);
    if (rc) return rc;
    if ((rc = toHostArgsCheck(ctx, N, mode, slabBegin, slabEnd, hostDst, bytes)) != DXRV_OK) return rc;
    if (toHostTransport(ctx, bytes) == 2u && !hostFillTwoPass())
    {
        hostFillBegin(hostDst, N, slabBegin, slabEnd - slabBegin);
        ctx->hostFillBegun = true;
    }
    rc = dxrv_build_bvh(ctx, vertices, numVerts, strideBytes, indices, numIndices, bound);
    if (rc == DXRV_OK) rc = dxrv_voxelize_to_host(ctx, N, mode, slabBegin, slabEnd, hostDst, bytes, chunks);
    if (ctx->hostFillBegun)   // the build failed, or the voxelize left before it took the pass over: release the pool
    {
        hostFillWait();
        ctx->hostFillBegun = false;
    }
    return rc;
}

int dxrv_set_read_back(dxrv_ctx* ctx, uint32_t transport)
{
    if (!ctx) return DXRV_ERR_INVALID_ARG;
    if (transport > DXRV_READ_BACK_SPARSE) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_set_read_back: unknown transport");
    ctx->readBack = transport;
    return DXRV_OK;
}

int dxrv_fetch_grid_sparse(dxrv_ctx* ctx, void* hostDst, size_t capacity, size_t* bytesWritten)
{
    if (!ctx || !hostDst || !bytesWritten) return DXRV_ERR_INVALID_ARG;
    *bytesWritten = 0;
    if (!ctx->haveGrid) return fail(ctx, DXRV_ERR_NO_GRID, "dxrv_fetch_grid_sparse: call dxrv_voxelize first");
    if (ctx->z1 == ctx->z0) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_fetch_grid_sparse: empty slab");
    NvtxRange range("dxrv fetch grid (sparse bricks, D2H)");
    DeviceGuard g(ctx->device);
    const SparseLayout L = sparseLayout(ctx->N, ctx->z1 - ctx->z0);
    const size_t countsOff = (L.maxBytes + 255) & ~(size_t)255;
    cudaError_t e = ensure(ctx->sparseBuf, ctx->sparseCap, countsOff + (size_t)L.numBlocks * sizeof(uint32_t));
    if (e != cudaSuccess) return cudaFail(ctx, e, "cudaMalloc(sparse bricks)");
    const uint32_t* grid = ctx->gridTarget ? ctx->gridTarget : ctx->gridOwned;
    ctx->launches += (uint64_t)launchSparseEncode(ctx->stream, grid, ctx->N, ctx->z0, ctx->z1, ctx->sparseBuf,
                                                  reinterpret_cast<uint32_t*>(ctx->sparseBuf + countsOff));
    DXRV_CUDA(cudaGetLastError());
    // the header (with the number of mixed bricks) first, then exactly the bytes that exist
    if (capacity < 64) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_fetch_grid_sparse: capacity below the header size");
    DXRV_CUDA(cudaMemcpyAsync(hostDst, ctx->sparseBuf, 64, cudaMemcpyDeviceToHost, ctx->stream));
    DXRV_CUDA(cudaStreamSynchronize(ctx->stream));
    const uint32_t numMixed = static_cast<const uint32_t*>(hostDst)[9];
    const size_t total = L.offPayload + (size_t)numMixed * 64;
    *bytesWritten = total;
    if (capacity < total) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_fetch_grid_sparse: capacity too small (needed size returned in *bytesWritten)");
    DXRV_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(hostDst) + 64, ctx->sparseBuf + 64, total - 64, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->lastD2hBytes = total;
    return checkDeviceError(ctx);
}

int dxrv_sparse_decode(const void* blob, size_t blobBytes, void* denseDst, size_t denseBytes)
{
    // pure host code (sparse_host.cpp): the inverse of the encoding in sparse.cu, on the host pool's threads
    SparseBlobView v;
    if (!denseDst || !sparseParse(blob, blobBytes, v)) return DXRV_ERR_INVALID_ARG;
    if (denseBytes != (size_t)(v.z1 - v.z0) * v.N * v.P * 4) return DXRV_ERR_INVALID_ARG;
    // one pass of the host pool over the dense grid (sparse_host.cpp); DXRV_HOST_FILL_DELAY_US (tests) publishes the blob
    // late, so that the pool has zeroed brick layers before it knows their contents and must expand them afterwards
    hostFillBegin(denseDst, v.N, v.z0, v.z1 - v.z0);
    if (const char* e = std::getenv("DXRV_HOST_FILL_DELAY_US"))
    {
        const long us = std::atol(e);
        if (us > 0 && us <= 1000000) std::this_thread::sleep_for(std::chrono::microseconds(us));
    }
    const bool published = hostFillPublish(v);
    const bool filled = hostFillWait();
    return published && filled ? DXRV_OK : DXRV_ERR_INVALID_ARG;
}

int dxrv_sparse_encode(const void* dense, size_t denseBytes, uint32_t N, uint32_t slabBegin, uint32_t slabEnd, void* blob, size_t capacity,
                       size_t* bytesWritten)
{
    // pure host code (sparse_host.cpp): the device encoder's bytes from a dense host grid
    if (!dense || !bytesWritten || N == 0 || N > 8192 || slabBegin >= slabEnd || slabEnd > N) return DXRV_ERR_INVALID_ARG;
    *bytesWritten = 0;
    if (denseBytes != (size_t)(slabEnd - slabBegin) * N * ((N + 31) / 32) * 4) return DXRV_ERR_INVALID_ARG;
    size_t bytes = 0;
    try
    {
        const bool ok = sparseEncode(static_cast<const uint32_t*>(dense), N, slabBegin, slabEnd, blob, capacity, bytes);
        *bytesWritten = bytes;
        return ok ? DXRV_OK : DXRV_ERR_INVALID_ARG;
    }
    catch (const std::bad_alloc&) { return DXRV_ERR_OOM; }
}

int dxrv_grid_device(dxrv_ctx* ctx, void** d_ptr, size_t* bytes)
{
    if (!ctx || !d_ptr || !bytes) return DXRV_ERR_INVALID_ARG;
    if (!ctx->haveGrid) return fail(ctx, DXRV_ERR_NO_GRID, "dxrv_grid_device: call dxrv_voxelize first");
    *d_ptr = ctx->gridTarget ? ctx->gridTarget : ctx->gridOwned;
    *bytes = slabWords(ctx->N, ctx->z0, ctx->z1) * sizeof(uint32_t);
    return DXRV_OK;
}

int dxrv_set_grid_target(dxrv_ctx* ctx, void* d_ptr, size_t bytes)
{
    if (!ctx) return DXRV_ERR_INVALID_ARG;
    if (d_ptr && (reinterpret_cast<uintptr_t>(d_ptr) & 15u)) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_set_grid_target: pointer must be 16-byte aligned");
    ctx->gridTarget = static_cast<uint32_t*>(d_ptr);
    ctx->gridTargetBytes = d_ptr ? bytes : 0;
    ctx->haveGrid = false;
    return DXRV_OK;
}

static size_t mipWords(uint32_t N, uint32_t layers, uint32_t level)
{
    const uint32_t n = N >> level, l = layers >> level;
    return (size_t)l * n * ((n + 31) / 32);
}

int dxrv_build_mips(dxrv_ctx* ctx, uint32_t* numLevels)
{
    if (!ctx || !numLevels) return DXRV_ERR_INVALID_ARG;
    if (!ctx->haveGrid) return fail(ctx, DXRV_ERR_NO_GRID, "dxrv_build_mips: call dxrv_voxelize first");
    DeviceGuard g(ctx->device);
    const uint32_t layers = ctx->z1 - ctx->z0;
    uint32_t levels = 1;
    size_t words = 0;
    while (((ctx->N >> (levels - 1)) & 1u) == 0u && ((layers >> (levels - 1)) & 1u) == 0u && (ctx->N >> levels) >= 1u && (layers >> levels) >= 1u)
    {
        words += mipWords(ctx->N, layers, levels);
        ++levels;
    }
    cudaError_t e = ensure(ctx->mips, ctx->mipCap, words * sizeof(uint32_t));
    if (e != cudaSuccess) return cudaFail(ctx, e, "cudaMalloc(mips)");
    const uint32_t* src = ctx->gridTarget ? ctx->gridTarget : ctx->gridOwned;
    uint32_t* dst = ctx->mips;
    for (uint32_t l = 1; l < levels; ++l)
    {
        launchMipReduce(ctx->stream, src, ctx->N >> (l - 1), layers >> (l - 1), dst);
        ctx->launches += 1;
        src = dst;
        dst += mipWords(ctx->N, layers, l);
    }
    DXRV_CUDA(cudaGetLastError());
    ctx->mipLevels = levels;
    *numLevels = levels;
    return DXRV_OK;
}

int dxrv_fetch_mip(dxrv_ctx* ctx, uint32_t level, void* hostDst, size_t bytes)
{
    if (!ctx || !hostDst) return DXRV_ERR_INVALID_ARG;
    if (!ctx->haveGrid || ctx->mipLevels == 0) return fail(ctx, DXRV_ERR_NO_GRID, "dxrv_fetch_mip: call dxrv_build_mips first");
    if (level < 1 || level >= ctx->mipLevels) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_fetch_mip: no such level (level 0 is dxrv_fetch_grid)");
    DeviceGuard g(ctx->device);
    const uint32_t layers = ctx->z1 - ctx->z0;
    const uint32_t* src = ctx->mips;
    for (uint32_t l = 1; l < level; ++l) src += mipWords(ctx->N, layers, l);
    const size_t need = mipWords(ctx->N, layers, level) * sizeof(uint32_t);
    if (bytes != need) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_fetch_mip: bytes does not match the level size");
    DXRV_CUDA(cudaMemcpyAsync(hostDst, src, need, cudaMemcpyDeviceToHost, ctx->stream));
    return checkDeviceError(ctx);
}

int dxrv_render_view(dxrv_ctx* ctx, uint32_t width, uint32_t height, const float screenToLocal[16], const float eye[3],
                     const float light[3], void* hostRGBA, size_t bytes)
{
    if (!ctx || !screenToLocal || !eye || !light || !hostRGBA) return DXRV_ERR_INVALID_ARG;
    if (!ctx->haveGrid) return fail(ctx, DXRV_ERR_NO_GRID, "dxrv_render_view: call dxrv_voxelize first");
    if (ctx->z0 != 0 || ctx->z1 != ctx->N) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_render_view: needs the full grid (slab [0, N))");
    if (width == 0 || height == 0 || width > 16384 || height > 16384) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_render_view: bad target size");
    const size_t need = (size_t)width * height * 4;
    if (bytes != need) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_render_view: bytes must be width * height * 4");
    DeviceGuard g(ctx->device);
    cudaError_t e = ensure(ctx->u8Temp, ctx->u8Cap, need);
    if (e != cudaSuccess) return cudaFail(ctx, e, "cudaMalloc(image)");
    const uint32_t* grid = ctx->gridTarget ? ctx->gridTarget : ctx->gridOwned;
    launchRaycastView(ctx->stream, grid, ctx->N, width, height, screenToLocal, eye, light, reinterpret_cast<uint32_t*>(ctx->u8Temp));
    ctx->launches += 1;
    DXRV_CUDA(cudaGetLastError());
    DXRV_CUDA(cudaMemcpyAsync(hostRGBA, ctx->u8Temp, need, cudaMemcpyDeviceToHost, ctx->stream));
    return checkDeviceError(ctx);
}

int dxrv_count_inside(dxrv_ctx* ctx, uint64_t* count)
{
    if (!ctx || !count) return DXRV_ERR_INVALID_ARG;
    if (!ctx->haveGrid) return fail(ctx, DXRV_ERR_NO_GRID, "dxrv_count_inside: call dxrv_voxelize first");
    DeviceGuard g(ctx->device);
    const uint32_t* grid = ctx->gridTarget ? ctx->gridTarget : ctx->gridOwned;
    launchPopcount(ctx->stream, grid, slabWords(ctx->N, ctx->z0, ctx->z1), ctx->dCount);
    ctx->launches += 1;
    unsigned long long c = 0;
    DXRV_CUDA(cudaMemcpyAsync(&c, ctx->dCount, sizeof(c), cudaMemcpyDeviceToHost, ctx->stream));
    int rc = checkDeviceError(ctx);
    *count = c;
    return rc;
}

int dxrv_get_info(dxrv_ctx* ctx, uint32_t what, uint64_t* value)
{
    if (!ctx || !value) return DXRV_ERR_INVALID_ARG;
    DeviceGuard g(ctx->device);
    switch (what)
    {
    case DXRV_INFO_NUM_TRIANGLES: *value = ctx->haveBvh ? ctx->mesh.numTris : 0; return DXRV_OK;
    case DXRV_INFO_NUM_NODES: *value = (ctx->haveBvh && ctx->mesh.numTris) ? 2ull * ctx->mesh.numTris - 1 : 0; return DXRV_OK;
    case DXRV_INFO_KERNEL_LAUNCHES: *value = ctx->launches; return DXRV_OK;
    case DXRV_INFO_SM_COUNT: *value = (uint64_t)ctx->smCount; return DXRV_OK;
    case DXRV_INFO_LAST_D2H_BYTES: *value = ctx->lastD2hBytes; return DXRV_OK;
    case DXRV_INFO_LAST_WALK_NS:
    case DXRV_INFO_LAST_FILL_NS:
    {
        if (!ctx->profValid) return fail(ctx, DXRV_ERR_NO_GRID, "dxrv_get_info: enable dxrv_set_profiling and run a MODE_PARITY voxelize first");
        DXRV_CUDA(cudaEventSynchronize(ctx->prof[2]));
        float ms = 0.0f;
        const int a = what == DXRV_INFO_LAST_WALK_NS ? 0 : 1;
        DXRV_CUDA(cudaEventElapsedTime(&ms, ctx->prof[a], ctx->prof[a + 1]));
        *value = (uint64_t)(ms * 1e6f + 0.5f);
        return DXRV_OK;
    }
    case DXRV_INFO_LAST_BUILD_NS:
    case DXRV_INFO_LAST_SORT_NS:
    {
        if (!ctx->profBuildValid) return fail(ctx, DXRV_ERR_NO_BVH, "dxrv_get_info: enable dxrv_set_profiling and build first");
        DXRV_CUDA(cudaEventSynchronize(ctx->profBuild[3]));
        float ms = 0.0f;
        if (what == DXRV_INFO_LAST_BUILD_NS) DXRV_CUDA(cudaEventElapsedTime(&ms, ctx->profBuild[0], ctx->profBuild[3]));
        else DXRV_CUDA(cudaEventElapsedTime(&ms, ctx->profBuild[1], ctx->profBuild[2]));
        *value = (uint64_t)(ms * 1e6f + 0.5f);
        return DXRV_OK;
    }
    case DXRV_INFO_CROSSINGS:
    {
        unsigned long long c = 0;
        DXRV_CUDA(cudaMemcpyAsync(&c, ctx->dCrossings, sizeof(c), cudaMemcpyDeviceToHost, ctx->stream));
        DXRV_CUDA(cudaStreamSynchronize(ctx->stream));
        *value = c;
        return DXRV_OK;
    }
    default: return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_get_info: unknown query");
    }
}

int dxrv_debug_read(dxrv_ctx* ctx, uint32_t what, void* hostDst, size_t bytes)
{
    if (!ctx || !hostDst) return DXRV_ERR_INVALID_ARG;
    if (!ctx->haveBvh) return fail(ctx, DXRV_ERR_NO_BVH, "dxrv_debug_read: no acceleration structure built");
    DeviceGuard g(ctx->device);
    const size_t T = ctx->mesh.numTris;
    const void* src = nullptr; size_t need = 0;
    if (what == DXRV_DBG_BINS_STATE)
    {
        if (bytes != 16) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_debug_read: size mismatch");
        uint32_t* out = static_cast<uint32_t*>(hostDst);
        if (!ctx->binsValid || !ctx->binsBuf) { out[0] = 0; out[1] = 1; out[2] = 0; out[3] = 0; return DXRV_OK; }   // LBVH walk only
        DXRV_CUDA(cudaMemcpyAsync(out, shaderBinsView(ctx->binsBuf, ctx->binsSizes).state, 12, cudaMemcpyDeviceToHost, ctx->stream));
        DXRV_CUDA(cudaStreamSynchronize(ctx->stream));
        out[3] = ctx->binsSizes.R;
        return DXRV_OK;
    }
    if ((what == DXRV_DBG_NODES || what == DXRV_DBG_ROOT_BOX) && !ctx->treeBuilt)
    {
        enqueueTree(ctx);
        DXRV_CUDA(cudaGetLastError());
        ctx->treeBuilt = true;
    }
    switch (what)
    {
    case DXRV_DBG_MORTON_SORTED: src = ctx->keysA; need = T * 4; break;
    case DXRV_DBG_PRIM_SORTED: src = ctx->valsA; need = T * 4; break;
    case DXRV_DBG_NODES: src = ctx->nodes; need = (T ? T - 1 : 0) * sizeof(BvhNode); break;
    case DXRV_DBG_TRIS: src = ctx->tris; need = T * sizeof(Tri48); break;
    case DXRV_DBG_ROOT_BOX: src = ctx->dRootBox; need = 6 * sizeof(float); break;
    default: return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_debug_read: unknown buffer");
    }
    if (bytes != need) return fail(ctx, DXRV_ERR_INVALID_ARG, "dxrv_debug_read: size mismatch");
    if (need) DXRV_CUDA(cudaMemcpyAsync(hostDst, src, need, cudaMemcpyDeviceToHost, ctx->stream));
    DXRV_CUDA(cudaStreamSynchronize(ctx->stream));
    return DXRV_OK;
}

int dxrv_debug_sort_pairs(dxrv_ctx* ctx, uint32_t* keys, uint32_t* values, uint32_t n)
{
    if (!ctx || (n && (!keys || !values))) return DXRV_ERR_INVALID_ARG;
    if (n == 0) return DXRV_OK;
    DeviceGuard g(ctx->device);
    uint32_t* buf = nullptr;
    void* temp = nullptr;
    DXRV_CUDA(cudaMalloc(reinterpret_cast<void**>(&buf), sizeof(uint32_t) * 4 * (size_t)n));
    cudaError_t e = cudaMalloc(&temp, SortTemp::bytesFor(n));
    if (e != cudaSuccess) { cudaFree(buf); return cudaFail(ctx, e, "cudaMalloc(sort temp)"); }
    uint32_t *kA = buf, *vA = buf + n, *kB = buf + 2 * (size_t)n, *vB = buf + 3 * (size_t)n;
    cudaMemcpyAsync(kA, keys, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(vA, values, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, ctx->stream);
    ctx->launches += (uint64_t)radixSortPairs(ctx->stream, temp, kA, vA, kB, vB, n, 4, false, nullptr);
    cudaMemcpyAsync(keys, kA, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(values, vA, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, ctx->stream);
    e = cudaStreamSynchronize(ctx->stream);
    cudaFree(buf); cudaFree(temp);
    if (e != cudaSuccess) return cudaFail(ctx, e, "radix sort");
    return DXRV_OK;
}

void* dxrv_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

void dxrv_host_free(void* p) { if (p) cudaFreeHost(p); }

int dxrv_ipc_export_grid(dxrv_ctx* ctx, size_t fullBytes, void* handle64, void** d_ptr)
{
    if (!ctx || !handle64 || !d_ptr || !fullBytes) return DXRV_ERR_INVALID_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DeviceGuard g(ctx->device);
    DXRV_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->gridCap < fullBytes || !ctx->gridOwned)
    {
        // exact, stand-alone allocation: IPC handles cover whole cudaMalloc allocations
        if (ctx->gridOwned) cudaFree(ctx->gridOwned);
        ctx->gridOwned = nullptr; ctx->gridCap = 0;
        DXRV_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->gridOwned), fullBytes));
        ctx->gridCap = fullBytes;
        ctx->haveGrid = false;
    }
    cudaIpcMemHandle_t h;
    DXRV_CUDA(cudaIpcGetMemHandle(&h, ctx->gridOwned));
    std::memcpy(handle64, &h, sizeof(h));
    *d_ptr = ctx->gridOwned;
    return DXRV_OK;
}

int dxrv_ipc_open(dxrv_ctx* ctx, const void* handle64, void** d_ptr)
{
    if (!ctx || !handle64 || !d_ptr) return DXRV_ERR_INVALID_ARG;
    DeviceGuard g(ctx->device);
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, sizeof(h));
    DXRV_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return DXRV_OK;
}

int dxrv_ipc_close(dxrv_ctx* ctx, void* d_ptr)
{
    if (!ctx || !d_ptr) return DXRV_ERR_INVALID_ARG;
    DeviceGuard g(ctx->device);
    DXRV_CUDA(cudaStreamSynchronize(ctx->stream));
    DXRV_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return DXRV_OK;
}

}  // extern "C"
