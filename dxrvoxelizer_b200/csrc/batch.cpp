// batch.cpp -- dxrv_voxelize_obj_batch (include/dxrv.h): a stream of distinct OBJ files -> one grid each.
//
// The reference loads ONE mesh in LoadAssets (Import, twice: DXRVoxelizer.cpp:190-199) and voxelizes it every frame; the
// streaming case (BASELINE config 5: 256 distinct meshes at 256^3) is that pair repeated per mesh.  Here the pair is a
// pipeline inside the library: loader threads parse the files (one thread per file, parseObjFast), every context
// (= one CUDA stream) is driven by its own host thread that takes meshes k, k + numCtx, ... in order -- upload, build,
// voxelize, read-back of that context's previous grid -- so the GPU always has several meshes in flight and no
// interpreter sits between the stages.  Pure host code over the public entry points; no device code of its own.
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/dxrv.h"
#include "obj_loader.h"

namespace dxrv
{
std::string& globalError();   // obj_capi.cpp: last error of the context-free entry points (thread local)
}

namespace
{
struct Slot
{
    std::unique_ptr<dxrv::ObjMesh> mesh;
    bool ready = false;      // parsed (mesh set) or failed (mesh null)
};

struct Batch
{
    std::mutex m;
    std::condition_variable loaded;     // a slot became ready
    std::condition_variable consumed;   // a mesh was taken: the loaders may run further ahead
    std::vector<Slot> slots;
    std::atomic<uint32_t> nextToLoad{0};
    uint32_t taken = 0;                 // meshes handed to the drivers so far
    uint32_t window = 0;                // how far the loaders may run ahead of `taken`
    int rc = DXRV_OK;                   // first failure
    std::string error;
    bool stop = false;

    void failWith(int code, const std::string& what)
    {
        std::lock_guard<std::mutex> g(m);
        if (rc == DXRV_OK) { rc = code; error = what; }
        stop = true;
        loaded.notify_all();
        consumed.notify_all();
    }
};
}  // namespace

extern "C" int dxrv_voxelize_obj_batch(dxrv_ctx* const* ctxs, uint32_t numCtx, const char* const* paths, uint32_t numMeshes, uint32_t N,
                                       uint32_t mode, void* hostGrids, size_t gridBytes, uint32_t loaderThreads, uint32_t* numTriangles)
{
    if (!ctxs || numCtx == 0 || (!paths && numMeshes)) { dxrv::globalError() = "dxrv_voxelize_obj_batch: null argument"; return DXRV_ERR_INVALID_ARG; }
    for (uint32_t s = 0; s < numCtx; ++s)
    {
        if (!ctxs[s]) { dxrv::globalError() = "dxrv_voxelize_obj_batch: null context"; return DXRV_ERR_INVALID_ARG; }
        for (uint32_t t = 0; t < s; ++t)
            if (ctxs[t] == ctxs[s]) { dxrv::globalError() = "dxrv_voxelize_obj_batch: the contexts must be distinct (a context is not thread-safe)"; return DXRV_ERR_INVALID_ARG; }
    }
    for (uint32_t k = 0; k < numMeshes; ++k)
        if (!paths[k]) { dxrv::globalError() = "dxrv_voxelize_obj_batch: null path"; return DXRV_ERR_INVALID_ARG; }
    const size_t P = (static_cast<size_t>(N) + 31) / 32;
    if (mode != DXRV_MODE_SHADER && mode != DXRV_MODE_PARITY)
    {
        dxrv::globalError() = "dxrv_voxelize_obj_batch: mode must be DXRV_MODE_SHADER or DXRV_MODE_PARITY (bit grids only)";
        return DXRV_ERR_INVALID_ARG;
    }
    if (N == 0 || N > 8192) { dxrv::globalError() = "dxrv_voxelize_obj_batch: N must be in [1, 8192]"; return DXRV_ERR_INVALID_ARG; }
    if (hostGrids && gridBytes != static_cast<size_t>(N) * N * P * 4)
    {
        dxrv::globalError() = "dxrv_voxelize_obj_batch: gridBytes must be N * N * ceil(N / 32) * 4 (DXRV_FORMAT_BITS)";
        return DXRV_ERR_INVALID_ARG;
    }
    if (numMeshes == 0) return DXRV_OK;
    try
    {
        if (loaderThreads == 0) loaderThreads = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
        loaderThreads = std::min(loaderThreads, numMeshes);

        Batch b;
        b.slots.resize(numMeshes);
        b.window = 2 * loaderThreads + 2 * numCtx;    // parsed meshes waiting for a stream: bounded, whatever the batch size

        auto loader = [&]() {
            while (true)
            {
                const uint32_t k = b.nextToLoad.fetch_add(1);
                if (k >= numMeshes) return;
                {
                    std::unique_lock<std::mutex> g(b.m);
                    b.consumed.wait(g, [&] { return b.stop || k < b.taken + b.window; });
                    if (b.stop) return;
                }
                std::unique_ptr<dxrv::ObjMesh> mesh(new dxrv::ObjMesh());
                std::string err;
                bool ok = false;
                try { ok = dxrv::loadObj(paths[k], *mesh, err, 1); }
                catch (const std::bad_alloc&) { b.failWith(DXRV_ERR_OOM, std::string("out of memory while loading ") + paths[k]); return; }
                catch (...) { err = "unexpected exception"; }
                if (!ok) { b.failWith(DXRV_ERR_IO, std::string(paths[k]) + ": " + err); return; }
                if (mesh->numIndices() == 0 || mesh->numVertices() == 0) { b.failWith(DXRV_ERR_IO, std::string(paths[k]) + ": no triangles"); return; }
                {
                    std::lock_guard<std::mutex> g(b.m);
                    b.slots[k].mesh = std::move(mesh);
                    b.slots[k].ready = true;
                }
                b.loaded.notify_all();
            }
        };

        auto driver = [&](uint32_t s) {
            dxrv_ctx* c = ctxs[s];
            auto ctxFail = [&](int code, const char* call, uint32_t k) {
                b.failWith(code, std::string(call) + " (" + paths[k] + "): " + dxrv_last_error(c));
            };
            bool havePrev = false;
            uint32_t prev = 0;
            for (uint32_t k = s; k < numMeshes; k += numCtx)
            {
                std::unique_ptr<dxrv::ObjMesh> mesh;
                {
                    std::unique_lock<std::mutex> g(b.m);
                    b.loaded.wait(g, [&] { return b.stop || b.slots[k].ready; });
                    if (b.stop) break;
                    mesh = std::move(b.slots[k].mesh);
                    ++b.taken;
                }
                b.consumed.notify_all();
                // this context's previous grid leaves before the next voxelize overwrites it; the other contexts keep the GPU busy
                if (havePrev && hostGrids)
                {
                    const int rc = dxrv_fetch_grid(c, static_cast<char*>(hostGrids) + static_cast<size_t>(prev) * gridBytes, gridBytes, DXRV_FORMAT_BITS);
                    if (rc != DXRV_OK) { ctxFail(rc, "dxrv_fetch_grid", prev); return; }
                }
                havePrev = false;
                // bound = NULL: derived on the GPU as Voxelizer::Init derives it (Voxelizer.cpp:52-57), inside the build
                int rc = dxrv_build_bvh(c, mesh->vertices.data(), mesh->numVertices(), mesh->stride, mesh->indices.data(), mesh->numIndices(), nullptr);
                if (rc != DXRV_OK) { ctxFail(rc, "dxrv_build_bvh", k); return; }
                if (numTriangles) numTriangles[k] = mesh->numIndices() / 3;
                mesh.reset();                             // the host arrays were only borrowed for the call
                rc = dxrv_voxelize(c, N, mode, 0, N);
                if (rc != DXRV_OK) { ctxFail(rc, "dxrv_voxelize", k); return; }
                havePrev = true; prev = k;
            }
            if (havePrev && hostGrids)
            {
                const int rc = dxrv_fetch_grid(c, static_cast<char*>(hostGrids) + static_cast<size_t>(prev) * gridBytes, gridBytes, DXRV_FORMAT_BITS);
                if (rc != DXRV_OK) { ctxFail(rc, "dxrv_fetch_grid", prev); return; }
            }
            else
            {
                const int rc = dxrv_synchronize(c);
                if (rc != DXRV_OK) b.failWith(rc, std::string("dxrv_synchronize: ") + dxrv_last_error(c));
            }
        };

        std::vector<std::thread> threads;
        threads.reserve(loaderThreads + numCtx);
        try
        {
            for (uint32_t t = 0; t < loaderThreads; ++t) threads.emplace_back(loader);
            for (uint32_t s = 1; s < numCtx; ++s) threads.emplace_back(driver, s);
            driver(0);                                    // the caller drives the first context
        }
        catch (...) { b.failWith(DXRV_ERR_OOM, "cannot start the pipeline's threads"); }
        {
            std::lock_guard<std::mutex> g(b.m);           // drivers are done or failed: release loaders still waiting for room
            if (b.rc != DXRV_OK) b.stop = true;
        }
        b.consumed.notify_all();
        for (auto& t : threads) t.join();
        if (b.rc != DXRV_OK) dxrv::globalError() = "dxrv_voxelize_obj_batch: " + b.error;
        return b.rc;
    }
    catch (const std::bad_alloc&) { dxrv::globalError() = "dxrv_voxelize_obj_batch: out of memory"; return DXRV_ERR_OOM; }
    catch (const std::exception& e) { dxrv::globalError() = std::string("dxrv_voxelize_obj_batch: ") + e.what(); return DXRV_ERR_UNSUPPORTED; }
}
