// CPU harness for dxrv_voxelize_obj_batch (csrc/batch.cpp): the pipeline's threads, ordering, bounded look-ahead and
// error paths against STUBS of the context entry points (no GPU, no CUDA): a stub context "voxelizes" a mesh into a grid
// that carries a hash of the mesh's vertex and index bytes, so every grid can be matched to its file.  Built and run by
// tests/test_batch_cpu.py (also under -fsanitize=thread / address by hand: profiles/r02c_sanitizer.txt).
//   g++ -std=c++17 -O1 -g -Idxrvoxelizer_b200/csrc tools/batch_mock.cpp dxrvoxelizer_b200/csrc/batch.cpp \
//       dxrvoxelizer_b200/csrc/obj_loader.cpp -o /tmp/batch_mock -lpthread && /tmp/batch_mock /tmp/some_dir
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../include/dxrv.h"
#include "obj_loader.h"

struct dxrv_ctx
{
    uint64_t sig = 0, gridSig = 0;
    uint32_t N = 0;
    bool haveBvh = false, haveGrid = false;
    int failBuildAt = -1, builds = 0;
    std::atomic<int> inCall{0};       // a context is not thread-safe: two calls at once are a bug of the pipeline
    std::string err;
};

namespace dxrv
{
std::string& globalError() { static thread_local std::string e; return e; }
}
static std::atomic<int> g_overlap{0};
struct CallGuard
{
    dxrv_ctx* c;
    explicit CallGuard(dxrv_ctx* ctx) : c(ctx) { if (c->inCall.fetch_add(1) != 0) ++g_overlap; }
    ~CallGuard() { c->inCall.fetch_sub(1); }
};
static uint64_t fnv(const void* p, size_t n, uint64_t h)
{
    const unsigned char* b = static_cast<const unsigned char*>(p);
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}
static uint64_t meshSig(const void* v, uint32_t nv, uint32_t stride, const uint32_t* idx, uint32_t ni)
{
    return fnv(idx, (size_t)ni * 4, fnv(v, (size_t)nv * stride, 1469598103934665603ull));
}

extern "C" {
const char* dxrv_last_error(const dxrv_ctx* c) { return c ? c->err.c_str() : dxrv::globalError().c_str(); }
int dxrv_build_bvh(dxrv_ctx* c, const void* v, uint32_t nv, uint32_t stride, const uint32_t* idx, uint32_t ni, const float* bound)
{
    CallGuard g(c);
    if (bound) { c->err = "the batch passes bound = NULL"; return DXRV_ERR_INVALID_ARG; }
    if (c->builds++ == c->failBuildAt) { c->err = "injected build failure"; return DXRV_ERR_CUDA; }
    c->sig = meshSig(v, nv, stride, idx, ni);
    std::this_thread::sleep_for(std::chrono::microseconds(50 + (c->sig & 127)));
    c->haveBvh = true;
    return DXRV_OK;
}
int dxrv_voxelize(dxrv_ctx* c, uint32_t N, uint32_t, uint32_t z0, uint32_t z1)
{
    CallGuard g(c);
    if (!c->haveBvh || z0 != 0 || z1 != N) { c->err = "voxelize without build / not the whole grid"; return DXRV_ERR_NO_BVH; }
    c->gridSig = c->sig; c->N = N; c->haveGrid = true; c->haveBvh = false;   // one build per voxelize in this pipeline
    return DXRV_OK;
}
int dxrv_fetch_grid(dxrv_ctx* c, void* dst, size_t bytes, uint32_t format)
{
    CallGuard g(c);
    if (!c->haveGrid || format != DXRV_FORMAT_BITS || bytes < 8) { c->err = "fetch without grid"; return DXRV_ERR_NO_GRID; }
    std::memset(dst, (int)(c->gridSig & 0xff), bytes);
    std::memcpy(dst, &c->gridSig, 8);
    c->haveGrid = false;                                                     // every grid is fetched exactly once
    return DXRV_OK;
}
int dxrv_synchronize(dxrv_ctx* c) { CallGuard g(c); return DXRV_OK; }
}

static void writeObj(const std::string& path, int seed, int nv, int nf)
{
    FILE* f = std::fopen(path.c_str(), "w");
    for (int i = 0; i < nv; ++i) std::fprintf(f, "v %d.%06d %d.5 -%d.25\n", i, (seed * 7919 + i * 31) % 1000000, seed, i % 7);
    for (int i = 0; i < nf; ++i) std::fprintf(f, "f %d %d %d\n", 1 + (i + seed) % nv, 1 + (i * 3 + 1) % nv, 1 + (i * 5 + 2) % nv);
    std::fclose(f);
}

#define CHECK(x) do { if (!(x)) { std::printf("FAILED line %d: %s\n", __LINE__, #x); return 1; } } while (0)

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    const std::string dir = argv[1];
    const uint32_t M = 97, N = 64;
    const size_t gridBytes = (size_t)N * N * 2 * 4;
    std::vector<std::string> files;
    std::vector<uint64_t> want;
    std::vector<uint32_t> wantTris;
    for (uint32_t k = 0; k < M; ++k)
    {
        files.push_back(dir + "/m" + std::to_string(k) + ".obj");
        writeObj(files.back(), (int)k, 20 + (int)k * 3, 30 + (int)k * 5);
        dxrv::ObjMesh m; std::string e;
        CHECK(dxrv::loadObj(files.back().c_str(), m, e, 1));
        want.push_back(meshSig(m.vertices.data(), m.numVertices(), m.stride, m.indices.data(), m.numIndices()));
        wantTris.push_back(m.numIndices() / 3);
    }
    std::vector<const char*> paths;
    for (auto& s : files) paths.push_back(s.c_str());
    std::vector<unsigned char> grids(M * gridBytes);

    // every mesh lands in its own grid, whatever the numbers of contexts and loader threads
    for (uint32_t numCtx : {1u, 3u, 4u, 8u})
        for (uint32_t loaders : {1u, 2u, 7u, 0u})
        {
            std::vector<dxrv_ctx> ctx(numCtx);
            std::vector<dxrv_ctx*> cp;
            for (auto& c : ctx) cp.push_back(&c);
            std::fill(grids.begin(), grids.end(), 0);
            std::vector<uint32_t> tris(M, 0);
            const int rc = dxrv_voxelize_obj_batch(cp.data(), numCtx, paths.data(), M, N, DXRV_MODE_PARITY, grids.data(), gridBytes, loaders, tris.data());
            CHECK(rc == DXRV_OK);
            for (uint32_t k = 0; k < M; ++k)
            {
                uint64_t got; std::memcpy(&got, &grids[k * gridBytes], 8);
                CHECK(got == want[k]);
                CHECK(grids[k * gridBytes + gridBytes - 1] == (unsigned char)(want[k] & 0xff));
                CHECK(tris[k] == wantTris[k]);
            }
            int builds = 0;
            for (auto& c : ctx) { builds += c.builds; CHECK(!c.haveGrid); }      // every grid was fetched
            CHECK(builds == (int)M);
        }
    // without a host buffer: voxelize only, the contexts end synchronised and keep their last grid
    {
        std::vector<dxrv_ctx> ctx(4);
        std::vector<dxrv_ctx*> cp;
        for (auto& c : ctx) cp.push_back(&c);
        CHECK(dxrv_voxelize_obj_batch(cp.data(), 4, paths.data(), M, N, DXRV_MODE_PARITY, nullptr, 0, 3, nullptr) == DXRV_OK);
        for (uint32_t s = 0; s < 4; ++s) { const uint32_t last = ((M - 1 - s) / 4) * 4 + s; CHECK(ctx[s].haveGrid && ctx[s].gridSig == want[last]); }
    }
    // fewer meshes than contexts; an empty batch
    {
        std::vector<dxrv_ctx> ctx(8);
        std::vector<dxrv_ctx*> cp;
        for (auto& c : ctx) cp.push_back(&c);
        CHECK(dxrv_voxelize_obj_batch(cp.data(), 8, paths.data(), 3, N, DXRV_MODE_PARITY, grids.data(), gridBytes, 0, nullptr) == DXRV_OK);
        uint64_t got; std::memcpy(&got, &grids[2 * gridBytes], 8);
        CHECK(got == want[2]);
        CHECK(dxrv_voxelize_obj_batch(cp.data(), 8, paths.data(), 0, N, DXRV_MODE_PARITY, grids.data(), gridBytes, 0, nullptr) == DXRV_OK);
    }
    // a file that does not exist: DXRV_ERR_IO, the message names it, nothing hangs
    {
        std::vector<dxrv_ctx> ctx(4);
        std::vector<dxrv_ctx*> cp;
        for (auto& c : ctx) cp.push_back(&c);
        std::vector<const char*> bad = paths;
        const std::string missing = dir + "/missing.obj";
        bad[40] = missing.c_str();
        CHECK(dxrv_voxelize_obj_batch(cp.data(), 4, bad.data(), M, N, DXRV_MODE_PARITY, grids.data(), gridBytes, 5, nullptr) == DXRV_ERR_IO);
        CHECK(std::strstr(dxrv_last_error(nullptr), "missing.obj") != nullptr);
    }
    // a context that fails: its code comes back with the context's message
    {
        std::vector<dxrv_ctx> ctx(4);
        std::vector<dxrv_ctx*> cp;
        for (auto& c : ctx) cp.push_back(&c);
        ctx[2].failBuildAt = 5;
        CHECK(dxrv_voxelize_obj_batch(cp.data(), 4, paths.data(), M, N, DXRV_MODE_PARITY, grids.data(), gridBytes, 2, nullptr) == DXRV_ERR_CUDA);
        CHECK(std::strstr(dxrv_last_error(nullptr), "injected build failure") != nullptr);
    }
    // argument checks
    {
        dxrv_ctx c; dxrv_ctx* one[2] = {&c, &c};
        CHECK(dxrv_voxelize_obj_batch(nullptr, 1, paths.data(), 1, N, 1, nullptr, 0, 0, nullptr) == DXRV_ERR_INVALID_ARG);
        CHECK(dxrv_voxelize_obj_batch(one, 0, paths.data(), 1, N, 1, nullptr, 0, 0, nullptr) == DXRV_ERR_INVALID_ARG);
        CHECK(dxrv_voxelize_obj_batch(one, 2, paths.data(), 1, N, 1, nullptr, 0, 0, nullptr) == DXRV_ERR_INVALID_ARG);   // the same context twice
        CHECK(dxrv_voxelize_obj_batch(one, 1, paths.data(), 1, N, 1, grids.data(), gridBytes - 4, 0, nullptr) == DXRV_ERR_INVALID_ARG);
        CHECK(dxrv_voxelize_obj_batch(one, 1, paths.data(), 1, 0, 1, nullptr, 0, 0, nullptr) == DXRV_ERR_INVALID_ARG);
        CHECK(dxrv_voxelize_obj_batch(one, 1, nullptr, 1, N, 1, nullptr, 0, 0, nullptr) == DXRV_ERR_INVALID_ARG);
        CHECK(dxrv_voxelize_obj_batch(one, 1, paths.data(), 1, N, DXRV_MODE_SHADER | DXRV_EMIT_TEXELS, nullptr, 0, 0, nullptr) == DXRV_ERR_INVALID_ARG);
        CHECK(dxrv_voxelize_obj_batch(one, 1, paths.data(), 1, N, 7, nullptr, 0, 0, nullptr) == DXRV_ERR_INVALID_ARG);
    }
    CHECK(g_overlap.load() == 0);
    std::printf("batch pipeline: all checks passed\n");
    return 0;
}
