// main.cpp -- headless CLI: the reference's `DXRVoxelizer.exe -mesh <path> [x y z scale]` without
// the window.  Argument grammar follows DXRVoxelizer::ParseCommandLineArgs
// (reference DXRVoxelizer.cpp:363-408): '-' or '/' prefixes, case-insensitive names, a value may
// start with '-' only when a digit or '.' follows; defaults Assets/bunny.obj and posScale (0,0,0,1)
// (DXRVoxelizer.cpp:36-37).  -warp / -uma select D3D adapters in the reference and are accepted and
// ignored.  New flags: -grid N (replaces #define GRID_SIZE 64), -mode shader|parity (default shader = the
// reference's function; parity = the column-parity fast path), -device k,
// -slab z0 z1, -frames n, -gpus k (z-slabs over k GPUs), -out file.bin (raw DXRV_FORMAT_BITS words),
// -view file.ppm|file.png (the reference's viewer pass, 1280 x 720; .png = its screenshot format), -batch list.txt [-streams k] (one grid per OBJ path
// of the list through dxrv_voxelize_obj_batch: k contexts per GPU, -gpus GPUs; -out prefix writes prefix00000.bin ...),
// -dryrun (print the parsed arguments and exit).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/dxrv.h"
#include "voxelizer_host.h"

namespace
{
std::string lower(std::string s)
{
    std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)std::tolower(c); });
    return s;
}

struct Args
{
    int argc;
    char** argv;
    bool matches(int i, const char* name) const
    {
        const char* a = argv[i];
        return (a[0] == '-' || a[0] == '/') && lower(a + 1) == lower(name);  // DXRVoxelizer.cpp:372-378
    }
    // One deviation from the reference: there every '/'-prefixed token is an option; on Linux that
    // would swallow absolute paths, so '/' only introduces an option when a known name follows.
    static bool knownOption(const char* name)
    {
        static const char* const names[] = {"warp", "uma", "mesh", "grid", "device", "frames", "out", "slab", "mode", "gpus", "view", "batch", "streams", "dryrun"};
        for (const char* n : names) if (lower(name) == n) return true;
        return false;
    }
    bool hasValue(int i) const
    {
        if (i + 1 >= argc) return false;
        const char* a = argv[i + 1];
        if (a[0] == '/') return !knownOption(a + 1);
        return a[0] != '-' || (a[1] >= '0' && a[1] <= '9') || a[1] == '.';
    }
};
// -batch: the streaming case.  Every path of the list file (one per line; blank lines and lines starting with '#'
// are skipped) becomes one grid; `streams` contexts on each of `gpus` GPUs share the work (mesh-parallel), the host
// grids of a round live in pinned memory (rounds of at most 2 GiB of grids).
int runBatch(const std::string& listPath, uint32_t grid, uint32_t mode, int device, int gpus, int streams, const std::string& outPrefix)
{
    FILE* lf = std::fopen(listPath.c_str(), "r");
    if (!lf) { std::fprintf(stderr, "cannot open %s\n", listPath.c_str()); return 1; }
    std::vector<std::string> files;
    char line[4096];
    while (std::fgets(line, sizeof line, lf))
    {
        std::string t = line;
        while (!t.empty() && (t.back() == '\n' || t.back() == '\r' || t.back() == ' ' || t.back() == '\t')) t.pop_back();
        size_t b = 0;
        while (b < t.size() && (t[b] == ' ' || t[b] == '\t')) ++b;
        if (b < t.size() && t[b] != '#') files.push_back(t.substr(b));
    }
    std::fclose(lf);
    if (files.empty()) { std::fprintf(stderr, "%s lists no meshes\n", listPath.c_str()); return 1; }
    if (grid == 0 || grid > 8192) { std::fprintf(stderr, "-grid must be in [1, 8192]\n"); return 2; }
    gpus = std::max(1, gpus); streams = std::max(1, std::min(streams, 64));

    std::vector<dxrv_ctx*> ctxs;
    auto cleanup = [&](void* pinned) { for (dxrv_ctx* c : ctxs) dxrv_destroy(c); if (pinned) dxrv_host_free(pinned); };
    for (int i = 0; i < gpus * streams; ++i)
    {
        dxrv_ctx* c = nullptr;
        if (dxrv_create(&c, device + i % gpus) != DXRV_OK) { std::fprintf(stderr, "Init failed: %s\n", dxrv_last_error(nullptr)); cleanup(nullptr); return 1; }
        ctxs.push_back(c);
    }
    const size_t words = (size_t)grid * grid * ((grid + 31) / 32), gridBytes = words * 4;
    const size_t perRound = std::min<size_t>(files.size(), std::max<size_t>(1, ((size_t)2 << 30) / gridBytes));
    void* pinned = dxrv_host_alloc(perRound * gridBytes);
    if (!pinned) { std::fprintf(stderr, "cannot allocate %zu bytes of pinned host memory\n", perRound * gridBytes); cleanup(nullptr); return 1; }

    using clock = std::chrono::steady_clock;
    unsigned long long inside = 0, triangles = 0;
    double seconds = 0.0;
    std::vector<uint32_t> tris(perRound);
    for (size_t first = 0; first < files.size(); first += perRound)
    {
        const size_t n = std::min(perRound, files.size() - first);
        std::vector<const char*> paths(n);
        for (size_t k = 0; k < n; ++k) paths[k] = files[first + k].c_str();
        const auto t0 = clock::now();
        const int rc = dxrv_voxelize_obj_batch(ctxs.data(), (uint32_t)ctxs.size(), paths.data(), (uint32_t)n, grid, mode, pinned, gridBytes, 0, tris.data());
        seconds += std::chrono::duration<double>(clock::now() - t0).count();
        if (rc != DXRV_OK) { std::fprintf(stderr, "Voxelize failed: %s\n", dxrv_last_error(nullptr)); cleanup(pinned); return 1; }
        for (size_t k = 0; k < n; ++k)
        {
            const uint32_t* g = static_cast<const uint32_t*>(pinned) + k * words;
            triangles += tris[k];
            for (size_t w = 0; w < words; ++w) inside += (unsigned long long)__builtin_popcount(g[w]);
            if (!outPrefix.empty())
            {
                char name[32];
                std::snprintf(name, sizeof name, "%05zu.bin", first + k);
                FILE* f = std::fopen((outPrefix + name).c_str(), "wb");
                if (!f) { std::fprintf(stderr, "cannot write %s%s\n", outPrefix.c_str(), name); cleanup(pinned); return 1; }
                std::fwrite(g, sizeof(uint32_t), words, f);
                std::fclose(f);
            }
        }
    }
    std::printf("{\"batch\": %zu, \"grid\": %u, \"mode\": \"%s\", \"gpus\": %d, \"contexts\": %zu, \"triangles\": %llu, \"inside\": %llu, "
                "\"seconds\": %.4f, \"meshes_per_s\": %.1f}\n",
                files.size(), grid, mode == DXRV_MODE_SHADER ? "shader" : "parity", gpus, ctxs.size(), triangles, inside, seconds,
                (double)files.size() / seconds);
    cleanup(pinned);
    return 0;
}
}  // namespace

int main(int argc, char** argv)
{
    std::string mesh = "Assets/bunny.obj", out, view, batch;
    int streams = 4;
    float posScale[4] = {0.0f, 0.0f, 0.0f, 1.0f};
    uint32_t grid = 64, slab0 = 0, slab1 = 0;
    int device = 0, frames = 1, gpus = 1;
    DXRVoxelizer::Mode mode = DXRVoxelizer::MODE_SHADER;   // Dragon.sh / TuringBowl.sh reproduce the reference's grid

    bool dryRun = false;
    Args a{argc, argv};
    for (int i = 1; i < argc; ++i)
    {
        if (a.matches(i, "warp") || a.matches(i, "uma")) continue;
        else if (a.matches(i, "dryrun")) dryRun = true;
        else if (a.matches(i, "mesh"))
        {
            if (a.hasValue(i)) mesh = argv[++i];
            for (int k = 0; k < 4; ++k)
                if (a.hasValue(i)) i += std::sscanf(argv[i + 1], "%f", &posScale[k]) == 1 ? 1 : 0;
        }
        else if (a.matches(i, "grid") && a.hasValue(i)) grid = (uint32_t)std::strtoul(argv[++i], nullptr, 10);
        else if (a.matches(i, "device") && a.hasValue(i)) device = std::atoi(argv[++i]);
        else if (a.matches(i, "frames") && a.hasValue(i)) frames = std::atoi(argv[++i]);
        else if (a.matches(i, "gpus") && a.hasValue(i)) gpus = std::atoi(argv[++i]);
        else if (a.matches(i, "out") && a.hasValue(i)) out = argv[++i];
        else if (a.matches(i, "view") && a.hasValue(i)) view = argv[++i];
        else if (a.matches(i, "batch") && a.hasValue(i)) batch = argv[++i];
        else if (a.matches(i, "streams") && a.hasValue(i)) streams = std::atoi(argv[++i]);
        else if (a.matches(i, "slab") && a.hasValue(i))
        {
            slab0 = (uint32_t)std::strtoul(argv[++i], nullptr, 10);
            if (a.hasValue(i)) slab1 = (uint32_t)std::strtoul(argv[++i], nullptr, 10);
        }
        else if (a.matches(i, "mode") && a.hasValue(i))
        {
            const std::string m = lower(argv[++i]);
            if (m == "shader") mode = DXRVoxelizer::MODE_SHADER;
            else if (m == "parity") mode = DXRVoxelizer::MODE_PARITY;
            else { std::fprintf(stderr, "unknown -mode %s (shader|parity)\n", m.c_str()); return 2; }
        }
    }

    if (dryRun)   // the parsed command line, nothing else (the argument grammar is part of the drop-in surface: tests/test_abi.py)
    {
        std::printf("{\"mesh\": \"%s\", \"posScale\": [%g, %g, %g, %g], \"grid\": %u, \"mode\": \"%s\", \"device\": %d, \"frames\": %d, "
                    "\"gpus\": %d, \"slab\": [%u, %u], \"out\": \"%s\", \"view\": \"%s\", \"batch\": \"%s\", \"streams\": %d}\n",
                    mesh.c_str(), posScale[0], posScale[1], posScale[2], posScale[3], grid, mode == DXRVoxelizer::MODE_SHADER ? "shader" : "parity",
                    device, frames, gpus, slab0, slab1, out.c_str(), view.c_str(), batch.c_str(), streams);
        return 0;
    }
    if (!batch.empty())
        return runBatch(batch, grid, mode == DXRVoxelizer::MODE_SHADER ? DXRV_MODE_SHADER : DXRV_MODE_PARITY, device, gpus, streams, out);

    DXRVoxelizer vox;
    vox.SetDevice(device);
    vox.SetGpuCount(gpus);
    vox.SetMode(mode);
    vox.SetSlab(slab0, slab1);
    using clock = std::chrono::steady_clock;
    auto t0 = clock::now();
    if (!vox.Init(mesh.c_str(), grid, posScale)) { std::fprintf(stderr, "Init failed: %s\n", vox.LastError()); return 1; }
    auto t1 = clock::now();
    uint64_t inside = 0;
    double best = 1e30;
    for (int f = 0; f < (frames < 1 ? 1 : frames); ++f)
    {
        auto s = clock::now();
        if (!vox.BuildAccelerationStructures() || !vox.Voxelize() || !vox.CountInside(inside))
        {
            std::fprintf(stderr, "Voxelize failed: %s\n", vox.LastError());
            return 1;
        }
        best = std::min(best, std::chrono::duration<double, std::milli>(clock::now() - s).count());
    }
    const double voxels = (double)vox.GridWords() * 32.0;
    std::printf("{\"mesh\": \"%s\", \"triangles\": %u, \"grid\": %u, \"mode\": \"%s\", \"bound\": [%g, %g, %g, %g], "
                "\"inside\": %llu, \"init_ms\": %.3f, \"build_plus_voxelize_ms\": %.4f, \"gvoxels_per_s\": %.2f}\n",
                mesh.c_str(), vox.NumTriangles(), grid, mode == DXRVoxelizer::MODE_SHADER ? "shader" : "parity",
                vox.Bound()[0], vox.Bound()[1], vox.Bound()[2], vox.Bound()[3], (unsigned long long)inside,
                std::chrono::duration<double, std::milli>(t1 - t0).count(), best, voxels / best * 1e-6);
    if (!out.empty())
    {
        const uint32_t* g = vox.Grid();
        if (!g) { std::fprintf(stderr, "fetch failed: %s\n", vox.LastError()); return 1; }
        FILE* f = std::fopen(out.c_str(), "wb");
        if (!f) { std::fprintf(stderr, "cannot write %s\n", out.c_str()); return 1; }
        std::fwrite(g, sizeof(uint32_t), vox.GridWords(), f);
        std::fclose(f);
    }
    if (!view.empty())
    {
        // what the reference's window shows (1280 x 720, Main.cpp:17), as a binary PPM
        std::vector<uint8_t> rgba;
        if (!vox.RenderView(1280, 720, rgba)) { std::fprintf(stderr, "view failed: %s\n", vox.LastError()); return 1; }
        const bool png = view.size() >= 4 && lower(view.substr(view.size() - 4)) == ".png";
        if (png)   // the reference's screenshot format (SaveImage, DXRVoxelizer.cpp:531-551: RGB PNG)
        {
            if (dxrv_save_image(view.c_str(), rgba.data(), 1280, 720, 1280 * 4, 3) != DXRV_OK) { std::fprintf(stderr, "cannot write %s\n", view.c_str()); return 1; }
        }
        else
        {
            FILE* f = std::fopen(view.c_str(), "wb");
            if (!f) { std::fprintf(stderr, "cannot write %s\n", view.c_str()); return 1; }
            std::fprintf(f, "P6\n1280 720\n255\n");
            for (size_t i = 0; i < rgba.size(); i += 4) std::fwrite(&rgba[i], 1, 3, f);
            std::fclose(f);
        }
    }
    return 0;
}
