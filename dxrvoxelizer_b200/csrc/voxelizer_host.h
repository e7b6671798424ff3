// voxelizer_host.h -- headless C++ face of the voxelization path, above the C ABI.
//
// Mirrors the slice of the reference's `Voxelizer` class that is on the hot path
// (Content/Voxelizer.h:10-24,90; Content/Voxelizer.cpp:30-79,351-369):
//   reference                                              here
//   ------------------------------------------------------ ------------------------------------
//   bool Init(cmdList, descTableLib, w, h, rtFmt, dsFmt,    bool Init(fileName, gridSize, posScale)
//             uploaders, pGeometry, fileName, posScale)       (device objects collapse into the ctx;
//                                                              gridSize replaces #define GRID_SIZE 64)
//   void voxelize(cmdList, frameIndex)   [protected]        bool Voxelize()
//   m_grids[frame] (R10G10B10A2 UAV)                        Grid() -> bit-packed occupancy
//   bool + XUSG_N_RETURN                                    bool + LastError()
// posScale only moves the volume in the reference's VIEWER (Voxelizer.cpp:84-87); it is accepted and
// kept so the Dragon.bat / TuringBowl.bat argument lists work unchanged, and does not touch the grid.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

struct dxrv_ctx;
struct dxrv_mesh;

class DXRVoxelizer
{
public:
    enum Mode : uint32_t { MODE_SHADER = 0, MODE_PARITY = 1 };

    DXRVoxelizer();
    ~DXRVoxelizer();
    DXRVoxelizer(const DXRVoxelizer&) = delete;
    DXRVoxelizer& operator=(const DXRVoxelizer&) = delete;

    void SetDevice(int cudaDevice) { m_device = cudaDevice; }
    // Shard the grid into z-slabs over `count` GPUs (devices m_device .. m_device+count-1), one context
    // each, mesh/BVH replicated.  Call before Init.  (New: the reference is single-GPU.)
    void SetGpuCount(int count) { m_gpus = count < 1 ? 1 : count; }
    // Default MODE_SHADER: what the reference's DispatchRays computes.  MODE_PARITY (column parity, a different
    // function: 17-69 voxels of 262 144 differ on the shipped meshes at 64^3) is an explicit opt-in.
    void SetMode(Mode mode) { m_mode = mode; }
    // z-slab [begin, end) computed by Voxelize(); end = 0 means the whole grid.
    void SetSlab(uint32_t begin, uint32_t end) { m_slabBegin = begin; m_slabEnd = end; }

    // Loads the OBJ exactly like ObjLoader::Import(fileName, true, true), derives the bound
    // (Voxelizer.cpp:52-57), uploads VB/IB and builds the acceleration structure.
    bool Init(const char* fileName, uint32_t gridSize, const float posScale[4] = nullptr);
    // Init from memory (vertices = interleaved {float3 pos; float3 nrm}, stride bytes); the arrays
    // must outlive this object (they are re-uploaded by BuildAccelerationStructures()).
    bool Init(const void* vertices, uint32_t numVerts, uint32_t stride, const uint32_t* indices,
              uint32_t numIndices, uint32_t gridSize);
    // Rebuild the acceleration structure from the mesh given to Init (the metric counts the build).
    bool BuildAccelerationStructures();
    // One DispatchRays(N, N*N, 1) worth of work.
    bool Voxelize();
    // Host copy of the slab (DXRV_FORMAT_BITS layout); fetched on demand, valid until the next Voxelize().
    const uint32_t* Grid();
    size_t GridWords() const;
    bool CountInside(uint64_t& count);
    // The reference's viewer pass (Render -> renderRayCast, Content/Voxelizer.cpp:371-399) into an RGBA8 image
    // with the reference's camera; single-GPU, full grid.
    bool RenderView(uint32_t width, uint32_t height, std::vector<uint8_t>& rgba);

    uint32_t GridSize() const { return m_gridSize; }
    uint32_t NumTriangles() const { return m_numIndices / 3; }
    const float* Bound() const { return m_bound; }
    const float* PosScale() const { return m_posScale; }
    dxrv_ctx* Context() const { return m_ctx; }
    const char* LastError() const { return m_error.c_str(); }

private:
    bool fail(const char* what);
    void computeSlabs(uint32_t begin, uint32_t end, int k);   // cost-balanced z-slab cut points of the k GPUs
    void slabOf(int g, uint32_t& z0, uint32_t& z1) const;

    dxrv_ctx* m_ctx = nullptr;            // context of the first GPU
    std::vector<dxrv_ctx*> m_more;        // contexts of GPUs 2..k
    std::vector<uint32_t> m_cuts;         // k + 1 slab cut points of the last Voxelize()
    int m_gpus = 1;
    dxrv_mesh* m_mesh = nullptr;
    const void* m_vertices = nullptr;
    const uint32_t* m_indices = nullptr;
    uint32_t m_numVerts = 0, m_stride = 0, m_numIndices = 0;
    int m_device = 0;
    Mode m_mode = MODE_SHADER;   // the reference's function (DXRVoxelizer.hlsl:58-85,132-140); MODE_PARITY is the opt-in fast path
    uint32_t m_gridSize = 64;  // GRID_SIZE, Voxelizer.cpp:8
    uint32_t m_slabBegin = 0, m_slabEnd = 0;
    float m_bound[4] = {0, 0, 0, 1};
    float m_posScale[4] = {0, 0, 0, 1};  // DXRVoxelizer.cpp:37
    std::vector<uint32_t> m_grid;
    bool m_gridFetched = false;
    bool m_commReady = false;   // NCCL communicators of the k contexts exist (dxrv_comm_init_all)
    bool m_fused = false;       // the last Voxelize() stored every slab into the first GPU's full grid
    std::string m_error;
};
