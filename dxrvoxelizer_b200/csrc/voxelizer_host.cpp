// voxelizer_host.cpp -- see voxelizer_host.h.  Thin: everything below goes through the C ABI.
#include "voxelizer_host.h"

#include "../../include/dxrv.h"

DXRVoxelizer::DXRVoxelizer() {}

DXRVoxelizer::~DXRVoxelizer()
{
    if (m_ctx) dxrv_destroy(m_ctx);
    if (m_mesh) dxrv_obj_free(m_mesh);
}

bool DXRVoxelizer::fail(const char* what)
{
    m_error = std::string(what) + ": " + dxrv_last_error(m_ctx);
    return false;
}

bool DXRVoxelizer::Init(const char* fileName, uint32_t gridSize, const float posScale[4])
{
    if (posScale) for (int i = 0; i < 4; ++i) m_posScale[i] = posScale[i];
    if (m_mesh) { dxrv_obj_free(m_mesh); m_mesh = nullptr; }
    // Load inputs (Voxelizer.cpp:45-49)
    if (dxrv_obj_load(fileName, &m_mesh) != DXRV_OK)
    {
        m_error = dxrv_last_error(nullptr);
        return false;
    }
    return Init(dxrv_obj_vertices(m_mesh), dxrv_obj_num_vertices(m_mesh), dxrv_obj_vertex_stride(m_mesh),
                dxrv_obj_indices(m_mesh), dxrv_obj_num_indices(m_mesh), gridSize);
}

bool DXRVoxelizer::Init(const void* vertices, uint32_t numVerts, uint32_t stride, const uint32_t* indices,
                        uint32_t numIndices, uint32_t gridSize)
{
    m_vertices = vertices; m_numVerts = numVerts; m_stride = stride;
    m_indices = indices; m_numIndices = numIndices;
    m_gridSize = gridSize;
    m_gridFetched = false;
    if (!m_ctx && dxrv_create(&m_ctx, m_device) != DXRV_OK)
    {
        m_error = dxrv_last_error(nullptr);
        return false;
    }
    return BuildAccelerationStructures();
}

bool DXRVoxelizer::BuildAccelerationStructures()
{
    if (!m_ctx || !m_vertices) { m_error = "Init has not been called"; return false; }
    // bound = NULL: extracted on the device as Voxelizer.cpp:52-57 does on the host
    if (dxrv_build_bvh(m_ctx, m_vertices, m_numVerts, m_stride, m_indices, m_numIndices, nullptr) != DXRV_OK)
        return fail("dxrv_build_bvh");
    if (dxrv_get_bound(m_ctx, m_bound) != DXRV_OK) return fail("dxrv_get_bound");
    return true;
}

bool DXRVoxelizer::Voxelize()
{
    if (!m_ctx) { m_error = "Init has not been called"; return false; }
    const uint32_t end = m_slabEnd ? m_slabEnd : m_gridSize;
    m_gridFetched = false;
    if (dxrv_voxelize(m_ctx, m_gridSize, m_mode, m_slabBegin, end) != DXRV_OK) return fail("dxrv_voxelize");
    return true;
}

size_t DXRVoxelizer::GridWords() const
{
    const uint32_t end = m_slabEnd ? m_slabEnd : m_gridSize;
    return static_cast<size_t>(end - m_slabBegin) * m_gridSize * ((m_gridSize + 31) / 32);
}

const uint32_t* DXRVoxelizer::Grid()
{
    if (!m_ctx) { m_error = "Init has not been called"; return nullptr; }
    if (!m_gridFetched)
    {
        m_grid.resize(GridWords());
        if (dxrv_fetch_grid(m_ctx, m_grid.data(), m_grid.size() * sizeof(uint32_t), DXRV_FORMAT_BITS) != DXRV_OK)
        {
            fail("dxrv_fetch_grid");
            return nullptr;
        }
        m_gridFetched = true;
    }
    return m_grid.data();
}

bool DXRVoxelizer::CountInside(uint64_t& count)
{
    if (!m_ctx) { m_error = "Init has not been called"; return false; }
    if (dxrv_count_inside(m_ctx, &count) != DXRV_OK) return fail("dxrv_count_inside");
    return true;
}
