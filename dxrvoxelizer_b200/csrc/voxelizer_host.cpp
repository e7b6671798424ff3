// voxelizer_host.cpp -- see voxelizer_host.h.  Thin: everything below goes through the C ABI.
#include "voxelizer_host.h"

#include "../../include/dxrv.h"

#include <algorithm>
#include <cmath>
#include <cstring>

DXRVoxelizer::DXRVoxelizer() {}

DXRVoxelizer::~DXRVoxelizer()
{
    if (m_ctx) dxrv_destroy(m_ctx);
    for (dxrv_ctx* c : m_more) dxrv_destroy(c);
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
    while (static_cast<int>(m_more.size()) + 1 < m_gpus)
    {
        dxrv_ctx* c = nullptr;
        if (dxrv_create(&c, m_device + 1 + static_cast<int>(m_more.size())) != DXRV_OK)
        {
            m_error = dxrv_last_error(nullptr);
            return false;
        }
        m_more.push_back(c);
    }
    return BuildAccelerationStructures();
}

bool DXRVoxelizer::BuildAccelerationStructures()
{
    if (!m_ctx || !m_vertices) { m_error = "Init has not been called"; return false; }
    // bound = NULL: extracted on the device as Voxelizer.cpp:52-57 does on the host
    if (m_more.empty())
    {
        if (dxrv_build_bvh(m_ctx, m_vertices, m_numVerts, m_stride, m_indices, m_numIndices, nullptr) != DXRV_OK)
            return fail("dxrv_build_bvh");
    }
    else
    {
        // One upload to the first GPU, NCCL broadcast to the others over NVLink (one group: this thread drives all
        // communicators), then every GPU builds the identical tree -- the builds are asynchronous and run concurrently.
        if (!m_commReady)
        {
            std::vector<dxrv_ctx*> all{m_ctx};
            all.insert(all.end(), m_more.begin(), m_more.end());
            if (dxrv_comm_init_all(all.data(), static_cast<int>(all.size())) != DXRV_OK) return fail("dxrv_comm_init_all");
            m_commReady = true;
        }
        if (dxrv_group_begin() != DXRV_OK) return fail("dxrv_group_begin");
        bool ok = dxrv_bcast_mesh(m_ctx, m_vertices, m_numVerts, m_stride, m_indices, m_numIndices, 0) == DXRV_OK;
        for (dxrv_ctx* c : m_more)
            ok = ok && dxrv_bcast_mesh(c, nullptr, m_numVerts, m_stride, nullptr, m_numIndices, 0) == DXRV_OK;
        if (dxrv_group_end() != DXRV_OK || !ok) return fail("dxrv_bcast_mesh");
        if (dxrv_build_bvh_replicated(m_ctx, nullptr) != DXRV_OK) return fail("dxrv_build_bvh_replicated");
        for (dxrv_ctx* c : m_more)
            if (dxrv_build_bvh_replicated(c, nullptr) != DXRV_OK)
            {
                m_error = std::string("dxrv_build_bvh_replicated: ") + dxrv_last_error(c);
                return false;
            }
    }
    if (dxrv_get_bound(m_ctx, m_bound) != DXRV_OK) return fail("dxrv_get_bound");
    return true;
}

// Cut points of the k z-slabs over [begin, end): contiguous, disjoint, balanced by cost.  Equal slabs are only
// balanced for a mesh that fills the grid evenly; the cost of layer z is modelled as 1 (its stores) +
// 1.2 * t(z) / mean(t), t(z) = triangles whose z extent overlaps the layer (the same model, measured on a
// B200, as dxrvoxelizer_b200/sharding.py balanced_slabs), and the cuts split the cumulative cost evenly.
void DXRVoxelizer::computeSlabs(uint32_t begin, uint32_t end, int k)
{
    m_cuts.assign(static_cast<size_t>(k) + 1, end);
    m_cuts[0] = begin;
    const uint32_t N = m_gridSize, layers = end - begin;
    if (k <= 1 || layers == 0) return;
    std::vector<double> cost(layers, 1.0);
    const uint32_t numTris = m_numIndices / 3;
    if (numTris && m_vertices && m_stride >= 12)
    {
        std::vector<double> diff(static_cast<size_t>(N) + 1, 0.0);
        const uint8_t* vb = static_cast<const uint8_t*>(m_vertices);
        for (uint32_t t = 0; t < numTris; ++t)
        {
            double zmin = 1e300, zmax = -1e300;
            for (int c = 0; c < 3; ++c)
            {
                const uint32_t vi = m_indices[3 * static_cast<size_t>(t) + c];
                if (vi >= m_numVerts) continue;
                float z;
                std::memcpy(&z, vb + static_cast<size_t>(m_stride) * vi + 8, sizeof z);
                const double zs = (static_cast<double>(z) - m_bound[2]) / m_bound[3];
                zmin = std::min(zmin, zs); zmax = std::max(zmax, zs);
            }
            if (zmin > zmax) continue;
            const double lo = std::floor((zmin + 1.0) * 0.5 * N - 0.5), hi = std::ceil((zmax + 1.0) * 0.5 * N - 0.5);
            const uint32_t l = static_cast<uint32_t>(std::min(std::max(lo, 0.0), N - 1.0)), h = static_cast<uint32_t>(std::min(std::max(hi, 0.0), N - 1.0));
            diff[l] += 1.0; diff[h + 1] -= 1.0;
        }
        std::vector<double> t(N);
        double run = 0.0, total = 0.0;
        for (uint32_t z = 0; z < N; ++z) { run += diff[z]; t[z] = run; total += run; }
        if (total > 0.0)
            for (uint32_t z = begin; z < end; ++z) cost[z - begin] += 1.2 * t[z] / (total / N);
    }
    double sum = 0.0;
    for (double c : cost) sum += c;
    double acc = 0.0;
    uint32_t z = 0;
    for (int g = 1; g < k; ++g)
    {
        const double target = sum * g / k;
        while (z < layers && acc + cost[z] <= target) acc += cost[z++];
        uint32_t cut = begin + z;
        cut = std::max(cut, m_cuts[g - 1] + (layers >= static_cast<uint32_t>(k) ? 1u : 0u));   // every GPU keeps a layer when there are enough
        cut = std::min(cut, end - std::min<uint32_t>(layers, static_cast<uint32_t>(k - g)));
        m_cuts[g] = std::max(cut, m_cuts[g - 1]);
    }
}

void DXRVoxelizer::slabOf(int g, uint32_t& z0, uint32_t& z1) const
{
    if (static_cast<size_t>(g) + 1 >= m_cuts.size()) { z0 = z1 = 0; return; }   // no Voxelize() yet
    z0 = m_cuts[static_cast<size_t>(g)];
    z1 = m_cuts[static_cast<size_t>(g) + 1];
}

bool DXRVoxelizer::Voxelize()
{
    if (!m_ctx) { m_error = "Init has not been called"; return false; }
    const uint32_t end = m_slabEnd ? m_slabEnd : m_gridSize;
    m_gridFetched = false;
    const int k = 1 + static_cast<int>(m_more.size());
    computeSlabs(m_slabBegin, end, k);
    // several GPUs, whole grid: every GPU's fill kernel stores its slab straight into the first GPU's full grid
    // (peer access over NVLink), so Grid() is ONE device-to-host copy; slab offsets must be 16-byte aligned
    m_fused = k > 1 && m_slabBegin == 0 && end == m_gridSize;
    const size_t layerBytes = static_cast<size_t>(m_gridSize) * ((m_gridSize + 31) / 32) * sizeof(uint32_t);
    for (int g = 0; g < k && m_fused; ++g)
    {
        uint32_t z0, z1;
        slabOf(g, z0, z1);
        if (z1 > z0 && (layerBytes * z0) % 16) m_fused = false;
    }
    for (int g = 0; g < k; ++g)
    {
        uint32_t z0, z1;
        slabOf(g, z0, z1);
        if (z0 == z1) continue;
        dxrv_ctx* c = g ? m_more[g - 1] : m_ctx;
        if (m_fused && dxrv_share_grid_target(c, m_ctx, m_gridSize, z0, z1) != DXRV_OK)
        {
            m_error = std::string("dxrv_share_grid_target: ") + dxrv_last_error(c);
            return false;
        }
        if (!m_fused && k > 1) dxrv_set_grid_target(c, nullptr, 0);
        if (dxrv_voxelize(c, m_gridSize, m_mode, z0, z1) != DXRV_OK)
        {
            m_error = std::string("dxrv_voxelize: ") + dxrv_last_error(c);
            return false;
        }
    }
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
        const size_t wordsPerLayer = static_cast<size_t>(m_gridSize) * ((m_gridSize + 31) / 32);
        const int k = 1 + static_cast<int>(m_more.size());
        if (m_fused)
        {
            for (dxrv_ctx* c : m_more)
                if (dxrv_synchronize(c) != DXRV_OK) { m_error = std::string("dxrv_synchronize: ") + dxrv_last_error(c); return nullptr; }
            if (dxrv_synchronize(m_ctx) != DXRV_OK) { fail("dxrv_synchronize"); return nullptr; }
            if (dxrv_fetch_full_grid(m_ctx, m_grid.data(), m_grid.size() * sizeof(uint32_t)) != DXRV_OK) { fail("dxrv_fetch_full_grid"); return nullptr; }
            m_gridFetched = true;
            return m_grid.data();
        }
        for (int g = 0; g < k; ++g)   // gather: every GPU's slab lands at its offset of the host grid
        {
            uint32_t z0, z1;
            slabOf(g, z0, z1);
            if (z0 == z1) continue;
            dxrv_ctx* c = g ? m_more[g - 1] : m_ctx;
            if (dxrv_fetch_grid(c, m_grid.data() + (z0 - m_slabBegin) * wordsPerLayer, (z1 - z0) * wordsPerLayer * sizeof(uint32_t),
                                DXRV_FORMAT_BITS) != DXRV_OK)
            {
                m_error = std::string("dxrv_fetch_grid: ") + dxrv_last_error(c);
                return nullptr;
            }
        }
        m_gridFetched = true;
    }
    return m_grid.data();
}

bool DXRVoxelizer::CountInside(uint64_t& count)
{
    if (!m_ctx) { m_error = "Init has not been called"; return false; }
    count = 0;
    const int k = 1 + static_cast<int>(m_more.size());
    for (int g = 0; g < k; ++g)
    {
        uint32_t z0, z1;
        slabOf(g, z0, z1);
        if (z0 == z1) continue;
        dxrv_ctx* c = g ? m_more[g - 1] : m_ctx;
        uint64_t part = 0;
        if (dxrv_count_inside(c, &part) != DXRV_OK)
        {
            m_error = std::string("dxrv_count_inside: ") + dxrv_last_error(c);
            return false;
        }
        count += part;
    }
    return true;
}

bool DXRVoxelizer::RenderView(uint32_t width, uint32_t height, std::vector<uint8_t>& rgba)
{
    if (!m_ctx) { m_error = "Init has not been called"; return false; }
    if (!m_more.empty()) { m_error = "RenderView needs the whole grid on one GPU"; return false; }
    float screenToLocal[16], eye[3], light[3];
    if (dxrv_default_view(m_bound, m_posScale, width, height, screenToLocal, eye, light) != DXRV_OK)
    {
        m_error = "dxrv_default_view: invalid arguments";
        return false;
    }
    rgba.resize(static_cast<size_t>(width) * height * 4);
    if (dxrv_render_view(m_ctx, width, height, screenToLocal, eye, light, rgba.data(), rgba.size()) != DXRV_OK)
        return fail("dxrv_render_view");
    return true;
}
