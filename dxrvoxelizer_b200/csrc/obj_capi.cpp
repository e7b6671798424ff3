// obj_capi.cpp -- C ABI for the mesh-input stage (include/dxrv.h, "mesh input").
#include <new>
#include <string>

#include "../../include/dxrv.h"
#include "obj_loader.h"

struct dxrv_mesh
{
    dxrv::ObjMesh mesh;
};

namespace dxrv
{
// last error of the context-free entry points (dxrv_create, dxrv_obj_load)
std::string& globalError()
{
    static thread_local std::string e;
    return e;
}
}  // namespace dxrv

extern "C" {

int dxrv_obj_load(const char* path, dxrv_mesh** out)
{
    if (!path || !out) { dxrv::globalError() = "dxrv_obj_load: null argument"; return DXRV_ERR_INVALID_ARG; }
    *out = nullptr;
    dxrv_mesh* m = new (std::nothrow) dxrv_mesh();
    if (!m) { dxrv::globalError() = "dxrv_obj_load: out of memory"; return DXRV_ERR_OOM; }
    std::string err;
    bool ok = false;
    try { ok = dxrv::loadObj(path, m->mesh, err); }
    catch (const std::bad_alloc&) { delete m; dxrv::globalError() = "dxrv_obj_load: out of memory"; return DXRV_ERR_OOM; }
    catch (...) { err = "unexpected exception"; }
    if (!ok) { delete m; dxrv::globalError() = "dxrv_obj_load: " + err; return DXRV_ERR_IO; }
    *out = m;
    return DXRV_OK;
}

int dxrv_obj_parse(const char* text, size_t size, dxrv_mesh** out)
{
    if (!text || !out) { dxrv::globalError() = "dxrv_obj_parse: null argument"; return DXRV_ERR_INVALID_ARG; }
    *out = nullptr;
    dxrv_mesh* m = new (std::nothrow) dxrv_mesh();
    if (!m) { dxrv::globalError() = "dxrv_obj_parse: out of memory"; return DXRV_ERR_OOM; }
    std::string err;
    bool ok = false;
    try
    {
        // as loadObj: the multi-threaded parser for well-formed text, the reference's token grammar for everything else
        ok = dxrv::parseObjFast(text, size, m->mesh, err, 0);
        if (!ok && err.empty()) ok = dxrv::parseObj(text, size, m->mesh, err);
    }
    catch (const std::bad_alloc&) { delete m; dxrv::globalError() = "dxrv_obj_parse: out of memory"; return DXRV_ERR_OOM; }
    catch (...) { ok = false; err = "unexpected exception"; }
    if (!ok) { delete m; dxrv::globalError() = "dxrv_obj_parse: " + err; return DXRV_ERR_IO; }
    *out = m;
    return DXRV_OK;
}

void dxrv_obj_free(dxrv_mesh* mesh) { delete mesh; }
uint32_t dxrv_obj_num_vertices(const dxrv_mesh* mesh) { return mesh ? mesh->mesh.numVertices() : 0; }
uint32_t dxrv_obj_num_indices(const dxrv_mesh* mesh) { return mesh ? mesh->mesh.numIndices() : 0; }
uint32_t dxrv_obj_vertex_stride(const dxrv_mesh* mesh) { return mesh ? mesh->mesh.stride : 0; }
const void* dxrv_obj_vertices(const dxrv_mesh* mesh) { return mesh ? mesh->mesh.vertices.data() : nullptr; }
const uint32_t* dxrv_obj_indices(const dxrv_mesh* mesh) { return mesh ? mesh->mesh.indices.data() : nullptr; }
void dxrv_obj_aabb(const dxrv_mesh* mesh, float out[6])
{
    if (!mesh || !out) return;
    for (int a = 0; a < 3; ++a) { out[a] = mesh->mesh.aabbMin[a]; out[3 + a] = mesh->mesh.aabbMax[a]; }
}
void dxrv_obj_bound(const dxrv_mesh* mesh, float out[4])
{
    if (mesh && out) mesh->mesh.bound(out);
}

}
