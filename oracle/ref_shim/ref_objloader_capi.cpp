/* TEST INFRASTRUCTURE ONLY -- C wrapper around the UNMODIFIED reference ObjLoader
 * (XUSG/Optional/XUSGObjLoader.{h,cpp}), built into oracle/_ref/libref_objloader.so by
 * oracle/Makefile.  Used by tests/ to prove the product loader (csrc/obj_loader.cpp) is
 * byte-identical, mirroring the call Voxelizer::Init makes:
 *   objLoader.Import(fileName, true, true)      -- Content/Voxelizer.cpp:46-47
 */
#include "XUSGObjLoader.h"

extern "C" {

struct RefMesh
{
    XUSG::ObjLoader* loader;
};

void* ref_obj_import(const char* path)
{
    auto* l = new XUSG::ObjLoader();
    if (!l->Import(path, true, true)) { delete l; return nullptr; }
    return l;
}

uint32_t ref_obj_num_vertices(void* h) { return static_cast<XUSG::ObjLoader*>(h)->GetNumVertices(); }
uint32_t ref_obj_num_indices(void* h) { return static_cast<XUSG::ObjLoader*>(h)->GetNumIndices(); }
uint32_t ref_obj_stride(void* h) { return static_cast<XUSG::ObjLoader*>(h)->GetVertexStride(); }
const uint8_t* ref_obj_vertices(void* h) { return static_cast<XUSG::ObjLoader*>(h)->GetVertices(); }
const uint32_t* ref_obj_indices(void* h) { return static_cast<XUSG::ObjLoader*>(h)->GetIndices(); }
void ref_obj_aabb(void* h, float out[6])
{
    const auto& a = static_cast<XUSG::ObjLoader*>(h)->GetAABB();
    out[0] = a.Min.x; out[1] = a.Min.y; out[2] = a.Min.z;
    out[3] = a.Max.x; out[4] = a.Max.y; out[5] = a.Max.z;
}
void ref_obj_free(void* h) { delete static_cast<XUSG::ObjLoader*>(h); }

}
