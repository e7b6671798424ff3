// obj_loader.h -- Wavefront OBJ ingestion for the voxelizer (host side, no CUDA).
//
// Produces exactly what the reference's XUSG::ObjLoader::Import(file, /*needNorm*/true,
// /*needAABB*/true, /*forDX*/true, /*swapYZ*/false) produces
// (XUSG/Optional/XUSGObjLoader.cpp:18-40), because that output *is* the input contract of the
// voxelization path (Content/Voxelizer.cpp:46-57,269-270):
//   * interleaved vertices {float3 pos; float3 nrm} (stride 24; +8 zero bytes when the file has
//     `vt` records, XUSGObjLoader.cpp:160),
//   * z negated (XUSGObjLoader.cpp:198,213), polygon fans (:263-297), 1-based / negative indices
//     resolved against the file's TOTAL record counts (:238,243),
//   * the whole index array reversed (:227),
//   * per-corner normal split when the file has `vn` (:300-335) or unit face normals summed per
//     vertex and normalised when it has none (:337-384),
//   * AABB over all vertices (:386-416).
// The implementation is a single-buffer tokenizer (the reference runs two fscanf passes); the
// byte-for-byte equality is enforced by tests/test_obj_loader.py against the reference loader
// compiled from /root/reference into oracle/_ref.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace dxrv
{
struct ObjMesh
{
    std::vector<uint8_t> vertices;  // numVertices * stride bytes
    std::vector<uint32_t> indices;  // 3 * numTriangles
    uint32_t stride = 24;
    float aabbMin[3] = {0, 0, 0};
    float aabbMax[3] = {0, 0, 0};

    uint32_t numVertices() const { return stride ? static_cast<uint32_t>(vertices.size() / stride) : 0; }
    uint32_t numIndices() const { return static_cast<uint32_t>(indices.size()); }
    // {cx, cy, cz, w}: Content/Voxelizer.cpp:52-57
    void bound(float out[4]) const;
};

// Returns false (and fills err) when the file cannot be read -- the reference returns false
// from Import when fopen fails (XUSGObjLoader.cpp:21-23).
// threads: parser threads for this file (0: DXRV_OBJ_THREADS, else one per core up to 16; a batch loader that parses many
// files at once passes 1).
bool loadObj(const char* path, ObjMesh& mesh, std::string& err, unsigned threads = 0);
// Same, from a memory buffer holding the OBJ text (follows the reference's fscanf grammar token by token).
bool parseObj(const char* text, size_t size, ObjMesh& mesh, std::string& err);
// Multi-threaded parser for well-formed files; returns false with an empty `err` when the text needs the
// exact grammar of parseObj instead (loadObj falls back automatically).  threads = 0: one per core, <= 16.
bool parseObjFast(const char* text, size_t size, ObjMesh& mesh, std::string& err, unsigned threads);
}  // namespace dxrv
