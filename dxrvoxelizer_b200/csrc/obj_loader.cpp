// obj_loader.cpp -- see obj_loader.h.  Compile with -ffp-contract=off: the normal arithmetic
// below must round exactly like the reference's scalar code (mul, add, sqrt, div; no FMA).
#include "obj_loader.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace dxrv
{
namespace
{
// Cursor over the OBJ text with the three scanning primitives the format needs.  They follow
// ISO fscanf semantics for "%s", "%f", "%lld" and a literal '/', because the reference's
// grammar is *defined* by its fscanf calls (e.g. a face continues across newlines for as long
// as the next token parses as an integer, XUSGObjLoader.cpp:258).
class Cursor
{
public:
    Cursor(const char* begin, const char* end) : p_(begin), end_(end) {}

    static bool isSpace(char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

    void skipSpace()
    {
        while (p_ < end_ && isSpace(*p_)) ++p_;
    }

    // "%s": next whitespace-delimited token; false at end of input.
    bool token(const char*& tok, size_t& len)
    {
        skipSpace();
        if (p_ >= end_) return false;
        tok = p_;
        while (p_ < end_ && !isSpace(*p_)) ++p_;
        len = static_cast<size_t>(p_ - tok);
        return true;
    }

    // fgets(buf, 256, f): consume up to 255 characters, stopping after a newline.
    void restOfLine()
    {
        for (int n = 0; n < 255 && p_ < end_; ++n)
            if (*p_++ == '\n') break;
    }

    // "%f" (the buffer is NUL-terminated, so strtof cannot run past end_).
    bool real(float& out)
    {
        skipSpace();
        if (p_ >= end_) return false;
        char* stop = nullptr;
        const float v = std::strtof(p_, &stop);
        if (stop == p_) return false;
        p_ = stop;
        out = v;
        return true;
    }

    // "%lld"
    bool integer(long long& out)
    {
        skipSpace();
        const char* q = p_;
        bool neg = false;
        if (q < end_ && (*q == '-' || *q == '+')) neg = (*q++ == '-');
        if (q >= end_ || *q < '0' || *q > '9') return false;
        unsigned long long v = 0;
        while (q < end_ && *q >= '0' && *q <= '9') v = v * 10 + static_cast<unsigned>(*q++ - '0');
        out = neg ? -static_cast<long long>(v) : static_cast<long long>(v);
        p_ = q;
        return true;
    }

    // a literal character in a scanf format: consumed only when it matches.
    bool literal(char c)
    {
        if (p_ < end_ && *p_ == c) { ++p_; return true; }
        return false;
    }

private:
    const char* p_;
    const char* end_;
};

struct Counts
{
    uint32_t positions = 0, texcoords = 0, normals = 0, triangles = 0;
};

// One corner reference "v", "v/vt", "v//vn" or "v/vt/vn".  Which separators are consumed is
// decided by whether the FILE has vt / vn records, not by the token (XUSGObjLoader.cpp:245-257);
// a failed sub-scan leaves `scratch` at its previous value, as the reference's `vi` does.
struct CornerReader
{
    bool hasTexc, hasNorm;
    uint32_t numPos, numTexc, numNorm;
    long long scratch = 0;

    static uint32_t resolve(long long i, uint32_t count)
    {
        return static_cast<uint32_t>(i < 0 ? i + count : i - 1);
    }

    // returns false when no leading integer is present (end of the face)
    bool read(Cursor& c, uint32_t& v, uint32_t& vt, uint32_t& vn, bool mustExist)
    {
        const bool got = c.integer(scratch);
        if (!got && !mustExist) return false;
        v = resolve(scratch, numPos);
        if (hasTexc)
        {
            if (c.literal('/')) c.integer(scratch);
            vt = resolve(scratch, numTexc);
        }
        else if (hasNorm) c.literal('/');
        if (hasNorm)
        {
            if (c.literal('/')) c.integer(scratch);
            vn = resolve(scratch, numNorm);
        }
        return true;
    }
};

template <bool kStore>
void scanFaces(Cursor& c, CornerReader& r, uint32_t& numTri, std::vector<uint32_t>* pIdx,
               std::vector<uint32_t>* pNrmIdx)
{
    uint32_t v[3] = {0, 0, 0}, t[3] = {0, 0, 0}, n[3] = {0, 0, 0};
    for (int i = 0; i < 3; ++i) r.read(c, v[i], t[i], n[i], true);
    auto emit = [&]() {
        if (kStore)
        {
            pIdx->insert(pIdx->end(), v, v + 3);
            if (r.hasNorm) pNrmIdx->insert(pNrmIdx->end(), n, n + 3);
        }
        ++numTri;
    };
    emit();
    // triangle fan: (v0, previous, new)
    uint32_t nv, nt = 0, nn = 0;
    while (r.read(c, nv, nt, nn, false))
    {
        v[1] = v[2]; n[1] = n[2];
        v[2] = nv;   n[2] = nn;
        emit();
    }
}

Counts countRecords(const char* text, size_t size)
{
    Counts k;
    Cursor c(text, text + size);
    const char* tok; size_t len;
    // Corner syntax does not matter for counting: treat every separator as optional.
    CornerReader r{true, true, 0, 0, 0};
    while (c.token(tok, len))
    {
        if (tok[0] == 'f') scanFaces<false>(c, r, k.triangles, nullptr, nullptr);
        else if (tok[0] == 'v')
        {
            const char kind = len > 1 ? tok[1] : '\0';
            if (kind == '\0') { ++k.positions; c.restOfLine(); }
            else if (kind == 't') { ++k.texcoords; c.restOfLine(); }
            else if (kind == 'n') { ++k.normals; c.restOfLine(); }
        }
        else c.restOfLine();
    }
    return k;
}

inline float* posOf(std::vector<uint8_t>& vb, uint32_t stride, uint32_t i)
{
    return reinterpret_cast<float*>(vb.data() + static_cast<size_t>(stride) * i);
}
inline float* nrmOf(std::vector<uint8_t>& vb, uint32_t stride, uint32_t i) { return posOf(vb, stride, i) + 3; }

inline void normalize3(float n[3])
{
    const float l = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    n[0] /= l; n[1] /= l; n[2] /= l;
}

// File supplied normals: the first (v, vn) pairing keeps the vertex, any later different vn for
// the same v appends a copy of the vertex (XUSGObjLoader.cpp:300-335).
void assignFileNormals(ObjMesh& m, const std::vector<float>& normals, const std::vector<uint32_t>& nrmIdx)
{
    if (normals.empty()) return;
    const uint32_t stride = m.stride;
    std::vector<uint32_t> owner(m.numVertices(), UINT32_MAX);
    for (size_t i = 0; i < m.indices.size(); ++i)
    {
        const uint32_t src = m.indices[i];
        uint32_t dst = src;
        if (owner[src] == nrmIdx[i]) continue;
        if (owner[src] != UINT32_MAX)
        {
            dst = m.numVertices();
            m.vertices.resize(m.vertices.size() + stride);
            std::memcpy(posOf(m.vertices, stride, dst), posOf(m.vertices, stride, src), stride);
            m.indices[i] = dst;
        }
        else owner[src] = nrmIdx[i];

        float n[3] = {normals[3 * size_t(nrmIdx[i])], normals[3 * size_t(nrmIdx[i]) + 1],
                      normals[3 * size_t(nrmIdx[i]) + 2]};
        normalize3(n);
        std::memcpy(nrmOf(m.vertices, stride, dst), n, sizeof(n));
    }
    m.vertices.shrink_to_fit();
}

// No normals in the file: unit (not area weighted) face normals, summed per vertex in triangle
// order, then normalised (XUSGObjLoader.cpp:337-384).  Runs on the already reversed indices.
void faceNormals(ObjMesh& m)
{
    const uint32_t stride = m.stride;
    const size_t numTri = m.indices.size() / 3;
    for (size_t k = 0; k < numTri; ++k)
    {
        const uint32_t* tri = &m.indices[3 * k];
        const float* p0 = posOf(m.vertices, stride, tri[0]);
        const float* p1 = posOf(m.vertices, stride, tri[1]);
        const float* p2 = posOf(m.vertices, stride, tri[2]);
        const float e1[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
        const float e2[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
        float n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2],
                      e1[0] * e2[1] - e1[1] * e2[0]};
        normalize3(n);
        for (int c = 0; c < 3; ++c)
        {
            float* dst = nrmOf(m.vertices, stride, tri[c]);
            dst[0] += n[0]; dst[1] += n[1]; dst[2] += n[2];
        }
    }
    const uint32_t numVert = m.numVertices();
    for (uint32_t i = 0; i < numVert; ++i) normalize3(nrmOf(m.vertices, stride, i));
}

void boundingBox(ObjMesh& m)
{
    const uint32_t numVert = m.numVertices();
    if (!numVert) return;
    const float* p = posOf(m.vertices, m.stride, 0);
    for (int a = 0; a < 3; ++a) m.aabbMin[a] = m.aabbMax[a] = p[a];
    for (uint32_t i = 1; i < numVert; ++i)
    {
        p = posOf(m.vertices, m.stride, i);
        for (int a = 0; a < 3; ++a)
        {
            if (p[a] < m.aabbMin[a]) m.aabbMin[a] = p[a];
            else if (p[a] > m.aabbMax[a]) m.aabbMax[a] = p[a];
        }
    }
}
}  // namespace

void ObjMesh::bound(float out[4]) const
{
    const float ex = aabbMax[0] - aabbMin[0], ey = aabbMax[1] - aabbMin[1], ez = aabbMax[2] - aabbMin[2];
    out[0] = (aabbMax[0] + aabbMin[0]) / 2.0f;
    out[1] = (aabbMax[1] + aabbMin[1]) / 2.0f;
    out[2] = (aabbMax[2] + aabbMin[2]) / 2.0f;
    out[3] = std::max(ex, std::max(ey, ez)) / 2.0f;
}

bool parseObj(const char* text, size_t size, ObjMesh& m, std::string& err)
{
    (void)err;
    const Counts k = countRecords(text, size);

    m = ObjMesh();
    m.stride = 24 + (k.texcoords ? 8 : 0);
    m.vertices.assign(static_cast<size_t>(m.stride) * k.positions, 0);
    m.indices.reserve(3 * static_cast<size_t>(k.triangles));

    std::vector<float> normals;          // xyz per `vn`
    std::vector<uint32_t> nrmIdx;        // per corner, parallel to m.indices
    normals.reserve(3 * static_cast<size_t>(k.normals));
    if (k.normals) nrmIdx.reserve(3 * static_cast<size_t>(k.triangles));

    CornerReader reader{k.texcoords != 0, k.normals != 0, k.positions, k.texcoords, k.normals};
    Cursor c(text, text + size);
    const char* tok; size_t len;
    uint32_t numPos = 0, numTri = 0;
    while (c.token(tok, len))
    {
        if (tok[0] == 'f') scanFaces<true>(c, reader, numTri, &m.indices, &nrmIdx);
        else if (tok[0] == 'v')
        {
            const char kind = len > 1 ? tok[1] : '\0';
            if (kind == '\0' && numPos < k.positions)
            {
                float* p = posOf(m.vertices, m.stride, numPos++);
                if (c.real(p[0]) && c.real(p[1])) c.real(p[2]);
                p[2] = -p[2];  // DX handedness
            }
            else if (kind == 'n')
            {
                float n[3] = {0, 0, 0};
                if (c.real(n[0]) && c.real(n[1])) c.real(n[2]);
                n[2] = -n[2];
                normals.insert(normals.end(), n, n + 3);
            }
        }
        else c.restOfLine();
    }

    // Out-of-range references would be out-of-bounds reads in the reference; reject instead.
    const uint32_t numVert = m.numVertices();
    for (uint32_t i : m.indices)
        if (i >= numVert) { err = "OBJ face references a vertex that does not exist"; return false; }
    for (uint32_t i : nrmIdx)
        if (3 * static_cast<size_t>(i) + 2 >= normals.size()) { err = "OBJ face references a normal that does not exist"; return false; }

    assignFileNormals(m, normals, nrmIdx);
    std::reverse(m.indices.begin(), m.indices.end());
    if (!k.normals) faceNormals(m);
    boundingBox(m);
    return true;
}

bool loadObj(const char* path, ObjMesh& mesh, std::string& err)
{
    FILE* f = std::fopen(path, "rb");
    if (!f) { err = std::string("cannot open ") + path; return false; }
    std::string text;
    char buf[1 << 16];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, n);
    std::fclose(f);
    return parseObj(text.c_str(), text.size(), mesh, err);
}
}  // namespace dxrv
