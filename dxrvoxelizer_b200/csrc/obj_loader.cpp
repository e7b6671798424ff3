// obj_loader.cpp -- see obj_loader.h.  Compile with -ffp-contract=off: the normal arithmetic
// below must round exactly like the reference's scalar code (mul, add, sqrt, div; no FMA).
#include "obj_loader.h"

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

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

// shared tail of both parsers: validation + the order-dependent post-import steps
static bool finishMesh(ObjMesh& m, const std::vector<float>& normals, const std::vector<uint32_t>& nrmIdx, bool fileHasNormals,
                       std::string& err)
{
    // Out-of-range references would be out-of-bounds reads in the reference; reject instead.
    const uint32_t numVert = m.numVertices();
    for (uint32_t i : m.indices)
        if (i >= numVert) { err = "OBJ face references a vertex that does not exist"; return false; }
    for (uint32_t i : nrmIdx)
        if (3 * static_cast<size_t>(i) + 2 >= normals.size()) { err = "OBJ face references a normal that does not exist"; return false; }

    assignFileNormals(m, normals, nrmIdx);
    std::reverse(m.indices.begin(), m.indices.end());
    if (!fileHasNormals) faceNormals(m);
    boundingBox(m);
    return true;
}

bool parseObj(const char* text, size_t size, ObjMesh& m, std::string& err)
{
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
    return finishMesh(m, normals, nrmIdx, k.normals != 0, err);
}

// ---- fast path ---------------------------------------------------------------------------------------
// Multi-threaded parser for WELL-FORMED files (one record per line; `v x y z`, `vn x y z`, `vt ...`,
// `f` with >= 3 corners whose syntax matches the file: v, v/vt, v//vn or v/vt/vn; plain decimal
// numbers; no line longer than 255 characters).  For such files the reference's fscanf grammar reduces
// to per-line parsing, so the text is cut at newlines and the chunks are parsed concurrently with
// std::from_chars (correctly rounded, like fscanf's %f).  Anything else -- a face continued on the
// next line, `vp` records, hexadecimal floats, ... -- makes it return false WITHOUT an error, and the
// caller falls back to parseObj, which follows the reference's grammar token by token.
namespace
{
struct Chunk
{
    const char* begin; const char* end;
    Counts counts;
    bool odd = false;
};

inline bool isBlank(char c) { return c == ' ' || (static_cast<unsigned char>(c - '\t') <= 4u && c != '\n'); }   // ' ', \t, \v, \f, \r

// first '\n' in [p, end), or end.  OBJ lines are ~30 bytes: a library memchr call per line cost as much as the line's
// numbers (10-15 ns of call and set-up); 16 bytes at a time inline is a few cycles.
inline const char* findNewline(const char* p, const char* end)
{
#if defined(__SSE2__)
    const __m128i nl = _mm_set1_epi8('\n');
    while (end - p >= 16)
    {
        const int mask = _mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128(reinterpret_cast<const __m128i*>(p)), nl));
        if (mask) return p + __builtin_ctz(static_cast<unsigned>(mask));
        p += 16;
    }
#endif
    while (p < end && *p != '\n') ++p;
    return p;
}

// Plain decimals without an exponent -- what OBJ writers emit -- are converted here: the digits as one integer w
// (exact in a double below 2^53), one correctly rounded double division by a power of ten (exact up to 10^22), one
// rounding to float.  Rounding twice differs from rounding once only when the double lands EXACTLY on the midpoint of
// two floats (its low 29 bits are 0x10000000): rounding to double is monotonic and the midpoint is a double, so any other
// double lies on the same side of it as the exact value.  Those, and everything else (exponents, more than 18 digits,
// inf / nan), go to std::from_chars below; both are correctly rounded like the reference's fscanf("%f").
inline bool decimalToFloat(const char* s, const char* end, const char*& stop, float& out)
{
    static const double kPow10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20,
                                      1e21, 1e22};
    bool neg = false;
    if (s < end && *s == '-') { neg = true; ++s; }
    uint64_t w = 0;
    int digits = 0, frac = 0;
    while (s < end && *s >= '0' && *s <= '9') { w = w * 10u + (uint64_t)(*s++ - '0'); ++digits; }
    if (s < end && *s == '.')
    {
        ++s;
        while (s < end && *s >= '0' && *s <= '9') { w = w * 10u + (uint64_t)(*s++ - '0'); ++digits; ++frac; }
    }
    if (digits == 0 || digits > 18 || frac > 22 || w >= (1ull << 53)) return false;
    if (s < end && !isBlank(*s)) return false;                                  // an exponent, a suffix: not here
    const double d = (double)w / kPow10[frac];
    uint64_t bits;
    std::memcpy(&bits, &d, sizeof bits);
    const uint32_t low = (uint32_t)(bits & 0x1fffffffu);
    if (low >= 0x0fffffffu && low <= 0x10000001u) return false;                 // (on or next to) a float midpoint
    const float f = (float)d;
    out = neg ? -f : f;
    stop = s;
    return true;
}

// The same conversion for the vertex fast path of pass B: leading spaces, an optional '-', digits [. digits], then a
// blank or the line's '\n' -- which the caller guarantees lies ahead, so no loop checks a bound.  Same acceptance rule as
// decimalToFloat (false = let the general code look at the line), same arithmetic.
inline bool plainDecimal(const char*& p, float& out)
{
    static const double kPow10[19] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18};
    const char* s = p;
    while (*s == ' ') ++s;
    const bool neg = *s == '-';
    s += neg;
    uint64_t w = 0;
    uint32_t d;
    const char* i0 = s;
    while ((d = static_cast<uint32_t>(static_cast<unsigned char>(*s)) - '0') <= 9u) { w = w * 10u + d; ++s; }
    int digits = static_cast<int>(s - i0), frac = 0;
    if (*s == '.')
    {
        const char* f0 = ++s;
        while ((d = static_cast<uint32_t>(static_cast<unsigned char>(*s)) - '0') <= 9u) { w = w * 10u + d; ++s; }
        frac = static_cast<int>(s - f0);
        digits += frac;
    }
    if (digits == 0 || digits > 18 || w >= (1ull << 53)) return false;
    if (!(*s == ' ' || *s == '\n' || *s == '\r')) return false;
    const double x = static_cast<double>(w) / kPow10[frac];
    uint64_t bits;
    std::memcpy(&bits, &x, sizeof bits);
    const uint32_t low = static_cast<uint32_t>(bits & 0x1fffffffu);
    if (low >= 0x0fffffffu && low <= 0x10000001u) return false;                 // (on or next to) a float midpoint
    const float f = static_cast<float>(x);
    out = neg ? -f : f;
    p = s;
    return true;
}

inline bool fastReal(const char*& p, const char* end, float& out)
{
    while (p < end && isBlank(*p)) ++p;
    const char* q = p;
    if (q < end && *q == '+') ++q;
    const char* d = q;
    if (d < end && *d == '-') ++d;
    if (d >= end || !((*d >= '0' && *d <= '9') || *d == '.')) return false;   // inf / nan / hex: not here
    {
        const char* stop = nullptr;
        if (decimalToFloat(q, end, stop, out)) { p = stop; return true; }
    }
    auto r = std::from_chars(q, end, out);
    if (r.ec != std::errc() || r.ptr == q) return false;
    if (r.ptr < end && !isBlank(*r.ptr)) return false;                          // e.g. 0x10, 1.5f
    p = r.ptr;
    return true;
}

inline bool fastInt(const char*& p, const char* end, long long& out)
{
    const char* q = p;
    bool neg = false;
    if (q < end && (*q == '-' || *q == '+')) neg = (*q++ == '-');
    if (q >= end || *q < '0' || *q > '9') return false;
    long long v = 0;
    while (q < end && *q >= '0' && *q <= '9') v = v * 10 + (*q++ - '0');
    out = neg ? -v : v;
    p = q;
    return true;
}

// Pass A of the fast path: COUNT the records of a chunk -- positions, normals, texture coordinates, triangles (a face of
// c corners = c - 2 triangles) -- without converting a single number.  Everything else (number syntax, corner syntax,
// trailing junk) is checked by pass B, which parses every token anyway and sends the file to the reference-grammar
// parser when something is unusual; the counts only have to be right for files pass B accepts, where a face's corners
// are exactly its blank-separated tokens.  (Pass A used to parse every number just to validate it: the same
// std::from_chars work twice.  C5 file, one thread, build container: 156 -> 274 MB/s; dragon.obj 181 -> 233 MB/s.)
// blank-separated tokens in [q, eol); q[-1] is a blank; bytes up to bufEnd may be read
inline uint32_t countTokens(const char* q, const char* eol, const char* bufEnd)
{
    uint32_t n = 0;
    bool prevBlank = true;
#if defined(__SSE2__)
    const __m128i space = _mm_set1_epi8(' '), lo = _mm_set1_epi8(8), hi = _mm_set1_epi8(14);
    while (q < eol && bufEnd - q >= 16)
    {
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(q));
        const __m128i blank = _mm_or_si128(_mm_cmpeq_epi8(c, space), _mm_and_si128(_mm_cmpgt_epi8(c, lo), _mm_cmplt_epi8(c, hi)));
        const uint32_t b = static_cast<uint32_t>(_mm_movemask_epi8(blank));          // ('\n' counts as a blank here: it ends the line)
        const uint32_t len = eol - q >= 16 ? 16u : static_cast<uint32_t>(eol - q);  // bytes of this block that belong to the line
        const uint32_t inLine = len == 16u ? 0xffffu : (1u << len) - 1u;
        const uint32_t starts = ~b & ((b << 1) | (prevBlank ? 1u : 0u)) & inLine;    // a non-blank right after a blank
        n += static_cast<uint32_t>(__builtin_popcount(starts));
        prevBlank = (b >> 15) & 1u;
        q += 16;
    }
    if (q >= eol) return n;
#endif
    for (; q < eol; ++q)
    {
        const bool blank = isBlank(*q);
        n += (!blank && prevBlank) ? 1u : 0u;
        prevBlank = blank;
    }
    return n;
}

void countChunk(Chunk& ch)
{
    const char* p = ch.begin;
    Counts k;
    while (p < ch.end)
    {
        const char* eol = findNewline(p, ch.end);
        if (eol - p > 255) { ch.odd = true; return; }
        // the records of a well-formed file start in column 0: decided from the first two or three bytes
        if (eol - p >= 2)
        {
            const char c0 = p[0], c1 = p[1];
            if (c0 == 'v')
            {
                if (isBlank(c1)) { ++k.positions; p = eol + 1; continue; }
                if ((c1 == 'n' || c1 == 't') && (eol - p == 2 || isBlank(p[2])))
                {
                    if (c1 == 'n') ++k.normals; else ++k.texcoords;
                    p = eol + 1;
                    continue;
                }
            }
            else if (c0 == 'f' && isBlank(c1))
            {
                const uint32_t corners = countTokens(p + 2, eol, ch.end);
                if (corners < 3) { ch.odd = true; return; }
                k.triangles += corners - 2;
                p = eol + 1;
                continue;
            }
        }
        const char* q = p;
        while (q < eol && isBlank(*q)) ++q;
        const char* tok = q;
        while (q < eol && !isBlank(*q)) ++q;
        const size_t len = static_cast<size_t>(q - tok);
        if (len == 0) { p = eol + 1; continue; }
        const char c0 = tok[0];
        if (c0 == 'v')
        {
            if (len == 1) ++k.positions;
            else if (len == 2 && tok[1] == 'n') ++k.normals;
            else if (len == 2 && tok[1] == 't') ++k.texcoords;
            else { ch.odd = true; return; }
        }
        else if (c0 == 'f')
        {
            if (len != 1) { ch.odd = true; return; }
            uint32_t corners = 0;
            while (true)
            {
                while (q < eol && isBlank(*q)) ++q;
                if (q >= eol) break;
                ++corners;
                while (q < eol && !isBlank(*q)) ++q;
            }
            if (corners < 3) { ch.odd = true; return; }
            k.triangles += corners - 2;
        }
        else if ((c0 >= '0' && c0 <= '9') || c0 == '-' || c0 == '+')
        {
            ch.odd = true; return;   // could be the continuation of a face on the previous line
        }
        p = eol + 1;
    }
    ch.counts = k;
}

// Pass B: parse and store a chunk (kStore = false: the same checks without storing; kept for the template's callers).
template <bool kStore>
void scanChunk(Chunk& ch, bool hasTexc, bool hasNorm, const Counts& total, uint8_t* vb, uint32_t stride, float* normals,
               uint32_t* indices, uint32_t* nrmIdx, Counts base)
{
    const char* p = ch.begin;
    Counts k;
    while (p < ch.end)
    {
        const char* eol = findNewline(p, ch.end);
        if (eol - p > 255) { ch.odd = true; return; }
        // The two records a mesh file is made of, in the form exporters write them -- "v x y z" with plain decimals and
        // "f a b c" with positive indices, from column 0 -- are read here without a bounds check per character: the
        // line's own '\n' (eol < ch.end) ends every loop.  Anything else about the line (other blanks, signs on indices,
        // slashes, polygons, exponents, trailing fields) leaves it to the general code below, which parses it again.
        if (eol < ch.end && eol - p >= 6)
        {
            const char r0 = p[0], r1 = p[1];
            if (r0 == 'f' && r1 == ' ')
            {
                // three corners "v", "v/vt", "v//vn" or "v/vt/vn" -- the form the FILE's records call for -- of positive indices
                const char* q = p + 2;
                uint32_t vIdx[3], nIdx[3] = {0, 0, 0};
                bool plain = true;
                auto index = [&q](uint32_t& out) {
                    const char* d0 = q;
                    uint32_t v = 0, d;
                    while ((d = static_cast<uint32_t>(static_cast<unsigned char>(*q)) - '0') <= 9u) { v = v * 10u + d; ++q; }
                    out = v - 1u;
                    return q != d0 && q - d0 <= 9 && v != 0;
                };
                for (int c = 0; c < 3 && plain; ++c)
                {
                    while (*q == ' ') ++q;
                    uint32_t t;
                    plain = index(vIdx[c]);
                    if (plain && hasTexc) plain = *q++ == '/' && index(t);
                    if (plain && hasNorm) plain = *q++ == '/' && (hasTexc || *q++ == '/') && index(nIdx[c]);
                    if (plain) plain = *q == ' ' || *q == '\r' || *q == '\n';
                }
                if (plain)
                {
                    while (*q == ' ' || *q == '\r') ++q;
                    if (q == eol)
                    {
                        if (kStore)
                        {
                            const size_t t = 3 * static_cast<size_t>(base.triangles + k.triangles);
                            indices[t] = vIdx[0]; indices[t + 1] = vIdx[1]; indices[t + 2] = vIdx[2];
                            if (hasNorm) { nrmIdx[t] = nIdx[0]; nrmIdx[t + 1] = nIdx[1]; nrmIdx[t + 2] = nIdx[2]; }
                        }
                        ++k.triangles;
                        p = eol + 1;
                        continue;
                    }
                }
            }
            else if (r0 == 'v' && (r1 == ' ' || (r1 == 'n' && p[2] == ' ')))
            {
                const bool isNormal = r1 == 'n';
                const char* q = p + (isNormal ? 3 : 2);
                float v[3];
                if (plainDecimal(q, v[0]) && plainDecimal(q, v[1]) && plainDecimal(q, v[2]))
                {
                    while (*q == ' ' || *q == '\r') ++q;
                    if (q == eol)
                    {
                        if (kStore)
                        {
                            float* dst = isNormal ? normals + 3 * static_cast<size_t>(base.normals + k.normals)
                                                  : reinterpret_cast<float*>(vb + static_cast<size_t>(stride) * (base.positions + k.positions));
                            dst[0] = v[0]; dst[1] = v[1]; dst[2] = -v[2];
                        }
                        if (isNormal) ++k.normals; else ++k.positions;
                        p = eol + 1;
                        continue;
                    }
                }
            }
        }
        const char* q = p;
        while (q < eol && isBlank(*q)) ++q;
        const char* tok = q;
        while (q < eol && !isBlank(*q)) ++q;
        const size_t len = static_cast<size_t>(q - tok);
        if (len == 0) { p = eol + 1; continue; }
        const char c0 = tok[0];
        if (c0 == 'v')
        {
            if (len == 1 || (len == 2 && tok[1] == 'n'))
            {
                float v[3];
                if (!(fastReal(q, eol, v[0]) && fastReal(q, eol, v[1]) && fastReal(q, eol, v[2]))) { ch.odd = true; return; }
                // whatever follows on the line (vertex colours, w) must not look like a record
                while (q < eol && isBlank(*q)) ++q;
                if (q < eol && (*q == 'f' || *q == 'v')) { ch.odd = true; return; }
                if (len == 1)
                {
                    if (kStore)
                    {
                        float* dst = reinterpret_cast<float*>(vb + static_cast<size_t>(stride) * (base.positions + k.positions));
                        dst[0] = v[0]; dst[1] = v[1]; dst[2] = -v[2];
                    }
                    ++k.positions;
                }
                else
                {
                    if (kStore)
                    {
                        float* dst = normals + 3 * static_cast<size_t>(base.normals + k.normals);
                        dst[0] = v[0]; dst[1] = v[1]; dst[2] = -v[2];
                    }
                    ++k.normals;
                }
            }
            else if (len == 2 && tok[1] == 't')
            {
                // the reference skips the record token by token; a sane vt line is numbers only
                while (q < eol && isBlank(*q)) ++q;
                if (q >= eol || !((*q >= '0' && *q <= '9') || *q == '-' || *q == '+' || *q == '.')) { ch.odd = true; return; }
                ++k.texcoords;
            }
            else { ch.odd = true; return; }
        }
        else if (c0 == 'f')
        {
            if (len != 1) { ch.odd = true; return; }
            uint32_t v0 = 0, n0 = 0, vPrev = 0, nPrev = 0, corners = 0;
            while (true)
            {
                while (q < eol && isBlank(*q)) ++q;
                if (q >= eol) break;
                long long vi = 0, ti = 0, ni = 0;
                if (!fastInt(q, eol, vi)) { ch.odd = true; return; }
                if (hasTexc)
                {
                    if (q >= eol || *q != '/') { ch.odd = true; return; }
                    ++q;
                    if (hasNorm && q < eol && *q == '/') { ch.odd = true; return; }   // "v//vn" in a file with vt: stale-value semantics
                    if (!fastInt(q, eol, ti)) { ch.odd = true; return; }
                }
                if (hasNorm)
                {
                    if (q >= eol || *q != '/') { ch.odd = true; return; }
                    ++q;
                    if (!hasTexc) { if (q >= eol || *q != '/') { ch.odd = true; return; } ++q; }
                    if (!fastInt(q, eol, ni)) { ch.odd = true; return; }
                }
                if (q < eol && !isBlank(*q)) { ch.odd = true; return; }
                const uint32_t v = CornerReader::resolve(vi, total.positions);
                const uint32_t n = hasNorm ? CornerReader::resolve(ni, total.normals) : 0u;
                ++corners;
                if (corners == 1) { v0 = v; n0 = n; }
                else if (corners >= 3)
                {
                    if (kStore)
                    {
                        const size_t t = 3 * static_cast<size_t>(base.triangles + k.triangles);
                        indices[t] = v0; indices[t + 1] = vPrev; indices[t + 2] = v;
                        if (hasNorm) { nrmIdx[t] = n0; nrmIdx[t + 1] = nPrev; nrmIdx[t + 2] = n; }
                    }
                    ++k.triangles;
                }
                vPrev = v; nPrev = n;
            }
            if (corners < 3) { ch.odd = true; return; }
        }
        else if ((c0 >= '0' && c0 <= '9') || c0 == '-' || c0 == '+')
        {
            ch.odd = true; return;   // could be the continuation of a face on the previous line
        }
        p = eol + 1;
    }
    ch.counts = k;
}
}  // namespace

bool parseObjFast(const char* text, size_t size, ObjMesh& m, std::string& err, unsigned threads)
{
    // DXRV_OBJ_THREADS: threads per FILE (default: up to 16).  A batch loader that parses many files at once (C5:
    // 256 meshes) sets it to 1 and parallelises over the files instead.
    if (threads == 0)
        if (const char* e = std::getenv("DXRV_OBJ_THREADS")) threads = (unsigned)std::max(0, std::atoi(e));
    if (threads == 0) threads = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    const size_t minChunk = 1u << 16;
    size_t numChunks = std::max<size_t>(1, std::min<size_t>(threads, size / minChunk));
    std::vector<Chunk> chunks(numChunks);
    const char* cursor = text;
    for (size_t i = 0; i < numChunks; ++i)
    {
        const char* target = (i + 1 == numChunks) ? text + size : text + size * (i + 1) / numChunks;
        if (target < cursor) target = cursor;
        if (i + 1 < numChunks)
        {
            const char* nl = static_cast<const char*>(std::memchr(target, '\n', static_cast<size_t>(text + size - target)));
            target = nl ? nl + 1 : text + size;
        }
        chunks[i].begin = cursor; chunks[i].end = target;
        cursor = target;
    }

    auto parallelFor = [&](auto&& body) {
        std::vector<std::thread> pool;
        for (size_t i = 1; i < numChunks; ++i) pool.emplace_back([&, i]() { body(i); });
        body(0);
        for (auto& t : pool) t.join();
    };

    // pass A: counts only (countChunk); pass B validates while it parses
    Counts total;
    parallelFor([&](size_t i) { countChunk(chunks[i]); });
    for (auto& ch : chunks) if (ch.odd) return false;
    std::vector<Counts> base(numChunks);
    for (size_t i = 0; i < numChunks; ++i)
    {
        base[i] = total;
        total.positions += chunks[i].counts.positions; total.texcoords += chunks[i].counts.texcoords;
        total.normals += chunks[i].counts.normals; total.triangles += chunks[i].counts.triangles;
    }
    const bool hasTexc = total.texcoords != 0, hasNorm = total.normals != 0;

    m = ObjMesh();
    m.stride = 24 + (hasTexc ? 8 : 0);
    m.vertices.assign(static_cast<size_t>(m.stride) * total.positions, 0);
    m.indices.assign(3 * static_cast<size_t>(total.triangles), 0);
    std::vector<float> normals(3 * static_cast<size_t>(total.normals));
    std::vector<uint32_t> nrmIdx(hasNorm ? 3 * static_cast<size_t>(total.triangles) : 0);

    // pass B: parse and store at the chunk's offsets
    parallelFor([&](size_t i) {
        scanChunk<true>(chunks[i], hasTexc, hasNorm, total, m.vertices.data(), m.stride, normals.data(), m.indices.data(),
                        nrmIdx.data(), base[i]);
    });
    for (auto& ch : chunks) if (ch.odd) return false;
    return finishMesh(m, normals, nrmIdx, hasNorm, err);
}

bool loadObj(const char* path, ObjMesh& mesh, std::string& err, unsigned threads)
{
    FILE* f = std::fopen(path, "rb");
    if (!f) { err = std::string("cannot open ") + path; return false; }
    // The whole text in ONE buffer of the file's size, filled by one read (a regular file says how long it is).  Appending
    // 64 KB pieces to a growing std::string copied the text again with every reallocation: 0.4-0.8 ms of the 2.5 ms a
    // 0.65 MB C5 file takes on one thread.  Anything that cannot say its size (a pipe) is still read piece by piece.
    std::unique_ptr<char[]> owned;
    std::string pieces;
    const char* text = nullptr;
    size_t size = 0;
    long fileSize = -1;
    if (std::fseek(f, 0, SEEK_END) == 0) { fileSize = std::ftell(f); std::rewind(f); }
    if (fileSize > 0)
    {
        owned.reset(new char[static_cast<size_t>(fileSize)]);
        size_t got;
        while (size < static_cast<size_t>(fileSize) && (got = std::fread(owned.get() + size, 1, static_cast<size_t>(fileSize) - size, f)) > 0) size += got;
        text = owned.get();
        if (size == static_cast<size_t>(fileSize) && std::fgetc(f) != EOF) { fileSize = -1; std::rewind(f); size = 0; }   // it grew meanwhile
    }
    if (fileSize <= 0)
    {
        char buf[1 << 16];
        size_t n;
        while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) pieces.append(buf, n);
        text = pieces.data(); size = pieces.size();
    }
    std::fclose(f);
    // well-formed files take the multi-threaded parser; anything unusual is parsed exactly like the reference does
    std::string fastErr;
    if (parseObjFast(text, size, mesh, fastErr, threads)) return true;
    if (!fastErr.empty()) { err = fastErr; return false; }
    return parseObj(text, size, mesh, err);
}
}  // namespace dxrv
