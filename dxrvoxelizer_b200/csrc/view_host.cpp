// view_host.cpp -- camera set-up of the reference's viewer, host side (double precision, then float).
//
// Restates DXRVoxelizer::LoadAssets (DXRVoxelizer.cpp:222-234: perspective FOV pi/4, near 1, far 1000;
// eye (8,12,-14), focus (0,4,0), up +y) and Voxelizer::UpdateFrame (Content/Voxelizer.cpp:81-106:
// world = S(w) T(c) S(posScale.w) T(posScale.xyz), light point (-10,45,-75), screen-to-local matrix) in the
// row-vector convention DirectXMath uses (p' = p * M).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <new>
#include <vector>

#include "../../include/dxrv.h"

namespace
{
struct M4 { double m[4][4]; };

M4 mul(const M4& a, const M4& b)
{
    M4 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
        {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += a.m[i][k] * b.m[k][j];
            r.m[i][j] = s;
        }
    return r;
}

M4 identity() { M4 r; std::memset(&r, 0, sizeof(r)); for (int i = 0; i < 4; ++i) r.m[i][i] = 1; return r; }
M4 scaling(double s) { M4 r = identity(); r.m[0][0] = r.m[1][1] = r.m[2][2] = s; return r; }
M4 translation(double x, double y, double z) { M4 r = identity(); r.m[3][0] = x; r.m[3][1] = y; r.m[3][2] = z; return r; }

bool inverse(const M4& a, M4& out)
{
    // Gauss-Jordan with partial pivoting on [a | I]
    double w[4][8];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) { w[i][j] = a.m[i][j]; w[i][4 + j] = (i == j) ? 1.0 : 0.0; }
    for (int c = 0; c < 4; ++c)
    {
        int p = c;
        for (int r = c + 1; r < 4; ++r) if (std::fabs(w[r][c]) > std::fabs(w[p][c])) p = r;
        if (w[p][c] == 0.0) return false;
        if (p != c) for (int j = 0; j < 8; ++j) { const double t = w[p][j]; w[p][j] = w[c][j]; w[c][j] = t; }
        const double d = w[c][c];
        for (int j = 0; j < 8; ++j) w[c][j] /= d;
        for (int r = 0; r < 4; ++r)
            if (r != c)
            {
                const double f = w[r][c];
                if (f != 0.0) for (int j = 0; j < 8; ++j) w[r][j] -= f * w[c][j];
            }
    }
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) out.m[i][j] = w[i][4 + j];
    return true;
}

void transformCoord(const double p[3], const M4& m, float out[3])
{
    double r[4];
    for (int j = 0; j < 4; ++j) r[j] = p[0] * m.m[0][j] + p[1] * m.m[1][j] + p[2] * m.m[2][j] + m.m[3][j];
    for (int j = 0; j < 3; ++j) out[j] = static_cast<float>(r[j] / r[3]);
}
}  // namespace

extern "C" int dxrv_default_view(const float bound[4], const float posScale[4], uint32_t width, uint32_t height,
                                 float screenToLocal[16], float eye[3], float light[3])
{
    if (!bound || !screenToLocal || !eye || !light || width == 0 || height == 0) return DXRV_ERR_INVALID_ARG;
    const float defaultPosScale[4] = {0.0f, 0.0f, 0.0f, 1.0f};   // DXRVoxelizer.cpp:37
    const float* ps = posScale ? posScale : defaultPosScale;

    // XMMatrixLookAtLH(eye, focus, up)
    const double e[3] = {8.0, 12.0, -14.0}, f[3] = {0.0, 4.0, 0.0};
    double z[3] = {f[0] - e[0], f[1] - e[1], f[2] - e[2]};
    const double zl = std::sqrt(z[0] * z[0] + z[1] * z[1] + z[2] * z[2]);
    for (double& c : z) c /= zl;
    double x[3] = {1.0 * z[2] - 0.0 * z[1], 0.0 * z[0] - 0.0 * z[2], 0.0 * z[1] - 1.0 * z[0]};   // cross(up, z), up = (0,1,0)
    const double xl = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    for (double& c : x) c /= xl;
    const double y[3] = {z[1] * x[2] - z[2] * x[1], z[2] * x[0] - z[0] * x[2], z[0] * x[1] - z[1] * x[0]};   // cross(z, x)
    M4 view = identity();
    for (int i = 0; i < 3; ++i) { view.m[i][0] = x[i]; view.m[i][1] = y[i]; view.m[i][2] = z[i]; }
    view.m[3][0] = -(x[0] * e[0] + x[1] * e[1] + x[2] * e[2]);
    view.m[3][1] = -(y[0] * e[0] + y[1] * e[1] + y[2] * e[2]);
    view.m[3][2] = -(z[0] * e[0] + z[1] * e[1] + z[2] * e[2]);

    // XMMatrixPerspectiveFovLH(pi/4, aspect, 1, 1000)
    const double fov = 0.785398163, zn = 1.0, zf = 1000.0, aspect = static_cast<double>(width) / static_cast<double>(height);
    const double h = 1.0 / std::tan(fov * 0.5), w = h / aspect;
    M4 proj; std::memset(&proj, 0, sizeof(proj));
    proj.m[0][0] = w; proj.m[1][1] = h; proj.m[2][2] = zf / (zf - zn); proj.m[2][3] = 1.0; proj.m[3][2] = -zn * zf / (zf - zn);

    const M4 world = mul(mul(mul(scaling(bound[3]), translation(bound[0], bound[1], bound[2])), scaling(ps[3])), translation(ps[0], ps[1], ps[2]));
    M4 worldI;
    if (!inverse(world, worldI)) return DXRV_ERR_INVALID_ARG;
    const M4 worldViewProj = mul(mul(world, view), proj);
    M4 toScreen; std::memset(&toScreen, 0, sizeof(toScreen));
    toScreen.m[0][0] = 0.5 * width; toScreen.m[1][1] = -0.5 * height; toScreen.m[2][2] = 1.0;
    toScreen.m[3][0] = 0.5 * width; toScreen.m[3][1] = 0.5 * height; toScreen.m[3][3] = 1.0;
    M4 screenToLocalM;
    if (!inverse(mul(worldViewProj, toScreen), screenToLocalM)) return DXRV_ERR_INVALID_ARG;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) screenToLocal[4 * i + j] = static_cast<float>(screenToLocalM.m[i][j]);
    const double lightPt[3] = {-10.0, 45.0, -75.0};
    transformCoord(lightPt, worldI, light);
    transformCoord(e, worldI, eye);
    return DXRV_OK;
}

// ---- dxrv_save_image: DXRVoxelizer::SaveImage (DXRVoxelizer.cpp:531-551) -- the screenshot the reference writes with
// stbi_write_png.  A PNG of 8-bit RGB (comp = 3, the reference's default) or RGBA (comp = 4) from an R8G8B8A8 buffer
// with a row pitch; the image data go into stored (uncompressed) deflate blocks: every decoder reads them, and the
// encoder is forty lines instead of a compressor.
namespace
{
uint32_t crc32Of(const uint8_t* p, size_t n, uint32_t crc)
{
    static uint32_t table[256];
    static const bool ready = [] {
        for (uint32_t i = 0; i < 256; ++i)
        {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xedb88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        return true;
    }();
    (void)ready;
    crc = ~crc;
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xffu] ^ (crc >> 8);
    return ~crc;
}
void putBE32(std::vector<uint8_t>& v, uint32_t x) { for (int s = 24; s >= 0; s -= 8) v.push_back(static_cast<uint8_t>(x >> s)); }
bool writeChunk(FILE* f, const char type[4], const std::vector<uint8_t>& data)
{
    std::vector<uint8_t> head;
    putBE32(head, static_cast<uint32_t>(data.size()));
    head.insert(head.end(), type, type + 4);
    uint32_t crc = crc32Of(reinterpret_cast<const uint8_t*>(type), 4, 0);
    if (!data.empty()) crc = crc32Of(data.data(), data.size(), crc);
    std::vector<uint8_t> tail;
    putBE32(tail, crc);
    return std::fwrite(head.data(), 1, head.size(), f) == head.size() && (data.empty() || std::fwrite(data.data(), 1, data.size(), f) == data.size()) &&
           std::fwrite(tail.data(), 1, tail.size(), f) == tail.size();
}
}  // namespace

extern "C" int dxrv_save_image(const char* fileName, const void* rgba, uint32_t width, uint32_t height, uint32_t rowPitchBytes, uint32_t comp)
{
    if (!fileName || !rgba || width == 0 || height == 0 || width > 32768 || height > 32768 || (comp != 3 && comp != 4) ||
        rowPitchBytes < width * 4u)
        return DXRV_ERR_INVALID_ARG;
    try
    {
        // scanlines: filter type 0 + comp bytes per pixel
        const size_t rowBytes = 1 + static_cast<size_t>(width) * comp;
        std::vector<uint8_t> raw(rowBytes * height);
        const uint8_t* src = static_cast<const uint8_t*>(rgba);
        for (uint32_t y = 0; y < height; ++y)
        {
            uint8_t* dst = &raw[rowBytes * y];
            *dst++ = 0;
            const uint8_t* s = src + static_cast<size_t>(rowPitchBytes) * y;
            for (uint32_t x = 0; x < width; ++x, s += 4)
                for (uint32_t k = 0; k < comp; ++k) *dst++ = s[k];
        }
        // zlib stream: header, stored blocks of at most 65535 bytes, Adler-32 of the raw data
        std::vector<uint8_t> z;
        z.reserve(raw.size() + raw.size() / 65535 * 5 + 16);
        z.push_back(0x78); z.push_back(0x01);
        uint32_t a = 1, b = 0;
        for (size_t off = 0; off < raw.size();)
        {
            const size_t len = std::min<size_t>(65535, raw.size() - off);
            z.push_back(off + len == raw.size() ? 1 : 0);
            z.push_back(static_cast<uint8_t>(len & 0xff)); z.push_back(static_cast<uint8_t>(len >> 8));
            z.push_back(static_cast<uint8_t>(~len & 0xff)); z.push_back(static_cast<uint8_t>((~len >> 8) & 0xff));
            z.insert(z.end(), raw.begin() + off, raw.begin() + off + len);
            for (size_t i = off; i < off + len; ++i) { a += raw[i]; if (a >= 65521u) a -= 65521u; b += a; if (b >= 65521u) b -= 65521u; }
            off += len;
        }
        putBE32(z, (b << 16) | a);
        std::vector<uint8_t> ihdr;
        putBE32(ihdr, width); putBE32(ihdr, height);
        ihdr.push_back(8); ihdr.push_back(comp == 3 ? 2 : 6); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
        FILE* f = std::fopen(fileName, "wb");
        if (!f) return DXRV_ERR_IO;
        static const uint8_t signature[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
        bool ok = std::fwrite(signature, 1, 8, f) == 8 && writeChunk(f, "IHDR", ihdr) && writeChunk(f, "IDAT", z) && writeChunk(f, "IEND", {});
        ok = (std::fclose(f) == 0) && ok;
        return ok ? DXRV_OK : DXRV_ERR_IO;
    }
    catch (const std::bad_alloc&) { return DXRV_ERR_OOM; }
}
