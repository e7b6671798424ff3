// view_host.cpp -- camera set-up of the reference's viewer, host side (double precision, then float).
//
// Restates DXRVoxelizer::LoadAssets (DXRVoxelizer.cpp:222-234: perspective FOV pi/4, near 1, far 1000;
// eye (8,12,-14), focus (0,4,0), up +y) and Voxelizer::UpdateFrame (Content/Voxelizer.cpp:81-106:
// world = S(w) T(c) S(posScale.w) T(posScale.xyz), light point (-10,45,-75), screen-to-local matrix) in the
// row-vector convention DirectXMath uses (p' = p * M).
#include <cmath>
#include <cstdint>
#include <cstring>

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
