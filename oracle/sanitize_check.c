/* ASan/UBSan driver for the CPU oracle (TEST INFRASTRUCTURE): both modes, both tiers, texels, slabs, ragged N, the
 * viewer pass, degenerate inputs.  Build + run: make -C oracle sanitize */
#include "dxrv_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static unsigned rng_state = 12345u;
static float frand(void) { rng_state = rng_state * 1664525u + 1013904223u; return (float)(rng_state >> 8) / 16777216.0f; }

int main(void)
{
    enum { NT = 300 };
    float* vb = (float*)malloc(sizeof(float) * 6 * (3 * NT + 2));
    uint32_t* ib = (uint32_t*)malloc(sizeof(uint32_t) * 3 * NT);
    for (int t = 0; t < NT; ++t)
    {
        const float cx = frand() * 2 - 1, cy = frand() * 2 - 1, cz = frand() * 2 - 1, s = powf(10.0f, -3.0f + 3.2f * frand());
        for (int c = 0; c < 3; ++c)
        {
            float* v = vb + 6 * (3 * t + c);
            v[0] = cx + s * (frand() * 2 - 1); v[1] = cy + s * (frand() * 2 - 1); v[2] = cz + s * (frand() * 2 - 1);
            v[3] = frand() - 0.5f; v[4] = frand() - 0.5f; v[5] = frand() - 0.5f;
            ib[3 * t + c] = (uint32_t)(3 * t + c);
        }
    }
    /* degenerate: a zero-area triangle and a duplicate */
    memcpy(vb + 6 * 3, vb + 6 * 4, 24); ib[6] = ib[3]; ib[7] = ib[4]; ib[8] = ib[5];
    float* corner = vb + 6 * 3 * NT;
    corner[0] = corner[1] = corner[2] = -1; corner[6] = corner[7] = corner[8] = 1;
    const uint32_t nv = 3 * NT + 2;
    float bound[4];
    oracle_bound(vb, nv, 24, bound);
    int bad = 0;
    const uint32_t sizes[] = {1, 5, 33, 40};
    for (unsigned si = 0; si < 4; ++si)
    {
        const uint32_t N = sizes[si], P = (N + 31) / 32;
        uint32_t* a = (uint32_t*)malloc(sizeof(uint32_t) * N * N * P);
        uint32_t* b = (uint32_t*)malloc(sizeof(uint32_t) * N * N * P);
        uint32_t* tex = (uint32_t*)malloc(sizeof(uint32_t) * N * N * N);
        for (uint32_t mode = 0; mode < 2; ++mode)
        {
            uint64_t cr = 0, odd = 0;
            bad |= oracle_voxelize(vb, nv, 24, ib, 3 * NT, NULL, N, mode, 0, N, ORACLE_TIER_BRUTE, 0, a, mode == 0 ? tex : NULL, &cr, &odd);
            bad |= oracle_voxelize(vb, nv, 24, ib, 3 * NT, bound, N, mode, 0, N, ORACLE_TIER_ACCEL, 2, b, NULL, &cr, &odd);
            if (memcmp(a, b, sizeof(uint32_t) * N * N * P)) { printf("tier mismatch N=%u mode=%u\n", N, mode); bad = 1; }
            if (N > 4)
            {
                bad |= oracle_voxelize(vb, nv, 24, ib, 3 * NT, NULL, N, mode, N / 3, N / 3 + 2, ORACLE_TIER_ACCEL, 0, b, NULL, NULL, NULL);
                if (memcmp(a + (size_t)(N / 3) * N * P, b, sizeof(uint32_t) * 2 * N * P)) { printf("slab mismatch N=%u mode=%u\n", N, mode); bad = 1; }
            }
        }
        if (N >= 33)
        {
            const float m[16] = {0.01f, 0, 0, 0, 0, -0.01f, 0, 0, 0, 0, 1, 0, -0.8f, 0.45f, -3, 1}, eye[3] = {0, 0, -3}, light[3] = {-1, 4, -7};
            uint32_t* img = (uint32_t*)malloc(sizeof(uint32_t) * 160 * 90);
            bad |= oracle_render_view(a, N, 160, 90, m, eye, light, img, 2);
            free(img);
        }
        free(a); free(b); free(tex);
    }
    /* invalid arguments must be refused, not crash */
    if (oracle_voxelize(vb, nv, 24, ib, 3 * NT, NULL, 8, 7, 0, 8, 1, 0, (uint32_t*)vb, NULL, NULL, NULL) == 0) bad = 1;
    if (oracle_voxelize(vb, nv, 24, ib, 3 * NT, NULL, 8, 1, 4, 4, 1, 0, (uint32_t*)vb, NULL, NULL, NULL) == 0) bad = 1;
    free(vb); free(ib);
    printf(bad ? "oracle sanitize check: FAILED\n" : "oracle sanitize check: ok (both modes, both tiers, slabs, viewer)\n");
    return bad;
}
