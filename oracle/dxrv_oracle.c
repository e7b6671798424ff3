/* dxrv_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY (see dxrv_oracle.h for the contract,
 * the reference file:line map and the "parity unpinned" note).
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -fPIC -shared  (NO -ffast-math; contraction off so
 * that every float operation below rounds once, exactly as Spec H says).
 */
#include "dxrv_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TMAX 10000.0f      /* DXRVoxelizer.hlsl:77 */
#define THRESHOLD 0.12f    /* DXRVoxelizer.hlsl:5  */

typedef struct { float x, y, z; } f3;

static inline float fminsel(float a, float b) { return (a < b) ? a : b; }
static inline float fmaxsel(float a, float b) { return (a > b) ? a : b; }
static inline float comp(const f3* v, int i) { return i == 0 ? v->x : (i == 1 ? v->y : v->z); }

/* ---- scene ------------------------------------------------------------------------------------ */

typedef struct
{
    uint32_t numTris;
    f3* a; f3* b; f3* c;       /* scene-space (normalised) corners per triangle            */
    f3* lo; f3* hi;            /* exact componentwise min / max of the corners              */
    const uint8_t* vertices;   /* original interleaved vertices (normals at byte offset 12) */
    uint32_t stride;
    const uint32_t* indices;
} Scene;

void oracle_bound(const void* vertices, uint32_t numVerts, uint32_t stride, float out[4])
{
    /* ObjLoader::computeAABB (XUSGObjLoader.cpp:386-416) then Voxelizer::Init (Voxelizer.cpp:52-57) */
    const uint8_t* vb = (const uint8_t*)vertices;
    float mn[3], mx[3];
    memcpy(mn, vb, 12);
    memcpy(mx, vb, 12);
    for (uint32_t i = 1; i < numVerts; ++i)
    {
        float p[3];
        memcpy(p, vb + (size_t)stride * i, 12);
        for (int k = 0; k < 3; ++k)
        {
            if (p[k] < mn[k]) mn[k] = p[k];
            else if (p[k] > mx[k]) mx[k] = p[k];
        }
    }
    const float ex = mx[0] - mn[0], ey = mx[1] - mn[1], ez = mx[2] - mn[2];
    out[0] = (mx[0] + mn[0]) / 2.0f;
    out[1] = (mx[1] + mn[1]) / 2.0f;
    out[2] = (mx[2] + mn[2]) / 2.0f;
    float m = ey > ez ? ey : ez;
    m = ex > m ? ex : m;
    out[3] = m / 2.0f;
}

static f3 to_scene(const uint8_t* vb, uint32_t stride, uint32_t i, const float bound[4])
{
    /* instance transform inverse(Scale(w)*Translate(c)) applied to the vertex: Voxelizer.cpp:304-310 */
    float p[3];
    memcpy(p, vb + (size_t)stride * i, 12);
    f3 r;
    r.x = (p[0] - bound[0]) / bound[3];
    r.y = (p[1] - bound[1]) / bound[3];
    r.z = (p[2] - bound[2]) / bound[3];
    return r;
}

static int scene_init(Scene* s, const void* vertices, uint32_t stride, const uint32_t* indices,
                      uint32_t numIndices, const float bound[4])
{
    memset(s, 0, sizeof(*s));
    s->numTris = numIndices / 3;
    s->vertices = (const uint8_t*)vertices;
    s->stride = stride;
    s->indices = indices;
    const size_t n = s->numTris ? s->numTris : 1;
    s->a = (f3*)malloc(n * sizeof(f3)); s->b = (f3*)malloc(n * sizeof(f3)); s->c = (f3*)malloc(n * sizeof(f3));
    s->lo = (f3*)malloc(n * sizeof(f3)); s->hi = (f3*)malloc(n * sizeof(f3));
    if (!s->a || !s->b || !s->c || !s->lo || !s->hi) return -1;
    for (uint32_t k = 0; k < s->numTris; ++k)
    {
        /* triangle k uses indices I[3k..3k+2]: DXRVoxelizer.hlsl:93-99 */
        const f3 a = to_scene(s->vertices, stride, indices[3 * k], bound);
        const f3 b = to_scene(s->vertices, stride, indices[3 * k + 1], bound);
        const f3 c = to_scene(s->vertices, stride, indices[3 * k + 2], bound);
        s->a[k] = a; s->b[k] = b; s->c[k] = c;
        s->lo[k].x = fminsel(fminsel(a.x, b.x), c.x); s->hi[k].x = fmaxsel(fmaxsel(a.x, b.x), c.x);
        s->lo[k].y = fminsel(fminsel(a.y, b.y), c.y); s->hi[k].y = fmaxsel(fmaxsel(a.y, b.y), c.y);
        s->lo[k].z = fminsel(fminsel(a.z, b.z), c.z); s->hi[k].z = fmaxsel(fmaxsel(a.z, b.z), c.z);
    }
    return 0;
}

static void scene_free(Scene* s)
{
    free(s->a); free(s->b); free(s->c); free(s->lo); free(s->hi);
}

static inline float centre(uint32_t i, uint32_t N)
{
    /* generateRay: (index + 0.5) / DispatchRaysDimensions().x * 2.0 - 1.0, DXRVoxelizer.hlsl:46 */
    return ((float)i + 0.5f) / (float)N * 2.0f - 1.0f;
}

/* ---- shared edge-function values (Spec H, MODE_SHADER step 2 / MODE_PARITY step 2) --------------- */

static inline void edge_values(float Ap, float Aq, float Bp, float Bq, float Cp, float Cq, float* U,
                               float* V, float* W)
{
    float u = Cp * Bq - Cq * Bp;
    float v = Ap * Cq - Aq * Cp;
    float w = Bp * Aq - Bq * Ap;
    if (u == 0.0f || v == 0.0f || w == 0.0f)
    {
        u = (float)((double)Cp * (double)Bq - (double)Cq * (double)Bp);
        v = (float)((double)Ap * (double)Cq - (double)Aq * (double)Cp);
        w = (float)((double)Bp * (double)Aq - (double)Bq * (double)Ap);
    }
    *U = u; *V = v; *W = w;
}

/* ---- MODE_SHADER --------------------------------------------------------------------------------- */

typedef struct
{
    f3 O, D, invD;
    int kx, ky, kz;
    float Sx, Sy, Sz;
} Ray;

typedef struct
{
    float tc;       /* clamped hit distance */
    uint32_t prim;  /* original triangle index */
    float bx, by;   /* barycentrics of vertices 1 and 2 */
    int valid;
} Hit;

/* returns 0 when the ray has no direction (centre voxel of an odd grid: normalize(0) is NaN) */
static int ray_init(Ray* r, uint32_t x, uint32_t y, uint32_t z, uint32_t N)
{
    f3 pos;
    pos.x = centre(x, N);
    pos.y = -centre(y, N);   /* "Invert Y for Y-up-style NDC", DXRVoxelizer.hlsl:49 */
    pos.z = centre(z, N);
    if (pos.x == 0.0f && pos.y == 0.0f && pos.z == 0.0f) return 0;
    const float len = sqrtf((pos.x * pos.x + pos.y * pos.y) + pos.z * pos.z);
    r->O = pos;
    r->D.x = pos.x / len; r->D.y = pos.y / len; r->D.z = pos.z / len;
    float inv[3] = { 1.0f / r->D.x, 1.0f / r->D.y, 1.0f / r->D.z };
    for (int i = 0; i < 3; ++i)
    {
        if (inv[i] > FLT_MAX) inv[i] = FLT_MAX;
        if (inv[i] < -FLT_MAX) inv[i] = -FLT_MAX;
    }
    r->invD.x = inv[0]; r->invD.y = inv[1]; r->invD.z = inv[2];

    const float ax = fabsf(r->D.x), ay = fabsf(r->D.y), az = fabsf(r->D.z);
    int kz = 0; float m = ax;
    if (ay > m) { kz = 1; m = ay; }
    if (az > m) { kz = 2; }
    int kx = (kz + 1) % 3, ky = (kx + 1) % 3;
    const float dz = comp(&r->D, kz);
    if (dz < 0.0f) { const int t = kx; kx = ky; ky = t; }
    r->kx = kx; r->ky = ky; r->kz = kz;
    r->Sx = comp(&r->D, kx) / dz;
    r->Sy = comp(&r->D, ky) / dz;
    r->Sz = 1.0f / dz;
    return 1;
}

/* Spec H step 1.  Returns 1 when the interval is non-empty. */
static inline int slab(const Ray* r, const f3* lo, const f3* hi, float* tinOut, float* toutOut)
{
    float tin = 0.0f, tout = TMAX;
    float t0, t1;
    t0 = (lo->x - r->O.x) * r->invD.x; t1 = (hi->x - r->O.x) * r->invD.x;
    tin = fmaxsel(fminsel(t0, t1), tin); tout = fminsel(fmaxsel(t0, t1), tout);
    t0 = (lo->y - r->O.y) * r->invD.y; t1 = (hi->y - r->O.y) * r->invD.y;
    tin = fmaxsel(fminsel(t0, t1), tin); tout = fminsel(fmaxsel(t0, t1), tout);
    t0 = (lo->z - r->O.z) * r->invD.z; t1 = (hi->z - r->O.z) * r->invD.z;
    tin = fmaxsel(fminsel(t0, t1), tin); tout = fminsel(fmaxsel(t0, t1), tout);
    *tinOut = tin; *toutOut = tout;
    return tin <= tout;
}

/* Spec H steps 1-4 for one (ray, triangle) pair; updates *best when (tc,k) is smaller. */
static inline void shader_pair(const Ray* r, const Scene* s, uint32_t k, Hit* best)
{
    float tin, tout;
    if (!slab(r, &s->lo[k], &s->hi[k], &tin, &tout)) return;

    const f3 A = { s->a[k].x - r->O.x, s->a[k].y - r->O.y, s->a[k].z - r->O.z };
    const f3 B = { s->b[k].x - r->O.x, s->b[k].y - r->O.y, s->b[k].z - r->O.z };
    const f3 C = { s->c[k].x - r->O.x, s->c[k].y - r->O.y, s->c[k].z - r->O.z };
    const float Akz = comp(&A, r->kz), Bkz = comp(&B, r->kz), Ckz = comp(&C, r->kz);
    const float Ax = comp(&A, r->kx) - r->Sx * Akz, Ay = comp(&A, r->ky) - r->Sy * Akz;
    const float Bx = comp(&B, r->kx) - r->Sx * Bkz, By = comp(&B, r->ky) - r->Sy * Bkz;
    const float Cx = comp(&C, r->kx) - r->Sx * Ckz, Cy = comp(&C, r->ky) - r->Sy * Ckz;

    float U, V, W;
    edge_values(Ax, Ay, Bx, By, Cx, Cy, &U, &V, &W);
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return;
    const float det = (U + V) + W;
    if (det == 0.0f) return;
    const float Az = r->Sz * Akz, Bz = r->Sz * Bkz, Cz = r->Sz * Ckz;
    const float t = ((U * Az + V * Bz) + W * Cz) / det;
    const float tc = fminsel(fmaxsel(t, tin), tout);
    if (!(tc > 0.0f && tc < TMAX)) return;
    if (!best->valid || tc < best->tc || (tc == best->tc && k < best->prim))
    {
        best->valid = 1; best->tc = tc; best->prim = k;
        best->bx = V / det; best->by = W / det;
    }
}

static inline uint32_t unorm10(float v)
{
    /* D3D FLOAT -> UNORM: NaN -> 0, clamp to [0,1], scale, +0.5, truncate */
    if (!(v > 0.0f)) return 0u;
    if (v > 1.0f) v = 1.0f;
    return (uint32_t)(v * 1023.0f + 0.5f);
}

/* closestHitMain (DXRVoxelizer.hlsl:132-140) + getInput (:90-119).  Returns inside. */
static int shade(const Ray* r, const Scene* s, const Hit* h, uint32_t* texel)
{
    float n0[3], n1[3], n2[3], nrm[3];
    memcpy(n0, s->vertices + (size_t)s->stride * s->indices[3 * h->prim] + 12, 12);
    memcpy(n1, s->vertices + (size_t)s->stride * s->indices[3 * h->prim + 1] + 12, 12);
    memcpy(n2, s->vertices + (size_t)s->stride * s->indices[3 * h->prim + 2] + 12, 12);
    for (int i = 0; i < 3; ++i) nrm[i] = (n0[i] + h->bx * (n1[i] - n0[i])) + h->by * (n2[i] - n0[i]);
    const float len = sqrtf((nrm[0] * nrm[0] + nrm[1] * nrm[1]) + nrm[2] * nrm[2]);
    const float nx = nrm[0] / len, ny = nrm[1] / len, nz = nrm[2] / len;
    const float d = (nx * r->D.x + ny * r->D.y) + nz * r->D.z;
    const int inside = d > THRESHOLD;
    if (texel) *texel = inside ? (unorm10(nx) | (unorm10(ny) << 10) | (unorm10(nz) << 20) | (3u << 30)) : 0u;
    return inside;
}

/* -- accelerated tier for MODE_SHADER: a plain top-down median-split BVH with exact boxes -------- */

typedef struct
{
    f3 lo, hi;
    uint32_t left, right;   /* children; leaf when count > 0 */
    uint32_t first, count;  /* range in order[] */
} CpuNode;

typedef struct
{
    CpuNode* nodes;
    uint32_t numNodes;
    uint32_t* order;
} CpuBvh;

static void box_of_range(const Scene* s, const uint32_t* order, uint32_t first, uint32_t count, f3* lo, f3* hi)
{
    *lo = s->lo[order[first]]; *hi = s->hi[order[first]];
    for (uint32_t i = 1; i < count; ++i)
    {
        const f3* l = &s->lo[order[first + i]]; const f3* h = &s->hi[order[first + i]];
        lo->x = fminsel(lo->x, l->x); lo->y = fminsel(lo->y, l->y); lo->z = fminsel(lo->z, l->z);
        hi->x = fmaxsel(hi->x, h->x); hi->y = fmaxsel(hi->y, h->y); hi->z = fmaxsel(hi->z, h->z);
    }
}

typedef struct { float key; uint32_t prim; } SortItem;
static int cmp_item(const void* a, const void* b)
{
    const SortItem* x = (const SortItem*)a; const SortItem* y = (const SortItem*)b;
    if (x->key < y->key) return -1;
    if (x->key > y->key) return 1;
    return (x->prim > y->prim) - (x->prim < y->prim);
}

static uint32_t bvh_build_rec(CpuBvh* bvh, const Scene* s, uint32_t first, uint32_t count, SortItem* tmp)
{
    const uint32_t me = bvh->numNodes++;
    CpuNode* n = &bvh->nodes[me];
    box_of_range(s, bvh->order, first, count, &n->lo, &n->hi);
    n->first = first; n->count = 0; n->left = n->right = 0;
    if (count <= 4) { n->count = count; return me; }
    const float ex = n->hi.x - n->lo.x, ey = n->hi.y - n->lo.y, ez = n->hi.z - n->lo.z;
    const int axis = (ex >= ey && ex >= ez) ? 0 : (ey >= ez ? 1 : 2);
    for (uint32_t i = 0; i < count; ++i)
    {
        const uint32_t k = bvh->order[first + i];
        tmp[i].prim = k;
        tmp[i].key = comp(&s->lo[k], axis) + comp(&s->hi[k], axis);
    }
    qsort(tmp, count, sizeof(SortItem), cmp_item);
    for (uint32_t i = 0; i < count; ++i) bvh->order[first + i] = tmp[i].prim;
    const uint32_t half = count / 2;
    const uint32_t l = bvh_build_rec(bvh, s, first, half, tmp);
    const uint32_t r = bvh_build_rec(bvh, s, first + half, count - half, tmp);
    bvh->nodes[me].left = l; bvh->nodes[me].right = r;
    return me;
}

static int bvh_build(CpuBvh* bvh, const Scene* s)
{
    memset(bvh, 0, sizeof(*bvh));
    if (!s->numTris) return 0;
    bvh->nodes = (CpuNode*)malloc(sizeof(CpuNode) * 2 * (size_t)s->numTris);
    bvh->order = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)s->numTris);
    SortItem* tmp = (SortItem*)malloc(sizeof(SortItem) * (size_t)s->numTris);
    if (!bvh->nodes || !bvh->order || !tmp) { free(tmp); return -1; }
    for (uint32_t i = 0; i < s->numTris; ++i) bvh->order[i] = i;
    bvh_build_rec(bvh, s, 0, s->numTris, tmp);
    free(tmp);
    return 0;
}

static void bvh_free(CpuBvh* bvh) { free(bvh->nodes); free(bvh->order); }

static void bvh_closest(const CpuBvh* bvh, const Scene* s, const Ray* r, Hit* best)
{
    if (!s->numTris) return;
    uint32_t stack[128];
    int sp = 0;
    stack[sp++] = 0;
    while (sp)
    {
        const CpuNode* n = &bvh->nodes[stack[--sp]];
        float tin, tout;
        if (!slab(r, &n->lo, &n->hi, &tin, &tout)) continue;
        if (best->valid && tin > best->tc) continue;   /* tc >= tin of every ancestor box */
        if (n->count)
        {
            for (uint32_t i = 0; i < n->count; ++i) shader_pair(r, s, bvh->order[n->first + i], best);
        }
        else
        {
            stack[sp++] = n->left;
            stack[sp++] = n->right;
        }
    }
}

/* ---- MODE_PARITY --------------------------------------------------------------------------------- */

static inline int sgn_d(double v) { return (v > 0.0) - (v < 0.0); }

/* exact sign of edge(P,Q) with the (+e,+e^2) tie rule */
static inline int edge_sign(float Pp, float Pq, float Qp, float Qq)
{
    int s = sgn_d((double)Pp * (double)Qq - (double)Pq * (double)Qp);
    if (s == 0)
    {
        if (Pq != Qq) s = (Pq > Qq) ? 1 : -1;
        else s = (Qp > Pp) - (Qp < Pp);
    }
    return s;
}

/* Spec H MODE_PARITY steps 1-3 for one (column, triangle) pair.  Returns 1 and the first toggled
 * voxel ix in [0,N] when the column's line crosses the triangle. */
static inline int parity_pair(const Scene* s, uint32_t k, float Y, float Z, uint32_t N, uint32_t* ixOut)
{
    const f3 a = s->a[k], b = s->b[k], c = s->c[k];
    const float Ap = a.y - Y, Aq = a.z - Z;
    const float Bp = b.y - Y, Bq = b.z - Z;
    const float Cp = c.y - Y, Cq = c.z - Z;
    const int sU = edge_sign(Cp, Cq, Bp, Bq);
    const int sV = edge_sign(Ap, Aq, Cp, Cq);
    const int sW = edge_sign(Bp, Bq, Ap, Aq);
    if (!(sU == sV && sV == sW && sU != 0)) return 0;
    float U, V, W;
    edge_values(Ap, Aq, Bp, Bq, Cp, Cq, &U, &V, &W);
    const float det = (U + V) + W;
    if (det == 0.0f) return 0;
    const float d = ((U * a.x + V * b.x) + W * c.x) / det;

    /* smallest x with centre(x) > d; centre() is monotone in x, so fix up an estimate */
    double g = floor(((double)d + 1.0) * 0.5 * (double)N + 0.5);
    if (!(g > 0.0)) g = 0.0;               /* also catches NaN */
    if (g > (double)N) g = (double)N;
    uint32_t ix = (uint32_t)g;
    while (ix > 0 && centre(ix - 1, N) > d) --ix;
    while (ix < N && !(centre(ix, N) > d)) ++ix;
    *ixOut = ix;
    return 1;
}

/* accelerated tier for MODE_PARITY: uniform bins over (y,z); a crossing implies the column lies in
 * the triangle's closed (y,z) box (all p or all q of one sign otherwise), and cell() is monotone. */
typedef struct
{
    uint32_t G;
    uint32_t* start;  /* G*G+1 */
    uint32_t* items;
} Bins;

static inline uint32_t cell_of(float v, uint32_t G)
{
    double c = floor(((double)v + 1.0) * 0.5 * (double)G);
    if (!(c > 0.0)) c = 0.0;
    if (c > (double)(G - 1)) c = (double)(G - 1);
    return (uint32_t)c;
}

static int bins_build(Bins* bn, const Scene* s, uint32_t N)
{
    memset(bn, 0, sizeof(*bn));
    uint32_t G = N < 16 ? 16 : (N > 1024 ? 1024 : N);
    bn->G = G;
    bn->start = (uint32_t*)calloc((size_t)G * G + 1, sizeof(uint32_t));
    if (!bn->start) return -1;
    for (int pass = 0; pass < 2; ++pass)
    {
        for (uint32_t k = 0; k < s->numTris; ++k)
        {
            const uint32_t y0 = cell_of(s->lo[k].y, G), y1 = cell_of(s->hi[k].y, G);
            const uint32_t z0 = cell_of(s->lo[k].z, G), z1 = cell_of(s->hi[k].z, G);
            for (uint32_t z = z0; z <= z1; ++z)
                for (uint32_t y = y0; y <= y1; ++y)
                {
                    if (pass == 0) bn->start[(size_t)z * G + y + 1]++;
                    else bn->items[bn->start[(size_t)z * G + y]++] = k;
                }
        }
        if (pass == 0)
        {
            for (size_t i = 0; i < (size_t)G * G; ++i) bn->start[i + 1] += bn->start[i];
            bn->items = (uint32_t*)malloc(sizeof(uint32_t) * (bn->start[(size_t)G * G] ? bn->start[(size_t)G * G] : 1));
            if (!bn->items) return -1;
        }
        else
        {
            /* undo the cursor advance: start[i] now holds the END of cell i */
            for (size_t i = (size_t)G * G; i > 0; --i) bn->start[i] = bn->start[i - 1];
            bn->start[0] = 0;
        }
    }
    return 0;
}

static void bins_free(Bins* bn) { free(bn->start); free(bn->items); }

/* ---- driver ---------------------------------------------------------------------------------------- */

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int oracle_voxelize(const void* vertices, uint32_t numVerts, uint32_t stride, const uint32_t* indices,
                    uint32_t numIndices, const float boundIn[4], uint32_t N, uint32_t mode, uint32_t z0,
                    uint32_t z1, int tier, int threads, uint32_t* outBits, uint32_t* outTexels,
                    uint64_t* crossingsOut, uint64_t* oddColumnsOut)
{
    if (!vertices || !indices || !outBits || N == 0 || z0 >= z1 || z1 > N || stride < 12 || numVerts == 0) return -1;
    if (mode == ORACLE_MODE_SHADER && stride < 24) return -1;
    if (mode != ORACLE_MODE_SHADER && mode != ORACLE_MODE_PARITY) return -1;
    for (uint32_t i = 0; i < numIndices; ++i) if (indices[i] >= numVerts) return -1;

    float bound[4];
    if (boundIn) memcpy(bound, boundIn, sizeof(bound));
    else oracle_bound(vertices, numVerts, stride, bound);

    Scene s;
    if (scene_init(&s, vertices, stride, indices, numIndices, bound)) { scene_free(&s); return -1; }

    CpuBvh bvh; memset(&bvh, 0, sizeof(bvh));
    Bins bins; memset(&bins, 0, sizeof(bins));
    int rc = 0;
    if (tier == ORACLE_TIER_ACCEL)
        rc = (mode == ORACLE_MODE_SHADER) ? bvh_build(&bvh, &s) : bins_build(&bins, &s, N);
    if (rc) { bvh_free(&bvh); bins_free(&bins); scene_free(&s); return -1; }

    const uint32_t P = (N + 31) / 32;
    const int64_t rows = (int64_t)(z1 - z0) * N;
    uint64_t crossings = 0, oddColumns = 0;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#else
    threads = 1;
#endif

#pragma omp parallel for schedule(dynamic, 16) num_threads(threads) reduction(+ : crossings, oddColumns)
    for (int64_t row = 0; row < rows; ++row)
    {
        const uint32_t z = z0 + (uint32_t)(row / N), y = (uint32_t)(row % N);
        uint32_t* words = outBits + (size_t)row * P;
        memset(words, 0, sizeof(uint32_t) * P);

        if (mode == ORACLE_MODE_SHADER)
        {
            for (uint32_t x = 0; x < N; ++x)
            {
                uint32_t texel = 0;
                Ray r;
                int inside = 0;
                if (ray_init(&r, x, y, z, N))
                {
                    Hit best; memset(&best, 0, sizeof(best));
                    if (tier == ORACLE_TIER_ACCEL) bvh_closest(&bvh, &s, &r, &best);
                    else for (uint32_t k = 0; k < s.numTris; ++k) shader_pair(&r, &s, k, &best);
                    if (best.valid) inside = shade(&r, &s, &best, &texel);
                }
                if (inside) words[x >> 5] |= 1u << (x & 31);
                if (outTexels) outTexels[(size_t)row * N + x] = texel;
            }
        }
        else
        {
            const float Y = -centre(y, N), Z = centre(z, N);
            uint32_t count = 0, ix;
            if (tier == ORACLE_TIER_ACCEL)
            {
                const size_t cell = (size_t)cell_of(Z, bins.G) * bins.G + cell_of(Y, bins.G);
                for (uint32_t i = bins.start[cell]; i < bins.start[cell + 1]; ++i)
                    if (parity_pair(&s, bins.items[i], Y, Z, N, &ix))
                    {
                        ++count;
                        if (ix < N) words[ix >> 5] ^= 1u << (ix & 31);
                    }
            }
            else
            {
                for (uint32_t k = 0; k < s.numTris; ++k)
                    if (parity_pair(&s, k, Y, Z, N, &ix))
                    {
                        ++count;
                        if (ix < N) words[ix >> 5] ^= 1u << (ix & 31);
                    }
            }
            crossings += count;
            oddColumns += count & 1u;
            /* toggles -> occupancy: inclusive prefix XOR along x */
            uint32_t carry = 0;
            for (uint32_t w = 0; w < P; ++w)
            {
                uint32_t v = words[w];
                v ^= v << 1; v ^= v << 2; v ^= v << 4; v ^= v << 8; v ^= v << 16;
                v ^= carry;
                carry = (v & 0x80000000u) ? 0xffffffffu : 0u;
                words[w] = v;
            }
            if (N & 31) words[P - 1] &= (1u << (N & 31)) - 1u;
        }
    }

    if (crossingsOut) *crossingsOut = crossings;
    if (oddColumnsOut) *oddColumnsOut = oddColumns;
    bvh_free(&bvh); bins_free(&bins); scene_free(&s);
    return 0;
}

/* ---- viewer pass ----------------------------------------------------------------
 * Restatement of Content/Shaders/PSRayCast.hlsl:61-187 over the bit grid: ScreenToLocal (:61-66),
 * ComputeStartPoint (:71-99), GetSample = trilinear LINEAR_CLAMP fetch of the alpha channel, min(d*8,16)
 * (:104-111), main (:117-187).  min16float evaluated as float.  The GPU kernel is compared with a
 * tolerance on the 8-bit image: normalize/sqrt precision and the sampler's filter weights are
 * implementation defined in the reference. */
static float occ_at(const uint32_t* bits, int N, int P, int x, int y, int z)
{
    x = x < 0 ? 0 : (x > N - 1 ? N - 1 : x);
    y = y < 0 ? 0 : (y > N - 1 ? N - 1 : y);
    z = z < 0 ? 0 : (z > N - 1 ? N - 1 : z);
    return (float)((bits[((size_t)z * N + y) * P + (x >> 5)] >> (x & 31)) & 1u);
}

static float get_sample(const uint32_t* bits, int N, int P, float tx, float ty, float tz)
{
    const float ux = tx * (float)N - 0.5f, uy = ty * (float)N - 0.5f, uz = tz * (float)N - 0.5f;
    const float x0 = floorf(ux), y0 = floorf(uy), z0 = floorf(uz);
    const float fx = ux - x0, fy = uy - y0, fz = uz - z0;
    const int ix = (int)x0, iy = (int)y0, iz = (int)z0;
    float c[2][2][2];
    for (int k = 0; k < 2; ++k) for (int j = 0; j < 2; ++j) for (int i = 0; i < 2; ++i) c[k][j][i] = occ_at(bits, N, P, ix + i, iy + j, iz + k);
    float cz[2];
    for (int k = 0; k < 2; ++k)
    {
        const float a = c[k][0][0] + fx * (c[k][0][1] - c[k][0][0]);
        const float b = c[k][1][0] + fx * (c[k][1][1] - c[k][1][0]);
        cz[k] = a + fy * (b - a);
    }
    const float d = cz[0] + fz * (cz[1] - cz[0]);
    return d * 8.0f < 16.0f ? d * 8.0f : 16.0f;
}

static float sat(float a) { return a < 0.0f ? 0.0f : (a > 1.0f ? 1.0f : a); }
static uint32_t unorm8(float a) { return (uint32_t)(sat(a) * 255.0f + 0.5f); }

int oracle_render_view(const uint32_t* bits, uint32_t N, uint32_t width, uint32_t height, const float m[16],
                       const float eye[3], const float light[3], uint32_t* image, int threads)
{
    if (!bits || !m || !eye || !light || !image || !N || !width || !height) return -1;
    const int P = (int)((N + 31) / 32);
    const float clear[3] = { 0.0f, 0.2f, 0.4f };
    const float maxDist = 2.0f * sqrtf(3.0f);
    const float stepScale = maxDist / 128.0f, lightStepScale = maxDist / 32.0f;
    const float ll = sqrtf(light[0] * light[0] + light[1] * light[1] + light[2] * light[2]);
    const float ls[3] = { light[0] / ll * lightStepScale, light[1] / ll * lightStepScale, light[2] / ll * lightStepScale };
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#else
    threads = 1;
#endif
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads)
    for (int64_t py = 0; py < (int64_t)height; ++py)
        for (uint32_t px = 0; px < width; ++px)
        {
            const float sx = (float)px + 0.5f, sy = (float)py + 0.5f;
            float h[4];
            for (int c = 0; c < 4; ++c) h[c] = sx * m[c] + sy * m[4 + c] + m[12 + c];
            float pos[3] = { h[0] / h[3], h[1] / h[3], h[2] / h[3] };
            float dir[3] = { pos[0] - eye[0], pos[1] - eye[1], pos[2] - eye[2] };
            const float dl = sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
            dir[0] /= dl; dir[1] /= dl; dir[2] /= dl;
            int hit = 1;
            if (!(fabsf(pos[0]) <= 1.0f && fabsf(pos[1]) <= 1.0f && fabsf(pos[2]) <= 1.0f))
            {
                float U = 3.402823466e+38f;
                hit = 0;
                for (int i = 0; i < 3; ++i)
                {
                    const float sg = dir[i] > 0.0f ? 1.0f : (dir[i] < 0.0f ? -1.0f : 0.0f);
                    const float u = (-sg - pos[i]) / dir[i];
                    if (u < 0.0f) continue;
                    const int j = (i + 1) % 3, k = (i + 2) % 3;
                    if (fabsf(dir[j] * u + pos[j]) > 1.0f) continue;
                    if (fabsf(dir[k] * u + pos[k]) > 1.0f) continue;
                    if (u < U) { U = u; hit = 1; }
                }
                for (int i = 0; i < 3; ++i)
                {
                    float v = dir[i] * U + pos[i];
                    pos[i] = v < -1.0f ? -1.0f : (v > 1.0f ? 1.0f : v);
                }
            }
            uint32_t* out = image + (size_t)py * width + px;
            if (!hit) { *out = unorm8(clear[0]) | (unorm8(clear[1]) << 8) | (unorm8(clear[2]) << 16); continue; }
            float transmit = 1.0f, scatter = 0.0f;
            for (int i = 0; i < 128; ++i)
            {
                if (fabsf(pos[0]) > 1.0f || fabsf(pos[1]) > 1.0f || fabsf(pos[2]) > 1.0f) break;
                const float density = get_sample(bits, (int)N, P, 0.5f * pos[0] + 0.5f, -0.5f * pos[1] + 0.5f, 0.5f * pos[2] + 0.5f);
                if (density > 0.01f)
                {
                    const float scaled = density * stepScale;
                    transmit *= sat(1.0f - scaled * 1.0f);
                    if (transmit < 0.01f) break;
                    float lt = 1.0f;
                    float lp[3] = { pos[0] + ls[0], pos[1] + ls[1], pos[2] + ls[2] };
                    for (int j = 0; j < 32; ++j)
                    {
                        if (fabsf(lp[0]) > 1.0f || fabsf(lp[1]) > 1.0f || fabsf(lp[2]) > 1.0f) break;
                        const float ld = get_sample(bits, (int)N, P, 0.5f * lp[0] + 0.5f, -0.5f * lp[1] + 0.5f, 0.5f * lp[2] + 0.5f);
                        lt *= sat(1.0f - 1.0f * lightStepScale * ld);
                        if (lt < 0.01f) break;
                        lp[0] += ls[0]; lp[1] += ls[1]; lp[2] += ls[2];
                    }
                    scatter += lt * transmit * scaled;
                }
                pos[0] += dir[0] * stepScale; pos[1] += dir[1] * stepScale; pos[2] += dir[2] * stepScale;
            }
            uint32_t rgba = 0xff000000u;
            for (int c = 0; c < 3; ++c)
            {
                float r = scatter * 0.8f + 0.2f;
                r = r + transmit * (clear[c] * clear[c] - r);
                rgba |= unorm8(sqrtf(r)) << (8 * c);
            }
            *out = rgba;
        }
    return 0;
}
