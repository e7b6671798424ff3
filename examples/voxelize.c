/* examples/voxelize.c -- the C ABI from plain C: what a maintainer's FFI (cgo, JNI, P/Invoke ...) binds.
 *
 *   gcc -std=c99 -Iinclude examples/voxelize.c -Ldxrvoxelizer_b200 -ldxrv -Wl,-rpath,$PWD/dxrvoxelizer_b200 -o /tmp/voxelize
 *   /tmp/voxelize mesh.obj [N] [view.png]
 *
 * Loads the mesh as ObjLoader::Import does, builds the acceleration structure, voxelizes with the reference's shader
 * function (DXRV_MODE_SHADER), counts the solid voxels and, optionally, writes the reference's viewer image as a PNG --
 * Voxelizer::Init + voxelize + Render of the reference (Content/Voxelizer.cpp:30-79,351-399), headless.
 */
#include <stdio.h>
#include <stdlib.h>

#include "dxrv.h"

#define CHECK(call, ctx)                                                                  \
    do {                                                                                  \
        int rc_ = (call);                                                                 \
        if (rc_ != DXRV_OK) {                                                             \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, dxrv_last_error(ctx));    \
            return 1;                                                                     \
        }                                                                                 \
    } while (0)

int main(int argc, char** argv)
{
    const char* path = argc > 1 ? argv[1] : "Assets/bunny.obj";
    const uint32_t N = argc > 2 ? (uint32_t)strtoul(argv[2], NULL, 10) : 64u;   /* GRID_SIZE, Voxelizer.cpp:8 */
    const char* png = argc > 3 ? argv[3] : NULL;
    dxrv_mesh* mesh = NULL;
    dxrv_ctx* ctx = NULL;
    uint64_t inside = 0;
    float bound[4];

    CHECK(dxrv_obj_load(path, &mesh), NULL);
    dxrv_obj_bound(mesh, bound);
    printf("%s: %u vertices, %u triangles, bound (%g, %g, %g; %g)\n", path, dxrv_obj_num_vertices(mesh), dxrv_obj_num_indices(mesh) / 3u,
           bound[0], bound[1], bound[2], bound[3]);
    CHECK(dxrv_create(&ctx, 0), NULL);
    CHECK(dxrv_build_bvh(ctx, dxrv_obj_vertices(mesh), dxrv_obj_num_vertices(mesh), dxrv_obj_vertex_stride(mesh), dxrv_obj_indices(mesh),
                         dxrv_obj_num_indices(mesh), NULL),
          ctx);
    dxrv_obj_free(mesh);                                   /* the host arrays were only borrowed for the call */
    CHECK(dxrv_voxelize(ctx, N, DXRV_MODE_SHADER, 0, N), ctx);
    CHECK(dxrv_count_inside(ctx, &inside), ctx);
    printf("%u^3 grid: %llu solid voxels\n", N, (unsigned long long)inside);
    if (png)
    {
        const uint32_t w = 1280, h = 720;                  /* the reference's window, Main.cpp:17 */
        float screenToLocal[16], eye[3], light[3];
        unsigned char* rgba = (unsigned char*)malloc((size_t)w * h * 4);
        if (!rgba) return 1;
        CHECK(dxrv_default_view(bound, NULL, w, h, screenToLocal, eye, light), NULL);
        CHECK(dxrv_render_view(ctx, w, h, screenToLocal, eye, light, rgba, (size_t)w * h * 4), ctx);
        CHECK(dxrv_save_image(png, rgba, w, h, w * 4, 3), NULL);
        free(rgba);
        printf("wrote %s\n", png);
    }
    dxrv_destroy(ctx);
    return 0;
}
