/* dxrv.h -- C ABI of the B200-native solid voxelizer (libdxrv.so).
 *
 * This is the drop-in boundary for the voxelization path of StarsX/DXRVoxelizer.  Every entry
 * point names the reference interface it replaces (paths relative to the reference tree,
 * DXRVoxelizer/...).  Plain pointers and sizes only; no C++/torch types cross this boundary.
 *
 * Conventions
 *   - every call returns DXRV_OK (0) or a negative dxrv_status; the message for the last
 *     failure on a context is available from dxrv_last_error().  Nothing throws or aborts
 *     across the ABI (reference: bool + XUSG_N_RETURN, XUSG/Core/XUSG.h:12-15).
 *   - a context is bound to one CUDA device and one stream; it is NOT thread-safe.  Distinct
 *     contexts are independent (one per GPU / per stream for batches).
 *   - host arrays are borrowed for the duration of the call only.  The context owns all
 *     device memory.  Calls are stream-ordered/asynchronous; dxrv_fetch_grid and
 *     dxrv_synchronize block.
 *   - there is NO CPU fallback: without a usable CUDA device dxrv_create fails.
 *
 * Grid layout (the reference's UAV is RWTexture3D grid[z][y][x], Voxelizer.cpp:62-67 and
 * DXRVoxelizer.hlsl:64-67,83-84; occupancy is its alpha channel, PSRayCast.hlsl:108):
 *   DXRV_FORMAT_BITS : uint32 word[((z - slabBegin) * N + y) * P + (x >> 5)], bit (x & 31),
 *                      P = (N + 31) / 32 words per x-row.  For N % 32 == 0 this is the
 *                      linear bit index ((z*N + y)*N + x).  x, y, z are the reference's
 *                      launch indices, so +y of the grid is -y of the scene (hlsl:49).
 *   DXRV_FORMAT_U8   : uint8 occ[((z - slabBegin) * N + y) * N + x] in {0,1}.
 *   DXRV_FORMAT_R10G10B10A2 : uint32 texel per voxel exactly as the reference's UAV holds it
 *                      (Normal.xyz, 1) -> UNORM 10/10/10/2, untouched texels 0.  Only after
 *                      a DXRV_MODE_SHADER | DXRV_EMIT_TEXELS voxelize.
 */
#ifndef DXRV_H
#define DXRV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DXRV_API __attribute__((visibility("default")))
#else
#define DXRV_API
#endif

typedef struct dxrv_ctx dxrv_ctx;
typedef struct dxrv_mesh dxrv_mesh;

enum dxrv_status
{
    DXRV_OK = 0,
    DXRV_ERR_INVALID_ARG = -1,
    DXRV_ERR_CUDA = -2,
    DXRV_ERR_NO_BVH = -3,   /* dxrv_voxelize before dxrv_build_bvh */
    DXRV_ERR_NO_GRID = -4,  /* dxrv_fetch_grid before dxrv_voxelize */
    DXRV_ERR_IO = -5,
    DXRV_ERR_OOM = -6,
    DXRV_ERR_UNSUPPORTED = -7
};

/* mode argument of dxrv_voxelize (low byte = algorithm, high bits = flags) */
enum dxrv_mode
{
    /* Exact restatement of Content/Shaders/DXRVoxelizer.hlsl: one radial ray per voxel,
     * closest hit, inside iff dot(normalize(interpolated normal), dir) > 0.12. */
    DXRV_MODE_SHADER = 0,
    /* One +x axis ray per (y,z) voxel column, watertight crossings, parity fill.  */
    DXRV_MODE_PARITY = 1,
    DXRV_MODE_MASK = 0xff,
    /* MODE_SHADER only: also produce the R10G10B10A2 texel grid (4 B / voxel). */
    DXRV_EMIT_TEXELS = 0x100
};

enum dxrv_format
{
    DXRV_FORMAT_BITS = 0,
    DXRV_FORMAT_U8 = 1,
    DXRV_FORMAT_R10G10B10A2 = 2
};

/* ---- context ------------------------------------------------------------------------- */

/* Replaces device/command-list acquisition (DXRVoxelizer.cpp:64-169,175-185). */
DXRV_API int dxrv_create(dxrv_ctx** out, int cuda_device);
DXRV_API void dxrv_destroy(dxrv_ctx* ctx);
/* ctx may be NULL: returns the message of the last failed dxrv_create / dxrv_obj_load. */
DXRV_API const char* dxrv_last_error(const dxrv_ctx* ctx);
/* Run all work of this context on an existing cudaStream_t (NULL = the context's own). */
DXRV_API int dxrv_set_stream(dxrv_ctx* ctx, void* cuda_stream);
/* Replaces WaitForGpu (DXRVoxelizer.cpp:485-493). */
DXRV_API int dxrv_synchronize(dxrv_ctx* ctx);

/* ---- acceleration structure ------------------------------------------------------------
 * Replaces Voxelizer::createVB/createIB (Content/Voxelizer.cpp:115-138) and
 * Voxelizer::buildAccelerationStructures (Content/Voxelizer.cpp:264-326): one triangle
 * geometry, float3 positions at `strideBytes` (normals at byte offset 12 when
 * strideBytes >= 24, as ObjLoader lays them out), uint32 indices; the instance transform
 * inverse(Scale(w) * Translate(c)) (Voxelizer.cpp:304-310) is applied as p' = (p - c) / w.
 * bound = {cx, cy, cz, w}; NULL => computed on the device exactly as Voxelizer.cpp:52-57
 * does from ObjLoader::computeAABB (min/max over ALL vertices).
 * The build is an LBVH: bounds -> 30-bit Morton keys -> onesweep radix sort -> sorted triangle
 * records and leaf boxes -> Karras hierarchy -> child boxes (range unions; atomic bottom-up refit
 * above 2^19 triangles).  RULE for the last two steps (the hierarchy): they run inside this call when
 * the previous dxrv_voxelize of the context traversed the hierarchy, and otherwise are deferred to the
 * first dxrv_voxelize that does.  MODE_SHADER traverses it (LBVH walk on small grids, and as the
 * overflow path of the direction bins).  MODE_PARITY does not by default: its candidates come from
 * triangle-parallel passes over the Morton-sorted triangle records (2-D binning into column tiles, or
 * the scatter path for fine meshes on coarse grids); DXRV_PARITY_CANDIDATES=walk selects the LBVH walk
 * (k_walk_columns) instead.  Grids are identical bit for bit whichever way the candidates are found.
 * The host variant copies the arrays to the device. */
DXRV_API int dxrv_build_bvh(dxrv_ctx* ctx, const void* vertices, uint32_t numVerts,
                            uint32_t strideBytes, const uint32_t* indices, uint32_t numIndices,
                            const float bound[4]);
/* Same, with vertices/indices already resident in device memory of this context's GPU
 * (they are read in place and must stay valid until the next build). */
DXRV_API int dxrv_build_bvh_device(dxrv_ctx* ctx, const void* d_vertices, uint32_t numVerts,
                                   uint32_t strideBytes, const uint32_t* d_indices,
                                   uint32_t numIndices, const float bound[4]);
/* {cx, cy, cz, w} used by the last build (synchronises). */
DXRV_API int dxrv_get_bound(dxrv_ctx* ctx, float bound[4]);

/* ---- voxelization ----------------------------------------------------------------------
 * Replaces Voxelizer::voxelize (Content/Voxelizer.cpp:351-369): DispatchRays(N, N*N, 1) of
 * raygenMain/closestHitMain/missMain.  N replaces the GRID_SIZE macro (Voxelizer.cpp:8).
 * Computes grid layers z in [slabBegin, slabEnd) (0 <= slabBegin <= slabEnd <= N; an empty slab
 * computes nothing and only lets the context take part in dxrv_gather_grid).  Every word of the
 * slab is written exactly once; no clear is needed. */
DXRV_API int dxrv_voxelize(dxrv_ctx* ctx, uint32_t N, uint32_t mode, uint32_t slabBegin,
                           uint32_t slabEnd);
/* Copy the slab computed by the last dxrv_voxelize to host memory (synchronises).
 * bytes must equal the slab size in `format`.  The reference has no read-back of the grid
 * (only of the back buffer, DXRVoxelizer.cpp:436,476); this is the headless replacement. */
DXRV_API int dxrv_fetch_grid(dxrv_ctx* ctx, void* hostDst, size_t bytes, uint32_t format);
/* dxrv_voxelize + dxrv_fetch_grid(DXRV_FORMAT_BITS) as one call that returns when hostDst holds the whole slab (dense BITS
 * layout, bytes = the slab size); the context then describes that slab exactly as after dxrv_voxelize.  Not available
 * with an external grid target.  How the grid travels is chosen with dxrv_set_read_back (or DXRV_TO_HOST=dense|sparse):
 *   DXRV_READ_BACK_DENSE   the slab is computed in `chunks` z sub-slabs (0 or 1 = no pipelining) and each is copied to
 *                          hostDst (pinned memory: dxrv_host_alloc) on a second stream while the next one is computed.
 *                          Floor: the dense grid over the PCIe link (128 MiB at 1024^3: 2.4 ms).
 *   DXRV_READ_BACK_SPARSE  the slab is encoded as DXRV_FORMAT_SPARSE_BRICKS on the device, the blob (a few MB) is copied
 *                          to pinned staging memory of the context and expanded into hostDst by a pool of host threads
 *                          (DXRV_HOST_THREADS, default all cores up to 32) in ONE pass of streaming stores: they start
 *                          zeroing hostDst from the outside of the slab inwards while the GPU is still computing and,
 *                          once the blob has arrived, write the remaining brick layers with their final contents (the
 *                          workers poll for the next call for DXRV_HOST_SPIN_US microseconds, default 2000, before they
 *                          sleep).  Floor: the host's memory write bandwidth.  A grid whose blob exceeds half the
 *                          dense size is copied densely after all.  hostDst is bit-identical either way.
 *   DXRV_READ_BACK_AUTO    (default) SPARSE for slabs of 8 MiB and more when the pool has at least 4 host threads, else DENSE. */
#define DXRV_READ_BACK_AUTO   0u
#define DXRV_READ_BACK_DENSE  1u
#define DXRV_READ_BACK_SPARSE 2u
DXRV_API int dxrv_set_read_back(dxrv_ctx* ctx, uint32_t transport);
DXRV_API int dxrv_voxelize_to_host(dxrv_ctx* ctx, uint32_t N, uint32_t mode, uint32_t slabBegin, uint32_t slabEnd,
                                   void* hostDst, size_t bytes, uint32_t chunks);
/* dxrv_build_bvh + dxrv_voxelize_to_host as ONE call -- a frame of a deforming mesh: upload, rebuild, voxelize, read
 * back (the reference rebuilds nothing per frame, Voxelizer.cpp:264-326 runs once; this is the call the "incl. BVH build"
 * metric times end to end).  Same arguments, results and errors as the two calls in sequence; the host arrays are
 * borrowed for the duration of the call.  With the compact transport the host threads start on hostDst before the
 * upload and the build instead of after them. */
DXRV_API int dxrv_voxelize_mesh_to_host(dxrv_ctx* ctx, const void* vertices, uint32_t numVerts, uint32_t strideBytes,
                                        const uint32_t* indices, uint32_t numIndices, const float bound[4], uint32_t N,
                                        uint32_t mode, uint32_t slabBegin, uint32_t slabEnd, void* hostDst, size_t bytes,
                                        uint32_t chunks);
/* DXRV_FORMAT_SPARSE_BRICKS: a lossless compact form of the slab's BITS grid for consumers that can take it (a solid
 * voxelization is almost all empty space and solid interior; the dense grid is already at the PCIe roofline).  Bricks
 * of 32 (x) x 4 (y) x 4 (z) voxels = 16 words of the dense grid; brick b = (bz * BY + by) * P + bx.  Blob:
 *   header  16 uint32: "DXRB", version 1, N, slabBegin, slabEnd, P, BY, BZ, bricks, mixed bricks, byte offset of the
 *           states, byte offset of the payload, brick dims 32, 4, 4, 0
 *   states  2 bits per brick, 16 per uint32 (brick b: bits 2 (b & 15) of word b >> 4): 0 empty, 1 full, 2 mixed
 *   payload 16 words per mixed brick {z = 4 bz + k {y = 4 by + j}}, in brick order
 * dxrv_fetch_grid_sparse encodes the slab of the last voxelize on the device and copies exactly the bytes that exist;
 * *bytesWritten receives their number (also when `capacity` is too small: DXRV_ERR_INVALID_ARG, nothing but the header
 * copied).  dxrv_sparse_decode (pure host code, on the same pool of host threads) expands a blob into the dense BITS
 * layout. */
DXRV_API int dxrv_fetch_grid_sparse(dxrv_ctx* ctx, void* hostDst, size_t capacity, size_t* bytesWritten);
DXRV_API int dxrv_sparse_decode(const void* blob, size_t blobBytes, void* denseDst, size_t denseBytes);
/* The encoder on the host (pure host code, same pool): the slab [slabBegin, slabEnd) of an N^3 BITS grid at `dense`
 * (denseBytes = layers * N * ceil(N/32) * 4) -> the same bytes the device encoder writes.  *bytesWritten receives the
 * blob's size -- also when `capacity` is too small (DXRV_ERR_INVALID_ARG, nothing written; blob may then be NULL). */
DXRV_API int dxrv_sparse_encode(const void* dense, size_t denseBytes, uint32_t N, uint32_t slabBegin, uint32_t slabEnd,
                                void* blob, size_t capacity, size_t* bytesWritten);
/* Device pointer / byte size of the slab's DXRV_FORMAT_BITS grid (valid until the next
 * voxelize with a different size, or destroy). */
DXRV_API int dxrv_grid_device(dxrv_ctx* ctx, void** d_ptr, size_t* bytes);
/* Make subsequent dxrv_voxelize calls write the BITS grid into caller-owned device memory
 * (may be peer-mapped memory of another GPU: the fill kernel's 128-bit stores then go over
 * NVLink, fusing the slab gather into the write).  NULL restores the internal grid. */
DXRV_API int dxrv_set_grid_target(dxrv_ctx* ctx, void* d_ptr, size_t bytes);
/* Occupancy pyramid of the slab of the last voxelize (BITS layout per level): level 0 is the grid itself,
 * a voxel of level l+1 is set when any of its 2x2x2 children in level l is.  Levels are built while the
 * grid size and the slab's layer count stay even; *numLevels receives how many exist (>= 1).  The reference
 * samples its grid texture at mip SHOW_MIP (Content/SharedConst.h:5, PSRayCast.hlsl:106-108) and never
 * builds the chain; this is the consumer-side format for empty-space skipping. */
DXRV_API int dxrv_build_mips(dxrv_ctx* ctx, uint32_t* numLevels);
/* Copy level `level` (1 <= level < numLevels; level 0 is dxrv_fetch_grid) to host memory.  bytes must be
 * layers_l * N_l * ceil(N_l / 32) * 4 with N_l = N >> level, layers_l = (slabEnd - slabBegin) >> level. */
DXRV_API int dxrv_fetch_mip(dxrv_ctx* ctx, uint32_t level, void* hostDst, size_t bytes);
/* ---- viewer pass (the step right after the path; headless) ---------------------------------------------
 * Camera of the reference (DXRVoxelizer.cpp:19-23,222-234) and the per-object constants of
 * Voxelizer::UpdateFrame (Content/Voxelizer.cpp:81-106) for a width x height target: screenToLocal in the
 * row-vector convention (p' = (x,y,z,1) * M, as PSRayCast.hlsl:63 uses it), eye and light point in the
 * grid's local space.  posScale may be NULL (0,0,0,1).  Pure host code. */
DXRV_API int dxrv_default_view(const float bound[4], const float posScale[4], uint32_t width, uint32_t height,
                               float screenToLocal[16], float eye[3], float light[3]);
/* Ray-march the FULL grid of the last dxrv_voxelize (slab [0, N)) exactly like PSRayCast.hlsl:117-187
 * (128 steps, 32 light steps, trilinear LINEAR_CLAMP sampling of the occupancy) into an R8G8B8A8 image
 * (row-major, y down, bytes = width * height * 4).  Synchronises. */
DXRV_API int dxrv_render_view(dxrv_ctx* ctx, uint32_t width, uint32_t height, const float screenToLocal[16],
                              const float eye[3], const float light[3], void* hostRGBA, size_t bytes);
/* DXRVoxelizer::SaveImage (DXRVoxelizer.cpp:531-551; the F11 screenshot, stbi_write_png there): an 8-bit PNG of
 * comp = 3 (RGB, the reference's default) or 4 (RGBA) channels from an R8G8B8A8 image with rowPitchBytes >= width * 4
 * -- e.g. the output of dxrv_render_view.  Pure host code. */
DXRV_API int dxrv_save_image(const char* fileName, const void* rgba, uint32_t width, uint32_t height, uint32_t rowPitchBytes,
                             uint32_t comp);
/* Number of set voxels in the slab of the last voxelize (device popcount; synchronises). */
DXRV_API int dxrv_count_inside(dxrv_ctx* ctx, uint64_t* count);

/* ---- introspection ----------------------------------------------------------------------- */
enum dxrv_info
{
    DXRV_INFO_NUM_TRIANGLES = 0,
    DXRV_INFO_NUM_NODES = 1,
    DXRV_INFO_KERNEL_LAUNCHES = 2, /* kernels launched by this context so far */
    DXRV_INFO_CROSSINGS = 3,       /* MODE_PARITY: surface crossings found by the last voxelize */
    DXRV_INFO_SM_COUNT = 4,
    DXRV_INFO_LAST_WALK_NS = 5,    /* device time of the last k_walk_columns (needs dxrv_set_profiling) */
    DXRV_INFO_LAST_FILL_NS = 6,    /* device time of the last k_trace_fill_columns                       */
    DXRV_INFO_LAST_BUILD_NS = 7,   /* device time of the last acceleration-structure build (needs dxrv_set_profiling) */
    DXRV_INFO_LAST_SORT_NS = 8,    /* ... of its onesweep radix-sort passes alone                        */
    DXRV_INFO_LAST_D2H_BYTES = 9   /* bytes the last dxrv_voxelize_to_host / dxrv_fetch_grid[_sparse] copied device -> host */
};
DXRV_API int dxrv_get_info(dxrv_ctx* ctx, uint32_t what, uint64_t* value);
/* Record CUDA events around the MODE_PARITY kernels of every dxrv_voxelize and around the phases of every build
 * (for roofline reporting; builds then run outside their CUDA graph). */
DXRV_API int dxrv_set_profiling(dxrv_ctx* ctx, int enable);

/* Read back an internal device buffer for tests (synchronises).  `what`: */
enum dxrv_debug_buffer
{
    DXRV_DBG_MORTON_SORTED = 0, /* uint32[T]   sorted Morton keys                          */
    DXRV_DBG_PRIM_SORTED = 1,   /* uint32[T]   original triangle index per sorted slot      */
    DXRV_DBG_NODES = 2,         /* 64 B * (T-1) internal nodes, see csrc/common.cuh           */
    DXRV_DBG_TRIS = 3,          /* 48 B * T    normalised triangles in sorted order         */
    DXRV_DBG_ROOT_BOX = 5,      /* float[6]    lo.xyz hi.xyz of the root                    */
    DXRV_DBG_BINS_STATE = 6     /* uint32[4]   MODE_SHADER direction bins of the last voxelize: entries, overflow
                                   flag (1 = the LBVH walk produced the grid), near-list length, cells per face edge */
};
DXRV_API int dxrv_debug_read(dxrv_ctx* ctx, uint32_t what, void* hostDst, size_t bytes);
/* Standalone key/value radix sort on the device (the LBVH's onesweep), for tests/bench:
 * sorts n (key, value) pairs given as host arrays, in place. */
DXRV_API int dxrv_debug_sort_pairs(dxrv_ctx* ctx, uint32_t* keys, uint32_t* values, uint32_t n);

/* ---- mesh input ---------------------------------------------------------------------------
 * Replaces XUSG::ObjLoader::Import(fileName, true, true) as called from Voxelizer::Init
 * (Content/Voxelizer.cpp:46-47; XUSG/Optional/XUSGObjLoader.cpp:18-40): byte-identical
 * interleaved {float3 pos; float3 nrm} vertices, uint32 indices (z flipped, index array
 * reversed), AABB.  Pure host code. */
DXRV_API int dxrv_obj_load(const char* path, dxrv_mesh** out);
/* The same from OBJ text in memory (`size` bytes, no terminator needed): a mesh that arrives over a socket or from an
 * archive.  Identical output to dxrv_obj_load of a file with these bytes. */
DXRV_API int dxrv_obj_parse(const char* text, size_t size, dxrv_mesh** out);
DXRV_API void dxrv_obj_free(dxrv_mesh* mesh);
DXRV_API uint32_t dxrv_obj_num_vertices(const dxrv_mesh* mesh);
DXRV_API uint32_t dxrv_obj_num_indices(const dxrv_mesh* mesh);
DXRV_API uint32_t dxrv_obj_vertex_stride(const dxrv_mesh* mesh);
DXRV_API const void* dxrv_obj_vertices(const dxrv_mesh* mesh);
DXRV_API const uint32_t* dxrv_obj_indices(const dxrv_mesh* mesh);
/* out = {min.x, min.y, min.z, max.x, max.y, max.z} (ObjLoader::GetAABB) */
DXRV_API void dxrv_obj_aabb(const dxrv_mesh* mesh, float out[6]);
/* out = {cx, cy, cz, w} as Voxelizer::Init derives it (Content/Voxelizer.cpp:52-57) */
DXRV_API void dxrv_obj_bound(const dxrv_mesh* mesh, float out[4]);

/* ---- a stream of distinct meshes (BASELINE config 5) -----------------------------------------
 * LoadAssets + voxelize per mesh (DXRVoxelizer.cpp:190-199, Voxelizer.cpp:351-369) as a pipeline inside
 * the library: `loaderThreads` host threads parse the OBJ files (0: one per core, <= 32; one thread per
 * file); context s -- driven by its own host thread -- takes meshes s, s + numCtx, ... in order:
 * dxrv_build_bvh (bound = NULL), dxrv_voxelize(N, mode, 0, N) and, when hostGrids is not NULL,
 * dxrv_fetch_grid of mesh k into hostGrids + k * gridBytes (DXRV_FORMAT_BITS; gridBytes =
 * N * N * ceil(N/32) * 4; pinned memory copies fastest).  mode is DXRV_MODE_SHADER or DXRV_MODE_PARITY
 * (bit grids; no flags).  The contexts must be distinct (typically 4 per GPU; they may sit on different
 * GPUs) and are left holding their last mesh.  numTriangles (optional): triangles of every mesh.  Parsed
 * meshes waiting for a stream are bounded, whatever the batch size.  Returns the first failure (the
 * message names the file; dxrv_last_error(NULL)); grids of meshes processed before it are valid. */
DXRV_API int dxrv_voxelize_obj_batch(dxrv_ctx* const* ctxs, uint32_t numCtx, const char* const* paths, uint32_t numMeshes,
                                     uint32_t N, uint32_t mode, void* hostGrids, size_t gridBytes, uint32_t loaderThreads,
                                     uint32_t* numTriangles);

/* ---- pinned host staging (optional; lets dxrv_build_bvh / dxrv_fetch_grid run at full
 * PCIe speed; replaces the upload heaps of Voxelizer.cpp:121,134) ------------------------- */
DXRV_API void* dxrv_host_alloc(size_t bytes);
DXRV_API void dxrv_host_free(void* p);

/* ---- CUDA IPC for the fused slab gather (one process per GPU) ----------------------------- */
/* Export a handle (64 bytes) for the internal full-size grid of `fullBytes` bytes, allocating
 * it if needed; another process opens it with dxrv_ipc_open and passes the pointer (plus its
 * slab offset) to dxrv_set_grid_target. */
DXRV_API int dxrv_ipc_export_grid(dxrv_ctx* ctx, size_t fullBytes, void* handle64, void** d_ptr);
DXRV_API int dxrv_ipc_open(dxrv_ctx* ctx, const void* handle64, void** d_ptr);
DXRV_API int dxrv_ipc_close(dxrv_ctx* ctx, void* d_ptr);

/* ---- multi-GPU: z-slab sharding over the GPUs of one box -----------------------------------------------------
 * New relative to the reference, which is single-GPU (XUSG/RayTracing/XUSGRayTracing.h:386 SetNodeMask is never
 * called); it replaces nothing there and extends the lower surface XUSGRayTracing.h:170-180,212-230,325-330
 * (BottomLevelAS::Build / TopLevelAS::Build / DispatchRays) to several devices: the mesh is replicated by NCCL
 * broadcast over NVLink, every GPU builds the identical LBVH (the sort is stable, the boxes are exact unions) and
 * fills its own z-slab (dxrv_voxelize slab arguments); slabs are gathered only when a full grid is requested.
 * NCCL is loaded with dlopen("libnccl.so.2") by the first call below; DXRV_ERR_UNSUPPORTED if it is not installed.
 * One process per GPU: rank 0 calls dxrv_comm_get_unique_id and ships the 128 bytes to the other ranks (any
 * out-of-band channel), every rank calls dxrv_comm_init.  One process, several GPUs: dxrv_comm_init_all, and
 * collective calls of several contexts from one thread go between dxrv_group_begin / dxrv_group_end. */
#define DXRV_COMM_ID_BYTES 128
DXRV_API int dxrv_comm_get_unique_id(void* id128);
DXRV_API int dxrv_comm_init(dxrv_ctx* ctx, const void* id128, int rank, int world);
DXRV_API int dxrv_comm_init_all(dxrv_ctx** ctxs, int count);
DXRV_API int dxrv_comm_destroy(dxrv_ctx* ctx);
DXRV_API int dxrv_group_begin(void);
DXRV_API int dxrv_group_end(void);
/* Broadcast count (<= 64) words from root's host array into every rank's (blocking; carries the mesh sizes). */
DXRV_API int dxrv_bcast_u32(dxrv_ctx* ctx, uint32_t* values, uint32_t count, int root);
/* Replicate root's mesh into context-owned device buffers of every rank: H2D on the root (host arrays borrowed
 * for the call), ncclBroadcast of vertices and indices.  numVerts / strideBytes / numIndices must be passed by
 * every rank; vertices / indices are read on the root only.  Collective, stream-ordered. */
DXRV_API int dxrv_bcast_mesh(dxrv_ctx* ctx, const void* vertices, uint32_t numVerts, uint32_t strideBytes,
                             const uint32_t* indices, uint32_t numIndices, int root);
/* dxrv_build_bvh on the mesh the last dxrv_bcast_mesh left in the context (Voxelizer.cpp:264-326 on every GPU). */
DXRV_API int dxrv_build_bvh_replicated(dxrv_ctx* ctx, const float bound[4]);
/* Gather the z-slabs of every rank's last dxrv_voxelize (same N, disjoint slabs) into the full N^3 BITS grid, in
 * device memory owned by the context, on `root` (ncclSend/ncclRecv) or on every rank when root < 0 (one
 * ncclBroadcast per slab).  Layers no rank computed are zero.  Collective; synchronises once (slab table). */
DXRV_API int dxrv_gather_grid(dxrv_ctx* ctx, int root);
/* Same with the slab table {z0, z1} of every rank supplied by the caller: no internal exchange and no host
 * synchronisation, so it may be called for several contexts inside one dxrv_group_begin / dxrv_group_end. */
DXRV_API int dxrv_gather_grid_slabs(dxrv_ctx* ctx, int root, const uint32_t* slabs);
DXRV_API int dxrv_full_grid_device(dxrv_ctx* ctx, void** d_ptr, size_t* bytes);
DXRV_API int dxrv_fetch_full_grid(dxrv_ctx* ctx, void* hostDst, size_t bytes);
/* Fused gather for contexts of ONE process: make ctx's next dxrv_voxelize(N, mode, slabBegin, slabEnd) store its
 * slab straight into `owner`'s full grid (peer access over NVLink is enabled; owner may be ctx itself), so the
 * full grid is complete on the owner when every context's stream has drained -- no collective, one D2H. */
DXRV_API int dxrv_share_grid_target(dxrv_ctx* ctx, dxrv_ctx* owner, uint32_t N, uint32_t slabBegin, uint32_t slabEnd);

#ifdef __cplusplus
}
#endif
#endif /* DXRV_H */
