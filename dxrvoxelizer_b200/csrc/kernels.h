// kernels.h -- host-callable launchers of the sm_100a kernels (internal to libdxrv.so).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace dxrv
{
struct MeshView
{
    const uint8_t* verts;     // device, interleaved, float3 position first
    uint32_t numVerts;
    uint32_t stride;          // bytes
    const uint32_t* indices;  // device
    uint32_t numTris;
};

// ---- lbvh.cu ------------------------------------------------------------------------------------
constexpr int kBoundsMaxBlocks = 1024;
// scratch: 6 * kBoundsMaxBlocks floats + 1 counter (zero before first use; self-resetting)
void launchBounds(cudaStream_t s, const MeshView& m, float* dBound, float* dPartials, uint32_t* dCounter);
void launchSetBound(cudaStream_t s, float cx, float cy, float cz, float w, float* dBound);
// keys = 30-bit Morton code >> keyShift; also accumulates the digit histograms of `numPasses` radix
// passes into hist (which must be zero).
void launchMorton(cudaStream_t s, const MeshView& m, const float* dBound, uint32_t* keys, uint32_t* vals,
                  uint32_t keyShift, int numPasses, uint32_t* hist, uint32_t* dErr);
// The whole build of a small mesh (bounds, keys, stable radix sort[, sorted triangle records]) in ONE cooperative kernel
// with grid barriers; same output as launchBounds/launchSetBound + launchMorton + radixSortPairs (+ the records of
// k_leaf_setup when tris != null; the box pyramid is NOT written).  fusedBuildPlan: does the mesh fit (<= 2 triangles
// per thread of one 1024-thread CTA per SM)?  scratch: fusedBuildScratchBytes(), zero before its first use.
// bnd: null = compute {c, w} from the vertices.  Returns false when nothing was launched (the caller takes the
// multi-kernel path).
constexpr int kMaxFusedPasses = 4;
constexpr uint32_t kFusedBuildMinTris = 49152;   // below: the multi-kernel build is as fast (measured)
size_t fusedBuildScratchBytes(int smCount);
bool fusedBuildPlan(uint32_t numTris, int smCount, uint32_t& ctas, uint32_t& rounds);
bool fusedBuildSupported(int device);
bool launchFusedBuild(cudaStream_t s, const MeshView& m, int smCount, const float* bnd, float* dBound, float* dPartials, void* scratch,
                      uint32_t* keysA, uint32_t* valsA, uint32_t* keysB, uint32_t* valsB, Tri48* tris, uint32_t keyShift, int numPasses, uint32_t* dErr);
constexpr int kMaxBoxLevels = 8;
// float4 entries needed for the leaf-box pyramid of numTris leaves
size_t boxPyramidFloat4s(uint32_t numTris);
// k_leaf_setup (+ k_box_level) beside k_hierarchy_topology, then k_node_boxes (small meshes: child boxes by range
// union) or k_refit_atomic (large meshes: bottom-up refit with atomics); returns the number of
// kernels launched.  refitScratch: 3 * numTris uint32, only touched when useAtomicRefit(numTris).
bool useAtomicRefit(uint32_t numTris);
// side (nullable): a second stream + two events; with it the leaf/pyramid kernels run beside the topology kernel.
struct SideStream { cudaStream_t stream; cudaEvent_t fork, join; };
// parts: kBuildLeaves = sorted triangle records + leaf boxes + box pyramid; kBuildTree = topology + child boxes (needs the
// leaves).  Both together run the two chains concurrently (fork/join).  The tree is only needed by consumers that
// TRAVERSE it (tile-path MODE_PARITY, MODE_SHADER); the scatter path reads the sorted triangle records alone.
constexpr int kBuildLeaves = 1, kBuildTree = 2;
int launchLeavesAndHierarchy(cudaStream_t s, const SideStream* side, const MeshView& m, const float* dBound, const uint32_t* sortedKeys,
                             const uint32_t* sortedPrims, BvhNode* nodes, Tri48* tris, float4* pyramidMem,
                             uint32_t* refitScratch, float* rootBox, uint32_t* dErr, int parts);

// ---- onesweep.cu --------------------------------------------------------------------------------
struct SortTemp
{
    uint32_t* hist;         // [4][256]   digit histograms -> exclusive digit bases
    uint32_t* tileCounter;  // [4]
    uint32_t* lookback;     // [4][tiles][256]
    uint32_t tilesCapacity;
    static size_t bytesFor(uint32_t n);
    static uint32_t tilesFor(uint32_t n);
};
// Zero the sort scratch (histograms, tile counters, look-back words); returns the histogram pointer
// ([4][256] digit counts) for a producer kernel to fill.
uint32_t* sortClearTemp(cudaStream_t s, void* tempBase, uint32_t n);
// Stable LSD radix sort of (key,value) pairs, `numPasses` passes of 8 bits from bit 0.  histReady: the
// digit counts are already in the scratch (sortClearTemp + producer); otherwise they are computed here.
// The result is in keysB/valsB when numPasses is odd (*resultInB), else in keysA/valsA.
// tempBase: device memory of SortTemp::bytesFor(n) bytes.  Returns the number of kernels launched.
int radixSortPairs(cudaStream_t s, void* tempBase, uint32_t* keysA, uint32_t* valsA, uint32_t* keysB,
                   uint32_t* valsB, uint32_t n, int numPasses, bool histReady, bool* resultInB);

// ---- trace_parity.cu ----------------------------------------------------------------------------
struct BvhView
{
    const BvhNode* nodes;   // null: no hierarchy (MODE_PARITY then bins its candidates instead of walking)
    const Tri48* tris;
    const float* rootBox;   // lo.xyz hi.xyz
    uint32_t numTris;
};
// MODE_PARITY: +x column rays for layers z in [z0,z1); writes every word of the slab exactly once.
// walkBuf: device scratch of parityScratchWords() uint32; its first parityScratchZeroWords() words
// must be zero the first time it is used (the kernels leave them zero again).
// Returns the number of kernels launched.
void parityTileCounts(uint32_t N, uint32_t z0, uint32_t z1, uint32_t& numTiles, uint32_t& candCap);
size_t parityScratchWords(uint32_t N, uint32_t z0, uint32_t z1);
size_t parityScratchZeroWords(uint32_t N);
int launchTraceFillColumns(cudaStream_t s, const BvhView& bvh, uint32_t N, uint32_t z0, uint32_t z1,
                           uint32_t* grid, uint32_t* walkBuf, unsigned long long* dCrossings, uint32_t* dErr,
                           cudaEvent_t* ev /* nullable: {before walk, between, after fill} */,
                           bool binsReady = false /* the candidate lists of exactly this grid / slab / structure are in walkBuf already */);

// ---- scatter_parity.cu --------------------------------------------------------------------------
// MODE_PARITY for meshes that are fine relative to the grid (useScatterParity): triangle-parallel scatter of
// toggle bits + in-place prefix pass; same result, bit for bit, as launchTraceFillColumns.
bool useScatterParity(uint32_t numTris, uint32_t N);
int launchScatterFillColumns(cudaStream_t s, const BvhView& bvh, uint32_t N, uint32_t z0, uint32_t z1, uint32_t* grid,
                             uint32_t* walkBuf /* the tile path's scratch, same contract */, unsigned long long* dCrossings,
                             cudaEvent_t* ev /* nullable, as above */);

// ---- shader_bins.cu / trace_shader.cu ---------------------------------------------------------------
// MODE_SHADER: one radial closest-hit ray per voxel (DXRVoxelizer.hlsl raygenMain/closestHitMain).
// Default path: direction bins (shader_bins.cu) built once per acceleration structure; the LBVH walk
// (trace_shader.cu) runs when the bins' device-side overflow flag is up.  Both launches are always enqueued;
// exactly one of the two kernels does the work (the other returns at once), so no host round trip is needed.
struct ShaderBinsView
{
    uint4* cells;         // [6 R^2] {first entry, count | unsorted flag, bits(max rmax), bits((max rmax)^2)}
    uint4* entries;       // [cap]   {bits(rmin), bits(rmax), triangle slot, bits(min rmin of this and all later entries)}
    uint32_t* cursors;    // [6 R^2] counts -> local exclusive offsets -> local end offsets
    uint32_t* blockSums;  // [numBlocks] exclusive base of every 2048-cell tile
    uint32_t* nearList;   // [nearCap] slots of the triangles closer than 1e-3 to the grid centre
    uint4* bigRects;      // [2 * bigCap] rectangles of more than 64 cells: {slot, face, iu0, nu}, {iv0, cells, bits(rmin), bits(rmax)}
    uint32_t* bigCells;   // [bigCap] cells whose lists are sorted by a whole warp
    uint32_t* state;      // [0] total entries, [1] overflow flag, [2] near count, [3] big rectangles, [4] big cells
    uint32_t R, cap, nearCap, bigCap;
};
struct ShaderBinsSizes
{
    uint32_t R, cap, nearCap, bigCap, numBlocks;
    size_t offCells, offEntries, offCursors, offBlockSums, offNear, offBigRects, offBigCells, offState, bytes;
};
uint32_t shaderBinsResolution(uint32_t numTris);
ShaderBinsSizes shaderBinsSizes(uint32_t numTris);
ShaderBinsView shaderBinsView(void* base, const ShaderBinsSizes& s);
// returns the number of kernels launched; forceOverflow: build nothing, raise the flag (LBVH walk only)
int launchShaderBinsBuild(cudaStream_t s, const BvhView& bvh, void* base, const ShaderBinsSizes& sz, bool forceOverflow);
// texels may be null.  verts/indices are the ORIGINAL buffers (normals at byte offset 12).
void launchTraceShaderBins(cudaStream_t s, const BvhView& bvh, const MeshView& m, uint32_t N, uint32_t z0, uint32_t z1,
                           uint32_t* grid, uint32_t* texels, uint32_t* dErr, void* base, const ShaderBinsSizes& sz,
                           float* centres /* device scratch, >= N floats: the voxel-centre table */);
void launchTraceShaderBvh(cudaStream_t s, const BvhView& bvh, const MeshView& m, uint32_t N, uint32_t z0, uint32_t z1,
                          uint32_t* grid, uint32_t* texels, uint32_t* dErr, const uint32_t* binsState /* nullable: always run */);

// ---- sparse.cu ------------------------------------------------------------------------------------------
// DXRV_FORMAT_SPARSE_BRICKS (include/dxrv.h): header | 2-bit brick states | 64-byte payload of the mixed bricks
struct SparseLayout
{
    uint32_t P, BY, BZ, numBricks, stateWords, numBlocks;
    size_t offStates, offPayload, maxBytes;
};
constexpr uint32_t kSparseBlockBricks = 256;   // bricks per block of the encoder; after the encode blockCounts[i] = rank of block i's first mixed brick
SparseLayout sparseLayout(uint32_t N, uint32_t layers);
// blob: device memory of sparseLayout().maxBytes; blockCounts: numBlocks words of scratch.  Returns kernels launched.
int launchSparseEncode(cudaStream_t s, const uint32_t* grid, uint32_t N, uint32_t z0, uint32_t z1, uint8_t* blob, uint32_t* blockCounts);

// ---- view.cu (headless port of the reference's viewer pass, PSRayCast.hlsl) -----------------------------
void launchRaycastView(cudaStream_t s, const uint32_t* grid, uint32_t N, uint32_t width, uint32_t height,
                       const float screenToLocal[16], const float eye[3], const float light[3], uint32_t* image);

// ---- misc (lbvh.cu) -----------------------------------------------------------------------------
void launchPopcount(cudaStream_t s, const uint32_t* words, size_t numWords, unsigned long long* dCount);
// dst = next level of the occupancy pyramid of src (Ns^2 x layersSrc voxels; Ns and layersSrc even)
void launchMipReduce(cudaStream_t s, const uint32_t* src, uint32_t Ns, uint32_t layersSrc, uint32_t* dst);
void launchBitsToU8(cudaStream_t s, const uint32_t* words, uint32_t N, uint32_t layers, uint8_t* out);
}  // namespace dxrv
