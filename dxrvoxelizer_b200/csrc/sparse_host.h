// sparse_host.h -- host side of DXRV_FORMAT_SPARSE_BRICKS (include/dxrv.h): parse a blob, expand it into the dense
// DXRV_FORMAT_BITS layout with the host pool's threads.  Internal to libdxrv.so.
#pragma once
#include <cstddef>
#include <cstdint>

namespace dxrv
{
struct SparseBlobView
{
    uint32_t N, z0, z1, P, BY, BZ, numBricks, numMixed;
    const uint32_t* states;
    const uint32_t* payload;
};
// false: not a blob of this format / truncated / inconsistent sizes
bool sparseParse(const void* blob, size_t blobBytes, SparseBlobView& v);
// Expand into dst (layers * N * P words).  dstIsZero: dst holds zeros already (hostZero below, typically issued while
// the GPU was still computing) and only the non-empty bricks are written; otherwise every word is written.
// false: the states disagree with the header's count of mixed bricks (nothing reliable was written).
bool sparseExpand(const SparseBlobView& v, uint32_t* dst, bool dstIsZero);
// The inverse on the host: dense BITS slab (layers * N * P words) -> blob, byte for byte what the device encoder
// (sparse.cu) writes.  bytes receives the blob's size; false when `capacity` is too small (nothing written but bytes).
bool sparseEncode(const uint32_t* dense, uint32_t N, uint32_t z0, uint32_t z1, void* blob, size_t capacity, size_t& bytes);
// Zero `bytes` at dst with the pool's threads: begin returns at once, wait blocks (host_pool.h: one batch at a time).
void hostZeroBegin(void* dst, size_t bytes);
void hostZeroWait();
// The dense grid of a slab (layers * N * P words at dst) in ONE pass of the pool: begin returns at once and the threads
// start zeroing from the outside of the grid inwards (the layers of the slab farthest from the grid's centre first); publish hands them the blob (as soon as it is on the host: the
// brick layers still to do are then written with their final contents); wait blocks until the grid is complete (brick
// layers zeroed before the publish are expanded last).  publish: false = the blob does not fit the grid / is
// inconsistent (nothing changes); wait: false = nothing was published (the grid holds zeros).
void hostFillBegin(void* dst, uint32_t N, uint32_t z0, uint32_t layers);   // the slab [z0, z0 + layers) of an N^3 grid
// blockRanks (optional): rank of the first mixed brick of every block of bricksPerBlock bricks, as the device encoder
// computed them -- used instead of counting the states when brick layers are whole blocks.
bool hostFillPublish(const SparseBlobView& v, const uint32_t* blockRanks = nullptr, uint32_t bricksPerBlock = 0);
bool hostFillWait();
}  // namespace dxrv
