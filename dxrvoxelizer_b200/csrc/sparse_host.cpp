// sparse_host.cpp -- see sparse_host.h.  The inverse of the encoder in sparse.cu, written for the host's memory
// system: every thread streams through whole z layers (contiguous 4-row runs), empty brick rows are one memset or
// nothing at all, and only the bricks the surface passes through are touched word by word.
#include "sparse_host.h"

#include <cstring>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <vector>

#include "host_pool.h"

namespace dxrv
{
namespace
{
inline uint32_t stateOf(const uint32_t* states, uint32_t b) { return (states[b >> 4] >> (2u * (b & 15u))) & 3u; }

// mixed bricks (state 2) among the bricks [b0, b1)
uint32_t countMixed(const uint32_t* states, uint32_t b0, uint32_t b1)
{
    uint32_t n = 0;
    uint32_t b = b0;
    while (b < b1 && (b & 15u)) { n += stateOf(states, b) == 2u; ++b; }
    for (; b + 16u <= b1; b += 16u)
    {
        const uint32_t w = states[b >> 4];
        n += (uint32_t)__builtin_popcount((w >> 1) & ~w & 0x55555555u);
    }
    for (; b < b1; ++b) n += stateOf(states, b) == 2u;
    return n;
}

bool anyNonEmpty(const uint32_t* states, uint32_t b0, uint32_t b1)
{
    uint32_t b = b0;
    while (b < b1 && (b & 15u)) { if (stateOf(states, b)) return true; ++b; }
    for (; b + 16u <= b1; b += 16u) if (states[b >> 4]) return true;
    for (; b < b1; ++b) if (stateOf(states, b)) return true;
    return false;
}
}  // namespace

bool sparseParse(const void* blob, size_t blobBytes, SparseBlobView& v)
{
    if (!blob || blobBytes < 64) return false;
    const uint32_t* h = static_cast<const uint32_t*>(blob);
    if (h[0] != 0x42525844u || h[1] != 1u || h[12] != 32u || h[13] != 4u || h[14] != 4u) return false;
    v.N = h[2]; v.z0 = h[3]; v.z1 = h[4]; v.P = h[5]; v.BY = h[6]; v.BZ = h[7]; v.numBricks = h[8]; v.numMixed = h[9];
    if (v.z1 <= v.z0 || v.N == 0 || v.P != (v.N + 31) / 32 || v.BY != (v.N + 3) / 4 || v.BZ != (v.z1 - v.z0 + 3) / 4 ||
        (uint64_t)v.numBricks != (uint64_t)v.P * v.BY * v.BZ)
        return false;
    const size_t offStates = h[10], offPayload = h[11];
    if (offStates < 64 || (offStates & 3u) || (offPayload & 3u) || offPayload < offStates + (size_t)((v.numBricks + 15) / 16) * 4 ||
        blobBytes < offPayload + (size_t)v.numMixed * 64)
        return false;
    v.states = reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(blob) + offStates);
    v.payload = reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(blob) + offPayload);
    return true;
}

bool sparseExpand(const SparseBlobView& v, uint32_t* dst, bool dstIsZero)
{
    const uint32_t N = v.N, P = v.P, BY = v.BY, BZ = v.BZ, layers = v.z1 - v.z0;
    const uint32_t perLayer = BY * P;   // bricks per brick layer
    // rank of the first mixed brick of every brick layer
    std::vector<uint32_t> base(BZ + 1u, 0u);
    hostParallelFor(BZ, [&](unsigned bz) { base[bz + 1u] = countMixed(v.states, bz * perLayer, (bz + 1u) * perLayer); });
    for (uint32_t bz = 0; bz < BZ; ++bz) base[bz + 1u] += base[bz];
    if (base[BZ] != v.numMixed) return false;
    const uint32_t tailMask = (N & 31u) ? ((1u << (N & 31u)) - 1u) : 0xffffffffu;
    hostParallelFor(BZ, [&](unsigned bz) {
        for (uint32_t k = 0; k < 4u; ++k)
        {
            const uint32_t z = 4u * bz + k;
            if (z >= layers) break;
            uint32_t rank = base[bz];
            for (uint32_t by = 0; by < BY; ++by)
            {
                const uint32_t b0 = (bz * BY + by) * P;
                const uint32_t rows = (4u * by + 4u <= N) ? 4u : N - 4u * by;
                uint32_t* out = dst + ((size_t)z * N + 4u * by) * P;   // `rows` consecutive rows of P words
                if (!anyNonEmpty(v.states, b0, b0 + P))
                {
                    if (!dstIsZero) std::memset(out, 0, sizeof(uint32_t) * rows * P);
                    continue;
                }
                for (uint32_t bx = 0; bx < P; ++bx)
                {
                    const uint32_t st = stateOf(v.states, b0 + bx);
                    if (st == 0u)
                    {
                        if (!dstIsZero) for (uint32_t j = 0; j < rows; ++j) out[j * P + bx] = 0u;
                    }
                    else if (st == 1u)
                    {
                        const uint32_t full = (bx == P - 1u) ? tailMask : 0xffffffffu;
                        for (uint32_t j = 0; j < rows; ++j) out[j * P + bx] = full;
                    }
                    else
                    {
                        const uint32_t* w = v.payload + (size_t)rank * 16u + 4u * k;
                        for (uint32_t j = 0; j < rows; ++j) out[j * P + bx] = w[j];
                        ++rank;
                    }
                }
            }
        }
    });
    return true;
}

// Zeroing with NON-TEMPORAL stores: an ordinary memset of a 1 MiB piece reads every line before it overwrites it
// (read-for-ownership), which halves the write bandwidth -- measured on the B200 box's 16 cores: 67 GB/s against
// 125 GB/s (glibc only switches to streaming stores for much larger blocks).
static void zeroStreaming(uint8_t* p, size_t n)
{
#if defined(__SSE2__)
    while (n && (reinterpret_cast<uintptr_t>(p) & 15u)) { *p++ = 0; --n; }
    const __m128i z = _mm_setzero_si128();
    size_t i = 0;
    for (; i + 64 <= n; i += 64)
    {
        _mm_stream_si128(reinterpret_cast<__m128i*>(p + i), z);
        _mm_stream_si128(reinterpret_cast<__m128i*>(p + i + 16), z);
        _mm_stream_si128(reinterpret_cast<__m128i*>(p + i + 32), z);
        _mm_stream_si128(reinterpret_cast<__m128i*>(p + i + 48), z);
    }
    _mm_sfence();
    if (i < n) std::memset(p + i, 0, n - i);
#else
    std::memset(p, 0, n);
#endif
}

void hostZeroBegin(void* dst, size_t bytes)
{
    const size_t chunk = 1u << 20;
    const unsigned tasks = (unsigned)((bytes + chunk - 1) / chunk);
    uint8_t* p = static_cast<uint8_t*>(dst);
    hostParallelBegin(tasks, [p, bytes, chunk](unsigned t) {
        const size_t a = (size_t)t * chunk, b = a + chunk < bytes ? a + chunk : bytes;
        zeroStreaming(p + a, b - a);
    });
}

void hostZeroWait() { hostParallelWait(); }
}  // namespace dxrv
