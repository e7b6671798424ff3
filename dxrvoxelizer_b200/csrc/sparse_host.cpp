// sparse_host.cpp -- see sparse_host.h.  The inverse of the encoder in sparse.cu, written for the host's memory
// system: every thread streams through whole z layers (contiguous 4-row runs), empty brick rows are one memset or
// nothing at all, and only the bricks the surface passes through are touched word by word.
#include "sparse_host.h"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include <vector>

#include "host_pool.h"

namespace dxrv
{
namespace
{
inline uint32_t stateOf(const uint32_t* states, uint32_t b) { return (states[b >> 4] >> (2u * (b & 15u))) & 3u; }

// mixed bricks among the 16 * words bricks of whole state words (the host's popcount instruction when it has one:
// without it the 0.5 MB of states of a 1024^3 grid take 0.3 ms on the thread that publishes the blob)
#if defined(__x86_64__)
__attribute__((target("popcnt"))) uint32_t countMixedWordsPopcnt(const uint32_t* w, uint32_t words, uint32_t* undefined)
{
    uint64_t n = 0, u = 0;
    uint32_t i = 0;
    for (; i + 2u <= words; i += 2u)
    {
        uint64_t v;
        std::memcpy(&v, w + i, 8);
        n += (uint64_t)__builtin_popcountll((v >> 1) & ~v & 0x5555555555555555ull);
        u |= (v >> 1) & v & 0x5555555555555555ull;
    }
    for (; i < words; ++i) { n += (uint64_t)__builtin_popcount((w[i] >> 1) & ~w[i] & 0x55555555u); u |= (w[i] >> 1) & w[i] & 0x55555555u; }
    if (u) *undefined = 1u;
    return (uint32_t)n;
}
#endif
uint32_t countMixedWords(const uint32_t* w, uint32_t words, uint32_t* undefined)   // *undefined = 1: a state 3 was seen
{
#if defined(__x86_64__)
    static const bool hw = __builtin_cpu_supports("popcnt");
    if (hw) return countMixedWordsPopcnt(w, words, undefined);
#endif
    uint32_t n = 0, u = 0;
    for (uint32_t i = 0; i < words; ++i) { n += (uint32_t)__builtin_popcount((w[i] >> 1) & ~w[i] & 0x55555555u); u |= (w[i] >> 1) & w[i] & 0x55555555u; }
    if (u) *undefined = 1u;
    return n;
}

// mixed bricks (state 2) among the bricks [b0, b1)
uint32_t countMixed(const uint32_t* states, uint32_t b0, uint32_t b1, uint32_t* undefined)
{
    uint32_t n = 0;
    uint32_t b = b0;
    while (b < b1 && (b & 15u)) { const uint32_t st = stateOf(states, b); n += st == 2u; if (st == 3u) *undefined = 1u; ++b; }
    if (b + 16u <= b1)
    {
        const uint32_t words = (b1 - b) >> 4;
        n += countMixedWords(states + (b >> 4), words, undefined);
        b += words << 4;
    }
    for (; b < b1; ++b) { const uint32_t st = stateOf(states, b); n += st == 2u; if (st == 3u) *undefined = 1u; }
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

// Host encoder: classify the bricks of every brick layer on the pool's threads (0 empty, 1 full, 2 mixed; rows and
// layers beyond the slab do not count against "full" and are stored as zeros), rank the mixed bricks, write header,
// packed states and payload.  Same bytes as the device encoder (tests/test_sparse.py: both against a numpy encoder).
bool sparseEncode(const uint32_t* dense, uint32_t N, uint32_t z0, uint32_t z1, void* blob, size_t capacity, size_t& bytes)
{
    const uint32_t P = (N + 31u) / 32u, layers = z1 - z0, BY = (N + 3u) / 4u, BZ = (layers + 3u) / 4u;
    const size_t numBricks = (size_t)P * BY * BZ, perLayer = (size_t)P * BY;
    const size_t stateWords = (numBricks + 15) / 16, offStates = 64, offPayload = (offStates + stateWords * 4 + 63) & ~(size_t)63;
    const uint32_t tailMask = (N & 31u) ? ((1u << (N & 31u)) - 1u) : 0xffffffffu;
    std::vector<uint8_t> state(numBricks);
    std::vector<size_t> mixedInLayer(BZ, 0);
    auto brickWords = [&](uint32_t bz, uint32_t by, uint32_t bx, uint32_t out[16], uint32_t& exists) {
        exists = 0;
        for (uint32_t k = 0; k < 4u; ++k)
            for (uint32_t j = 0; j < 4u; ++j)
            {
                const uint32_t z = 4u * bz + k, y = 4u * by + j;
                const bool ex = z < layers && y < N;
                out[4u * k + j] = ex ? dense[((size_t)z * N + y) * P + bx] : 0u;
                exists |= (ex ? 1u : 0u) << (4u * k + j);
            }
    };
    hostParallelFor(BZ, [&](unsigned bz) {
        size_t mixed = 0;
        for (uint32_t by = 0; by < BY; ++by)
            for (uint32_t bx = 0; bx < P; ++bx)
            {
                uint32_t w[16], exists;
                brickWords(bz, by, bx, w, exists);
                const uint32_t fullWord = bx == P - 1u ? tailMask : 0xffffffffu;
                bool any = false, full = true;
                for (uint32_t i = 0; i < 16u; ++i)
                {
                    any = any || w[i] != 0u;
                    if ((exists >> i) & 1u) full = full && w[i] == fullWord;
                }
                const uint8_t st = !any ? 0 : (full ? 1 : 2);
                state[(size_t)bz * perLayer + (size_t)by * P + bx] = st;
                mixed += st == 2;
            }
        mixedInLayer[bz] = mixed;
    });
    std::vector<size_t> firstRank(BZ + 1, 0);
    for (uint32_t bz = 0; bz < BZ; ++bz) firstRank[bz + 1] = firstRank[bz] + mixedInLayer[bz];
    const size_t numMixed = firstRank[BZ];
    bytes = offPayload + numMixed * 64;
    if (!blob || capacity < bytes) return false;
    uint8_t* out = static_cast<uint8_t*>(blob);
    std::memset(out, 0, offPayload);
    const uint32_t header[16] = {0x42525844u, 1u, N, z0, z1, P, BY, BZ, (uint32_t)numBricks, (uint32_t)numMixed, (uint32_t)offStates, (uint32_t)offPayload, 32u, 4u, 4u, 0u};
    std::memcpy(out, header, sizeof header);
    uint32_t* sw = reinterpret_cast<uint32_t*>(out + offStates);
    for (size_t b = 0; b < numBricks; ++b) sw[b >> 4] |= (uint32_t)state[b] << (2u * (b & 15u));
    uint32_t* payload = reinterpret_cast<uint32_t*>(out + offPayload);
    hostParallelFor(BZ, [&](unsigned bz) {
        size_t rank = firstRank[bz];
        for (uint32_t by = 0; by < BY; ++by)
            for (uint32_t bx = 0; bx < P; ++bx)
                if (state[(size_t)bz * perLayer + (size_t)by * P + bx] == 2)
                {
                    uint32_t exists;
                    brickWords(bz, by, bx, payload + 16 * rank, exists);
                    ++rank;
                }
    });
    return true;
}

// brick layer bz of the blob into dst, word by word (dstIsZero: only the non-empty bricks are written)
static void expandLayer(const SparseBlobView& v, uint32_t* dst, uint32_t bz, uint32_t firstRank, bool dstIsZero)
{
    const uint32_t N = v.N, P = v.P, BY = v.BY, layers = v.z1 - v.z0;
    const uint32_t tailMask = (N & 31u) ? ((1u << (N & 31u)) - 1u) : 0xffffffffu;
    for (uint32_t k = 0; k < 4u; ++k)
    {
        const uint32_t z = 4u * bz + k;
        if (z >= layers) break;
        uint32_t rank = firstRank;
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
}

bool sparseExpand(const SparseBlobView& v, uint32_t* dst, bool dstIsZero)
{
    const uint32_t BZ = v.BZ;
    const uint32_t perLayer = v.BY * v.P;   // bricks per brick layer
    // rank of the first mixed brick of every brick layer
    std::vector<uint32_t> base(BZ + 1u, 0u);
    std::vector<uint32_t> undefinedState(BZ, 0u);
    hostParallelFor(BZ, [&](unsigned bz) { base[bz + 1u] = countMixed(v.states, bz * perLayer, (bz + 1u) * perLayer, &undefinedState[bz]); });
    for (uint32_t bz = 0; bz < BZ; ++bz) { base[bz + 1u] += base[bz]; if (undefinedState[bz]) return false; }
    if (base[BZ] != v.numMixed) return false;
    hostParallelFor(BZ, [&](unsigned bz) { expandLayer(v, dst, bz, base[bz], dstIsZero); });
    return true;
}

// Zeroing without read-for-ownership.  Measured on a B200 box's 16 cores, bare loops over 128 MiB (tools/host_zero_bw.c):
// streaming stores of 16 / 32 / 64 bytes 184 / 188 / 187 GB/s, `rep stosb` 172 GB/s, glibc memset 161 GB/s -- but
// INSIDE the end-to-end call, methods alternated within one process (tools/e2e_ab.py; separate processes differ by more
// than the methods do): `rep stosb` 0.708 ms and memset 0.699 ms (glibc takes `rep stosb` at these sizes) against
// 0.784 / 0.779 / 0.817 ms for the 16 / 32 / 64-byte streaming stores; the same order for the bunny.  Fast strings
// write whole lines without ownership reads and leave the memory system more room for the other threads than a
// stream of non-temporal stores does.  `rep stosb` is the default; DXRV_HOST_ZERO = stosb | sse2 | avx2 | avx512 |
// memset overrides it (looked at again at the start of every pass).
enum class ZeroMethod { Memset, Sse2, Avx2, Avx512, Stosb };

#if defined(__x86_64__)
__attribute__((target("avx2"))) static void zeroAvx2(uint8_t* p, size_t n)   // p 32-byte aligned, n a multiple of 64
{
    const __m256i z = _mm256_setzero_si256();
    for (size_t i = 0; i < n; i += 64)
    {
        _mm256_stream_si256(reinterpret_cast<__m256i*>(p + i), z);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(p + i + 32), z);
    }
    _mm_sfence();
}
__attribute__((target("avx512f"))) static void zeroAvx512(uint8_t* p, size_t n)   // p 64-byte aligned, n a multiple of 64
{
    const __m512i z = _mm512_setzero_si512();
    for (size_t i = 0; i < n; i += 64) _mm512_stream_si512(reinterpret_cast<__m512i*>(p + i), z);
    _mm_sfence();
}
static void zeroSse2(uint8_t* p, size_t n)   // p 16-byte aligned, n a multiple of 64
{
    const __m128i z = _mm_setzero_si128();
    for (size_t i = 0; i < n; i += 64)
    {
        _mm_stream_si128(reinterpret_cast<__m128i*>(p + i), z);
        _mm_stream_si128(reinterpret_cast<__m128i*>(p + i + 16), z);
        _mm_stream_si128(reinterpret_cast<__m128i*>(p + i + 32), z);
        _mm_stream_si128(reinterpret_cast<__m128i*>(p + i + 48), z);
    }
    _mm_sfence();
}
static void zeroStosb(uint8_t* p, size_t n) { __asm__ volatile("rep stosb" : "+D"(p), "+c"(n) : "a"(0) : "memory"); }

static ZeroMethod pickZeroMethod()
{
    __builtin_cpu_init();
    ZeroMethod m = ZeroMethod::Stosb;
    if (const char* e = std::getenv("DXRV_HOST_ZERO"))
    {
        if (!std::strcmp(e, "stosb")) m = ZeroMethod::Stosb;
        else if (!std::strcmp(e, "avx512") && __builtin_cpu_supports("avx512f")) m = ZeroMethod::Avx512;
        else if (!std::strcmp(e, "avx2") && __builtin_cpu_supports("avx2")) m = ZeroMethod::Avx2;
        else if (!std::strcmp(e, "sse2")) m = ZeroMethod::Sse2;
        else if (!std::strcmp(e, "memset")) m = ZeroMethod::Memset;
    }
    return m;
}
#endif

static std::atomic<int> gZeroMethod{-1};
// (the environment is looked at again at the start of every pass: tools A/B the methods inside one process, on one box)
static void refreshZeroMethod()
{
#if defined(__x86_64__)
    gZeroMethod.store((int)pickZeroMethod(), std::memory_order_relaxed);
#endif
}

static void zeroStreaming(uint8_t* p, size_t n)
{
#if defined(__x86_64__)
    int mi = gZeroMethod.load(std::memory_order_relaxed);
    if (mi < 0) { refreshZeroMethod(); mi = gZeroMethod.load(std::memory_order_relaxed); }
    const ZeroMethod method = (ZeroMethod)mi;
    if (method == ZeroMethod::Memset) { std::memset(p, 0, n); return; }
    if (method == ZeroMethod::Stosb) { zeroStosb(p, n); return; }
    while (n && (reinterpret_cast<uintptr_t>(p) & 63u)) { *p++ = 0; --n; }
    const size_t body = n & ~(size_t)63;
    if (method == ZeroMethod::Avx512) zeroAvx512(p, body);
    else if (method == ZeroMethod::Avx2) zeroAvx2(p, body);
    else zeroSse2(p, body);
    if (body < n) std::memset(p + body, 0, n - body);
#else
    std::memset(p, 0, n);
#endif
}

void hostZeroBegin(void* dst, size_t bytes)
{
    refreshZeroMethod();
    const size_t chunk = 1u << 20;
    const unsigned tasks = (unsigned)((bytes + chunk - 1) / chunk);
    uint8_t* p = static_cast<uint8_t*>(dst);
    hostParallelBegin(tasks, [p, bytes, chunk](unsigned t) {
        const size_t a = (size_t)t * chunk, b = a + chunk < bytes ? a + chunk : bytes;
        zeroStreaming(p + a, b - a);
    });
}

void hostZeroWait() { hostParallelWait(); }

// ---- the caller's dense grid in ONE pass ---------------------------------------------------------------------------
// Zeroing the whole grid and then expanding the blob into it touches the lines of every non-empty brick twice, the
// second time with an ownership read (they were written around the caches), and the expansion runs after the zeroing:
// measured on a B200 box (dragon 1024^3, 16 threads): GPU part 0.25 ms, zeroing done at 0.76 ms, expansion 0.33 ms more.
// Here the pass starts as plain zeroing -- nothing else can be done before the GPU has produced the blob -- from the
// OUTSIDE of the slab inwards (a mesh is normalised to the grid's centre: its outermost layers are the likeliest to be
// empty), and as soon as the blob is published the brick layers still to do are written with their final contents:
// empty row runs as streaming zeros, the others composed in a small buffer (zeros, full words, payload rows) and
// streamed out as whole lines.  Brick layers that were zeroed before the blob arrived are expanded afterwards the old
// way; for the dragon none of them holds a brick.
namespace
{
void streamOut(uint32_t* dst, const uint32_t* src, size_t words)
{
#if defined(__x86_64__)
    if (((reinterpret_cast<uintptr_t>(dst) | (words * 4u)) & 63u) == 0u)
    {
        for (size_t i = 0; i < words; i += 16)
        {
            const __m128i* q = reinterpret_cast<const __m128i*>(src + i);
            __m128i* d = reinterpret_cast<__m128i*>(dst + i);
            _mm_stream_si128(d, _mm_loadu_si128(q)); _mm_stream_si128(d + 1, _mm_loadu_si128(q + 1));
            _mm_stream_si128(d + 2, _mm_loadu_si128(q + 2)); _mm_stream_si128(d + 3, _mm_loadu_si128(q + 3));
        }
        return;
    }
#endif
    std::memcpy(dst, src, words * 4u);
}

struct FillState
{
    uint32_t* dst = nullptr;
    uint32_t N = 0, P = 0, BY = 0, BZ = 0, layers = 0, group = 1, numTasks = 0;
    std::atomic<const SparseBlobView*> blob{nullptr};
    SparseBlobView view{};
    std::vector<uint32_t> base;          // rank of the first mixed brick of every brick layer
    std::vector<uint8_t> zeroedOnly;     // per group: its brick layers were zeroed before the blob arrived
    std::vector<uint32_t> order;         // task -> group of brick layers, farthest from the grid's centre first
    bool active = false;
    std::mutex owner;                    // held from hostFillBegin to hostFillWait: one pass at a time per process
} gFill;

#if defined(__x86_64__)
// The same with AVX-512, for grids whose rows are whole groups of 16 words (N a multiple of 512): one state word describes
// the 16 bricks under one 64-byte line of each of the run's rows, so a line is built in a register -- all-ones lanes for
// the full bricks, a masked gather of one payload word per mixed brick (lane x's brick has rank `first rank of the word +
// mixed bricks before x`: an expand of consecutive ranks into the mixed lanes) -- and streamed out whole.  No staging
// buffer, no per-brick loop: 30 ns per non-empty run instead of 74 (dragon 1024^3, one thread).
__attribute__((target("avx512f,bmi2,popcnt")))
void fillLayerAvx512(const SparseBlobView& v, uint32_t* dst, uint32_t bz, uint32_t firstRank)
{
    const uint32_t N = v.N, P = v.P, BY = v.BY, layers = v.z1 - v.z0, groups = P >> 4;
    const uint32_t* states = v.states + (((size_t)bz * BY * P) >> 4);       // P % 16 == 0: every run starts a state word
    const __m512i iota = _mm512_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    const __m512i zero = _mm512_setzero_si512();
    const int* payload = reinterpret_cast<const int*>(v.payload);
    // (streaming stores: ordinary 64-byte stores measured 0.20 ms slower per 1024^3 grid -- they pull the lines through the caches)
    auto put = [](uint32_t* p, __m512i v) __attribute__((target("avx512f"))) { _mm512_stream_si512(reinterpret_cast<__m512i*>(p), v); };
    for (uint32_t k = 0; k < 4u; ++k)
    {
        const uint32_t z = 4u * bz + k;
        if (z >= layers) break;
        uint32_t rank = firstRank;
        uint32_t* out = dst + (size_t)z * N * P;                             // row y of this layer: out + y * P
        const uint32_t* sw = states;
        uint8_t* zeroFrom = nullptr;                                         // pending run of empty row runs (contiguous in the grid)
        for (uint32_t by = 0; by < BY; ++by, out += 4u * P, sw += groups)
        {
            uint32_t any = 0;
            for (uint32_t g = 0; g < groups; ++g) any |= sw[g];
            if (any == 0u)
            {
                if (!zeroFrom) zeroFrom = reinterpret_cast<uint8_t*>(out);
                continue;
            }
            if (zeroFrom) { zeroStreaming(zeroFrom, (size_t)(reinterpret_cast<uint8_t*>(out) - zeroFrom)); zeroFrom = nullptr; }
            for (uint32_t g = 0; g < groups; ++g)
            {
                uint32_t* line = out + 16u * g;
                const uint32_t w = sw[g];
                if (w == 0u)
                {
                    put(line, zero); put(line + P, zero); put(line + 2u * P, zero); put(line + 3u * P, zero);
                    continue;
                }
                const uint32_t lo = _pext_u32(w, 0x55555555u), hi = _pext_u32(w, 0xaaaaaaaau);
                const __mmask16 full = (__mmask16)(lo & ~hi), mixed = (__mmask16)hi;      // (state 3 was rejected at publish)
                const __m512i fullv = _mm512_maskz_set1_epi32(full, -1);
                // word index of lane x's payload row 4k + 0: (rank of its brick) * 16 + 4k
                const __m512i ranks = _mm512_maskz_expand_epi32(mixed, _mm512_add_epi32(iota, _mm512_set1_epi32((int)rank)));
                const __m512i idx = _mm512_add_epi32(_mm512_slli_epi32(ranks, 4), _mm512_set1_epi32((int)(4u * k)));
                for (uint32_t j = 0; j < 4u; ++j)
                {
                    const __m512i val = mixed ? _mm512_mask_i32gather_epi32(fullv, mixed, _mm512_add_epi32(idx, _mm512_set1_epi32((int)j)), payload, 4) : fullv;
                    put(line + j * P, val);
                }
                rank += (uint32_t)__builtin_popcount(hi);
            }
        }
        if (zeroFrom) zeroStreaming(zeroFrom, (size_t)(reinterpret_cast<uint8_t*>(out) - zeroFrom));
    }
    _mm_sfence();
}
#endif

// brick layer bz with its final contents, every word written exactly once
void fillLayer(const SparseBlobView& v, uint32_t* dst, uint32_t bz, uint32_t firstRank)
{
    const uint32_t N = v.N, P = v.P, BY = v.BY, layers = v.z1 - v.z0;
    const uint32_t tailMask = (N & 31u) ? ((1u << (N & 31u)) - 1u) : 0xffffffffu;
    const size_t wordsPerLayer = (size_t)N * P;
    if (!anyNonEmpty(v.states, bz * BY * P, (bz + 1u) * BY * P))     // an empty brick layer: one run of zeros
    {
        const size_t z0 = (size_t)4u * bz, z1 = std::min<size_t>(z0 + 4u, layers);
        zeroStreaming(reinterpret_cast<uint8_t*>(dst + z0 * wordsPerLayer), (z1 - z0) * wordsPerLayer * 4u);
        return;
    }
#if defined(__x86_64__)
    static const bool wide = [] {
        const char* e = std::getenv("DXRV_HOST_FILL");
        return !(e && !std::strcmp(e, "scalar")) && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("bmi2") && __builtin_cpu_supports("popcnt");
    }();
    // (N % 512 == 0: whole state words per run, four rows per brick, no partial last word; the payload may hold 2^28 words)
    if (wide && (N & 511u) == 0u && (reinterpret_cast<uintptr_t>(dst) & 63u) == 0u && v.numMixed < (1u << 27))
    {
        fillLayerAvx512(v, dst, bz, firstRank);
        return;
    }
#endif
    std::vector<uint32_t> tmpStore;
    uint32_t tmpStack[4 * 64 + 16];                                  // rows of up to 64 words (N <= 2048) on the stack
    uint32_t* tmp = tmpStack;
    if (4u * P > 4u * 64u) { tmpStore.resize(4u * (size_t)P + 16u); tmp = tmpStore.data(); }
    tmp = reinterpret_cast<uint32_t*>((reinterpret_cast<uintptr_t>(tmp) + 63u) & ~(uintptr_t)63u);
    for (uint32_t k = 0; k < 4u; ++k)
    {
        const uint32_t z = 4u * bz + k;
        if (z >= layers) break;
        uint32_t rank = firstRank;
        uint8_t* zeroFrom = nullptr;                                 // pending run of empty rows (contiguous in the grid)
        uint8_t* zeroTo = nullptr;
        for (uint32_t by = 0; by < BY; ++by)
        {
            const uint32_t b0 = (bz * BY + by) * P;
            const uint32_t rows = (4u * by + 4u <= N) ? 4u : N - 4u * by;
            uint32_t* out = dst + ((size_t)z * N + 4u * by) * P;     // `rows` consecutive rows of P words
            const size_t runWords = (size_t)rows * P;
            if (!anyNonEmpty(v.states, b0, b0 + P))
            {
                if (!zeroFrom) zeroFrom = reinterpret_cast<uint8_t*>(out);
                zeroTo = reinterpret_cast<uint8_t*>(out + runWords);
                continue;
            }
            if (zeroFrom) { zeroStreaming(zeroFrom, (size_t)(zeroTo - zeroFrom)); zeroFrom = nullptr; }
            std::memset(tmp, 0, runWords * 4u);
            for (uint32_t bx = 0; bx < P;)
            {
                const uint32_t b = b0 + bx;
                if ((b & 15u) == 0u && bx + 16u <= P)
                {
                    // a whole state word: only its non-empty bricks are visited (lowest first: the payload is in brick order)
                    const uint32_t w = v.states[b >> 4];
                    for (uint32_t m = (w | (w >> 1)) & 0x55555555u; m; m &= m - 1u)
                    {
                        const uint32_t sh = (uint32_t)__builtin_ctz(m), x = bx + (sh >> 1);
                        if (((w >> sh) & 3u) == 1u)
                        {
                            const uint32_t full = (x == P - 1u) ? tailMask : 0xffffffffu;
                            for (uint32_t j = 0; j < rows; ++j) tmp[j * P + x] = full;
                        }
                        else
                        {
                            const uint32_t* pw = v.payload + (size_t)rank * 16u + 4u * k;
                            for (uint32_t j = 0; j < rows; ++j) tmp[j * P + x] = pw[j];
                            ++rank;
                        }
                    }
                    bx += 16u;
                    continue;
                }
                const uint32_t st = stateOf(v.states, b);
                if (st == 1u)
                {
                    const uint32_t full = (bx == P - 1u) ? tailMask : 0xffffffffu;
                    for (uint32_t j = 0; j < rows; ++j) tmp[j * P + bx] = full;
                }
                else if (st != 0u)
                {
                    const uint32_t* w = v.payload + (size_t)rank * 16u + 4u * k;
                    for (uint32_t j = 0; j < rows; ++j) tmp[j * P + bx] = w[j];
                    ++rank;
                }
                ++bx;
            }
            streamOut(out, tmp, runWords);
        }
        if (zeroFrom) zeroStreaming(zeroFrom, (size_t)(zeroTo - zeroFrom));
    }
#if defined(__x86_64__)
    _mm_sfence();
#endif
}
}  // namespace

void hostFillBegin(void* dst, uint32_t N, uint32_t z0, uint32_t layers)
{
    FillState& f = gFill;
    f.owner.lock();                      // (a second context's call waits here until the first one's pass is complete)
    refreshZeroMethod();
    f.dst = static_cast<uint32_t*>(dst);
    f.N = N; f.P = (N + 31u) / 32u; f.BY = (N + 3u) / 4u; f.BZ = (layers + 3u) / 4u; f.layers = layers;
    // a task = a group of brick layers of about 512 KiB
    const size_t layerBytes = (size_t)4u * N * f.P * 4u;
    f.group = (uint32_t)std::max<size_t>(1u, (512u << 10) / std::max<size_t>(layerBytes, 1u));
    f.numTasks = (f.BZ + f.group - 1u) / f.group;
    f.blob.store(nullptr, std::memory_order_relaxed);
    f.zeroedOnly.assign(f.numTasks, 0);
    // From the outside of the GRID inwards (a z-slab of a multi-GPU run has the mesh at one of its ends, not in its
    // middle): groups ordered by the distance of their middle layer from the grid's centre plane, farthest first.
    f.order.resize(f.numTasks);
    for (uint32_t g = 0; g < f.numTasks; ++g) f.order[g] = g;
    {
        const uint32_t group = f.group, BZ = f.BZ;
        const double centre = 0.5 * (double)N;
        auto dist = [=](uint32_t g) {
            const double mid = (double)z0 + 2.0 * ((double)g * group + (double)std::min((g + 1u) * group, BZ));   // 4 * (bz0 + bz1) / 2
            return mid > centre ? mid - centre : centre - mid;
        };
        std::stable_sort(f.order.begin(), f.order.end(), [&](uint32_t a, uint32_t b) { return dist(a) > dist(b); });
    }
    f.active = true;
    hostParallelBegin(f.numTasks, [](unsigned t) {
        FillState& s = gFill;
        const uint32_t g = s.order[t];
        const uint32_t bz0 = g * s.group, bz1 = std::min(bz0 + s.group, s.BZ);
        const SparseBlobView* v = s.blob.load(std::memory_order_acquire);
        if (!v)
        {
            const size_t wordsPerLayer = (size_t)s.N * s.P;
            const size_t z0 = (size_t)4u * bz0, z1 = std::min<size_t>((size_t)4u * bz1, s.layers);
            zeroStreaming(reinterpret_cast<uint8_t*>(s.dst + z0 * wordsPerLayer), (z1 - z0) * wordsPerLayer * 4u);
            s.zeroedOnly[g] = 1;
            return;
        }
        for (uint32_t bz = bz0; bz < bz1; ++bz) fillLayer(*v, s.dst, bz, s.base[bz]);
    });
}

bool hostFillPublish(const SparseBlobView& v, const uint32_t* blockRanks, uint32_t bricksPerBlock)
{
    FillState& f = gFill;
    if (!f.active || v.N != f.N || v.z1 - v.z0 != f.layers || v.BZ != f.BZ || v.BY != f.BY || v.P != f.P) return false;
    const uint32_t perLayer = v.BY * v.P;
    f.base.assign(f.BZ + 1u, 0u);
    if (blockRanks && bricksPerBlock && perLayer % bricksPerBlock == 0u)
    {
        // the encoder's own exclusive ranks per block of bricks (sparse.cu: k_brick_scan), brick layers being whole blocks:
        // nothing to count here (reading the 0.5 MB of states of a 1024^3 grid on this thread took 0.06 - 0.4 ms while the
        // pool's streaming stores saturate the memory system)
        const uint32_t blocksPerLayer = perLayer / bricksPerBlock;
        for (uint32_t bz = 0; bz < f.BZ; ++bz) f.base[bz] = blockRanks[(size_t)bz * blocksPerLayer];
        f.base[f.BZ] = v.numMixed;
        for (uint32_t bz = 0; bz < f.BZ; ++bz) if (f.base[bz] > f.base[bz + 1u]) return false;
    }
    else
    {
        uint32_t undefinedState = 0;   // (state 3 does not exist: such a blob would read payload it does not have)
        for (uint32_t bz = 0; bz < f.BZ; ++bz) f.base[bz + 1u] = f.base[bz] + countMixed(v.states, bz * perLayer, (bz + 1u) * perLayer, &undefinedState);
        if (f.base[f.BZ] != v.numMixed || undefinedState) return false;
    }
    f.view = v;
    f.blob.store(&f.view, std::memory_order_release);
    return true;
}

bool hostFillWait()
{
    FillState& f = gFill;
    hostParallelWait();
    f.active = false;
    const SparseBlobView* v = f.blob.load(std::memory_order_acquire);
    if (!v) { f.owner.unlock(); return false; }   // nothing was published: the grid holds zeros only
    // the brick layers that were zeroed before the blob arrived: expand the ones that hold something
    std::vector<uint32_t> todo;
    const uint32_t perLayer = v->BY * v->P;
    for (uint32_t g = 0; g < f.numTasks; ++g)
        if (f.zeroedOnly[g])
            for (uint32_t bz = g * f.group; bz < std::min((g + 1u) * f.group, f.BZ); ++bz)
                if (anyNonEmpty(v->states, bz * perLayer, (bz + 1u) * perLayer)) todo.push_back(bz);
    if (!todo.empty())
        hostParallelFor((unsigned)todo.size(), [&](unsigned i) { expandLayer(*v, f.dst, todo[i], f.base[todo[i]], true); });
    f.blob.store(nullptr, std::memory_order_relaxed);
    f.owner.unlock();
    return true;
}
}  // namespace dxrv
