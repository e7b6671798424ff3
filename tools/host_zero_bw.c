// Host write bandwidth into a 128 MiB buffer by method and thread count (development aid for the compact read-back
// transport of dxrv_voxelize_to_host: its floor is how fast the host threads can zero the dense grid).
//   gcc -O2 -fopenmp tools/host_zero_bw.c -o /tmp/host_zero_bw && /tmp/host_zero_bw
#define _GNU_SOURCE
#include <immintrin.h>
#include <omp.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static void z_memset(char* p, size_t n) { memset(p, 0, n); }
static void z_sse2(char* p, size_t n)
{
    const __m128i z = _mm_setzero_si128();
    for (size_t i = 0; i + 64 <= n; i += 64)
    {
        _mm_stream_si128((__m128i*)(p + i), z); _mm_stream_si128((__m128i*)(p + i + 16), z);
        _mm_stream_si128((__m128i*)(p + i + 32), z); _mm_stream_si128((__m128i*)(p + i + 48), z);
    }
    _mm_sfence();
}
__attribute__((target("avx2"))) static void z_avx2(char* p, size_t n)
{
    const __m256i z = _mm256_setzero_si256();
    for (size_t i = 0; i + 64 <= n; i += 64) { _mm256_stream_si256((__m256i*)(p + i), z); _mm256_stream_si256((__m256i*)(p + i + 32), z); }
    _mm_sfence();
}
__attribute__((target("avx512f"))) static void z_avx512(char* p, size_t n)
{
    const __m512i z = _mm512_setzero_si512();
    for (size_t i = 0; i + 256 <= n; i += 256)
    {
        _mm512_stream_si512((__m512i*)(p + i), z); _mm512_stream_si512((__m512i*)(p + i + 64), z);
        _mm512_stream_si512((__m512i*)(p + i + 128), z); _mm512_stream_si512((__m512i*)(p + i + 192), z);
    }
    _mm_sfence();
}
static void z_stosb(char* p, size_t n) { __asm__ volatile("rep stosb" : "+D"(p), "+c"(n) : "a"(0) : "memory"); }
static void z_stosq(char* p, size_t n) { size_t q = n / 8; __asm__ volatile("rep stosq" : "+D"(p), "+c"(q) : "a"(0) : "memory"); }

typedef void (*zfn)(char*, size_t);
int main(void)
{
    const size_t n = 128u << 20;
    char* p = aligned_alloc(1 << 21, n);
    memset(p, 1, n);
    printf("cores online: %d  avx2 %d avx512f %d\n", omp_get_num_procs(), __builtin_cpu_supports("avx2"), __builtin_cpu_supports("avx512f"));
    struct { const char* name; zfn f; int ok; } m[] = {
        {"memset", z_memset, 1}, {"sse2 nt", z_sse2, 1}, {"avx2 nt", z_avx2, __builtin_cpu_supports("avx2")},
        {"avx512 nt", z_avx512, __builtin_cpu_supports("avx512f")}, {"rep stosb", z_stosb, 1}, {"rep stosq", z_stosq, 1}};
    const int ks[] = {4, 8, 12, 16, 24, 32};
    for (unsigned mi = 0; mi < sizeof m / sizeof m[0]; ++mi)
    {
        if (!m[mi].ok) continue;
        for (unsigned ki = 0; ki < sizeof ks / sizeof ks[0]; ++ki)
            for (int piece = 0; piece < 2; ++piece)   // 0: one contiguous share per thread, 1: interleaved 1 MiB pieces
            {
                const int k = ks[ki];
                double best = 1e9;
                for (int rep = 0; rep < 7; ++rep)
                {
                    memset(p, 1, 1 << 20);
                    const double t0 = now();
#pragma omp parallel num_threads(k)
                    {
                        const int i = omp_get_thread_num();
                        if (!piece) { const size_t sh = (n / k) & ~(size_t)4095, a = sh * i, b = (i == k - 1) ? n : sh * (i + 1); m[mi].f(p + a, b - a); }
                        else for (size_t a = (size_t)i << 20; a < n; a += (size_t)k << 20) m[mi].f(p + a, 1 << 20);
                    }
                    const double t = now() - t0;
                    if (t < best) best = t;
                }
                printf("%-10s threads %2d %s: %.3f ms  %.1f GB/s\n", m[mi].name, k, piece ? "1 MiB pieces" : "contiguous  ", best * 1e3, n / best * 1e-9);
            }
    }
    return p[12345] == 7;
}
