// Host write bandwidth into a 128 MiB buffer with k threads (development aid: is a compressed read-back + host
// expansion faster than the dense copy over PCIe?).  gcc -O2 -fopenmp tools/host_fill_bw.c -o /tmp/host_fill_bw
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
int main(void)
{
    const size_t n = 128u << 20;
    char* p = aligned_alloc(4096, n);
    memset(p, 1, n);
    printf("cores online: %d\n", omp_get_num_procs());
    for (int k = 1; k <= 64; k *= 2)
    {
        double best = 1e9;
        for (int rep = 0; rep < 5; ++rep)
        {
            const double t0 = now();
#pragma omp parallel num_threads(k)
            {
                const int i = omp_get_thread_num();
                const size_t a = n / k * i, b = (i == k - 1) ? n : n / k * (i + 1);
                memset(p + a, rep & 1, b - a);
            }
            const double t = now() - t0;
            if (t < best) best = t;
        }
        printf("threads %2d: %.3f ms  %.1f GB/s\n", k, best * 1e3, n / best * 1e-9);
    }
    return p[12345] == 7;
}
