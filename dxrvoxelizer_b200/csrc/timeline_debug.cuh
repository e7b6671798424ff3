// timeline_debug.cuh -- development aid, compiled only with -DDXRV_TIMELINE (never in the shipped
// library): every CTA of an instrumented kernel records {start, end, role, SM} from %globaltimer, and
// the launcher prints per-role and per-SM summaries when DXRV_DBG_TIMELINE is set.  This is how the
// work ordering / split thresholds of k_trace_fill_columns were tuned (tools/timeline.sh).
#pragma once
#ifdef DXRV_TIMELINE
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace dxrv
{
constexpr uint32_t kTlMaxBlocks = 65536;
__device__ unsigned long long gTimeline[4 * kTlMaxBlocks];
__device__ unsigned long long gPhase[4 * kTlMaxBlocks];   // up to four intermediate stamps per CTA (0 = not reached)
struct TimelineScope
{
    unsigned long long t0;
    uint32_t role;
    __device__ static unsigned long long now() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
    __device__ TimelineScope() : t0(now()), role(0)
    {
        if (threadIdx.x == 0 && blockIdx.x < kTlMaxBlocks)
            for (int k = 0; k < 4; ++k) gPhase[4 * (size_t)blockIdx.x + k] = 0ull;
    }
    __device__ static void stamp(int k) { if (threadIdx.x == 0 && blockIdx.x < kTlMaxBlocks) gPhase[4 * (size_t)blockIdx.x + k] = now(); }
    __device__ ~TimelineScope()
    {
        __syncthreads();
        if (threadIdx.x == 0 && blockIdx.x < kTlMaxBlocks)
        {
            uint32_t sm;
            asm volatile("mov.u32 %0, %smid;" : "=r"(sm));
            unsigned long long* e = gTimeline + 4 * (size_t)blockIdx.x;
            e[0] = t0; e[1] = now(); e[2] = role; e[3] = sm;
        }
    }
};
inline void timelineReport(cudaStream_t s, uint32_t numBlocks, const char* const* roleNames, int numRoles)
{
    if (!std::getenv("DXRV_DBG_TIMELINE")) return;
    cudaStreamSynchronize(s);
    const uint32_t nb = std::min(numBlocks, kTlMaxBlocks);
    std::vector<unsigned long long> h(4 * (size_t)nb);
    cudaMemcpyFromSymbol(h.data(), gTimeline, h.size() * sizeof(unsigned long long));
    std::vector<unsigned long long> hp(4 * (size_t)nb);
    cudaMemcpyFromSymbol(hp.data(), gPhase, hp.size() * sizeof(unsigned long long));
    unsigned long long t0 = ~0ull;
    for (uint32_t b = 0; b < nb; ++b) t0 = std::min(t0, h[4 * b]);
    for (int role = 1; role < numRoles; ++role)
    {
        double s0 = 1e30, s1 = 0, e0 = 1e30, e1 = 0, dur = 0; int n = 0;
        for (uint32_t b = 0; b < nb; ++b)
            if ((int)h[4 * b + 2] == role)
            {
                const double a = (h[4 * b] - t0) * 1e-3, e = (h[4 * b + 1] - t0) * 1e-3;
                s0 = std::min(s0, a); s1 = std::max(s1, a); e0 = std::min(e0, e); e1 = std::max(e1, e); dur += e - a; ++n;
            }
        if (n) std::printf("  %-14s n=%5d start %.1f..%.1f end %.1f..%.1f avg dur %.2f us\n", roleNames[role], n, s0, s1, e0, e1, dur / n);
        // phases: start -> stamp 0 -> stamp 1 -> ... -> end, averaged over the CTAs of the role that reached every stamp
        double ph[5] = {0, 0, 0, 0, 0}; int np = 0;
        for (uint32_t b = 0; b < nb; ++b)
            if ((int)h[4 * b + 2] == role && hp[4 * b] && hp[4 * b + 1] && hp[4 * b + 2])
            {
                ph[0] += (hp[4 * b] - h[4 * b]) * 1e-3; ph[1] += (hp[4 * b + 1] - hp[4 * b]) * 1e-3; ph[2] += (hp[4 * b + 2] - hp[4 * b + 1]) * 1e-3;
                ph[3] += (h[4 * b + 1] - hp[4 * b + 2]) * 1e-3; ++np;
            }
        if (np) std::printf("      phases (n=%d): to item known %.2f | ids + zero + sync %.2f | trace %.2f | write-out %.2f us\n", np, ph[0] / np, ph[1] / np, ph[2] / np, ph[3] / np);
    }
    std::vector<double> fin(1024, 0.0);
    for (uint32_t b = 0; b < nb; ++b)
        if (h[4 * b + 2] >= 1 && h[4 * b + 2] <= 4) { double& f = fin[h[4 * b + 3] & 1023]; f = std::max(f, (h[4 * b + 1] - t0) * 1e-3); }
    std::vector<double> f;
    for (double v : fin) if (v > 0) f.push_back(v);
    std::sort(f.begin(), f.end());
    if (!f.empty())
        std::printf("  SM finish (roles 1..4): min %.1f p25 %.1f med %.1f p75 %.1f max %.1f us (%zu SMs)\n", f[0], f[f.size() / 4], f[f.size() / 2],
                    f[3 * f.size() / 4], f.back(), f.size());
}
}  // namespace dxrv
#define DXRV_TL_SCOPE() TimelineScope tlScope
#define DXRV_TL_ROLE(r) tlScope.role = (r)
#define DXRV_TL_STAMP(k) TimelineScope::stamp(k)
#define DXRV_TL_REPORT(s, nb, names, n) timelineReport(s, nb, names, n)
#else
#define DXRV_TL_SCOPE() ((void)0)
#define DXRV_TL_ROLE(r) ((void)0)
#define DXRV_TL_STAMP(k) ((void)0)
#define DXRV_TL_REPORT(s, nb, names, n) ((void)0)
#endif
