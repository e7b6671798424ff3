// host_pool.cpp -- see host_pool.h
#include "host_pool.h"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <vector>

namespace dxrv
{
namespace
{
struct Pool
{
    std::vector<std::thread> workers;
    std::mutex m;                       // guards the batch description and `generation`
    std::condition_variable wake, done;
    std::function<void(unsigned)> fn;
    std::atomic<unsigned> numTasks{0}, finished{0};
    std::atomic<unsigned long long> state{0};
    unsigned long generation = 0, current = 0;
    bool busy = false, stop = false;
    std::atomic<unsigned long> published{0};   // == generation once a batch is fully described (read without the mutex)
    std::atomic<bool> stopping{false};
    long spinUs = 2000;                 // a worker polls this long for the next batch before it blocks (DXRV_HOST_SPIN_US)
    std::mutex callers;                 // one batch at a time

    explicit Pool(unsigned n)
    {
        if (const char* e = std::getenv("DXRV_HOST_SPIN_US")) { const long v = std::atol(e); if (v >= 0 && v <= 1000000) spinUs = v; }
        for (unsigned i = 0; i + 1 < n; ++i) workers.emplace_back([this] { loop(); });
    }
    ~Pool()
    {
        {
            std::lock_guard<std::mutex> g(m);
            stop = true;
            stopping.store(true, std::memory_order_release);
        }
        wake.notify_all();
        for (auto& t : workers) t.join();
    }
    // Tasks are handed out through ONE 64-bit word {generation, next index}: a worker that is late for its batch
    // (still on its way into drain() when the batch completes and the next one is set up) sees the other generation
    // and leaves, instead of taking a task with a stale task count.
    void drain(unsigned long gen)
    {
        const unsigned n = numTasks.load(std::memory_order_acquire);
        for (;;)
        {
            unsigned long long st = state.load(std::memory_order_acquire);
            unsigned t;
            for (;;)
            {
                if ((unsigned long)(st >> 32) != (gen & 0xfffffffful)) return;
                t = (unsigned)(st & 0xffffffffull);
                if (t >= n) return;
                if (state.compare_exchange_weak(st, st + 1, std::memory_order_acq_rel, std::memory_order_acquire)) break;
            }
            fn(t);
            if (finished.fetch_add(1, std::memory_order_acq_rel) + 1 == n)
            {
                std::lock_guard<std::mutex> g(m);
                busy = false;
                done.notify_all();
            }
        }
    }
    void loop()
    {
        // A caller that voxelizes every frame starts a batch every millisecond or so: a worker that went to sleep on
        // the condition variable after every batch would spend the first ~0.1 ms of the next one waking up (futex,
        // the convoy on the mutex, a core that has dropped into an idle state).  So a worker first POLLS for the next
        // batch for spinUs microseconds and only then blocks; an idle process costs nothing after that.
        unsigned long seen = 0;
        for (;;)
        {
            bool found = false;
            if (seen != 0 && spinUs > 0)
            {
                const auto until = std::chrono::steady_clock::now() + std::chrono::microseconds(spinUs);
                for (unsigned polls = 0;; ++polls)
                {
                    if (stopping.load(std::memory_order_acquire)) return;
                    const unsigned long g = published.load(std::memory_order_acquire);
                    if (g != seen) { seen = g; found = true; break; }
#if defined(__x86_64__)
                    __builtin_ia32_pause();
#endif
                    if ((polls & 63u) == 63u && std::chrono::steady_clock::now() >= until) break;
                }
            }
            if (!found)
            {
                std::unique_lock<std::mutex> g(m);
                wake.wait(g, [&] { return stop || generation != seen; });
                if (stop) return;
                seen = generation;
            }
            drain(seen);
        }
    }
    void begin(unsigned n, const std::function<void(unsigned)>& f)
    {
        callers.lock();
        {
            std::lock_guard<std::mutex> g(m);
            ++generation;
            state.store(0xffffffffull << 32);   // (no generation is 2^32 - 1 at this point: nobody takes a task while the batch is described)
            fn = f; finished.store(0);
            numTasks.store(n);
            busy = n != 0;
            current = generation;
            state.store((unsigned long long)(generation & 0xfffffffful) << 32);
            published.store(generation, std::memory_order_release);
        }
        wake.notify_all();
    }
    void wait()
    {
        drain(current);   // the caller helps
        {
            std::unique_lock<std::mutex> g(m);
            done.wait(g, [&] { return !busy; });
        }
        callers.unlock();
    }
};

unsigned poolSize()
{
    unsigned n = std::thread::hardware_concurrency();
    if (n == 0) n = 4;
    if (n > 32) n = 32;
    if (const char* e = std::getenv("DXRV_HOST_THREADS"))
    {
        const int v = std::atoi(e);
        if (v >= 1 && v <= 256) n = (unsigned)v;
    }
    return n;
}

Pool& pool()
{
    static Pool p(poolSize());
    return p;
}
}  // namespace

unsigned hostPoolThreads() { return (unsigned)pool().workers.size() + 1u; }

void hostParallelBegin(unsigned numTasks, const std::function<void(unsigned)>& fn) { pool().begin(numTasks, fn); }

void hostParallelWait() { pool().wait(); }

void hostParallelFor(unsigned numTasks, const std::function<void(unsigned)>& fn)
{
    hostParallelBegin(numTasks, fn);
    hostParallelWait();
}
}  // namespace dxrv
