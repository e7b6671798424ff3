// host_pool.h -- a small persistent thread pool for the host side of the read-back (internal to libdxrv.so).
//
// The dense bit grid of a 1024^3 slab is 128 MiB: over one PCIe Gen5 x16 link that is 2.4 ms, and most of it is
// zeros.  dxrv_voxelize_to_host can bring the grid back as DXRV_FORMAT_SPARSE_BRICKS (a few MB) and expand it into
// the caller's dense buffer with these threads; the expansion is bound by the host's memory write bandwidth
// (~195 GB/s on the 16 cores of a B200 box: 0.69 ms), which is more than three times the link's.  Workers poll for the
// next batch for DXRV_HOST_SPIN_US microseconds (default 2000) before they block: a caller that voxelizes every frame
// finds them awake (waking 15 sleepers costs the first ~0.1 ms of a batch).
#pragma once
#include <cstddef>
#include <functional>

namespace dxrv
{
// Number of worker threads the pool runs with (DXRV_HOST_THREADS, default min(hardware threads, 32)).
unsigned hostPoolThreads();
// Run fn(task) for task = 0 .. numTasks-1 on the pool's threads (the caller takes part); returns when all are done.
void hostParallelFor(unsigned numTasks, const std::function<void(unsigned)>& fn);
// The same, asynchronous: returns a ticket at once; hostPoolWait(ticket) blocks until the tasks are done.  One batch
// at a time per process (a second begin waits for the first).
void hostParallelBegin(unsigned numTasks, const std::function<void(unsigned)>& fn);
void hostParallelWait();
}  // namespace dxrv
