#!/usr/bin/env python3
"""bench.py -- headline benchmark of the voxelization hot path (BASELINE.json: "Gvoxels/s end-to-end
incl. BVH build at 1/2/4/8 B200; ms per 1024^3 grid").

A step = one pass of the hot path over one mesh: LBVH build (bounds, Morton, onesweep sort, Karras
hierarchy, refit) + MODE_PARITY trace/fill of the bit-packed grid.  Workload: the reference's
Stanford dragon (Bin/Assets/dragon.obj via Dragon.bat) at 1024^3 voxels PER GPU:
  N=1      1024^3 grid, one GPU does all of it (BASELINE configs[2] at one GPU).
  N=2,4,8  z-slab sharding, weak scaling: the grid grows to 1280^3 / 1664^3 / 2048^3 so every rank
           still fills ~1024^3 voxels (its own z-slab); the mesh/BVH is replicated, no data-path
           collective in the timed region.  ("zslab_1024" in the JSON additionally reports the
           strong-scaling number: ONE 1024^3 grid split into N slabs.)
`value`      inputs (vertex/index buffers) resident in HBM; CUDA-event time, max over ranks.
`e2e`        same metric through the C ABI with HOST buffers: H2D of the mesh, build, voxelize and
             D2H of the rank's slab inside the timed region (wall clock around synchronising calls).
`roofline`   dominant kernel (k_trace_fill_columns): algorithmic bytes / its CUDA-event time.
`cpu_baseline` the CPU oracle (a port: the reference itself is DXR/Windows-only) on the host cores.
`--impl reference` times that same oracle as the reference arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "gvoxels_per_s_incl_bvh_build"
UNIT = "Gvoxel/s"
GRID_FOR_GPUS = {1: 1024, 2: 1280, 4: 1664, 8: 2048}   # ~1024^3 voxels per GPU


def grid_for(n_gpus):
    if n_gpus in GRID_FOR_GPUS:
        return GRID_FOR_GPUS[n_gpus]
    n = int(round(1024 * n_gpus ** (1.0 / 3.0) / 32.0)) * 32
    while n % n_gpus:
        n += 32
    return n


def slab_of(rank, world, N):
    return N * rank // world, N * (rank + 1) // world


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.05] or [r for _, r in self.rows]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    """All host cores this process may use (torchrun pins OMP_NUM_THREADS=1, so ask the OS instead)."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def load_workload():
    import dxrvoxelizer_b200 as d
    return d.load_obj(d.asset_path("dragon.obj"))


# ---------------------------------------------------------------------------------------------------
def run_reference(args):
    """Reference arm: the reference's own implementation cannot run here (Windows + D3D12/DXR +
    binary-only XUSG), so this times the CPU oracle port of its algorithm with all host threads on
    the same workload.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    mesh = load_workload()
    world = args.gpus
    N = grid_for(world)
    threads = host_threads()
    # bounded sample: the central z-slab one GPU owns (N/world layers ~ 1024^3 voxels; the whole grid
    # at N=1), ~0.5 s per step on 8 cores
    layers = max(1, N // world)
    z0 = (N - layers) // 2
    for _ in range(args.warmup):
        oracle.voxelize(mesh.vertices, mesh.indices, N, oracle.MODE_PARITY, z0=z0, z1=z0 + layers, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.voxelize(mesh.vertices, mesh.indices, N, oracle.MODE_PARITY, z0=z0, z1=z0 + layers, threads=threads)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    value = layers * N * N / dt * 1e-9
    sample = "central z-slab of %d layers of the %d^3 dragon grid per step (MODE_PARITY, own yz-bin build included)" % (layers, N)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "reference asset dragon.obj (Stanford dragon, 100k triangles)",
        "config": workload_config(N, world, mesh),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(N, world, mesh):
    return {"workload": "dragon.obj %d^3 MODE_PARITY, LBVH rebuilt every step, z-slab per GPU" % N,
            "grid": N, "voxels_per_gpu_mean": N * N * N // world, "slabs": "one z-slab per GPU, cut points balance stores + triangles per layer",
            "triangles": mesh.num_triangles,
            "vertices": mesh.num_vertices, "mode": "parity", "parallelism": "zslab%d" % world,
            "l2": "flushed between timed steps (256 MiB device write, untimed)"}


# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import dxrvoxelizer_b200 as d
    from dxrvoxelizer_b200 import _lib as L

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    N = grid_for(world)
    stream = torch.cuda.Stream()
    vox = d.Voxelizer(local)
    vox.set_stream(stream.cuda_stream)

    # ---- inputs: rank 0 loads the OBJ; the mesh is replicated by NCCL broadcast over NVLink ----
    if rank == 0:
        mesh = load_workload()
        meta = torch.tensor([mesh.num_vertices, mesh.stride, mesh.indices.size], dtype=torch.int64, device="cuda")
    else:
        mesh, meta = None, torch.zeros(3, dtype=torch.int64, device="cuda")
    if world > 1:
        dist.broadcast(meta, 0)
    nv, stride, ni = (int(x) for x in meta.tolist())
    with numa_local(local):
        h_vb = torch.empty(nv * stride, dtype=torch.uint8).pin_memory()
        h_ib = torch.empty(ni, dtype=torch.int32).pin_memory()
    if rank == 0:
        h_vb.copy_(torch.from_numpy(mesh.vertex_bytes))
        h_ib.copy_(torch.from_numpy(mesh.indices.view(np.int32)))
    d_vb = h_vb.cuda()
    d_ib = h_ib.cuda()
    if world > 1:
        dist.broadcast(d_vb, 0)
        dist.broadcast(d_ib, 0)
        h_vb.copy_(d_vb)
        h_ib.copy_(d_ib)
    torch.cuda.synchronize()
    T = ni // 3
    # z-slab of this rank.  Equal slabs are badly balanced for a real mesh (the dragon is thin along z: the ranks
    # owning its layers would do all the crossing tests), so the cut points balance a cost model instead:
    # stores + triangles overlapping the layer (dxrvoxelizer_b200.sharding.balanced_slabs, host-side, untimed,
    # identical on every rank because the mesh is replicated).
    from dxrvoxelizer_b200.sharding import balanced_slabs
    host_mesh = d.Mesh(h_vb.numpy(), h_ib.numpy().view(np.uint32), stride)
    z0, z1 = balanced_slabs(host_mesh, N, world)[rank]

    slab_bytes = (z1 - z0) * N * ((N + 31) // 32) * 4
    with numa_local(local):
        h_grid = torch.empty(slab_bytes, dtype=torch.uint8).pin_memory()
        h_grid.zero_()   # touch the pages while the thread still sits next to the GPU
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def step_resident():
        vox.build_bvh_device(d_vb.data_ptr(), nv, stride, d_ib.data_ptr(), ni)
        vox.voxelize(N, d.MODE_PARITY, z0, z1)

    def step_e2e():
        if world > 1:
            # replicate the host mesh of rank 0: H2D on rank 0, NCCL broadcast, build from device memory
            with torch.cuda.stream(stream):
                if rank == 0:
                    d_vb.copy_(h_vb, non_blocking=True)
                    d_ib.copy_(h_ib, non_blocking=True)
                dist.broadcast(d_vb, 0)
                dist.broadcast(d_ib, 0)
            vox.build_bvh_device(d_vb.data_ptr(), nv, stride, d_ib.data_ptr(), ni)
        else:
            vox.build_bvh_host_ptr(h_vb.data_ptr(), nv, stride, h_ib.data_ptr(), ni)
        vox.voxelize(N, d.MODE_PARITY, z0, z1)
        vox.fetch_into(h_grid.data_ptr(), slab_bytes)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # ---- device-resident arm --------------------------------------------------------------------
    for _ in range(max(3, args.warmup)):
        step_resident()
    vox.synchronize()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    launches0 = vox.info(L.INFO_KERNEL_LAUNCHES)
    barrier()
    t_wall0 = time.time()
    for i in range(args.steps):
        with torch.cuda.stream(stream):
            flush.fill_(i & 0xff)                      # evict L2 (untimed)
        ev[i][0].record(stream)
        vox.build_bvh_device(d_vb.data_ptr(), nv, stride, d_ib.data_ptr(), ni)
        ev[i][1].record(stream)
        vox.voxelize(N, d.MODE_PARITY, z0, z1)
        ev[i][2].record(stream)
    barrier()
    t_wall1 = time.time()
    launches = vox.info(L.INFO_KERNEL_LAUNCHES) - launches0
    vox.synchronize()
    build_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / args.steps
    trace_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / args.steps
    step_ms = sum(e[0].elapsed_time(e[2]) for e in ev) / args.steps
    crossings = vox.info(L.INFO_CROSSINGS)

    # ---- the dominant kernel on its own: CUDA events recorded by the library on the launching stream
    # right around k_walk_columns / k_trace_fill_columns, same step sequence and L2 flush as above ----
    vox.set_profiling(True)
    walk_ns = fill_ns = 0
    prof_steps = min(args.steps, 50)
    for i in range(prof_steps):
        with torch.cuda.stream(stream):
            flush.fill_(i & 0xff)
        vox.build_bvh_device(d_vb.data_ptr(), nv, stride, d_ib.data_ptr(), ni)
        vox.voxelize(N, d.MODE_PARITY, z0, z1)
        walk_ns += vox.info(L.INFO_LAST_WALK_NS)
        fill_ns += vox.info(L.INFO_LAST_FILL_NS)
    vox.set_profiling(False)
    walk_ms, fill_ms = walk_ns / prof_steps * 1e-6, fill_ns / prof_steps * 1e-6

    # ---- end-to-end arm (host buffers through the C ABI) ----------------------------------------
    for _ in range(3):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    clocks = sampler.stop(t_wall0, time.time()) if rank == 0 else None
    # where the end-to-end time goes (separate short loop with a synchronisation between compute and read-back)
    up_ms = down_ms = 0.0
    if world == 1:
        reps = min(args.steps, 20)
        for _ in range(reps):
            ta = time.perf_counter()
            vox.build_bvh_host_ptr(h_vb.data_ptr(), nv, stride, h_ib.data_ptr(), ni)
            vox.voxelize(N, d.MODE_PARITY, z0, z1)
            vox.synchronize()
            tb = time.perf_counter()
            vox.fetch_into(h_grid.data_ptr(), slab_bytes)
            tc = time.perf_counter()
            up_ms += (tb - ta) * 1e3 / reps
            down_ms += (tc - tb) * 1e3 / reps

    # ---- strong-scaling side number: ONE 1024^3 grid split into `world` slabs ---------------------
    zs_ms = None
    if world > 1:
        a, b = balanced_slabs(host_mesh, 1024, world)[rank]
        def step_1024():
            vox.build_bvh_device(d_vb.data_ptr(), nv, stride, d_ib.data_ptr(), ni)
            vox.voxelize(1024, d.MODE_PARITY, a, b)
        for _ in range(3):
            step_1024()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step_1024()
        e1.record(stream)
        barrier()
        zs_ms = e0.elapsed_time(e1) / args.steps

    # ---- max over ranks ---------------------------------------------------------------------------
    fill_ms_rank0 = fill_ms   # the roofline of the kernel is a per-GPU figure: rank 0's launches against rank 0's bytes
    times = torch.tensor([step_ms, build_ms, trace_ms, e2e_ms, zs_ms or 0.0, walk_ms, fill_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    step_ms, build_ms, trace_ms, e2e_ms, zs_ms, walk_ms, fill_ms = times.tolist()

    if rank == 0:
        total_voxels = float(N) ** 3
        peak, peak_src = measured_peak()
        # algorithmic bytes of one k_trace_fill_columns launch on one GPU: the slab
        # of the bit grid written once + every scene-space triangle (48 B) read once.  (The BVH nodes
        # are read by k_walk_columns, which is latency bound and reported under phases_ms.)
        alg_bytes = slab_bytes + 48 * T
        achieved = alg_bytes / (fill_ms_rank0 * 1e-3) * 1e-9
        cpu = cpu_baseline(mesh, N, world)
        out = {
            "metric": METRIC, "value": total_voxels / (step_ms * 1e-3) * 1e-9, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "reference asset dragon.obj (Stanford dragon, 100k triangles); no synthetic substitution needed",
            "config": workload_config(N, world, mesh),
            "phases_ms": {"bvh_build": build_ms, "voxelize": trace_ms, "k_walk_columns": walk_ms, "k_trace_fill_columns": fill_ms},
            "ms_per_1024_cubed_grid": step_ms if world == 1 else zs_ms,
            "e2e": {"value": total_voxels / (e2e_ms * 1e-3) * 1e-9, "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(nv * stride + ni * 4), "d2h_bytes_per_step": int(N * N * ((N + 31) // 32) * 4),
                    "timing": "wall clock around synchronising C-ABI calls, max over ranks",
                    "phases_ms": ({"h2d_mesh_build_voxelize": up_ms, "d2h_grid": down_ms} if world == 1 else None)},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "k_trace_fill_columns", "bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(),
                         "algorithmic_bytes_per_launch": int(alg_bytes), "kernel_ms": fill_ms_rank0, "peak_source": peak_src,
                         "frac_of_nominal_8000_GBps": achieved / 8000.0,
                         "timing": "cudaEventRecord on the launching stream around the kernel, mean of %d launches" % prof_steps},
            "cpu_baseline": cpu,
            "clocks": clocks,
            "crossings": int(crossings),
        }
        if world > 1:
            out["zslab_1024"] = {"ms_per_1024_cubed_grid": zs_ms, "gvoxels_per_s": 1024.0 ** 3 / (zs_ms * 1e-3) * 1e-9,
                                 "scaling": "strong"}
        print(json.dumps(out))
    vox.close()
    if world > 1:
        dist.destroy_process_group()


class numa_local:
    """Allocate pinned host buffers on the NUMA node next to GPU `index`: the thread is moved onto the CPUs NVML
    reports for that GPU while the pages are allocated and touched, then gets its original affinity back (the CPU
    baseline must see every core).  A D2H copy into far memory crosses the socket interconnect and loses bandwidth;
    with 8 ranks reading back at once it is the difference between the PCIe links and one saturated socket link.
    Best effort: without NVML (or on a single-node box) it does nothing."""
    def __init__(self, index):
        self.index, self.saved = index, None

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
            allowed = os.sched_getaffinity(0)
            if cpus & allowed:
                self.saved = allowed
                os.sched_setaffinity(0, cpus & allowed)
        except Exception:
            self.saved = None
        return self

    def __exit__(self, *exc):
        if self.saved is not None:
            os.sched_setaffinity(0, self.saved)
        return False


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    ncu --set full capture (profiles/), or null when no capture is recorded."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get("k_trace_fill_columns_dram_bytes_per_launch")
    except Exception:
        return None


def cpu_baseline(mesh, N, world):
    """The CPU oracle (a port of the reference's algorithm) on the box's host cores, bounded sample."""
    import oracle
    threads = host_threads()
    layers = max(1, N // world)
    z0 = (N - layers) // 2
    oracle.voxelize(mesh.vertices, mesh.indices, N, oracle.MODE_PARITY, z0=z0, z1=z0 + layers, threads=threads)
    best = 1e30
    t_all = time.perf_counter()
    for _ in range(3):
        t = time.perf_counter()
        oracle.voxelize(mesh.vertices, mesh.indices, N, oracle.MODE_PARITY, z0=z0, z1=z0 + layers, threads=threads)
        best = min(best, time.perf_counter() - t)
        if time.perf_counter() - t_all > 25:
            break
    return {"value": layers * N * N / best * 1e-9, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "central z-slab of %d layers of the %d^3 dragon grid, MODE_PARITY, own acceleration build included, best of 3" % (layers, N)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
