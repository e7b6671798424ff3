#!/usr/bin/env python3
"""bench.py -- benchmark of the voxelization hot path (BASELINE.json: "Gvoxels/s end-to-end incl. BVH build at
1/2/4/8 B200; ms per 1024^3 grid").

Default workload (--config c3, BASELINE.json configs[2]): the reference's Stanford dragon (Bin/Assets/dragon.obj via
Dragon.bat) at ONE 1024^3 bit-packed grid, MODE_PARITY, the LBVH rebuilt every step.
  N=1      one GPU fills the whole grid.
  N=2,4,8  STRONG scaling: the same 1024^3 grid split into N cost-balanced z-slabs, one per rank; the mesh is
           replicated (NCCL broadcast), every rank builds the identical LBVH, no collective in the timed region.
           (`weak_scaling` in the JSON is a side number: the grid grown to 1280^3/1664^3/2048^3.)
A step = acceleration-structure build + trace/fill of the rank's slab.  What the build holds: bounds, Morton keys,
onesweep radix sort, Morton-sorted scene-space triangle records -- and the Karras hierarchy + node boxes WHEN the
consumer traverses them (include/dxrv.h states the rule).  MODE_SHADER and the LBVH-walk candidate path do;
MODE_PARITY's default path finds its candidates by triangle-parallel 2-D binning and does not, so the headline step
builds no hierarchy.  `phases_ms.step_full_lbvh_walk` is the same step WITH the full LBVH (hierarchy + boxes rebuilt
every step, candidates by the warp-cooperative walk; DXRV_PARITY_CANDIDATES=walk), gated against the oracle too.

`mismatched_voxels`  correctness GATE, run before anything is timed: every rank fetches its slab and XORs it against
             the CPU oracle's grid of the same layers; the sum over ranks must be 0 or the bench exits non-zero.
             (`shader_mismatched_voxels`: the same for MODE_SHADER on a few layers of the slab.)
`value`      inputs (vertex/index buffers) resident in HBM; CUDA-event time, max over ranks.
`e2e`        same metric through the C ABI with HOST buffers: H2D of the mesh, build, voxelize and D2H of the
             rank's slab inside the timed region (wall clock around synchronising calls, max over ranks).
`roofline`   dominant kernel (k_trace_fill_columns) on rank 0: algorithmic bytes / its CUDA-event time.
`phases_ms.shader_1024`  MODE_SHADER (the reference's own function: radial closest-hit rays) on the same slab,
             direction bins rebuilt every step.
`cpu_baseline` (N=1) the CPU oracle (a port: the reference itself is DXR/Windows-only) on all host cores.
`--impl reference` times that same oracle as the reference arm; it never imports the product package.
`--config c4` (synthetic 16.8 M-triangle torus knot at 512^3, build-dominated) and `--config c5` (256 distinct
meshes, OBJ text -> grid, at 256^3, mesh-parallel) are BASELINE.json configs[3] and [4].
"""
import argparse
import json
import lzma
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "gvoxels_per_s_incl_bvh_build"
UNIT = "Gvoxel/s"
GRID = 1024
WEAK_GRID = {2: 1280, 4: 1664, 8: 2048}   # ~1024^3 voxels per GPU (side number only)


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


def popcount(a):
    return int(np.unpackbits(np.ascontiguousarray(a).reshape(-1).view(np.uint8)).sum())


# ---- mesh loading without the product (reference arm) ---------------------------------------------------------
def unpack_asset(name):
    """assets/<name>.xz -> a temp file (the product's dxrvoxelizer_b200.assets does the same; the reference arm
    must not import the product package)."""
    out = os.path.join(tempfile.gettempdir(), "dxrv_bench_%d_%s" % (os.getuid(), name))
    if not os.path.exists(out):
        with open(os.path.join(ROOT, "assets", name + ".xz"), "rb") as f:
            raw = lzma.decompress(f.read())
        tmp = out + ".tmp%d" % os.getpid()
        with open(tmp, "wb") as f:
            f.write(raw)
        os.replace(tmp, out)
    return out


def load_mesh_without_product(name):
    """(vertices float32 [nv, stride/4], indices uint32) through the REFERENCE's own ObjLoader compiled into
    oracle/_ref (XUSG/Optional/XUSGObjLoader.cpp, unmodified); positions-only numpy parse when that is absent."""
    import oracle
    path = unpack_asset(name)
    if oracle.ref_loader_available():
        vb, ib, stride, _ = oracle.ref_load_obj(path)
        return vb.view(np.float32).reshape(-1, stride // 4), ib, "oracle/_ref ObjLoader (the reference's own loader)"
    pos, tri = [], []
    with open(path) as f:
        for line in f:
            if line.startswith("v "):
                x, y, z = line.split()[1:4]
                pos.append((float(x), float(y), -float(z)))          # z flip, XUSGObjLoader.cpp:198
            elif line.startswith("f "):
                c = [int(t.split("/")[0]) - 1 for t in line.split()[1:]]
                for k in range(1, len(c) - 1):
                    tri.append((c[0], c[k], c[k + 1]))               # fan, XUSGObjLoader.cpp:263-297
    idx = np.asarray(tri, np.uint32).reshape(-1)[::-1].copy()         # whole index array reversed, :227
    return np.asarray(pos, np.float32), idx, "numpy positions-only parse (oracle/_ref absent)"


# ---- reference arm ----------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own implementation cannot run here (Windows + D3D12/DXR + binary-only XUSG), so this times
    the CPU oracle port of the path with all host threads on the same workload: the WHOLE 1024^3 dragon grid per
    step at any N (strong scaling: the job does not grow with N).  The algorithm timed is MODE_PARITY, the same one
    the GPU arm's headline uses (the oracle's MODE_SHADER rate is reported beside it).  Rank 0 only; the product
    package is never imported here."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    verts, idx, how = load_mesh_without_product("dragon.obj")
    world = args.gpus
    threads = host_threads()
    N = GRID
    for _ in range(max(1, args.warmup)):
        oracle.voxelize(verts, idx, N, oracle.MODE_PARITY, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.voxelize(verts, idx, N, oracle.MODE_PARITY, threads=threads)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    value = float(N) ** 3 / dt * 1e-9
    # the reference's actual function (radial closest hit): a bounded sample of 4 central layers
    shader_rate = float("nan")
    if verts.shape[1] >= 6:                                      # (needs vertex normals)
        t1 = time.perf_counter()
        oracle.voxelize(verts, idx, N, oracle.MODE_SHADER, z0=N // 2 - 2, z1=N // 2 + 2, threads=threads)
        shader_rate = 4.0 * N * N / (time.perf_counter() - t1) * 1e-9
    sample = "the whole %d^3 dragon grid per step, MODE_PARITY (column-parity algorithm), own yz-bin build included" % N
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(1, args.warmup), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "reference asset dragon.obj (Stanford dragon, 100k triangles)",
        "config": workload_config(N, world, idx.size // 3, verts.shape[0]),
        "mesh_loader": how,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "algorithm": "MODE_PARITY (column parity); the reference's shader algorithm (MODE_SHADER) runs at "
                                      "%.3f Gvoxel/s on the same cores (4 central layers sampled)" % shader_rate,
                         "shader_gvoxels_per_s": shader_rate},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(N, world, tris, verts):
    return {"workload": "dragon.obj %d^3 MODE_PARITY, acceleration structure rebuilt every step, one grid split into %d z-slab(s)" % (N, world),
            "acceleration_structure": "bounds + Morton keys + radix sort + sorted triangle records + per-tile candidate bins, all inside the timed "
                                      "step; the Karras hierarchy is built on demand and this mode's default path does not traverse it "
                                      "(phases_ms.step_full_lbvh_walk = the step with hierarchy + LBVH walk)",
            "grid": N, "voxels_total": N * N * N, "slabs": "one z-slab per GPU, cut points balance stores + triangles per layer",
            "triangles": int(tris), "vertices": int(verts), "mode": "parity", "parallelism": "zslab%d" % world,
            "l2": "flushed between timed steps (256 MiB device write, untimed)"}


# ---- our arm -----------------------------------------------------------------------------------------------------
class Rig:
    """Process-group, device and context plumbing shared by the configs."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import dxrvoxelizer_b200 as d
        self.torch, self.dist, self.d = torch, dist, d
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, self.world))
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.stream = torch.cuda.Stream()
        self.vox = d.Voxelizer(self.local)
        self.vox.set_stream(self.stream.cuda_stream)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce_max(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def reduce_sum(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.tolist()

    def flush_l2(self, i=0):
        with self.torch.cuda.stream(self.stream):
            self.flush.fill_(i & 0xff)

    def replicate_mesh(self, mesh):
        """rank 0's mesh -> pinned host + device copies on every rank (NCCL broadcast over NVLink)."""
        torch, dist = self.torch, self.dist
        if self.rank == 0:
            meta = torch.tensor([mesh.num_vertices, mesh.stride, mesh.indices.size], dtype=torch.int64, device="cuda")
        else:
            meta = torch.zeros(3, dtype=torch.int64, device="cuda")
        if self.world > 1:
            dist.broadcast(meta, 0)
        nv, stride, ni = (int(x) for x in meta.tolist())
        with numa_local(self.local):
            h_vb = torch.empty(nv * stride, dtype=torch.uint8).pin_memory()
            h_ib = torch.empty(ni, dtype=torch.int32).pin_memory()
        if self.rank == 0:
            h_vb.copy_(torch.from_numpy(mesh.vertex_bytes))
            h_ib.copy_(torch.from_numpy(mesh.indices.view(np.int32)))
        d_vb, d_ib = h_vb.cuda(), h_ib.cuda()
        if self.world > 1:
            dist.broadcast(d_vb, 0)
            dist.broadcast(d_ib, 0)
            h_vb.copy_(d_vb)
            h_ib.copy_(d_ib)
        torch.cuda.synchronize()
        host_mesh = self.d.Mesh(h_vb.numpy(), h_ib.numpy().view(np.uint32), stride)
        return host_mesh, (h_vb, h_ib, d_vb, d_ib, nv, stride, ni)

    def close(self):
        self.vox.close()
        if self.world > 1:
            self.dist.destroy_process_group()


GATE_REF = {}


def gate_slab(rig, host_mesh, N, mode, z0, z1, layers=None):
    """XOR-popcount of this rank's slab against the CPU oracle (the checker; nothing here is timed).
    layers: None = every layer of the slab, else an iterable of single layers (MODE_SHADER: the oracle needs
    ~0.1 s per 1024^2 layer)."""
    import oracle
    threads = max(1, host_threads() // rig.world)
    rig.vox.voxelize(N, mode, z0, z1)
    got = rig.vox.fetch_bits()
    mism = 0
    if layers is None:
        ref = oracle.voxelize(host_mesh.vertices, host_mesh.indices, N, mode, z0=z0, z1=z1, threads=threads)["bits"]
        GATE_REF[mode] = ref
        mism = popcount(got ^ ref)
        checked = z1 - z0
    else:
        checked = 0
        for z in layers:
            ref = oracle.voxelize(host_mesh.vertices, host_mesh.indices, N, mode, z0=z, z1=z + 1, threads=threads)["bits"]
            mism += popcount(got[z - z0:z - z0 + 1] ^ ref)
            checked += 1
    return mism, checked


def run_c3(args):
    import dxrvoxelizer_b200 as d
    from dxrvoxelizer_b200 import _lib as L
    from dxrvoxelizer_b200.sharding import balanced_slabs
    rig = Rig(args)
    torch, vox, stream, rank, world = rig.torch, rig.vox, rig.stream, rig.rank, rig.world
    N = GRID
    # host threads of the library's read-back pool (dxrv_voxelize_to_host, sparse transport): the ranks of one box share its cores
    host_pool_threads = max(2, min(32, host_threads() // world))
    os.environ.setdefault("DXRV_HOST_THREADS", str(host_pool_threads))
    host_pool_threads = int(os.environ["DXRV_HOST_THREADS"])
    mesh = d.load_obj(d.asset_path("dragon.obj")) if rank == 0 else None
    host_mesh, (h_vb, h_ib, d_vb, d_ib, nv, stride, ni) = rig.replicate_mesh(mesh)
    T = ni // 3
    # z-slab of this rank.  Equal slabs are badly balanced for a real mesh (the dragon is thin along z), so the cut
    # points balance a cost model instead (dxrvoxelizer_b200.sharding.balanced_slabs: host-side, untimed, identical
    # on every rank because the mesh is replicated).
    z0, z1 = balanced_slabs(host_mesh, N, world)[rank]
    P = (N + 31) // 32
    slab_bytes = (z1 - z0) * N * P * 4

    def build():
        vox.build_bvh_device(d_vb.data_ptr(), nv, stride, d_ib.data_ptr(), ni)

    # ---- correctness gate (before any timing) ---------------------------------------------------------------
    build()
    mism, checked = gate_slab(rig, host_mesh, N, d.MODE_PARITY, z0, z1)
    gate_ref = GATE_REF[d.MODE_PARITY]                           # the oracle's slab: also checks the e2e arm's host buffer
    # MODE_SHADER runs on the same cost-balanced slabs (measured on 8 GPUs: 2.7 ms against 4.5 ms with equal slabs --
    # the layers through the mesh are the expensive ones for the radial rays too)
    sz0, sz1 = z0, z1
    mid = (sz0 + sz1) // 2
    shader_layers = sorted({sz0, mid, min(mid + 1, sz1 - 1), sz1 - 1})
    smism, schecked = gate_slab(rig, host_mesh, N, d.MODE_SHADER, sz0, sz1, layers=shader_layers)
    mism_total, smism_total, checked_total, schecked_total = rig.reduce_sum([mism, smism, checked, schecked])
    if mism_total != 0 or smism_total != 0:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "error": "GPU grid differs from the CPU oracle", "n_gpus": world,
                              "mismatched_voxels": int(mism_total), "shader_mismatched_voxels": int(smism_total)}))
        rig.close()
        raise SystemExit(3)

    sampler = ClockSampler(rig.local)
    if rank == 0:
        sampler.start()

    # ---- device-resident arm --------------------------------------------------------------------------------
    def step_resident():
        build()
        vox.voxelize(N, d.MODE_PARITY, z0, z1)

    warm = max(3, args.warmup)
    for _ in range(warm):
        step_resident()
    vox.synchronize()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    launches0 = vox.info(L.INFO_KERNEL_LAUNCHES)
    rig.barrier()
    t_wall0 = time.time()
    for i in range(args.steps):
        rig.flush_l2(i)                                # evict L2 (untimed)
        ev[i][0].record(stream)
        build()
        ev[i][1].record(stream)
        vox.voxelize(N, d.MODE_PARITY, z0, z1)
        ev[i][2].record(stream)
    rig.barrier()
    launches = vox.info(L.INFO_KERNEL_LAUNCHES) - launches0
    vox.synchronize()
    build_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / args.steps
    trace_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / args.steps
    step_ms = sum(e[0].elapsed_time(e[2]) for e in ev) / args.steps
    crossings = vox.info(L.INFO_CROSSINGS)

    # ---- the dominant kernel on its own: CUDA events recorded by the library on the launching stream right
    # around k_walk_columns / k_trace_fill_columns, same step sequence and L2 flush as above -------------------
    vox.set_profiling(True)
    walk_ns = fill_ns = 0
    prof_steps = min(args.steps, 50)
    for i in range(prof_steps):
        rig.flush_l2(i)
        step_resident()
        walk_ns += vox.info(L.INFO_LAST_WALK_NS)
        fill_ns += vox.info(L.INFO_LAST_FILL_NS)
    vox.set_profiling(False)
    walk_ms, fill_ms = walk_ns / prof_steps * 1e-6, fill_ns / prof_steps * 1e-6

    # ---- the same step with the FULL LBVH: hierarchy + node boxes rebuilt every step, candidates by the LBVH walk ---
    # (a side number: what the step costs when the consumer traverses the tree; gated like the headline)
    os.environ["DXRV_PARITY_CANDIDATES"] = "walk"
    for _ in range(3):
        step_resident()                                # (the first voxelize builds the tree on demand, later builds include it)
    full_bits = vox.fetch_bits()
    full_mism = popcount(full_bits ^ gate_ref)
    del full_bits
    full_steps = min(args.steps, 50)
    fev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(full_steps)]
    rig.barrier()
    for i in range(full_steps):
        rig.flush_l2(i)
        fev[i][0].record(stream)
        step_resident()
        fev[i][1].record(stream)
    rig.barrier()
    vox.synchronize()
    full_ms = sum(e[0].elapsed_time(e[1]) for e in fev) / full_steps
    del os.environ["DXRV_PARITY_CANDIDATES"]
    for _ in range(2):
        step_resident()                                # back to the default path (no hierarchy wanted)
    vox.synchronize()

    # ---- MODE_SHADER on the same slab (the reference's own function), bins rebuilt every step ------------------
    shader_steps = max(2, min(args.steps, 5))
    for _ in range(2):
        build(); vox.voxelize(N, d.MODE_SHADER, sz0, sz1)
    rig.barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record(stream)
    for _ in range(shader_steps):
        build(); vox.voxelize(N, d.MODE_SHADER, sz0, sz1)
    s1.record(stream)
    rig.barrier()
    shader_ms = s0.elapsed_time(s1) / shader_steps

    # ---- end-to-end arm (host buffers through the C ABI) ------------------------------------------------------
    # When the grid goes back to the host the slabs are cut by BYTES, not by compute: a layer costs ~0.06 us to fill
    # and 2.4 us to copy (128 KiB over PCIe), so equal slabs minimise the longest read-back (the compute-balanced
    # cuts of the device-resident arm give the sparse ends of the dragon a quarter of the grid each).
    if world > 1:
        from dxrvoxelizer_b200.sharding import proportional_slabs
        import oracle
        # The GPUs of one box do not all see the same host link: measured on this pool's 8-GPU boxes, GPU 0 reads back at
        # 53 GB/s while GPUs 1-3 get 13-17 GB/s and GPUs 4-7 ~25 GB/s when all eight copy at once.  Untimed calibration
        # under the real pattern: equal slabs, every rank voxelizes, all ranks start their read-back together, best of 5;
        # the slabs are then cut in proportion to the measured rates.
        from dxrvoxelizer_b200.sharding import slab_range
        c0, c1 = slab_range(rank, world, N)
        cal_bytes = (c1 - c0) * N * P * 4
        with numa_local(rig.local):
            cal_h = torch.empty(cal_bytes, dtype=torch.uint8).pin_memory()
            cal_h.zero_()
        build()
        best = 1e30
        for _ in range(5):
            vox.voxelize(N, d.MODE_PARITY, c0, c1)
            vox.synchronize()
            rig.barrier()
            tcal = time.perf_counter()
            vox.fetch_into(cal_h.data_ptr(), cal_bytes)
            best = min(best, time.perf_counter() - tcal)
        rates = [None] * world
        rig.dist.all_gather_object(rates, cal_bytes / best * 1e-9)
        dz0, dz1 = proportional_slabs(N, rates)[rank]     # dense copy: cut by the measured link rates
        ez0, ez1 = c0, c1                                 # default transport (compact blob + host expansion): equal cuts
        cal_rates = [round(float(r), 2) for r in rates]
        del cal_h
        e_ref = oracle.voxelize(host_mesh.vertices, host_mesh.indices, N, oracle.MODE_PARITY, z0=ez0, z1=ez1,
                                threads=max(1, host_threads() // world))["bits"]
        d_ref = oracle.voxelize(host_mesh.vertices, host_mesh.indices, N, oracle.MODE_PARITY, z0=dz0, z1=dz1,
                                threads=max(1, host_threads() // world))["bits"]
    else:
        ez0, ez1, e_ref, cal_rates = z0, z1, gate_ref, None
        dz0, dz1, d_ref = z0, z1, gate_ref
    e_bytes = (ez1 - ez0) * N * P * 4
    d_bytes = (dz1 - dz0) * N * P * 4
    with numa_local(rig.local):
        h_grid = torch.empty(max(e_bytes, d_bytes), dtype=torch.uint8).pin_memory()
        h_grid.zero_()

    def upload_and_build():
        if world > 1:
            # replicate the host mesh of rank 0: H2D on rank 0, NCCL broadcast, build from device memory
            with torch.cuda.stream(stream):
                if rank == 0:
                    d_vb.copy_(h_vb, non_blocking=True)
                    d_ib.copy_(h_ib, non_blocking=True)
                rig.dist.broadcast(d_vb, 0)
                rig.dist.broadcast(d_ib, 0)
            build()
        else:
            vox.build_bvh_host_ptr(h_vb.data_ptr(), nv, stride, h_ib.data_ptr(), ni)

    def step_e2e():
        # the public call with its default transport (DXRV_READ_BACK_AUTO: at this size the slab comes back as a
        # DXRV_FORMAT_SPARSE_BRICKS blob and is expanded into the dense host grid by the library's host threads)
        if world == 1:
            # upload + build + voxelize + read-back as the library's ONE call for a frame (dxrv_voxelize_mesh_to_host): the host
            # threads start on the grid buffer before the upload and the build instead of after them
            vox.voxelize_mesh_to_host(h_vb.data_ptr(), nv, stride, h_ib.data_ptr(), ni, N, d.MODE_PARITY, ez0, ez1, h_grid.data_ptr(), e_bytes, chunks=8)
            return
        upload_and_build()
        vox.voxelize_to_host(N, d.MODE_PARITY, ez0, ez1, h_grid.data_ptr(), e_bytes, chunks=8)

    def step_e2e_dense():
        # the same call, grid copied densely: pipelined in 8 z sub-slabs (D2H of chunk k beside the fill of chunk k + 1)
        upload_and_build()
        vox.voxelize_to_host(N, d.MODE_PARITY, dz0, dz1, h_grid.data_ptr(), d_bytes, chunks=8)

    from dxrvoxelizer_b200 import _lib as L
    vox.set_read_back(L.READ_BACK_DENSE)
    for _ in range(3):
        step_e2e_dense()
    dense_mism = popcount(h_grid.numpy()[:d_bytes].view(np.uint32) ^ d_ref.reshape(-1))
    rig.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e_dense()
    rig.barrier()
    dense_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    vox.set_read_back(L.READ_BACK_AUTO)
    h_grid.fill_(0xA5)            # the expansion must produce every byte, zeros included
    for _ in range(warm):         # (the host threads of the pool were asleep during the dense-copy arm: the first steps after
        step_e2e()                #  that run at 1.2 ms while the cores wake up and clock up, the steady state is 0.77 ms)
    e2e_d2h = vox.info(L.INFO_LAST_D2H_BYTES)
    e2e_mism = popcount(h_grid.numpy()[:e_bytes].view(np.uint32) ^ e_ref.reshape(-1)) + dense_mism
    e2e_mism_total, e2e_d2h_total = rig.reduce_sum([e2e_mism, e2e_d2h])
    if e2e_mism_total != 0:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "error": "end-to-end host grid differs from the CPU oracle", "n_gpus": world,
                              "mismatched_voxels": int(e2e_mism_total)}))
        rig.close()
        raise SystemExit(3)
    rig.barrier()
    t0 = time.perf_counter()
    dbg = []
    for _ in range(args.steps):
        ta = time.perf_counter()
        step_e2e()
        dbg.append((time.perf_counter() - ta) * 1e3)
    rig.barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    if os.environ.get("DXRV_BENCH_DEBUG"):
        sys.stderr.write("e2e steps ms: %s\n" % " ".join("%.2f" % v for v in dbg))
    # ---- the same end-to-end step with the compact read-back format (DXRV_FORMAT_SPARSE_BRICKS, lossless): side number
    with numa_local(rig.local):
        h_sparse = torch.empty(e_bytes + e_bytes // 64 + (1 << 16), dtype=torch.uint8).pin_memory()

    def step_e2e_sparse():
        upload_and_build()
        vox.voxelize(N, d.MODE_PARITY, ez0, ez1)
        return vox.fetch_sparse_into(h_sparse.data_ptr(), h_sparse.numel())

    for _ in range(3):
        sparse_bytes = step_e2e_sparse()
    sparse_mism = popcount(d.sparse_decode(h_sparse.numpy()[:sparse_bytes]) ^ e_ref)
    rig.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e_sparse()
    rig.barrier()
    sparse_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    clocks = sampler.stop(t_wall0, time.time()) if rank == 0 else None
    # where the end-to-end time goes (separate short loop with a synchronisation between compute and read-back)
    reps = min(args.steps, 20)
    up_ms = down_ms = 0.0
    for _ in range(reps):
        ta = time.perf_counter()
        if world == 1:
            vox.build_bvh_host_ptr(h_vb.data_ptr(), nv, stride, h_ib.data_ptr(), ni)
        else:
            build()
        vox.voxelize(N, d.MODE_PARITY, ez0, ez1)
        vox.synchronize()
        tb = time.perf_counter()
        vox.fetch_into(h_grid.data_ptr(), e_bytes)
        tc = time.perf_counter()
        up_ms += (tb - ta) * 1e3 / reps
        down_ms += (tc - tb) * 1e3 / reps
    d2h_gbs = e_bytes / (down_ms * 1e-3) * 1e-9 if down_ms > 0 else 0.0

    # ---- weak-scaling side number: the grid grown so that every rank still fills ~1024^3 voxels -----------------
    weak = None
    if world in WEAK_GRID:
        Nw = WEAK_GRID[world]
        a, b = balanced_slabs(host_mesh, Nw, world)[rank]
        for _ in range(3):
            build(); vox.voxelize(Nw, d.MODE_PARITY, a, b)
        rig.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            build(); vox.voxelize(Nw, d.MODE_PARITY, a, b)
        e1.record(stream)
        rig.barrier()
        weak = (Nw, e0.elapsed_time(e1) / args.steps)

    # ---- max over ranks ---------------------------------------------------------------------------------------
    fill_ms_rank0 = fill_ms   # the roofline of the kernel is a per-GPU figure: rank 0's launches against rank 0's bytes
    step_ms, build_ms, trace_ms, e2e_ms, walk_ms, fill_ms, shader_ms, weak_ms, sparse_ms, dense_ms, full_ms = rig.reduce_max(
        [step_ms, build_ms, trace_ms, e2e_ms, walk_ms, fill_ms, shader_ms, weak[1] if weak else 0.0, sparse_ms, dense_ms, full_ms])
    sparse_bytes_total, sparse_mism_total, full_mism_total = rig.reduce_sum([sparse_bytes, sparse_mism, full_mism])
    per_rank = None
    if world > 1:
        gathered = [None] * world
        rig.dist.all_gather_object(gathered, {"rank": rank, "slab": [z0, z1], "e2e_slab": [ez0, ez1], "dense_copy_slab": [dz0, dz1], "d2h_gbs": round(d2h_gbs, 2), "d2h_ms": round(down_ms, 4),
                                              "fill_kernel_ms": round(fill_ms_rank0, 5)})
        per_rank = gathered

    if rank == 0:
        total_voxels = float(N) ** 3
        peak, peak_src = measured_peak()
        # algorithmic bytes of one k_trace_fill_columns launch on one GPU: the slab of the bit
        # grid written once + every scene-space triangle (48 B) read once.
        alg_bytes = slab_bytes + 48 * T
        achieved = alg_bytes / (fill_ms_rank0 * 1e-3) * 1e-9
        out = {
            "metric": METRIC, "value": total_voxels / (step_ms * 1e-3) * 1e-9, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "reference asset dragon.obj (Stanford dragon, 100k triangles); no synthetic substitution needed",
            "config": workload_config(N, world, T, nv),
            "mismatched_voxels": int(mism_total), "gate": {"parity_layers_checked": int(checked_total), "shader_layers_checked": int(schecked_total),
                                                            "shader_mismatched_voxels": int(smism_total),
                                                            "how": "XOR-popcount of every rank's fetched slab against the CPU oracle before timing"},
            "phases_ms": {"bvh_build": build_ms, "voxelize": trace_ms, "k_walk_columns": walk_ms, "k_trace_fill_columns": fill_ms,
                          "shader_1024": shader_ms, "step_full_lbvh_walk": full_ms},
            "full_lbvh": {"ms_per_step": full_ms, "gvoxels_per_s": total_voxels / (full_ms * 1e-3) * 1e-9, "mismatched_voxels": int(full_mism_total),
                          "what": "the same step with the whole LBVH rebuilt every step (bounds, Morton, three radix passes, leaves, Karras "
                                  "hierarchy, node boxes) and the candidates found by the warp-cooperative LBVH walk instead of the binning"},
            "shader": {"ms_per_1024_cubed_grid_incl_build_and_bins": shader_ms, "grays_per_s": total_voxels / (shader_ms * 1e-3) * 1e-9,
                       "what": "MODE_SHADER (DXRVoxelizer.hlsl radial closest hit), same z-slabs, LBVH + direction bins rebuilt every step"},
            "ms_per_1024_cubed_grid": step_ms,
            "e2e": {"value": total_voxels / (e2e_ms * 1e-3) * 1e-9, "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(nv * stride + ni * 4), "d2h_bytes_per_step": int(e2e_d2h_total),
                    "host_grid_bytes_per_step": int(N * N * P * 4),
                    "timing": "wall clock around synchronising C-ABI calls, max over ranks",
                    "transport": "DXRV_READ_BACK_AUTO = %s" % ("sparse bricks + one-pass host grid (slab >= 8 MiB, pool of >= 4 host threads)"
                                                               if host_pool_threads >= 4 and e_bytes >= (8 << 20) else
                                                               "dense copy (fewer than 4 host threads per rank, or a slab below 8 MiB)"),
                    "how": "dxrv_voxelize_mesh_to_host (N = 1; = dxrv_build_bvh from host arrays + dxrv_voxelize_to_host, which N > 1 calls "
                           "separately after the NCCL broadcast of the mesh) into a pinned host buffer that ends up holding the DENSE "
                           "128 MiB bit grid (filled with 0xA5 beforehand, checked against the oracle).  Default transport of the call: "
                           "the slab is encoded as DXRV_FORMAT_SPARSE_BRICKS on the GPU, `d2h_bytes_per_step` cross PCIe, and the "
                           "library's host threads (%d per rank) write the dense layout in ONE pass of streaming stores: they start "
                           "zeroing the buffer from the outside of the slab inwards while the GPU is still computing and, once the blob "
                           "is on the host, write the remaining brick layers with their final contents; the floor is the host's "
                           "memory write bandwidth (128 MiB / 0.69 ms on these 16 cores), not the link.  "
                           "`dense_copy` = the same call with DXRV_READ_BACK_DENSE (8 z sub-slabs, D2H of chunk k beside the "
                           "fill of chunk k+1: the PCIe floor, `phases_ms.d2h_grid` measured unpipelined; N > 1: its slabs are cut in "
                           "proportion to every rank's measured read-back rate)" % host_pool_threads,
                    "dense_copy": {"value": total_voxels / (dense_ms * 1e-3) * 1e-9, "unit": UNIT, "ms_per_step": dense_ms,
                                   "d2h_bytes_per_step": int(N * N * P * 4)},
                    "mismatched_voxels": int(e2e_mism_total), "readback_calibration_gbs": cal_rates,
                    "sparse_bricks": {"value": total_voxels / (sparse_ms * 1e-3) * 1e-9, "unit": UNIT, "ms_per_step": sparse_ms,
                                      "d2h_bytes_per_step": int(sparse_bytes_total), "mismatched_voxels_after_decode": int(sparse_mism_total),
                                      "what": "same step, read-back as DXRV_FORMAT_SPARSE_BRICKS (lossless: header + 2-bit brick states + the "
                                              "mixed 32x4x4 bricks; dxrv_fetch_grid_sparse / dxrv_sparse_decode) -- a side number for consumers "
                                              "that can take the compact form and skip the host expansion"},
                    "phases_ms": {"h2d_mesh_build_voxelize": up_ms, "d2h_grid": down_ms}, "d2h_gbs_rank0": d2h_gbs,
                    "per_rank": per_rank},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "k_trace_fill_columns", "bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(world),
                         "algorithmic_bytes_per_launch": int(alg_bytes), "kernel_ms": fill_ms_rank0, "peak_source": peak_src,
                         "frac_of_nominal_8000_GBps": achieved / 8000.0,
                         "timing": "cudaEventRecord on the launching stream around the kernel, mean of %d launches" % prof_steps},
            "cpu_baseline": cpu_baseline(host_mesh, N) if world == 1 else None,
            "clocks": clocks,
            "crossings": int(crossings),
        }
        if weak:
            out["weak_scaling"] = {"grid": weak[0], "ms_per_step": weak_ms, "gvoxels_per_s": float(weak[0]) ** 3 / (weak_ms * 1e-3) * 1e-9,
                                   "scaling": "weak", "note": "side number: ~1024^3 voxels per GPU; not the BASELINE.json config"}
        print(json.dumps(out))
    rig.close()


class numa_local:
    """Allocate pinned host buffers on the NUMA node next to GPU `index`: the thread is moved onto the CPUs NVML
    reports for that GPU while the pages are allocated and touched, then gets its original affinity back (the CPU
    baseline must see every core).  Best effort: without NVML (or on a single-node box) it does nothing."""
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


def ncu_traffic(world):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    ncu --set full capture of THIS configuration (profiles/roofline_traffic.json: dragon 1024^3, one GPU);
    null at N > 1, where the slab differs and no capture exists."""
    if world != 1:
        return None
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get("k_trace_fill_columns_dram_bytes_per_launch")
    except Exception:
        return None


def cpu_baseline(mesh, N):
    """The CPU oracle (a port of the path) on the box's host cores: the whole grid, best of 3."""
    import oracle
    threads = host_threads()
    oracle.voxelize(mesh.vertices, mesh.indices, N, oracle.MODE_PARITY, threads=threads)
    best = 1e30
    t_all = time.perf_counter()
    for _ in range(3):
        t = time.perf_counter()
        oracle.voxelize(mesh.vertices, mesh.indices, N, oracle.MODE_PARITY, threads=threads)
        best = min(best, time.perf_counter() - t)
        if time.perf_counter() - t_all > 25:
            break
    t = time.perf_counter()
    oracle.voxelize(mesh.vertices, mesh.indices, N, oracle.MODE_SHADER, z0=N // 2 - 2, z1=N // 2 + 2, threads=threads)
    shader_rate = 4.0 * N * N / (time.perf_counter() - t) * 1e-9
    return {"value": float(N) ** 3 / best * 1e-9, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "the whole %d^3 dragon grid, MODE_PARITY (the column-parity algorithm the GPU headline uses), own acceleration build "
                      "included, best of 3" % N,
            "shader_gvoxels_per_s": shader_rate,
            "shader_sample": "MODE_SHADER (the reference's radial closest-hit algorithm), 4 central layers of the same grid"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=["c3", "c4", "c5"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "c3":
        run_c3(args)
    else:
        import bench_configs
        (bench_configs.run_c4 if args.config == "c4" else bench_configs.run_c5)(args, Rig, ClockSampler, measured_peak, host_threads, popcount)


if __name__ == "__main__":
    main()
