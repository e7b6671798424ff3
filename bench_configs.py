"""bench.py --config c4 / c5: BASELINE.json configs[3] (synthetic watertight meshes of millions of triangles at 512^3,
the BVH-build-dominated regime) and configs[4] (a batch of 256 distinct meshes at 256^3, mesh-parallel, the streaming
case).  Same JSON contract as the default config (bench.py), same gate: every grid reported on is first XORed against
the CPU oracle.  Only imported by bench.py."""
import json
import os
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))


# ------------------------------------------------------------------------------------------------------------------
def run_c4(args, Rig, ClockSampler, measured_peak, host_threads, popcount):
    """Torus knot 4096 x 2048 = 16 777 216 triangles, N = 512, MODE_PARITY.
    A step = LBVH build + voxelize of the rank's z-slab.  `value` = triangles built per second (every rank builds
    the whole tree: the build is replicated, so this is a per-GPU rate, not an aggregate)."""
    import dxrvoxelizer_b200 as d
    from dxrvoxelizer_b200 import _lib as L, meshes
    from dxrvoxelizer_b200.sharding import balanced_slabs
    import oracle
    rig = Rig(args)
    torch, vox, stream, rank, world = rig.torch, rig.vox, rig.stream, rig.rank, rig.world
    N = 512
    nu, nv_ = (4096, 2048) if not os.environ.get("DXRV_C4_SMALL") else (1024, 512)
    t0 = time.time()
    mesh = meshes.torus_knot(nu, nv_, normals=False) if rank == 0 else None
    gen_s = time.time() - t0
    host_mesh, (h_vb, h_ib, d_vb, d_ib, nv, stride, ni) = rig.replicate_mesh(mesh)
    T = ni // 3
    z0, z1 = balanced_slabs(host_mesh, N, world)[rank]

    def build():
        vox.build_bvh_device(d_vb.data_ptr(), nv, stride, d_ib.data_ptr(), ni)

    # ---- gate: a few z-slabs of this rank's part against the oracle (it needs seconds per slab at 16.8 M triangles)
    build()
    vox.voxelize(N, d.MODE_PARITY, z0, z1)
    got = vox.fetch_bits()
    threads = max(1, host_threads() // world)
    mism = checked = 0
    span = z1 - z0
    for a in sorted({z0, z0 + span // 2 - 2, z1 - 4}):
        a = max(z0, min(a, z1 - 4)) if span >= 4 else z0
        b = min(a + 4, z1)
        ref = oracle.voxelize(host_mesh.vertices, host_mesh.indices, N, oracle.MODE_PARITY, z0=a, z1=b, threads=threads)["bits"]
        mism += popcount(got[a - z0:b - z0] ^ ref)
        checked += b - a
    mism_total, checked_total = rig.reduce_sum([mism, checked])
    if mism_total != 0:
        if rank == 0:
            print(json.dumps({"metric": "gtris_per_s_lbvh_build", "error": "GPU grid differs from the CPU oracle", "mismatched_voxels": int(mism_total)}))
        rig.close()
        raise SystemExit(3)

    sampler = ClockSampler(rig.local)
    if rank == 0:
        sampler.start()
    steps = max(3, min(args.steps, 20))
    warm = max(3, min(args.warmup, 5))
    for _ in range(warm):
        build(); vox.voxelize(N, d.MODE_PARITY, z0, z1)
    vox.synchronize()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    launches0 = vox.info(L.INFO_KERNEL_LAUNCHES)
    rig.barrier()
    t_wall0 = time.time()
    for i in range(steps):
        ev[i][0].record(stream)
        build()
        ev[i][1].record(stream)
        vox.voxelize(N, d.MODE_PARITY, z0, z1)
        ev[i][2].record(stream)
    rig.barrier()
    launches = vox.info(L.INFO_KERNEL_LAUNCHES) - launches0
    build_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / steps
    vox_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / steps
    step_ms = sum(e[0].elapsed_time(e[2]) for e in ev) / steps
    # the sort passes alone (events recorded by the library around them; inputs = 402 MB of keys + values, > L2)
    vox.set_profiling(True)
    sort_ns = bld_ns = 0
    for _ in range(min(steps, 10)):
        build()
        sort_ns += vox.info(L.INFO_LAST_SORT_NS)
        bld_ns += vox.info(L.INFO_LAST_BUILD_NS)
    vox.set_profiling(False)
    sort_ms = sort_ns / min(steps, 10) * 1e-6
    # end to end: host mesh -> device, build, voxelize, slab back
    P = (N + 31) // 32
    slab_bytes = (z1 - z0) * N * P * 4
    h_grid = torch.empty(slab_bytes, dtype=torch.uint8).pin_memory()

    def step_e2e():
        if world > 1:
            with torch.cuda.stream(stream):
                if rank == 0:
                    d_vb.copy_(h_vb, non_blocking=True); d_ib.copy_(h_ib, non_blocking=True)
                rig.dist.broadcast(d_vb, 0); rig.dist.broadcast(d_ib, 0)
            build()
        else:
            vox.build_bvh_host_ptr(h_vb.data_ptr(), nv, stride, h_ib.data_ptr(), ni)
        vox.voxelize(N, d.MODE_PARITY, z0, z1)
        vox.fetch_into(h_grid.data_ptr(), slab_bytes)
    for _ in range(2):
        step_e2e()
    rig.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step_e2e()
    rig.barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / steps
    clocks = sampler.stop(t_wall0, time.time()) if rank == 0 else None
    step_ms, build_ms, vox_ms, sort_ms, e2e_ms = rig.reduce_max([step_ms, build_ms, vox_ms, sort_ms, e2e_ms])
    if rank == 0:
        peak, peak_src = measured_peak()
        passes = 2   # 16-bit keys when no consumer traverses the hierarchy (scatter path); 4 with one
        sort_bytes = 16.0 * T * passes
        build_bytes = 228.0 * T
        # as run (no hierarchy: the scatter path does not traverse one): indices 12T and vertices 12V read by k_morton and
        # again by k_leaf_setup, keys + values 8T written, the radix passes, sorted indices 4T read, records 48T written
        as_run_bytes = 2.0 * (12.0 * T + 12.0 * nv) + 8.0 * T + sort_bytes + 4.0 * T + 48.0 * T
        # CPU baseline: the oracle's own acceleration build + 16 central layers (bounded sample)
        tc = time.perf_counter()
        oracle.voxelize(host_mesh.vertices, host_mesh.indices, N, oracle.MODE_PARITY, z0=N // 2 - 8, z1=N // 2 + 8, threads=host_threads())
        cpu_s = time.perf_counter() - tc
        out = {
            "metric": "gtris_per_s_lbvh_build", "value": T / (build_ms * 1e-3) * 1e-9, "unit": "Gtri/s", "n_gpus": world,
            "steps": steps, "warmup": warm, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "replicas (the build is replicated; z-slabs split only the voxelize)",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic: torus knot %dx%d, seed 1234 (generated in %.1f s)" % (nu, nv_, gen_s),
            "config": {"workload": "torus knot %d triangles at %d^3 MODE_PARITY, LBVH rebuilt every step" % (T, N), "grid": N, "triangles": T,
                       "vertices": nv, "parallelism": "zslab%d" % world, "l2": "inputs (402 MB of mesh, 268 MB of keys) exceed the 126 MB L2"},
            "mismatched_voxels": int(mism_total), "gate": {"layers_checked": int(checked_total)},
            "phases_ms": {"bvh_build": build_ms, "onesweep_sort": sort_ms, "voxelize": vox_ms},
            "gvoxels_per_s_incl_bvh_build": float(N) ** 3 / (step_ms * 1e-3) * 1e-9,
            "e2e": {"value": T / (e2e_ms * 1e-3) * 1e-9, "unit": "Gtri/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(nv * stride + ni * 4),
                    "d2h_bytes_per_step": int(N * N * P * 4), "timing": "wall clock around synchronising C-ABI calls, max over ranks"},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "k_onesweep_pass_big x %d" % passes, "bound": "hbm", "achieved": sort_bytes / (sort_ms * 1e-3) * 1e-9, "peak": peak, "unit": "GB/s",
                         "frac": sort_bytes / (sort_ms * 1e-3) * 1e-9 / peak, "traffic": None, "algorithmic_bytes_per_launch": int(16 * T),
                         "kernel_ms": sort_ms / passes, "peak_source": peak_src,
                         "build": {"algorithmic_bytes": int(as_run_bytes), "achieved": as_run_bytes / (build_ms * 1e-3) * 1e-9,
                                   "frac": as_run_bytes / (build_ms * 1e-3) * 1e-9 / peak,
                                   "what": "unique bytes of the build AS RUN (%.0f B per triangle: no hierarchy is built for a consumer that does not "
                                           "traverse one); SURVEY.md section 8d's 228 B per triangle includes 128 B of hierarchy traffic and would read "
                                           "%.2f" % (as_run_bytes / T, build_bytes / (build_ms * 1e-3) * 1e-9 / peak)}},
            "cpu_baseline": {"value": 16.0 * N * N / cpu_s * 1e-9, "unit": "Gvoxel/s", "cores": host_threads(), "kind": "port",
                             "sample": "oracle MODE_PARITY, own acceleration build + 16 central layers of the %d^3 grid: %.2f s" % (N, cpu_s)},
            "clocks": clocks,
        }
        print(json.dumps(out))
    rig.close()


# ------------------------------------------------------------------------------------------------------------------
def write_obj(path, mesh):
    """OBJ text as an exporter writes it: `v x y z` and `f a b c` (1-based).  The loader flips z and reverses the
    index array (XUSGObjLoader.cpp:198,227), which is part of what is timed."""
    pos = mesh.vertices[:, :3]
    tri = mesh.indices.reshape(-1, 3).astype(np.int64) + 1
    with open(path, "w") as f:
        f.write("# synthetic icosphere\n")
        f.write("".join("v %.6f %.6f %.6f\n" % (p[0], p[1], p[2]) for p in pos))
        f.write("".join("f %d %d %d\n" % (t[0], t[1], t[2]) for t in tri))


def run_c5(args, Rig, ClockSampler, measured_peak, host_threads, popcount):
    """256 distinct icosphere(5) meshes (20 480 triangles each, per-mesh seeded displacement + rotation) arriving as
    OBJ TEXT, voxelized at 256^3 MODE_PARITY; mesh-parallel: rank r takes meshes r, r + world, ...; 4 contexts
    (streams) per GPU, loader threads feed them.  A step = the rank's whole share: parse, H2D, LBVH build, voxelize,
    D2H of every 2 MiB grid.  `value` = meshes per second over all ranks."""
    import ctypes
    from concurrent.futures import ThreadPoolExecutor
    import dxrvoxelizer_b200 as d
    from dxrvoxelizer_b200 import _lib as L, meshes
    import oracle
    rig = Rig(args)
    torch, rank, world = rig.torch, rig.rank, rig.world
    N, n_mesh, n_streams = 256, int(os.environ.get("DXRV_C5_MESHES", "256")), int(os.environ.get("DXRV_C5_STREAMS", "4"))
    tmp = os.path.join(tempfile.gettempdir(), "dxrv_c5_%d" % os.getuid())
    os.makedirs(tmp, exist_ok=True)
    mine = list(range(rank, n_mesh, world))
    paths = {}
    text_bytes = 0
    for i in mine:                                   # input generation (untimed)
        p = os.path.join(tmp, "ico5_%03d.obj" % i)
        if not os.path.exists(p):
            write_obj(p + ".tmp%d" % rank, meshes.icosphere(5, seed=i, rotate=True, normals=False))
            os.replace(p + ".tmp%d" % rank, p)
        paths[i] = p
        text_bytes += os.path.getsize(p)
    ctxs = [d.Voxelizer(rig.local) for _ in range(n_streams)]
    P = (N + 31) // 32
    grid_bytes = N * N * P * 4
    h_grids = torch.empty((len(mine), grid_bytes), dtype=torch.uint8).pin_memory()
    os.environ["DXRV_OBJ_THREADS"] = "1"          # one thread per file, the pool parallelises over the files
    loaders = ThreadPoolExecutor(max_workers=max(2, host_threads() // world))

    drivers = ThreadPoolExecutor(max_workers=n_streams)

    def run_share(fetch):
        """parse (thread pool) -> build + voxelize on stream k % n_streams -> optional D2H; returns the loaded meshes.
        Every stream (= context) is driven by its own host thread: a context's calls -- upload (waits for the copy), build,
        voxelize, read-back of its previous grid -- take ~0.25 ms of host time per mesh, and issued from ONE thread for all
        four streams they were what bounded the step (63 ms for 256 meshes, with 16 parser threads idle most of the time).
        Distinct contexts are independent (include/dxrv.h); ctypes releases the GIL during the calls."""
        futures = [loaders.submit(d.load_obj, paths[i]) for i in mine]
        loaded = [None] * len(futures)

        def drive(s):
            c = ctxs[s]
            prev = None                                    # mesh whose grid is still on this stream
            for k in range(s, len(futures), n_streams):
                m = futures[k].result()
                if fetch and prev is not None:             # read back the previous grid of THIS stream only: the other
                    c.fetch_into(h_grids[prev].data_ptr(), grid_bytes)   # streams keep the GPU busy meanwhile
                c.build_bvh(m)
                c.voxelize(N, d.MODE_PARITY)
                prev = k
                loaded[k] = m
            if fetch and prev is not None:
                c.fetch_into(h_grids[prev].data_ptr(), grid_bytes)
            c.synchronize()

        for f in [drivers.submit(drive, s) for s in range(n_streams)]:
            f.result()
        return loaded

    # The product's own pipeline (dxrv_voxelize_obj_batch, csrc/batch.cpp): the same stages -- loader threads, one driver
    # thread per context, read-back of a context's previous grid before its next build -- without an interpreter between
    # them.  This is the timed path; run_share above (Python threads over the single-mesh calls) stays as a side number.
    my_paths = [paths[i] for i in mine]
    n_loaders = max(2, host_threads() // world)

    def run_batch():
        d.voxelize_obj_batch(ctxs, my_paths, N, d.MODE_PARITY, out_ptr=h_grids.data_ptr(), loader_threads=n_loaders)

    # ---- gate: EVERY grid of this rank against the oracle (grids produced by the timed path) ------------------------
    loaded = run_share(True)
    h_grids.zero_()
    run_batch()
    threads = max(1, host_threads() // world)
    mism = 0
    for k, m in enumerate(loaded):
        ref = oracle.voxelize(m.vertices, m.indices, N, oracle.MODE_PARITY, threads=threads)["bits"]
        mism += popcount(h_grids[k].numpy().view(np.uint32).reshape(N, N, P) ^ ref)
    mism_total, = rig.reduce_sum([mism])
    if mism_total != 0:
        if rank == 0:
            print(json.dumps({"metric": "meshes_per_s", "error": "GPU grid differs from the CPU oracle", "mismatched_voxels": int(mism_total)}))
        rig.close()
        raise SystemExit(3)

    sampler = ClockSampler(rig.local)
    if rank == 0:
        sampler.start()
    steps = max(3, min(args.steps, 10))
    for _ in range(2):
        run_share(True)
    rig.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        run_share(True)
    rig.barrier()
    py_ms = (time.perf_counter() - t0) * 1e3 / steps
    warm = max(3, min(args.warmup, 20))     # the host cores clock up over the first few hundred milliseconds of parsing
    for _ in range(warm):
        run_batch()
    launches0 = sum(c.info(L.INFO_KERNEL_LAUNCHES) for c in ctxs)
    rig.barrier()
    t_wall0 = time.time()
    t0 = time.perf_counter()
    for _ in range(steps):
        run_batch()
    rig.barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / steps
    launches = sum(c.info(L.INFO_KERNEL_LAUNCHES) for c in ctxs) - launches0
    # device-resident variant: meshes already parsed and uploaded, no read-back
    dev = [(torch.from_numpy(m.vertex_bytes).cuda(), torch.from_numpy(m.indices.view(np.int32)).cuda()) for m in loaded]
    torch.cuda.synchronize()

    def run_resident():
        for k, m in enumerate(loaded):
            c = ctxs[k % n_streams]
            c.build_bvh_device(dev[k][0].data_ptr(), m.num_vertices, m.stride, dev[k][1].data_ptr(), m.indices.size)
            c.voxelize(N, d.MODE_PARITY)
        for c in ctxs:
            c.synchronize()
    for _ in range(2):
        run_resident()
    rig.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        run_resident()
    rig.barrier()
    res_ms = (time.perf_counter() - t0) * 1e3 / steps
    clocks = sampler.stop(t_wall0, time.time()) if rank == 0 else None
    # loader alone: the product's parser against the reference's own fscanf loader (oracle/_ref), same files
    os.environ["DXRV_OBJ_THREADS"] = "0"          # the loader on its own: all its threads on one file
    tl = time.perf_counter()
    for i in mine[:16]:
        d.load_obj(paths[i])
    fast_s = (time.perf_counter() - tl) / max(1, len(mine[:16]))
    os.environ["DXRV_OBJ_THREADS"] = "1"
    tl = time.perf_counter()
    for i in mine[:16]:
        d.load_obj(paths[i])
    fast1_s = (time.perf_counter() - tl) / max(1, len(mine[:16]))
    ref_s = None
    if rank == 0 and oracle.ref_loader_available():
        tl = time.perf_counter()
        for i in mine[:8]:
            oracle.ref_load_obj(paths[i])
        ref_s = (time.perf_counter() - tl) / max(1, len(mine[:8]))
    e2e_ms, res_ms, py_ms = rig.reduce_max([e2e_ms, res_ms, py_ms])
    if rank == 0:
        mb = text_bytes / len(mine) * 1e-6
        # CPU baseline: the whole path on the host for 8 meshes (reference loader + oracle), all cores
        tc = time.perf_counter()
        for i in mine[:8]:
            if oracle.ref_loader_available():
                vb, ib, st, _ = oracle.ref_load_obj(paths[i])
                v = vb.view(np.float32).reshape(-1, st // 4)
            else:
                mm = d.load_obj(paths[i]); v, ib = mm.vertices, mm.indices
            oracle.voxelize(v, ib, N, oracle.MODE_PARITY, threads=host_threads())
        cpu_rate = len(mine[:8]) / (time.perf_counter() - tc)
        out = {
            "metric": "meshes_per_s", "value": n_mesh / (e2e_ms * 1e-3), "unit": "mesh/s", "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": e2e_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic: %d x icosphere(5) = 20480 triangles each, seed = mesh index, random rotation, as OBJ text" % n_mesh,
            "config": {"workload": "%d distinct meshes, OBJ text -> %d^3 MODE_PARITY grid, mesh-parallel over %d GPU(s), %d streams per GPU" % (n_mesh, N, world, n_streams),
                       "grid": N, "meshes": n_mesh, "triangles_per_mesh": 20480, "parallelism": "mesh-parallel x%d" % world,
                       "l2": "every step touches %d distinct meshes and grids (%.0f MB per GPU)" % (len(mine), len(mine) * (grid_bytes + 1.2e6) * 1e-6)},
            "mismatched_voxels": int(mism_total), "gate": {"grids_checked": n_mesh},
            "gvoxels_per_s": n_mesh * float(N) ** 3 / (e2e_ms * 1e-3) * 1e-9,
            "device_resident": {"meshes_per_s": n_mesh / (res_ms * 1e-3), "ms_per_step": res_ms,
                                "what": "meshes already parsed and resident in HBM, no read-back: LBVH build + voxelize only"},
            "e2e": {"value": n_mesh / (e2e_ms * 1e-3), "unit": "mesh/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(n_mesh * 20480 * 1.5 * 24 / 2 + n_mesh * 20480 * 12),
                    "d2h_bytes_per_step": int(n_mesh * grid_bytes), "timing": "wall clock, OBJ text on disk (page cache) -> grids in pinned host memory",
                    "call": "dxrv_voxelize_obj_batch: %d loader threads + %d contexts per rank, one C-ABI call per step" % (n_loaders, n_streams)},
            "python_driven": {"meshes_per_s": n_mesh / (py_ms * 1e-3), "ms_per_step": py_ms,
                              "what": "the same stages as Python threads over dxrv_obj_load / dxrv_build_bvh / dxrv_voxelize / dxrv_fetch_grid (the timed path of earlier lines)"},
            "gpu_launches": int(launches),
            "loader": {"parseObjFast_MBps": mb / fast_s, "ms_per_mesh": fast_s * 1e3, "obj_text_MB_per_mesh": mb,
                       "parseObjFast_single_thread_MBps": mb / fast1_s,
                       "reference_fscanf_loader_MBps": (mb / ref_s) if ref_s else None, "reference_ms_per_mesh": ref_s * 1e3 if ref_s else None,
                       "what": "dxrv_obj_load (product, byte-identical output) vs XUSGObjLoader.cpp compiled into oracle/_ref, same files"},
            "roofline": None,
            "cpu_baseline": {"value": cpu_rate, "unit": "mesh/s", "cores": host_threads(), "kind": "reference" if oracle.ref_loader_available() else "port",
                             "sample": "8 meshes: the reference's own ObjLoader (oracle/_ref) + the oracle's MODE_PARITY voxelization, all cores"},
            "clocks": clocks,
        }
        print(json.dumps(out))
    for c in ctxs:
        c.close()
    rig.close()
