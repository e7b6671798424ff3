#!/usr/bin/env python3
"""BASELINE.json configs 2, 4 and 5 on one GPU (development / reporting aid; bench.py is the contract).
  C2  TuringBowl at 512^3 (and 256^3), both modes, vs the OpenMP oracle
  C4  synthetic meshes (icosphere k=8, k=9, torus knot 4096x2048) at 512^3, MODE_PARITY: build-dominated
  C5  batch of 256 distinct icosphere(5) meshes at 256^3 on several streams (meshes/s)
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import dxrvoxelizer_b200 as d
from dxrvoxelizer_b200 import meshes, _lib as L


def timeit(fn, stream, iters=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    which = sys.argv[1:] or ["c2", "c4", "c5"]
    out = {}
    s = torch.cuda.Stream()
    vox = d.Voxelizer(0)
    vox.set_stream(s.cuda_stream)
    if "c2" in which:
        import oracle
        m = d.load_obj(d.asset_path("TuringBowl.obj"))
        vb = torch.from_numpy(m.vertex_bytes).cuda(); ib = torch.from_numpy(m.indices.view(np.int32)).cuda()
        build = lambda: vox.build_bvh_device(vb.data_ptr(), m.num_vertices, m.stride, ib.data_ptr(), m.indices.size)
        for N in (256, 512):
            for mode, name in ((d.MODE_PARITY, "parity"), (d.MODE_SHADER, "shader")):
                t = timeit(lambda: (build(), vox.voxelize(N, mode)), s, iters=5)
                t0 = time.perf_counter()
                layers = N if mode == d.MODE_PARITY else 16
                oracle.voxelize(m.vertices, m.indices, N, mode, z0=(N - layers) // 2, z1=(N - layers) // 2 + layers)
                tc = (time.perf_counter() - t0) * N / layers
                out["c2_bowl_%d_%s" % (N, name)] = {"gpu_ms_incl_build": t, "gvox_s": N ** 3 / t * 1e-6,
                                                    "oracle_ms_est": tc * 1e3, "oracle_threads": oracle.max_threads()}
                print("C2 bowl N=%d %s: GPU %.3f ms (%.1f Gvox/s)  oracle ~%.0f ms" % (N, name, t, N ** 3 / t * 1e-6, tc * 1e3), flush=True)
    if "c4" in which:
        for name, gen in (("ico8_1.3M", lambda: meshes.icosphere(8, normals=False)),
                          ("ico9_5.2M", lambda: meshes.icosphere(9, normals=False)),
                          ("knot_16.8M", lambda: meshes.torus_knot(4096, 2048, normals=False))):
            t0 = time.time(); m = gen(); tg = time.time() - t0
            vb = torch.from_numpy(m.vertex_bytes).cuda(); ib = torch.from_numpy(m.indices.view(np.int32)).cuda()
            T = m.num_triangles
            build = lambda: vox.build_bvh_device(vb.data_ptr(), m.num_vertices, m.stride, ib.data_ptr(), m.indices.size)
            tb = timeit(build, s, iters=5)
            tv = timeit(lambda: vox.voxelize(512, d.MODE_PARITY), s, iters=5)
            vox.synchronize()
            out["c4_" + name] = {"triangles": T, "build_ms": tb, "voxelize_512_ms": tv, "mtris_per_s": T / tb * 1e-3,
                                 "build_GBps_at_228B_per_tri": 228.0 * T / tb * 1e-6, "inside": vox.count_inside(), "crossings": vox.info(L.INFO_CROSSINGS)}
            print("C4 %s T=%d (gen %.1fs): build %.3f ms (%.0f Mtri/s, %.0f GB/s @228B/tri)  voxelize 512^3 %.3f ms  inside=%d" %
                  (name, T, tg, tb, T / tb * 1e-3, 228.0 * T / tb * 1e-6, tv, out["c4_" + name]["inside"]), flush=True)
            del vb, ib, m
    if "c5" in which:
        n_mesh, n_streams, N = 256, 4, 256
        ms = [meshes.icosphere(5, seed=i, rotate=True, normals=False) for i in range(n_mesh)]
        streams = [torch.cuda.Stream() for _ in range(n_streams)]
        ctxs = [d.Voxelizer(0) for _ in range(n_streams)]
        for c, st in zip(ctxs, streams):
            c.set_stream(st.cuda_stream)
        dev = [(torch.from_numpy(m.vertex_bytes).cuda(), torch.from_numpy(m.indices.view(np.int32)).cuda()) for m in ms]
        torch.cuda.synchronize()

        def run_all():
            for i, m in enumerate(ms):
                c = ctxs[i % n_streams]
                c.build_bvh_device(dev[i][0].data_ptr(), m.num_vertices, m.stride, dev[i][1].data_ptr(), m.indices.size)
                c.voxelize(N, d.MODE_PARITY)
        run_all(); torch.cuda.synchronize()
        t0 = time.perf_counter(); run_all(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        out["c5_batch256"] = {"meshes_per_s": n_mesh / dt, "gvox_s": n_mesh * N ** 3 / dt * 1e-9, "streams": n_streams, "ms_total": dt * 1e3}
        print("C5 256 x icosphere(5) at 256^3 on %d streams: %.1f ms total, %.0f meshes/s, %.1f Gvox/s" % (n_streams, dt * 1e3, n_mesh / dt, n_mesh * N ** 3 / dt * 1e-9), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/configs.json", "w"), indent=1)


if __name__ == "__main__":
    main()
