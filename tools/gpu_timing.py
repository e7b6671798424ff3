#!/usr/bin/env python3
"""Quick phase timing on the GPU box (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import dxrvoxelizer_b200 as d
from dxrvoxelizer_b200 import meshes


def timeit(fn, stream, iters=20, warm=3):
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
    torch.cuda.init()
    s = torch.cuda.Stream()
    vox = d.Voxelizer(0)
    vox.set_stream(s.cuda_stream)
    cases = [("dragon", d.load_obj(d.asset_path("dragon.obj"))), ("bowl", d.load_obj(d.asset_path("TuringBowl.obj")))]
    if len(sys.argv) > 1:
        cases.append(("ico8", meshes.icosphere(8)))
    for name, m in cases:
        vb = torch.from_numpy(m.vertex_bytes).cuda()
        ib = torch.from_numpy(m.indices.view(np.int32)).cuda()
        build = lambda: vox.build_bvh_device(vb.data_ptr(), m.num_vertices, m.stride, ib.data_ptr(), m.indices.size)
        t_build = timeit(build, s)
        print("%s T=%d build %.1f us" % (name, m.num_triangles, t_build * 1e3))
        for N in (256, 512, 1024, 2048):
            t = timeit(lambda: vox.voxelize(N, d.MODE_PARITY), s, iters=10)
            print("  parity N=%d: %.1f us  (%.1f Gvox/s, grid %.0f GB/s) crossings=%d" % (N, t * 1e3, N ** 3 / t * 1e-6, N ** 3 / 8 / t * 1e-6, vox.info(3)))
        for N in (64, 128, 256):
            t = timeit(lambda: vox.voxelize(N, d.MODE_SHADER), s, iters=3, warm=1)
            print("  shader N=%d: %.1f us  (%.2f Grays/s)" % (N, t * 1e3, N ** 3 / t * 1e-6))
    vox.close()


if __name__ == "__main__":
    main()
