#!/usr/bin/env python3
"""MODE_SHADER timing on the GPU box (development aid; bench.py is the contract): direction bins vs LBVH walk."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import dxrvoxelizer_b200 as d
from dxrvoxelizer_b200 import _lib as L


def timeit(fn, stream, iters=5, warm=1):
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
    sizes = [int(a) for a in sys.argv[1:]] or [256, 512, 1024]
    s = torch.cuda.Stream()
    vox = d.Voxelizer(0)
    vox.set_stream(s.cuda_stream)
    for name in ("dragon.obj", "TuringBowl.obj", "bunny.obj"):
        m = d.load_obj(d.asset_path(name))
        vb = torch.from_numpy(m.vertex_bytes).cuda(); ib = torch.from_numpy(m.indices.view(np.int32)).cuda()
        build = lambda: vox.build_bvh_device(vb.data_ptr(), m.num_vertices, m.stride, ib.data_ptr(), m.indices.size)
        for path in ("bins", "bvh"):
            os.environ["DXRV_SHADER_PATH"] = path
            for N in sizes:
                if path == "bvh" and N > 512:
                    continue
                t_first = timeit(lambda: (build(), vox.voxelize(N, d.MODE_SHADER)), s, iters=3)
                t_again = timeit(lambda: vox.voxelize(N, d.MODE_SHADER), s, iters=3)
                st = vox.debug_read(L.DBG_BINS_STATE, np.uint32, 4)
                print("%s %s N=%d: build+bins+trace %.3f ms, trace only %.3f ms (%.2f Grays/s) inside=%d bins: entries=%d overflow=%d near=%d R=%d" %
                      (name, path, N, t_first, t_again, N ** 3 / t_again * 1e-6, vox.count_inside(), st[0], st[1], st[2], st[3]), flush=True)
    vox.close()


if __name__ == "__main__":
    main()
