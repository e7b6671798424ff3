#!/usr/bin/env python3
"""Does a coarser Morton key (fewer radix passes) hurt the consumers when the mesh arrives in RANDOM triangle order?
(development aid) icosphere(9) = 5.2 M triangles, triangles shuffled; build + MODE_PARITY voxelize at 512^3 and 1024^3."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dxrvoxelizer_b200 as d
from dxrvoxelizer_b200 import meshes
def timeit(fn, stream, iters=10, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters): fn()
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
s = torch.cuda.Stream(); vox = d.Voxelizer(0); vox.set_stream(s.cuda_stream)
for name, m in (("ico9", meshes.icosphere(9, normals=False)), ("dragon", d.load_obj(d.asset_path("dragon.obj")))):
    rng = np.random.default_rng(1)
    tri = m.indices.reshape(-1, 3)[rng.permutation(m.num_triangles)]
    vb = torch.from_numpy(m.vertex_bytes).cuda(); ib = torch.from_numpy(np.ascontiguousarray(tri).reshape(-1).view(np.int32)).cuda()
    build = lambda: vox.build_bvh_device(vb.data_ptr(), m.num_vertices, m.stride, ib.data_ptr(), ib.numel())
    for N in (512, 1024):
        tb = timeit(build, s); tv = timeit(lambda: vox.voxelize(N, d.MODE_PARITY), s)
        print("%s shuffled, passes=%s N=%d: build %.3f ms voxelize %.3f ms inside=%d" % (name, os.environ.get("DXRV_KEY_PASSES", "default"), N, tb, tv, vox.count_inside()), flush=True)
