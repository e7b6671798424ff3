#!/usr/bin/env python3
"""Device time of the build inside a build + voxelize step, fused (one cooperative kernel) against multi-kernel
(development aid; the host's launch rate hides anything below ~25 us per call, hence the voxelize behind it)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dxrvoxelizer_b200 as d
from dxrvoxelizer_b200 import meshes
s = torch.cuda.Stream()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
for name, m in (("dragon", d.load_obj(d.asset_path("dragon.obj"))), ("bowl", d.load_obj(d.asset_path("TuringBowl.obj"))), ("ico5", meshes.icosphere(5)), ("knot200k", meshes.torus_knot(1000, 100))):
    for fused in ("1", "0"):
        os.environ["DXRV_FUSED_BUILD"] = fused
        vox = d.Voxelizer(0); vox.set_stream(s.cuda_stream)
        vb = torch.from_numpy(m.vertex_bytes).cuda(); ib = torch.from_numpy(m.indices.view(np.int32)).cuda()
        build = lambda: vox.build_bvh_device(vb.data_ptr(), m.num_vertices, m.stride, ib.data_ptr(), m.indices.size)
        t = []
        for with_build in (True, False):
            for _ in range(5):
                build(); vox.voxelize(N, d.MODE_PARITY)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(50):
                if with_build: build()
                vox.voxelize(N, d.MODE_PARITY)
            e1.record(s); torch.cuda.synchronize()
            t.append(e0.elapsed_time(e1) * 20)
        print("%s T=%d fused=%s: step %.1f us, voxelize alone %.1f us -> build %.1f us" % (name, m.num_triangles, fused, t[0], t[1], t[0] - t[1]))
        vox.close()
