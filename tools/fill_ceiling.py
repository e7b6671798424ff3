#!/usr/bin/env python3
"""Write-path ceiling of k_trace_fill_columns: voxelize an EMPTY mesh (every tile streams zeros)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dxrvoxelizer_b200 as d
from dxrvoxelizer_b200 import meshes, _lib as L
s = torch.cuda.Stream()
vox = d.Voxelizer(0); vox.set_stream(s.cuda_stream)
c = meshes.cube()
empty = d.Mesh(c.vertex_bytes, np.zeros(0, np.uint32), c.stride)
dragon = d.load_obj(d.asset_path("dragon.obj"))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, m in (("empty", empty), ("dragon", dragon)):
    vox.build_bvh(m)
    vox.set_profiling(True)
    for N in (1024, 2048):
        w = f = 0
        for i in range(12):
            with torch.cuda.stream(s):
                flush.fill_(i)
            vox.voxelize(N, d.MODE_PARITY)
            if i >= 2:
                w += vox.info(L.INFO_LAST_WALK_NS); f += vox.info(L.INFO_LAST_FILL_NS)
        print("%s N=%d: walk %.1f us, fill %.1f us -> %.0f GB/s grid write" % (name, N, w / 10e3, f / 10e3, N ** 3 / 8 / (f / 10) ))
    vox.set_profiling(False)
