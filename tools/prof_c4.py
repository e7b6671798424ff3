#!/usr/bin/env python3
"""ncu driver for the build-dominated regime: torus knot with 16.8 M triangles (or icosphere k), one build + one voxelize."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dxrvoxelizer_b200 as d
from dxrvoxelizer_b200 import meshes
which = sys.argv[1] if len(sys.argv) > 1 else "knot"
m = meshes.torus_knot(4096, 2048, normals=False) if which == "knot" else meshes.icosphere(int(which), normals=False)
vb = torch.from_numpy(m.vertex_bytes).cuda(); ib = torch.from_numpy(m.indices.view(np.int32)).cuda()
v = d.Voxelizer(0)
for _ in range(2):
    v.build_bvh_device(vb.data_ptr(), m.num_vertices, m.stride, ib.data_ptr(), m.indices.size)
    v.voxelize(512, d.MODE_PARITY)
v.synchronize()
print("T", m.num_triangles, "inside", v.count_inside())
