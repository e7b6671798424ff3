#!/usr/bin/env python3
"""Build-phase timing at C4 scale (development aid): mean build / sort time over 10 builds of the 16.8 M-triangle knot."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dxrvoxelizer_b200 as d
from dxrvoxelizer_b200 import meshes, _lib as L
m = meshes.torus_knot(4096, 2048, normals=False)
vb = torch.from_numpy(m.vertex_bytes).cuda(); ib = torch.from_numpy(m.indices.view(np.int32)).cuda()
v = d.Voxelizer(0)
v.set_profiling(True)
bs = ss = 0
for i in range(12):
    v.build_bvh_device(vb.data_ptr(), m.num_vertices, m.stride, ib.data_ptr(), m.indices.size)
    if i >= 2:
        bs += v.info(L.INFO_LAST_BUILD_NS); ss += v.info(L.INFO_LAST_SORT_NS)
print("%s: build %.1f us  sort %.1f us (%.1f per pass, %.0f GB/s)" % (os.environ.get("TAG", ""), bs / 10e3, ss / 10e3, ss / 40e3, 16.0 * m.num_triangles * 4 / (ss / 10) ))
