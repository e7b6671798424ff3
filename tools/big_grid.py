#!/usr/bin/env python3
"""Large grids: dragon at 2048^3 / 4096^3 (8 GiB bit grid), kernel times and checks against the oracle on a slab."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dxrvoxelizer_b200 as d
from dxrvoxelizer_b200 import _lib as L
import oracle
s = torch.cuda.Stream()
vox = d.Voxelizer(0); vox.set_stream(s.cuda_stream)
m = d.load_obj(d.asset_path("dragon.obj"))
vox.build_bvh(m)
for N in (2048, 4096):
    vox.set_profiling(True)
    w = f = 0
    for i in range(5):
        vox.voxelize(N, d.MODE_PARITY)
        if i >= 1:
            w += vox.info(L.INFO_LAST_WALK_NS); f += vox.info(L.INFO_LAST_FILL_NS)
    vox.set_profiling(False)
    inside = vox.count_inside()
    print("N=%d: walk %.1f us, fill %.1f us -> %.0f GB/s grid write (%.1f%% of 6463), inside fraction %.4f" %
          (N, w / 4e3, f / 4e3, N ** 3 / 8 / (f / 4), 100 * N ** 3 / 8 / (f / 4) / 6463, inside / N ** 3), flush=True)
    z0 = N // 2
    vox.voxelize(N, d.MODE_PARITY, z0, z0 + 8)
    ref = oracle.voxelize(m.vertices, m.indices, N, 1, z0=z0, z1=z0 + 8)
    mism = int(np.unpackbits((vox.fetch_bits() ^ ref["bits"]).view(np.uint8)).sum())
    print("   slab [%d,%d) vs oracle: mismatched voxels %d, crossings %d / %d" % (z0, z0 + 8, mism, vox.info(L.INFO_CROSSINGS), ref["crossings"]), flush=True)
