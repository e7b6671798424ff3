#!/usr/bin/env python3
"""Fill-kernel time of one z-slab (profiling events of the C ABI).  usage: slab_fill_time.py N z0 z1 [asset] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dxrvoxelizer_b200 as d
N, z0, z1 = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
name = sys.argv[4] if len(sys.argv) > 4 else "dragon.obj"
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 60
m = d.load_obj(d.asset_path(name))
v = d.Voxelizer(0)
v.build_bvh(m)
v.set_profiling(True)
fill = []
for _ in range(reps):
    v.voxelize(N, d.MODE_PARITY, z0, z1)
    v.synchronize()
    fill.append(v.info(6) * 1e-3)
print("N=%d [%d,%d) %s split=%s: fill mean %.2f min %.1f us | inside %d" % (N, z0, z1, name, os.environ.get("DXRV_DBG_SPLIT", "auto"), np.mean(fill[5:]), min(fill[5:]), v.count_inside()))
