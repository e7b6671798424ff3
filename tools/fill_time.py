#!/usr/bin/env python3
"""Walk / fill kernel times (profiling events of the C ABI) for one mesh and grid size.  usage: fill_time.py [N] [asset] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dxrvoxelizer_b200 as d

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
name = sys.argv[2] if len(sys.argv) > 2 else "dragon.obj"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 60
m = d.load_obj(d.asset_path(name))
v = d.Voxelizer(0)
v.build_bvh(m)
v.set_profiling(True)
walk, fill = [], []
for _ in range(reps):
    v.voxelize(N, d.MODE_PARITY)
    v.synchronize()
    walk.append(v.info(5) * 1e-3)
    fill.append(v.info(6) * 1e-3)
walk, fill = walk[5:], fill[5:]
print("N=%d %s walk mean %.2f min %.1f us | fill mean %.2f min %.1f us | inside %d crossings %d" % (
    N, name, np.mean(walk), min(walk), np.mean(fill), min(fill), v.count_inside(), v.info(3)))
