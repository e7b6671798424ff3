#!/usr/bin/env python3
"""Minimal driver for ncu: build the dragon LBVH, then voxelize (MODE_PARITY) a few times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dxrvoxelizer_b200 as d

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
name = sys.argv[3] if len(sys.argv) > 3 else "dragon.obj"
z0 = int(sys.argv[4]) if len(sys.argv) > 4 else 0
z1 = int(sys.argv[5]) if len(sys.argv) > 5 else N
m = d.load_obj(d.asset_path(name))
v = d.Voxelizer(0)
for _ in range(reps):
    v.build_bvh(m)
    v.voxelize(N, d.MODE_PARITY, z0, z1)
v.synchronize()
print("inside", v.count_inside(), "crossings", v.info(3))
