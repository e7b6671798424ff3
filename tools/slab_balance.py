#!/usr/bin/env python3
"""Per-slab voxelize time on ONE GPU for the z-slab partitions bench.py would use on `world` GPUs
(development aid for the cost model of sharding.balanced_slabs).  usage: slab_balance.py [N] [world] [weights...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dxrvoxelizer_b200 as d
from dxrvoxelizer_b200 import sharding

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
weights = [float(w) for w in sys.argv[3:]] or [0.0, 1.5, 3.0, 5.0, 8.0]
m = d.load_obj(d.asset_path("dragon.obj"))
s = torch.cuda.Stream()
v = d.Voxelizer(0); v.set_stream(s.cuda_stream)
v.build_bvh(m)

def t_slab(z0, z1, iters=10):
    for _ in range(3): v.voxelize(N, d.MODE_PARITY, z0, z1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(iters): v.voxelize(N, d.MODE_PARITY, z0, z1)
    e1.record(s); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

for w in weights:
    slabs = sharding.balanced_slabs(m, N, world, compute_weight=w) if w > 0 else [sharding.slab_range(r, world, N) for r in range(world)]
    ts = [t_slab(z0, z1) for z0, z1 in slabs]
    print("weight %4.1f: max %.1f us  mean %.1f  slabs %s  times %s" % (w, max(ts), np.mean(ts), [z1 - z0 for z0, z1 in slabs], ["%.0f" % t for t in ts]))
