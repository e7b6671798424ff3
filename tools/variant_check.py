#!/usr/bin/env python3
"""A/B aid: for each (asset, N) voxelize in MODE_PARITY, XOR the whole grid against the oracle, and print the candidate /
fill kernel times (profiling events of the C ABI).  usage: variant_check.py [asset:N ...]   (on the GPU box)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dxrvoxelizer_b200 as d
import oracle

cases = [a.split(":") for a in sys.argv[1:]] or [["dragon.obj", "1024"], ["TuringBowl.obj", "1024"], ["bunny.obj", "1024"], ["dragon.obj", "2048"]]
v = d.Voxelizer(0)
for name, N in cases:
    N = int(N)
    m = d.load_obj(d.asset_path(name))
    v.build_bvh(m)
    v.set_profiling(False)
    v.voxelize(N, d.MODE_PARITY)
    got = v.fetch_bits()
    ref = oracle.voxelize(m.vertices, m.indices, N, d.MODE_PARITY)["bits"]
    mism = int(np.unpackbits((got ^ ref).view(np.uint8)).sum())
    v.set_profiling(True)
    cand, fill = [], []
    for _ in range(45):
        v.voxelize(N, d.MODE_PARITY)
        v.synchronize()
        cand.append(v.info(5) * 1e-3)
        fill.append(v.info(6) * 1e-3)
    cand, fill = cand[5:], fill[5:]
    print("%-14s N=%4d mismatched_voxels=%d | candidates mean %.2f us | fill mean %.2f median %.2f min %.2f us" % (
        name, N, mism, np.mean(cand), np.mean(fill), np.median(fill), min(fill)), flush=True)
v.close()
