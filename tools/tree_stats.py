#!/usr/bin/env python3
"""Depth statistics of the LBVH built for a shipped mesh (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dxrvoxelizer_b200 as d
from dxrvoxelizer_b200 import _lib as L
for name in ("dragon.obj", "TuringBowl.obj"):
    m = d.load_obj(d.asset_path(name)); T = m.num_triangles
    v = d.Voxelizer(0); v.build_bvh(m)
    nodes = v.debug_read(L.DBG_NODES, np.uint32, (T - 1) * 16).reshape(T - 1, 16)
    child = nodes[:, 12:14]; leaf = (child & 0x80000000) != 0; idx = (child & 0x7fffffff).astype(np.int64)
    depth = np.zeros(T - 1, np.int64); leafdepth = np.zeros(T, np.int64)
    frontier = np.array([0])
    while frontier.size:
        nxt = []
        for c in range(2):
            ch = idx[frontier, c]; isl = leaf[frontier, c]
            leafdepth[ch[isl]] = depth[frontier[isl]] + 1
            depth[ch[~isl]] = depth[frontier[~isl]] + 1
            nxt.append(ch[~isl])
        frontier = np.concatenate(nxt)
    print(name, "T", T, "max inner depth", depth.max(), "leaf depth mean %.1f p99 %d max %d" % (leafdepth.mean(), np.percentile(leafdepth, 99), leafdepth.max()))
    v.close()
