#!/usr/bin/env python3
"""End-to-end step (host mesh -> dense host grid) through dxrv_voxelize_to_host, per transport (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dxrvoxelizer_b200 as d
from dxrvoxelizer_b200 import _lib as L
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
m = d.load_obj(d.asset_path(sys.argv[2] if len(sys.argv) > 2 else "dragon.obj"))
vox = d.Voxelizer(0)
nbytes = N * N * ((N + 31) // 32) * 4
h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
vb = torch.from_numpy(m.vertex_bytes.copy()).pin_memory(); ib = torch.from_numpy(m.indices.view(np.int32).copy()).pin_memory()
only = sys.argv[3] if len(sys.argv) > 3 else None
for name, tr in (("dense", L.READ_BACK_DENSE), ("sparse", L.READ_BACK_SPARSE)):
    if only and name != only:
        continue
    vox.set_read_back(tr)
    def step():
        if os.environ.get("E2E_ONE_CALL", "1") == "1":
            vox.voxelize_mesh_to_host(vb.data_ptr(), m.num_vertices, m.stride, ib.data_ptr(), m.indices.size, N, d.MODE_PARITY, 0, N, h.data_ptr(), nbytes, chunks=8)
            return
        vox.build_bvh_host_ptr(vb.data_ptr(), m.num_vertices, m.stride, ib.data_ptr(), m.indices.size)
        vox.voxelize_to_host(N, d.MODE_PARITY, 0, N, h.data_ptr(), nbytes, chunks=8)
    for _ in range(15): step()
    t = []
    for _ in range(30):
        t0 = time.perf_counter(); step(); t.append((time.perf_counter() - t0) * 1e3)
    print(" ".join("%s=%s" % (k, v) for k, v in sorted(os.environ.items()) if k.startswith("DXRV_") or k.startswith("E2E_")) or "(defaults)", end=" ")
    print("%s: e2e mean %.3f ms, min %.3f ms -> %.0f Gvoxel/s" % (name, np.mean(t), min(t), N ** 3 / np.mean(t) * 1e-6))
