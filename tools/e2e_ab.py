#!/usr/bin/env python3
"""A/B of host-side settings of the end-to-end call INSIDE one process (one box, one buffer placement): rounds of
`steps` calls per setting, the settings taken in turn.  usage: e2e_ab.py "VAR=a,VAR=b;VAR2=c,..." [N] [asset] [rounds] [steps]
(a setting is a ;-separated list of VAR=value; `-` = no variables).  Only variables the library re-reads per call work
(DXRV_HOST_ZERO, E2E_ONE_CALL)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dxrvoxelizer_b200 as d
from dxrvoxelizer_b200 import _lib as L
settings = [dict(kv.split("=") for kv in s.split(";") if kv != "-") for s in sys.argv[1].split(",")]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
m = d.load_obj(d.asset_path(sys.argv[3] if len(sys.argv) > 3 else "dragon.obj"))
rounds = int(sys.argv[4]) if len(sys.argv) > 4 else 5
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 40
vox = d.Voxelizer(0)
nbytes = N * N * ((N + 31) // 32) * 4
h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
vb = torch.from_numpy(m.vertex_bytes.copy()).pin_memory(); ib = torch.from_numpy(m.indices.view(np.int32).copy()).pin_memory()
vox.set_read_back(L.READ_BACK_SPARSE)
def step():
    if os.environ.get("E2E_ONE_CALL", "1") == "1":
        vox.voxelize_mesh_to_host(vb.data_ptr(), m.num_vertices, m.stride, ib.data_ptr(), m.indices.size, N, d.MODE_PARITY, 0, N, h.data_ptr(), nbytes, chunks=8)
    else:
        vox.build_bvh_host_ptr(vb.data_ptr(), m.num_vertices, m.stride, ib.data_ptr(), m.indices.size)
        vox.voxelize_to_host(N, d.MODE_PARITY, 0, N, h.data_ptr(), nbytes, chunks=8)
for _ in range(30): step()
res = [[] for _ in settings]
for r in range(rounds):
    for i, st in enumerate(settings):
        saved = {k: os.environ.get(k) for k in st}
        os.environ.update(st)
        for _ in range(5): step()
        t = []
        for _ in range(steps):
            t0 = time.perf_counter(); step(); t.append((time.perf_counter() - t0) * 1e3)
        res[i].append(float(np.mean(t)))
        for k, v in saved.items():
            if v is None: os.environ.pop(k, None)
            else: os.environ[k] = v
for st, r in zip(settings, res):
    print("%-40s mean of rounds %.3f ms | rounds %s" % (";".join("%s=%s" % kv for kv in st.items()) or "-", np.mean(r), " ".join("%.3f" % v for v in r)))
