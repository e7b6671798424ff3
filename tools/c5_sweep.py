"""C5 (256 OBJ files -> 256^3 grids on one GPU): dxrv_voxelize_obj_batch over (contexts, loader threads), with and without
the read-back -- which stage bounds the step.  python tools/c5_sweep.py  (prints one line per setting)"""
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import dxrvoxelizer_b200 as d  # noqa: E402
from dxrvoxelizer_b200 import meshes  # noqa: E402
from bench_configs import write_obj  # noqa: E402

N, n_mesh = 256, 256
tmp = os.path.join(tempfile.gettempdir(), "dxrv_c5_%d" % os.getuid())
os.makedirs(tmp, exist_ok=True)
paths = []
for i in range(n_mesh):
    p = os.path.join(tmp, "ico5_%03d.obj" % i)
    if not os.path.exists(p):
        write_obj(p, meshes.icosphere(5, seed=i, rotate=True, normals=False))
    paths.append(p)
P = (N + 31) // 32
h = torch.empty((n_mesh, N * N * P * 4), dtype=torch.uint8).pin_memory()
ctxs = [d.Voxelizer(0) for _ in range(16)]
cores = os.cpu_count()
for streams, loaders, fetch in [(4, 16, True), (8, 16, True), (16, 16, True), (8, 12, True), (8, 24, True), (6, 14, True),
                                (4, 16, False), (8, 16, False), (1, 16, False), (4, 4, True), (4, 8, True)]:
    best = 1e9
    for it in range(5):
        t0 = time.perf_counter()
        d.voxelize_obj_batch(ctxs[:streams], paths, N, d.MODE_PARITY, out_ptr=h.data_ptr(), loader_threads=loaders, fetch=fetch)
        best = min(best, time.perf_counter() - t0)
    print("contexts %2d loaders %2d fetch %d: %.2f ms per 256 meshes = %.0f meshes/s (%d cores)" % (streams, loaders, fetch, best * 1e3, n_mesh / best, cores), flush=True)
