"""dxrv_voxelize_obj_batch on the GPU: OBJ text in, one grid per mesh out, identical to loading and voxelizing every
mesh on its own and to the CPU oracle (BASELINE config 5's path: parse -> upload -> build -> voxelize -> read-back)."""
import numpy as np
import pytest

import dxrvoxelizer_b200 as d
from bench_configs import write_obj
from conftest import popcount

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def obj_files(tmp_path_factory, meshes_mod):
    root = tmp_path_factory.mktemp("batch")
    paths = []
    for i in range(13):                                   # not a multiple of the number of contexts
        p = root / ("ico_%02d.obj" % i)
        write_obj(str(p), meshes_mod.icosphere(2 + i % 3, seed=100 + i, rotate=True, normals=False))
        paths.append(str(p))
    return paths


@pytest.mark.parametrize("num_ctx,loaders,N,mode", [(4, 0, 96, d.MODE_PARITY), (1, 2, 64, d.MODE_PARITY), (3, 5, 48, d.MODE_SHADER)])
def test_batch_equals_one_mesh_at_a_time_and_the_oracle(vox, obj_files, oracle_mod, num_ctx, loaders, N, mode):
    ctxs = [d.Voxelizer(0) for _ in range(num_ctx)]
    try:
        grids, tris = d.voxelize_obj_batch(ctxs, obj_files, N, mode, loader_threads=loaders)
        assert grids.shape == (len(obj_files), N, N, (N + 31) // 32)
        for k, p in enumerate(obj_files):
            m = d.load_obj(p)
            assert tris[k] == m.num_triangles
            vox.build_bvh(m)
            vox.voxelize(N, mode)
            assert np.array_equal(grids[k], vox.fetch_bits()), k
            ref = oracle_mod.voxelize(m.vertices, m.indices, N, mode)["bits"]
            assert popcount(grids[k] ^ ref) == 0, k
        # the contexts are left holding their last mesh
        last = {s: max(k for k in range(len(obj_files)) if k % num_ctx == s) for s in range(num_ctx)}
        for s, c in enumerate(ctxs):
            assert np.array_equal(c.fetch_bits(), grids[last[s]])
        # voxelize only (no host buffer)
        none, tris2 = d.voxelize_obj_batch(ctxs, obj_files, N, mode, fetch=False)
        assert none is None and np.array_equal(tris, tris2)
        for s, c in enumerate(ctxs):
            assert np.array_equal(c.fetch_bits(), grids[last[s]])
    finally:
        for c in ctxs:
            c.close()


def test_batch_reports_the_file_that_failed(obj_files, tmp_path):
    ctxs = [d.Voxelizer(0) for _ in range(2)]
    try:
        bad = list(obj_files)
        bad[5] = str(tmp_path / "missing.obj")
        with pytest.raises(d.DxrvError) as e:
            d.voxelize_obj_batch(ctxs, bad, 64, d.MODE_PARITY)
        assert e.value.code == -5 and "missing.obj" in str(e.value)
        # the contexts stay usable
        grids, _ = d.voxelize_obj_batch(ctxs, obj_files[:3], 64, d.MODE_PARITY)
        assert grids.any()
    finally:
        for c in ctxs:
            c.close()
