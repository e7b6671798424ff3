"""The BASELINE.json configurations at their STATED sizes, each against the CPU oracle:
C2 TuringBowl 512^3 MODE_PARITY (full grid), C4 5.24 M and 16.8 M triangles at 512^3, C5 256 distinct
icosphere(5) meshes at 256^3 on four streams (every grid checked)."""
import numpy as np
import pytest

import dxrvoxelizer_b200 as d
from conftest import popcount
from dxrvoxelizer_b200 import _lib as L

pytestmark = pytest.mark.gpu


def test_c2_turingbowl_512_parity_full_grid(vox, assets, oracle_mod):
    m = assets("TuringBowl.obj")
    vox.build_bvh(m)
    vox.voxelize(512, d.MODE_PARITY)
    ref = oracle_mod.voxelize(m.vertices, m.indices, 512, oracle_mod.MODE_PARITY)
    assert popcount(vox.fetch_bits() ^ ref["bits"]) == 0
    assert vox.info(L.INFO_CROSSINGS) == ref["crossings"] and ref["odd_columns"] == 0


@pytest.mark.parametrize("which", ["icosphere9_5.24M", "knot_16.8M"])
def test_c4_millions_of_triangles_at_512(vox, meshes_mod, oracle_mod, which):
    """Build-dominated regime at its stated size: 30-bit keys, four big-tile radix passes, atomic refit, scatter
    voxelization.  The oracle checks z-slabs (it needs seconds per slab at this triangle count)."""
    if which.startswith("ico"):
        m = meshes_mod.icosphere(9, seed=1234, normals=False)
        assert m.num_triangles == 5242880
    else:
        m = meshes_mod.torus_knot(4096, 2048, normals=False)
        assert m.num_triangles == 16777216
    N = 512
    vox.build_bvh(m)
    vox.voxelize(N, d.MODE_PARITY)
    got = vox.fetch_bits()
    keys = vox.debug_read(L.DBG_MORTON_SORTED, np.uint32, m.num_triangles)
    assert (np.diff(keys.astype(np.int64)) >= 0).all()
    inside = vox.count_inside()
    assert 0 < inside < N ** 3
    for z0 in (3, N // 2 - 2, N - 40):
        ref = oracle_mod.voxelize(m.vertices, m.indices, N, oracle_mod.MODE_PARITY, z0=z0, z1=z0 + 4)
        assert ref["odd_columns"] == 0
        assert popcount(got[z0:z0 + 4] ^ ref["bits"]) == 0, z0
    # the slab API on the same tree gives the same layers
    vox.voxelize(N, d.MODE_PARITY, 250, 262)
    assert np.array_equal(vox.fetch_bits(), got[250:262])


def test_c5_batch_of_256_meshes_at_256(meshes_mod, oracle_mod):
    """C5's actual shape: 256 x icosphere(5) (20 480 triangles each, seed = mesh index, random rotation) at 256^3,
    four contexts/streams on one GPU, all grids in flight before the first is read back; every grid checked."""
    ctxs = [d.Voxelizer(0) for _ in range(4)]
    N = 256
    for base in range(0, 256, 16):
        ms = [meshes_mod.icosphere(5, seed=i, rotate=True, normals=False) for i in range(base, base + 16)]
        grids = []
        for k in range(0, 16, 4):
            for j, c in enumerate(ctxs):
                c.build_bvh(ms[k + j])
                c.voxelize(N, d.MODE_PARITY)
            grids += [c.fetch_bits() for c in ctxs]
        for m, g in zip(ms, grids):
            assert m.num_triangles == 20480
            assert popcount(g ^ oracle_mod.voxelize(m.vertices, m.indices, N, oracle_mod.MODE_PARITY)["bits"]) == 0
    for c in ctxs:
        c.close()
