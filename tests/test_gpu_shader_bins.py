"""MODE_SHADER has two traversal kernels: the direction bins (csrc/shader_bins.cu, the default: an exact
accelerator for the reference's radial ray family, DXRVoxelizer.hlsl:44-53) and the general LBVH walk
(csrc/trace_shader.cu, the device-side overflow path).  Both must equal the CPU oracle bit for bit -- grid AND
R10G10B10A2 texels -- on the shipped meshes, on soups that exercise every special case of the binning (huge
rectangles, slivers, triangles through the grid centre -> near list, near-list overflow, entry-budget overflow),
for any bin resolution."""
import numpy as np
import pytest

import dxrvoxelizer_b200 as d
from conftest import popcount

pytestmark = pytest.mark.gpu

SHIPPED = ["dragon.obj", "bunny.obj", "TuringBowl.obj"]


def _shader(vox, mesh, N, z0=0, z1=None, bound=None, texels=True):
    vox.build_bvh(mesh, bound=bound)
    vox.voxelize(N, d.MODE_SHADER, z0, z1, texels=texels)
    return vox.fetch_bits(), (vox.fetch_texels() if texels else None)


def _both_paths(vox, monkeypatch, mesh, N, ref, z0=0, z1=None, bound=None):
    for path in ("bins", "bvh"):
        monkeypatch.setenv("DXRV_SHADER_PATH", path)
        bits, tex = _shader(vox, mesh, N, z0, z1, bound)
        assert popcount(bits ^ ref["bits"]) == 0, path
        assert np.array_equal(tex, ref["texels"]), path
    monkeypatch.delenv("DXRV_SHADER_PATH")


@pytest.mark.parametrize("name", SHIPPED)
@pytest.mark.parametrize("N", [64, 128])
def test_bins_and_bvh_equal_oracle_on_shipped_meshes(vox, assets, oracle_mod, monkeypatch, name, N):
    m = assets(name)
    ref = oracle_mod.voxelize(m.vertices, m.indices, N, oracle_mod.MODE_SHADER, texels=True)
    _both_paths(vox, monkeypatch, m, N, ref)


@pytest.mark.parametrize("R", [8, 64, 1024])
def test_any_bin_resolution_gives_the_same_grid(vox, assets, oracle_mod, monkeypatch, R):
    m = assets("bunny.obj")
    ref = oracle_mod.voxelize(m.vertices, m.indices, 64, oracle_mod.MODE_SHADER, texels=True)
    monkeypatch.setenv("DXRV_SHADER_BINS_R", str(R))
    monkeypatch.setenv("DXRV_SHADER_PATH", "bins")               # (N < 160 would take the LBVH walk by default)
    bits, tex = _shader(vox, m, 64)
    assert popcount(bits ^ ref["bits"]) == 0 and np.array_equal(tex, ref["texels"])


def _soup(seed, n, with_normals=True):
    """Open soup of triangles of every size and position (no structure a cull could rely on): specks, slivers,
    sheets spanning the scene, triangles through and next to the grid centre."""
    rng = np.random.default_rng(seed)
    centre = rng.uniform(-1, 1, size=(n, 1, 3))
    centre[: n // 8] *= 0.02                                     # a cluster around the grid centre
    size = 10.0 ** rng.uniform(-3.5, 0.3, size=(n, 1, 1))
    tri = centre + size * rng.uniform(-1, 1, size=(n, 3, 3))
    sliver = rng.random(n) < 0.15
    tri[sliver, 2] = tri[sliver, 0] + (tri[sliver, 1] - tri[sliver, 0]) * rng.uniform(0, 1, size=(int(sliver.sum()), 1)) \
        + 1e-6 * rng.uniform(-1, 1, size=(int(sliver.sum()), 3))
    tri[0] = [[-0.3, -0.2, 0.0], [0.4, -0.1, 0.0], [0.0, 0.5, 0.0]]        # contains the grid centre
    tri[1] = [[-0.5, 1e-4, -0.5], [0.5, 1e-4, -0.5], [0.0, 1e-4, 0.7]]     # 1e-4 from it
    pos = np.concatenate([tri.reshape(-1, 3), [[-1, -1, -1], [1, 1, 1]]]).astype(np.float32)
    nrm = rng.normal(size=pos.shape).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    return d.Mesh.from_arrays(pos, np.arange(3 * n, dtype=np.uint32).reshape(n, 3), nrm)


@pytest.mark.parametrize("seed,n,N", [(1, 300, 33), (2, 300, 64), (3, 2000, 48), (4, 60, 96)])
def test_triangle_soup_both_paths(vox, oracle_mod, monkeypatch, seed, n, N):
    m = _soup(seed, n)
    ref = oracle_mod.voxelize(m.vertices, m.indices, N, oracle_mod.MODE_SHADER, texels=True)
    assert popcount(ref["bits"]) > 0
    _both_paths(vox, monkeypatch, m, N, ref)


def test_soup_against_brute_force_oracle(vox, oracle_mod, monkeypatch):
    """Same, against the oracle's own ground truth (every ray against every triangle)."""
    m = _soup(7, 200)
    ref = oracle_mod.voxelize(m.vertices, m.indices, 40, oracle_mod.MODE_SHADER, texels=True, tier=oracle_mod.TIER_BRUTE)
    _both_paths(vox, monkeypatch, m, 40, ref)


def test_near_list_overflow_falls_back_on_the_device(vox, meshes_mod, oracle_mod, monkeypatch):
    """More than 4096 triangles within 1e-3 of the grid centre: the near list overflows, the bins raise their
    device flag and the LBVH walk produces the grid (no host round trip) -- still the oracle's bits."""
    from dxrvoxelizer_b200 import _lib as L
    monkeypatch.setenv("DXRV_SHADER_PATH", "bins")
    ball = meshes_mod.icosphere(4, seed=2)                       # 5120 triangles
    v = ball.vertices.copy()
    v[:, :3] *= 4e-4
    far = meshes_mod.icosphere(3, seed=5)
    nv = v.shape[0]
    verts = np.concatenate([v, far.vertices])
    idx = np.concatenate([ball.indices.reshape(-1, 3), far.indices.reshape(-1, 3) + nv]).astype(np.uint32)
    m = d.Mesh(verts, idx, ball.stride)
    bound = [0.0, 0.0, 0.0, 1.25]                                # pin the grid centre onto the small ball
    ref = oracle_mod.voxelize(m.vertices, m.indices, 64, oracle_mod.MODE_SHADER, texels=True, bound=bound)
    bits, tex = _shader(vox, m, 64, bound=bound)
    assert popcount(bits ^ ref["bits"]) == 0 and np.array_equal(tex, ref["texels"])
    st = vox.debug_read(L.DBG_BINS_STATE, np.uint32, 4)
    assert st[1] == 1 and st[2] > 4096                           # overflow flag up, near list over its capacity


def test_entry_budget_overflow_falls_back_on_the_device(vox, oracle_mod, monkeypatch):
    """Thousands of scene-sized triangles: every one covers a large part of every cube-map face, the lists would
    need far more than the entry budget -> flag -> LBVH walk."""
    from dxrvoxelizer_b200 import _lib as L
    monkeypatch.setenv("DXRV_SHADER_PATH", "bins")
    rng = np.random.default_rng(11)
    n = 3000
    tri = rng.uniform(-1, 1, size=(n, 3, 3))
    pos = np.concatenate([tri.reshape(-1, 3), [[-1, -1, -1], [1, 1, 1]]]).astype(np.float32)
    nrm = rng.normal(size=pos.shape).astype(np.float32)
    m = d.Mesh.from_arrays(pos, np.arange(3 * n, dtype=np.uint32).reshape(n, 3), nrm)
    ref = oracle_mod.voxelize(m.vertices, m.indices, 32, oracle_mod.MODE_SHADER, texels=True)
    bits, tex = _shader(vox, m, 32)
    assert popcount(bits ^ ref["bits"]) == 0 and np.array_equal(tex, ref["texels"])
    st = vox.debug_read(L.DBG_BINS_STATE, np.uint32, 4)
    assert st[1] == 1 and st[0] > (1 << 20)                      # more entries than the budget


@pytest.mark.parametrize("name,N,slabs", [("dragon.obj", 256, (100, 127, 128, 200)), ("dragon.obj", 1024, (511, 600)),
                                          ("bunny.obj", 512, (255, 256, 400))])
def test_slabs_at_larger_grids(vox, assets, oracle_mod, name, N, slabs):
    """C3's MODE_SHADER side at its real size (and the sizes between): single layers against the oracle."""
    m = assets(name)
    vox.build_bvh(m)
    for z0 in slabs:
        vox.voxelize(N, d.MODE_SHADER, z0, z0 + 1)
        ref = oracle_mod.voxelize(m.vertices, m.indices, N, oracle_mod.MODE_SHADER, z0=z0, z1=z0 + 1)
        assert popcount(vox.fetch_bits() ^ ref["bits"]) == 0, z0


def test_bins_are_rebuilt_for_every_acceleration_structure(vox, assets, meshes_mod, oracle_mod, monkeypatch):
    """The bins are cached per build: a new mesh (even of the same size) must not see the old lists."""
    monkeypatch.setenv("DXRV_SHADER_PATH", "bins")
    a, b = meshes_mod.icosphere(4, seed=1), meshes_mod.icosphere(4, seed=2, rotate=True)
    for m in (a, b, a):
        bits, _ = _shader(vox, m, 64, texels=False)
        assert popcount(bits ^ oracle_mod.voxelize(m.vertices, m.indices, 64, oracle_mod.MODE_SHADER)["bits"]) == 0
        vox.voxelize(64, d.MODE_SHADER, 10, 20)                  # second voxelize on the same build: cached bins
        assert np.array_equal(vox.fetch_bits(), bits[10:20])


def test_default_path_by_grid_size(vox, assets, oracle_mod):
    """Small grids take the LBVH walk (the bins would cost more to build than 64^3 rays cost to trace), large ones
    build the bins, and later small voxelizes on the same acceleration structure reuse them."""
    from dxrvoxelizer_b200 import _lib as L
    m = assets("bunny.obj")
    ref64 = oracle_mod.voxelize(m.vertices, m.indices, 64, oracle_mod.MODE_SHADER)["bits"]
    vox.build_bvh(m)
    vox.voxelize(64, d.MODE_SHADER)
    assert vox.debug_read(L.DBG_BINS_STATE, np.uint32, 4)[3] == 0            # no bins built
    assert popcount(vox.fetch_bits() ^ ref64) == 0
    vox.voxelize(192, d.MODE_SHADER, 90, 92)
    st = vox.debug_read(L.DBG_BINS_STATE, np.uint32, 4)
    assert st[3] >= 8 and st[1] == 0 and st[0] > 0                           # bins built, no overflow
    ref = oracle_mod.voxelize(m.vertices, m.indices, 192, oracle_mod.MODE_SHADER, z0=90, z1=92)["bits"]
    assert popcount(vox.fetch_bits() ^ ref) == 0
    vox.voxelize(64, d.MODE_SHADER)                                          # reuses the bins
    assert popcount(vox.fetch_bits() ^ ref64) == 0
