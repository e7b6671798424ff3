"""The CPU oracle checked against itself and against what little the reference pins.

The reference has no tests and no golden grids: "parity unpinned" for
the voxel grid.  What can be checked without a GPU: the two tiers of the oracle agree (brute force
is the oracle's own ground truth), the metamorphic properties of watertight meshes hold, analytic
cases come out exactly, and the counts recorded when the oracle was first validated are stable.
"""
import numpy as np
import pytest

from conftest import popcount

# inside-voxel counts (MODE_SHADER, MODE_PARITY) at N=64, the reference's GRID_SIZE (Voxelizer.cpp:8),
# recorded from this oracle.  The surveyor's independent Moller-Trumbore prototype (SURVEY.md 8c,
# BASELINE.md section 4; "indicative, not golden") found 14477/14529, 52303/52356 and 11753/11772:
# identical except 2 voxels of TuringBowl in MODE_SHADER, where the two intersection routines differ.
GOLDEN_64 = {"dragon.obj": (14477, 14529), "bunny.obj": (52303, 52356), "TuringBowl.obj": (11755, 11772)}


@pytest.mark.parametrize("name", sorted(GOLDEN_64))
def test_counts_at_reference_grid_size(name, assets, oracle_mod):
    m = assets(name)
    shader = oracle_mod.voxelize(m.vertices, m.indices, 64, oracle_mod.MODE_SHADER)
    parity = oracle_mod.voxelize(m.vertices, m.indices, 64, oracle_mod.MODE_PARITY)
    assert popcount(shader["bits"]) == GOLDEN_64[name][0]
    assert popcount(parity["bits"]) == GOLDEN_64[name][1]
    assert parity["odd_columns"] == 0          # watertight => every column crosses an even number of times
    # the two modes are different functions: report-level agreement only
    assert popcount(shader["bits"] ^ parity["bits"]) < 0.001 * 64 ** 3


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("N", [16, 31, 32])
def test_accelerated_tier_equals_brute_force(mode, N, meshes_mod, oracle_mod):
    for m in (meshes_mod.icosphere(2, seed=3), meshes_mod.torus_knot(48, 12, seed=5)):
        a = oracle_mod.voxelize(m.vertices, m.indices, N, mode, tier=oracle_mod.TIER_ACCEL, texels=(mode == 0))
        b = oracle_mod.voxelize(m.vertices, m.indices, N, mode, tier=oracle_mod.TIER_BRUTE, texels=(mode == 0))
        assert np.array_equal(a["bits"], b["bits"])
        if mode == 0:
            assert np.array_equal(a["texels"], b["texels"])
        else:
            assert a["crossings"] == b["crossings"] and a["odd_columns"] == 0


def test_accelerated_tier_equals_brute_force_on_dragon_slab(assets, oracle_mod):
    m = assets("dragon.obj")
    for mode in (0, 1):
        a = oracle_mod.voxelize(m.vertices, m.indices, 32, mode, z0=12, z1=16, tier=1)
        b = oracle_mod.voxelize(m.vertices, m.indices, 32, mode, z0=12, z1=16, tier=0)
        assert np.array_equal(a["bits"], b["bits"]) and popcount(a["bits"]) > 0


def test_slabs_concatenate_to_full_grid(assets, oracle_mod):
    m = assets("bunny.obj")
    for mode in (0, 1):
        full = oracle_mod.voxelize(m.vertices, m.indices, 32, mode)["bits"]
        parts = [oracle_mod.voxelize(m.vertices, m.indices, 32, mode, z0=z, z1=z + 8)["bits"] for z in range(0, 32, 8)]
        assert np.array_equal(np.concatenate(parts, 0), full)


def test_cube_is_exact(meshes_mod, oracle_mod):
    import dxrvoxelizer_b200 as d
    m = meshes_mod.cube(0.75)
    N = 32
    c = (np.arange(N, dtype=np.float32) + np.float32(0.5)) / np.float32(N) * np.float32(2) - np.float32(1)
    inside1d = np.abs(c) < 0.75
    want = inside1d[:, None, None] & inside1d[None, :, None] & inside1d[None, None, :]
    for mode in (0, 1):
        got = oracle_mod.voxelize(m.vertices, m.indices, N, mode, bound=[0, 0, 0, 1])["bits"]
        assert np.array_equal(d.unpack_bits(got, N).astype(bool), want)


def test_ray_through_shared_edges_and_vertices_counts_once(oracle_mod, meshes_mod):
    """Columns passing exactly through mesh edges / vertices: the (+e,+e^2) tie rule must keep every
    column's crossing count even and the fill identical to a slightly shifted cube."""
    import dxrvoxelizer_b200 as d
    N = 16
    # cube corners ON voxel-centre coordinates: faces' diagonals and edges pass through column centres
    h = float((np.float32(11.5) / np.float32(16)) * np.float32(2) - np.float32(1))  # centre(11)
    m = meshes_mod.cube(h)
    r = oracle_mod.voxelize(m.vertices, m.indices, N, oracle_mod.MODE_PARITY, bound=[0, 0, 0, 1], tier=0)
    assert r["odd_columns"] == 0
    occ = d.unpack_bits(r["bits"], N)
    # a column is either empty or one solid run
    runs = np.abs(np.diff(np.pad(occ.astype(np.int8), ((0, 0), (0, 0), (1, 1))), axis=-1)).sum(-1)
    assert set(np.unique(runs)) <= {0, 2}
    assert np.array_equal(r["bits"], oracle_mod.voxelize(m.vertices, m.indices, N, 1, bound=[0, 0, 0, 1], tier=1)["bits"])


def test_odd_grid_centre_voxel_has_no_ray(meshes_mod, oracle_mod):
    import dxrvoxelizer_b200 as d
    m = meshes_mod.icosphere(1, seed=2)
    occ = d.unpack_bits(oracle_mod.voxelize(m.vertices, m.indices, 15, oracle_mod.MODE_SHADER)["bits"], 15)
    assert occ[7, 7, 7] == 0      # normalize(0) is NaN: miss (hlsl:52)
    assert occ[7, 7, 8] == 1


def test_bound_matches_reference_formula(assets, oracle_mod):
    m = assets("TuringBowl.obj")
    assert np.array_equal(oracle_mod.bound(m.vertices), m.bound)


def test_invalid_arguments(oracle_mod, meshes_mod):
    m = meshes_mod.cube()
    with pytest.raises(ValueError):
        oracle_mod.voxelize(m.vertices, m.indices, 8, 7)
    with pytest.raises(ValueError):
        oracle_mod.voxelize(m.vertices, m.indices, 8, 1, z0=4, z1=4)
    with pytest.raises(ValueError):
        oracle_mod.voxelize(m.vertices[:, :3].copy(), m.indices, 8, 0)   # MODE_SHADER needs normals
