"""Committed golden fixtures (tests/golden/oracle_grids_64.npz, made by tests/golden/make_oracle_counts.py from
the oracle at the reference's GRID_SIZE 64): the oracle must keep reproducing them (CPU), and the CUDA
path must match them bit for bit (GPU) -- independently of the live oracle comparison."""
import json
import os

import numpy as np
import pytest

import dxrvoxelizer_b200 as d
from conftest import popcount

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "oracle_grids_64.npz"))
COUNTS = json.load(open(os.path.join(HERE, "golden", "oracle_counts_64.json")))["meshes"]
MESHES = ["dragon", "bunny", "TuringBowl"]


@pytest.mark.parametrize("name", MESHES)
def test_oracle_reproduces_golden_grids(name, assets, oracle_mod):
    m = assets(name + ".obj")
    for mode, tag in ((0, "shader"), (1, "parity")):
        got = oracle_mod.voxelize(m.vertices, m.indices, 64, mode, texels=(mode == 0))
        assert np.array_equal(got["bits"], GOLD["%s_%s" % (name, tag)])
        assert popcount(got["bits"]) == COUNTS[name + ".obj"]["inside_" + tag]
        if mode == 0 and name == "dragon":
            assert np.array_equal(got["texels"], GOLD["dragon_texels"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", MESHES)
def test_gpu_matches_golden_grids(name, assets, vox):
    m = assets(name + ".obj")
    vox.build_bvh(m)
    for mode, tag in ((d.MODE_SHADER, "shader"), (d.MODE_PARITY, "parity")):
        vox.voxelize(64, mode, texels=(mode == d.MODE_SHADER))
        assert popcount(vox.fetch_bits() ^ GOLD["%s_%s" % (name, tag)]) == 0
        if mode == d.MODE_SHADER and name == "dragon":
            assert np.array_equal(vox.fetch_texels(), GOLD["dragon_texels"])
