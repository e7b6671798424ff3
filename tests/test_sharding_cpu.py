"""Multi-GPU host logic on CPU: world_size-2 `gloo` job, no GPU.  The slab
partition, the mesh broadcast and the gather are exercised with the ORACLE standing in for the
per-rank kernel (tests may use the oracle; the product path never does)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from dxrvoxelizer_b200.sharding import slab_range, slab_words

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("N", [1, 7, 64, 100, 1024, 1664])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_slabs_partition_the_grid(N, world):
    ranges = [slab_range(r, world, N) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == N
    for (a0, a1), (b0, b1) in zip(ranges[:-1], ranges[1:]):
        assert a1 == b0 and a0 <= a1                      # contiguous, disjoint, in order
    sizes = [b - a for a, b in ranges]
    assert max(sizes) - min(sizes) <= 1                    # balanced
    assert sum(slab_words(N, a, b) for a, b in ranges) == slab_words(N, 0, N)
    with pytest.raises(ValueError):
        slab_range(world, world, N)


WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np
import torch.distributed as dist
import dxrvoxelizer_b200 as d
from dxrvoxelizer_b200 import meshes
from dxrvoxelizer_b200.sharding import slab_range, broadcast_mesh, gather_slabs
import oracle

dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=int(sys.argv[2]))
rank, world = dist.get_rank(), dist.get_world_size()
N = 40                                                   # 40 layers over 3 ranks: ragged slabs
src = meshes.icosphere(3, seed=5) if rank == 0 else None  # only rank 0 has the mesh
mesh, vb, ib = broadcast_mesh(src, 0)
assert mesh.num_triangles == 1280 and mesh.stride == 24
for mode in (0, 1):
    z0, z1 = slab_range(rank, world, N)
    mine = oracle.voxelize(mesh.vertices, mesh.indices, N, mode, z0=z0, z1=z1)["bits"]   # stand-in for the kernel
    full = gather_slabs(mine, N, world)
    want = oracle.voxelize(mesh.vertices, mesh.indices, N, mode)["bits"]
    assert full.shape == want.shape and np.array_equal(full, want), (rank, mode)
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


@pytest.mark.parametrize("world", [2, 3])
def test_broadcast_and_gather_over_gloo(world, tmp_path):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "port": port})
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(world)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, out
        assert "rank %d ok" % r in out


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_balanced_slabs_cover_grid_and_balance_the_dragon(world, assets):
    from dxrvoxelizer_b200.sharding import balanced_slabs
    m = assets("dragon.obj")
    N = 1024
    slabs = balanced_slabs(m, N, world)
    assert slabs[0][0] == 0 and slabs[-1][1] == N and len(slabs) == world
    assert all(a1 == b0 and a0 < a1 for (a0, a1), (b0, b1) in zip(slabs[:-1], slabs[1:]))
    if world > 1:
        # triangles per layer (same proxy the function uses): the heaviest rank must carry clearly less than
        # with equal slabs, where the few ranks owning the dragon's thin z range do all the work
        b = m.bound
        z = (m.vertices[:, 2][m.indices.reshape(-1, 3)] - b[2]) / b[3]
        layer = np.clip(((z.mean(1) + 1) * 0.5 * N).astype(int), 0, N - 1)
        hist = np.bincount(layer, minlength=N).astype(float)
        cost = 1.0 + 1.5 * hist / hist.mean()
        def worst(parts):
            return max(cost[a:b].sum() for a, b in parts)
        equal = [slab_range(r, world, N) for r in range(world)]
        assert worst(slabs) <= worst(equal) * (1.05 if world == 2 else 0.8)
        assert worst(slabs) <= 1.35 * cost.sum() / world


def test_proportional_slabs_cover_the_grid():
    from dxrvoxelizer_b200.sharding import proportional_slabs
    for N, w in ((1024, [53, 16.7, 16.7, 16.7, 25, 25, 25, 25]), (33, [1, 1]), (5, [1, 0, 3]), (1, [2, 2, 2, 2]), (64, [0, 0])):
        sl = proportional_slabs(N, w)
        assert sl[0][0] == 0 and sl[-1][1] == N and all(a[1] == b[0] for a, b in zip(sl, sl[1:])) and all(a <= b for a, b in sl)
    sl = proportional_slabs(1024, [53, 16.7, 16.7, 16.7, 25, 25, 25, 25])
    assert sl[0][1] - sl[0][0] > 2.5 * (sl[1][1] - sl[1][0])
