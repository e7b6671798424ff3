"""Multi-GPU path on real devices: z-slab sharding over NCCL, both gather flavours.  Needs >= 2 GPUs
(skipped on a 1-GPU box; the driver's scaling run and `gpurun --gpus 2` exercise it)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
import dxrvoxelizer_b200 as d
from dxrvoxelizer_b200.sharding import ShardedVoxelizer
import oracle
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
sv = ShardedVoxelizer(rank)                      # dxrv_comm_init inside
mesh = d.load_obj(d.asset_path("bunny.obj")) if rank == 0 else None
nv, stride, ni = sv.replicate_and_build(mesh)    # dxrv_bcast_u32 + dxrv_bcast_mesh + dxrv_build_bvh_replicated
ref_mesh = d.load_obj(d.asset_path("bunny.obj"))
assert (nv, stride, ni) == (ref_mesh.num_vertices, ref_mesh.stride, ref_mesh.indices.size)
for N, mode, balanced in ((96, d.MODE_PARITY, False), (96, d.MODE_PARITY, True), (100, d.MODE_SHADER, True), (33, d.MODE_PARITY, False)):
    want = oracle.voxelize(ref_mesh.vertices, ref_mesh.indices, N, mode)["bits"]
    sv.voxelize(N, mode, balanced=balanced)
    full = sv.gather(-1)                         # dxrv_gather_grid: every rank gets the grid (one ncclBroadcast per slab)
    assert np.array_equal(full, want), ("all-gather", N, mode, balanced)
    sv.voxelize(N, mode, balanced=balanced)
    full = sv.gather(world - 1)                  # ncclSend / ncclRecv to the last rank
    assert (full is None) == (rank != world - 1)
    if full is not None:
        assert np.array_equal(full, want), ("gather to root", N, mode, balanced)
    t = sv.gather_nccl()
    assert np.array_equal(t.cpu().numpy().view(np.uint32), want), ("device view", N)
# more ranks than layers: ranks beyond N get empty slabs and still take part in the gather
want = oracle.voxelize(ref_mesh.vertices, ref_mesh.indices, 1, 1)["bits"]
sv.slabs = None
sv.voxelize(1, d.MODE_PARITY, balanced=True)
assert np.array_equal(sv.gather(-1), want)
# fused gather across processes: every rank's fill kernel stores straight into rank 0's grid over NVLink
N = 96
want = oracle.voxelize(ref_mesh.vertices, ref_mesh.indices, N, 1)["bits"]
assert sv.setup_peer_gather(N, owner=0)
dist.barrier()
sv.vox.voxelize(N, d.MODE_PARITY, *sv.slabs[rank])
sv.vox.synchronize()
dist.barrier()
if rank == 0:
    torch.cuda.synchronize()
    base = sv._peer[3]
    t = torch.as_tensor(d.sharding._DevicePtr(base, N * N * ((N + 31) // 32)), device="cuda:0")
    got = t.cpu().numpy().view(np.uint32).reshape(N, N, -1)
    assert np.array_equal(got, want), "peer gather"
dist.barrier()
sv2 = ShardedVoxelizer(rank)
sv2.replicate_and_build(mesh)
assert sv2.setup_peer_gather(35, owner=0) is False      # 280-byte layers, cut at z = 17: unaligned slab offset -> caller falls back to gather()
sv2.close()
sv.close()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_zslab_sharding_two_gpus(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout


def test_cli_two_gpus_matches_one(tmp_path):
    """The C++ host class with SetGpuCount(2): cost-balanced z-slabs on two contexts, gathered on the host --
    the grid dump must equal the single-GPU one byte for byte (and `-gpus 1` is checked against the oracle
    in test_gpu_voxelize.py::test_cli_matches_oracle)."""
    import numpy as np
    import torch
    import dxrvoxelizer_b200 as d
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    exe = os.path.join(ROOT, "dxrvoxelizer_b200", "dxrvoxelizer")
    outs = []
    for gpus in (1, 2):
        out = tmp_path / ("grid%d.bin" % gpus)
        r = subprocess.run([exe, "-mesh", d.asset_path("dragon.obj"), "-grid", "256", "-mode", "parity", "-gpus", str(gpus), "-out", str(out)],
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        outs.append(np.fromfile(out, np.uint32))
    assert outs[0].size == 256 * 256 * 8 and np.array_equal(outs[0], outs[1])


def test_c_abi_one_process_two_gpus(oracle_mod=None):
    """The multi-GPU entry points for ONE process driving several GPUs, through ctypes: dxrv_comm_init_all, grouped
    dxrv_bcast_mesh, dxrv_build_bvh_replicated, slabs, grouped dxrv_gather_grid; then the fused form
    (dxrv_share_grid_target: the second GPU's fill kernel stores into the first GPU's grid over NVLink)."""
    import ctypes
    import numpy as np
    import torch
    import dxrvoxelizer_b200 as d
    from dxrvoxelizer_b200 import _lib as L
    import oracle
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    lib = L.lib()
    a, b = d.Voxelizer(0), d.Voxelizer(1)
    arr = (ctypes.c_void_p * 2)(a._h, b._h)
    assert lib.dxrv_comm_init_all(arr, 2) == 0, lib.dxrv_last_error(a._h)
    m = d.load_obj(d.asset_path("dragon.obj"))
    assert lib.dxrv_group_begin() == 0
    a.bcast_mesh(m, m.num_vertices, m.stride, m.indices.size, 0)
    b.bcast_mesh(None, m.num_vertices, m.stride, m.indices.size, 0)
    assert lib.dxrv_group_end() == 0
    a.build_bvh_replicated(); b.build_bvh_replicated()
    N = 128
    want = oracle.voxelize(m.vertices, m.indices, N, 1)["bits"]
    a.voxelize(N, d.MODE_PARITY, 0, 70); b.voxelize(N, d.MODE_PARITY, 70, N)
    assert lib.dxrv_group_begin() == 0
    a.gather_grid(0, [(0, 70), (70, N)]); b.gather_grid(0, [(0, 70), (70, N)])
    assert lib.dxrv_group_end() == 0
    assert np.array_equal(a.fetch_full_grid(N), want)
    # fused: both slabs land in a's full grid, no collective
    a.share_grid_target(a, N, 0, 64); b.share_grid_target(a, N, 64, N)
    a.voxelize(N, d.MODE_PARITY, 0, 64); b.voxelize(N, d.MODE_PARITY, 64, N)
    b.synchronize(); a.synchronize()
    assert np.array_equal(a.fetch_full_grid(N), want)
    b.close(); a.close()
