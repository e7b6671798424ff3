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
sv = ShardedVoxelizer(rank)
mesh = d.load_obj(d.asset_path("bunny.obj")) if rank == 0 else None
m = sv.replicate_and_build(mesh)
N = 96
want = oracle.voxelize(m.vertices, m.indices, N, 1)["bits"]
sv.voxelize(N)
full = sv.gather_nccl().cpu().numpy().view(np.uint32)
assert np.array_equal(full, want), "nccl gather"
# fused gather: every rank's fill kernel stores straight into rank 0's grid over NVLink
sv.setup_peer_gather(N, owner=0)
dist.barrier()
sv.voxelize(N)
sv.vox.synchronize()
dist.barrier()
if rank == 0:
    sv.vox._shape = (N, N, (N + 31) // 32)
    got = np.empty(sv.vox._shape, np.uint32)
    import ctypes
    torch.cuda.synchronize()
    base = sv._peer[3]
    t = torch.as_tensor(d.sharding._DevicePtr(base, N * N * ((N + 31) // 32)), device="cuda:0")
    got = t.cpu().numpy().view(np.uint32).reshape(N, N, -1)
    assert np.array_equal(got, want), "peer gather"
dist.barrier()
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
