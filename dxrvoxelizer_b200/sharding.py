"""Multi-GPU host logic: z-slab sharding of one grid over the ranks of a torch.distributed job
(one process per GPU), mesh replication by broadcast, optional gather of the slabs.

The voxelization path shards without any collective in the data path: rays
run along x, a z-slab owns whole rays and its output is one contiguous byte range of the grid.  The
mesh (a few MB) is replicated once; every rank builds the identical LBVH (the radix sort is stable
and deterministic) and fills its own slab.  Slabs are gathered only when a full grid is requested:
  * gather="nccl"  all_gather_into_tensor of the device slabs (equal slabs) / all_gather (ragged);
  * gather="peer"  the owner exports its full-size grid through CUDA IPC, every rank's fill kernel
                   then stores straight into the owner's memory over NVLink (the gather is fused
                   into the 128-bit stores of k_trace_fill_columns; no collective at all).
torch is used for plumbing only (process group, device tensors as buffers).
"""
import numpy as np

from . import _lib as L
from .voxelizer import Mesh, Voxelizer


def slab_range(rank, world, N):
    """Layers [z0, z1) owned by `rank`: contiguous, disjoint, covering [0, N), sizes differ by <= 1."""
    if not (0 <= rank < world) or N < 1:
        raise ValueError("bad rank/world/N")
    return N * rank // world, N * (rank + 1) // world


def balanced_slabs(mesh, N, world, bound=None, compute_weight=1.2):
    """Cost-balanced z-slabs: [(z0, z1)] * world, contiguous and covering [0, N).

    Equal slabs are only balanced for a mesh that fills the grid evenly; a real mesh is thin along some
    axis (the dragon occupies a quarter of the z range), so the ranks owning its layers do all the
    crossing tests while the others only stream zeros.  The cost of layer z is modelled as
    1 (the stores) + compute_weight * t(z) / mean(t), t(z) = triangles whose z extent overlaps the layer,
    and the cut points split the cumulative cost evenly (compute_weight measured with tools/slab_balance.py:
    per-slab kernel times of the 2048^3 dragon on one B200).  Pure host code (numpy), deterministic: every
    rank computes the same partition from the replicated mesh."""
    if world < 1 or N < 1:
        raise ValueError("bad world/N")
    if world == 1:
        return [(0, N)]
    pos = mesh.vertices[:, :3].astype(np.float64)
    if bound is None:
        mn, mx = pos.min(0), pos.max(0)
        bound = np.concatenate([(mx + mn) / 2, [(mx - mn).max() / 2]])
    tri = mesh.indices.reshape(-1, 3)
    z = (pos[:, 2][tri] - bound[2]) / bound[3]                      # scene z of the three corners
    lo = np.clip(np.floor((z.min(1) + 1.0) * 0.5 * N - 0.5).astype(np.int64), 0, N - 1)
    hi = np.clip(np.ceil((z.max(1) + 1.0) * 0.5 * N - 0.5).astype(np.int64), 0, N - 1)
    diff = np.zeros(N + 1, np.float64)
    np.add.at(diff, lo, 1.0)
    np.add.at(diff, hi + 1, -1.0)
    t = np.cumsum(diff[:N])
    cost = 1.0 + (compute_weight * t / t.mean() if t.sum() > 0 else 0.0)
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    cuts = [0]
    for r in range(1, world):
        target = cum[-1] * r / world
        zc = int(np.searchsorted(cum, target))
        zc = max(cuts[-1] + 1, min(zc, N - (world - r)))             # every rank keeps at least one layer
        cuts.append(zc)
    cuts.append(N)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def slab_words(N, z0, z1):
    return (z1 - z0) * N * ((N + 31) // 32)


def broadcast_mesh(mesh, src=0, device=None, group=None):
    """Replicate rank `src`'s mesh on every rank.  Works with the gloo backend on CPU tensors (tests)
    and with NCCL on CUDA tensors (device = torch.device("cuda", k)): the vertex/index buffers then
    travel GPU-to-GPU over NVLink and are returned as device tensors as well.
    Returns (Mesh, vertex_tensor, index_tensor)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    dev = device if device is not None else torch.device("cpu")
    meta = torch.zeros(3, dtype=torch.int64, device=dev)
    if rank == src:
        meta = torch.tensor([mesh.num_vertices, mesh.stride, mesh.indices.size], dtype=torch.int64, device=dev)
    dist.broadcast(meta, src, group=group)
    nv, stride, ni = (int(v) for v in meta.tolist())
    if rank == src:
        vb = torch.from_numpy(mesh.vertex_bytes.copy()).to(dev)
        ib = torch.from_numpy(mesh.indices.view(np.int32).copy()).to(dev)
    else:
        vb = torch.empty(nv * stride, dtype=torch.uint8, device=dev)
        ib = torch.empty(ni, dtype=torch.int32, device=dev)
    dist.broadcast(vb, src, group=group)
    dist.broadcast(ib, src, group=group)
    out = Mesh(vb.cpu().numpy(), ib.cpu().numpy().view(np.uint32), stride)
    return out, vb, ib


def gather_slabs(local_slab, N, world, group=None):
    """all_gather of per-rank slabs (numpy uint32 [(z1-z0), N, P]) into the full grid, on every rank.
    Host-side variant used by the gloo tests and the ragged case."""
    import torch
    import torch.distributed as dist
    P = (N + 31) // 32
    sizes = [slab_words(N, *slab_range(r, world, N)) for r in range(world)]
    # all_gather wants equal sizes: ragged slabs (N % world != 0) are padded to the largest one
    biggest = max(sizes)
    bufs = [torch.empty(biggest, dtype=torch.int32) for _ in sizes]
    mine = torch.zeros(biggest, dtype=torch.int32)
    mine[: sizes[dist.get_rank(group)]] = torch.from_numpy(np.ascontiguousarray(local_slab).reshape(-1).view(np.int32))
    dist.all_gather(bufs, mine, group=group)
    return np.concatenate([b.numpy().view(np.uint32)[:s] for b, s in zip(bufs, sizes)]).reshape(N, N, P)


class ShardedVoxelizer:
    """One rank's share of a z-slab sharded voxelization (GPU path; needs NCCL + CUDA)."""

    def __init__(self, local_device, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = torch.device("cuda", local_device)
        self.vox = Voxelizer(local_device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.vox.set_stream(self.stream.cuda_stream)
        self._peer = None

    def close(self):
        if self._peer is not None and self._peer[0] is not None and self.rank != self._peer[2]:
            self.vox.ipc_close(self._peer[0])
        self.vox.close()

    def replicate_and_build(self, mesh, src=0, bound=None):
        """Broadcast the mesh over NCCL and build the LBVH from the device-resident copy."""
        m, vb, ib = broadcast_mesh(mesh, src, self.device, self.group)
        self.torch.cuda.synchronize(self.device)
        self._vb, self._ib, self.mesh = vb, ib, m
        self.vox.build_bvh_device(vb.data_ptr(), m.num_vertices, m.stride, ib.data_ptr(), m.indices.size, bound)
        return m

    def voxelize(self, N, mode=L.MODE_PARITY):
        z0, z1 = slab_range(self.rank, self.world, N)
        self.N, self.z0, self.z1 = N, z0, z1
        self.vox.voxelize(N, mode, z0, z1)

    def local_slab(self):
        return self.vox.fetch_bits()

    def gather_nccl(self):
        """Full grid on every rank as a device tensor (int32 view of the uint32 words)."""
        torch, dist = self.torch, self.dist
        N, P = self.N, (self.N + 31) // 32
        self.vox.synchronize()
        ptr, nbytes = self.vox.grid_device()
        sizes = [slab_words(N, *slab_range(r, self.world, N)) for r in range(self.world)]
        biggest = max(sizes)
        mine = torch.zeros(biggest, dtype=torch.int32, device=self.device)
        # wrap the context's grid without a copy through the CUDA array interface
        view = torch.as_tensor(_DevicePtr(ptr, sizes[self.rank]), device=self.device)
        mine[: sizes[self.rank]].copy_(view)
        gathered = torch.empty(biggest * self.world, dtype=torch.int32, device=self.device)
        dist.all_gather_into_tensor(gathered, mine, group=self.group)
        if len(set(sizes)) == 1:
            full = gathered
        else:  # ragged slabs were padded to the largest one
            full = torch.cat([gathered[r * biggest: r * biggest + sizes[r]] for r in range(self.world)])
        return full.view(N, N, P)

    def setup_peer_gather(self, N, owner=0):
        """Fused gather: rank `owner` exports its full-size grid (CUDA IPC); every rank aims its fill
        kernel at its slab inside that allocation.  Call once per (N, owner); then voxelize()."""
        full_bytes = slab_words(N, 0, N) * 4
        handle = [None]
        if self.rank == owner:
            h, base = self.vox.ipc_export_grid(full_bytes)
            handle[0] = h
        else:
            base = None
        self.dist.broadcast_object_list(handle, src=owner, group=self.group)
        if self.rank != owner:
            base = self.vox.ipc_open(handle[0])
        z0, z1 = slab_range(self.rank, self.world, N)
        off = slab_words(N, 0, z0) * 4
        self.vox.set_grid_target(base + off, slab_words(N, z0, z1) * 4)
        self._peer = (base if self.rank != owner else None, full_bytes, owner, base)
        return base


class _DevicePtr:
    """Minimal __cuda_array_interface__ wrapper around a raw device pointer (int32 elements)."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<i4", "data": (ptr, False), "version": 2}
