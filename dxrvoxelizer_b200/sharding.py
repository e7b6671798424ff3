"""Multi-GPU host logic: z-slab sharding of one grid over the ranks of a torch.distributed job
(one process per GPU), mesh replication by broadcast, optional gather of the slabs.

The voxelization path shards without any collective in the data path: rays
run along x, a z-slab owns whole rays and its output is one contiguous byte range of the grid.  The
mesh (a few MB) is replicated once; every rank builds the identical LBVH (the radix sort is stable
and deterministic) and fills its own slab.  Slabs are gathered only when a full grid is requested:
  * gather()       dxrv_gather_grid inside the library: ncclSend/Recv to a root, or one ncclBroadcast per slab;
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
    if N <= world:
        # more ranks than layers: one layer each, the rest get empty slabs (dxrv_voxelize accepts them)
        return [(min(r, N), min(r + 1, N)) for r in range(world)]
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


def proportional_slabs(N, weights):
    """[(z0, z1)] * len(weights): contiguous slabs covering [0, N) with layer counts proportional to `weights` (e.g. every
    rank's measured read-back bandwidth, when the slabs go back to the host and bytes bound the step).  Deterministic."""
    w = np.maximum(np.asarray(weights, np.float64), 0.0)
    if w.size < 1 or N < 1:
        raise ValueError("bad weights/N")
    if w.sum() <= 0:
        w = np.ones_like(w)
    cuts = np.floor(np.concatenate([[0.0], np.cumsum(w)]) / w.sum() * N + 0.5).astype(np.int64)
    cuts[0], cuts[-1] = 0, N
    cuts = np.maximum.accumulate(np.clip(cuts, 0, N))
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(w.size)]


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
    """One rank's share of a z-slab sharded voxelization: one process per GPU.  Everything on the data path goes
    through the C ABI (dxrv_comm_init / dxrv_bcast_mesh / dxrv_build_bvh_replicated / dxrv_voxelize /
    dxrv_gather_grid): NCCL lives inside libdxrv.so.  torch.distributed is used once, to ship the 128-byte NCCL
    unique id (any out-of-band channel would do), and by nothing else here."""

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
        uid = [Voxelizer.comm_unique_id() if self.rank == 0 else None]
        dist.broadcast_object_list(uid, src=0, group=group)
        self.vox.comm_init(uid[0], self.rank, self.world)
        self._peer = None
        self.slabs = None

    def close(self):
        if self._peer is not None and self._peer[0] is not None and self.rank != self._peer[2]:
            self.vox.ipc_close(self._peer[0])
        self.vox.close()

    def replicate_and_build(self, mesh, src=0, bound=None):
        """Broadcast rank src's host mesh over NCCL (inside the library) and build the LBVH on every rank.
        Returns (num_vertices, stride, num_indices)."""
        hdr = np.zeros(3, np.uint32)
        if self.rank == src:
            hdr[:] = (mesh.num_vertices, mesh.stride, mesh.indices.size)
        hdr = self.vox.bcast_u32(hdr, src)
        self.vox.bcast_mesh(mesh if self.rank == src else None, int(hdr[0]), int(hdr[1]), int(hdr[2]), src)
        self.vox.build_bvh_replicated(bound)
        self.mesh = mesh if self.rank == src else None
        self._src = src
        return tuple(int(x) for x in hdr)

    def plan_slabs(self, N, balanced=True):
        """[(z0, z1)] * world, identical on every rank: cost-balanced cuts computed by the rank that holds the host
        mesh and broadcast as world + 1 words (equal slabs when balanced=False or no host mesh exists)."""
        cuts = np.zeros(self.world + 1, np.uint32)
        if self.rank == getattr(self, "_src", 0):
            if balanced and getattr(self, "mesh", None) is not None:
                sl = balanced_slabs(self.mesh, N, self.world)
            else:
                sl = [slab_range(r, self.world, N) for r in range(self.world)]
            cuts[:] = [sl[0][0]] + [b for _, b in sl]
        cuts = self.vox.bcast_u32(cuts, getattr(self, "_src", 0))
        self.slabs = [(int(cuts[r]), int(cuts[r + 1])) for r in range(self.world)]
        return self.slabs

    def voxelize(self, N, mode=L.MODE_PARITY, balanced=False):
        if self.slabs is None or getattr(self, "N", None) != N or getattr(self, "_balanced", None) != balanced:
            self.plan_slabs(N, balanced)
            self._balanced = balanced
        self.N = N
        self.z0, self.z1 = self.slabs[self.rank]
        self.vox.voxelize(N, mode, self.z0, self.z1)     # (an empty slab -- more ranks than layers -- computes nothing)

    def local_slab(self):
        return self.vox.fetch_bits()

    def gather(self, root=-1):
        """dxrv_gather_grid: the full grid on `root` (ncclSend/Recv) or on every rank (root < 0); returns it as a
        numpy array on the ranks that hold it, None elsewhere."""
        self.vox.gather_grid(root)
        if root < 0 or root == self.rank:
            return self.vox.fetch_full_grid(self.N)
        return None

    def gather_nccl(self):
        """Full grid on every rank as a device tensor (int32 view of the uint32 words)."""
        self.vox.gather_grid(-1)
        ptr, nbytes = self.vox.full_grid_device()
        self.vox.synchronize()
        t = self.torch.as_tensor(_DevicePtr(ptr, nbytes // 4), device=self.device)
        return t.view(self.N, self.N, (self.N + 31) // 32)

    def setup_peer_gather(self, N, owner=0):
        """Fused gather across processes: rank `owner` exports its full-size grid (CUDA IPC); every rank aims its fill
        kernel at its slab inside that allocation (the 128-bit stores go over NVLink).  Returns False -- and changes
        nothing -- when some slab's byte offset is not 16-byte aligned (odd N with an odd cut: N = 33, 63, ...): use
        gather() then."""
        self.plan_slabs(N, balanced=False)
        self._balanced = False
        self.N = N
        if any((slab_words(N, 0, a) * 4) % 16 for a, b in self.slabs if b > a):
            return False
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
        z0, z1 = self.slabs[self.rank]
        if z1 > z0:
            self.vox.set_grid_target(base + slab_words(N, 0, z0) * 4, slab_words(N, z0, z1) * 4)
        self._peer = (base if self.rank != owner else None, full_bytes, owner, base)
        return True


class _DevicePtr:
    """Minimal __cuda_array_interface__ wrapper around a raw device pointer (int32 elements)."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<i4", "data": (ptr, False), "version": 2}
