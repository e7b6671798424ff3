"""Python face of the voxelization path: a thin mirror of the C++ host class
(csrc/voxelizer_host.h), which itself mirrors the reference's ``Voxelizer::Init`` / ``voxelize``
(Content/Voxelizer.h:16-22,90).  Everything goes through the C ABI (include/dxrv.h); numpy is only
used to hold host buffers.
"""
import ctypes

import numpy as np

from . import _lib as L


class Mesh:
    """Output of the OBJ loader (== XUSG::ObjLoader::Import(file, true, true))."""

    def __init__(self, vertices, indices, stride, aabb=None, bound=None):
        self.vertex_bytes = np.ascontiguousarray(vertices).view(np.uint8).reshape(-1)
        self.indices = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
        self.stride = int(stride)
        self.aabb = aabb
        self.bound = bound

    @property
    def num_vertices(self):
        return self.vertex_bytes.size // self.stride

    @property
    def num_triangles(self):
        return self.indices.size // 3

    @property
    def vertices(self):
        """float32 view [numVerts, stride/4]: columns 0..2 position, 3..5 normal."""
        return self.vertex_bytes.view(np.float32).reshape(self.num_vertices, self.stride // 4)

    @staticmethod
    def from_arrays(positions, triangles, normals=None):
        pos = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
        if normals is None:
            vb = pos
        else:
            vb = np.concatenate([pos, np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)], axis=1)
        vb = np.ascontiguousarray(vb)
        return Mesh(vb, triangles, vb.shape[1] * 4)


def load_obj(path=None, text=None):
    """dxrv_obj_load(path), or dxrv_obj_parse of OBJ text (bytes) already in memory."""
    lib = L.lib()
    h = ctypes.c_void_p()
    if text is not None:
        data = bytes(text)
        rc = lib.dxrv_obj_parse(data, len(data), ctypes.byref(h))
    else:
        rc = lib.dxrv_obj_load(str(path).encode(), ctypes.byref(h))
    if rc != L.OK:
        raise L.DxrvError(rc, lib.dxrv_last_error(None).decode())
    try:
        nv, ni, st = lib.dxrv_obj_num_vertices(h), lib.dxrv_obj_num_indices(h), lib.dxrv_obj_vertex_stride(h)
        vb = np.ctypeslib.as_array(ctypes.cast(lib.dxrv_obj_vertices(h), ctypes.POINTER(ctypes.c_uint8)), (nv * st,)).copy() \
            if nv else np.zeros(0, np.uint8)
        ib = np.ctypeslib.as_array(ctypes.cast(lib.dxrv_obj_indices(h), ctypes.POINTER(ctypes.c_uint32)), (ni,)).copy() \
            if ni else np.zeros(0, np.uint32)
        aabb = np.zeros(6, np.float32)
        bound = np.zeros(4, np.float32)
        lib.dxrv_obj_aabb(h, aabb.ctypes.data)
        lib.dxrv_obj_bound(h, bound.ctypes.data)
    finally:
        lib.dxrv_obj_free(h)
    return Mesh(vb, ib, st, aabb, bound)


def default_view(bound, width=1280, height=720, pos_scale=None):
    """(screenToLocal[4,4], eye[3], light[3]) of the reference's camera for this bound (dxrv_default_view)."""
    lib = L.lib()
    b = np.ascontiguousarray(bound, dtype=np.float32)
    ps = None if pos_scale is None else np.ascontiguousarray(pos_scale, dtype=np.float32)
    m, eye, light = np.zeros(16, np.float32), np.zeros(3, np.float32), np.zeros(3, np.float32)
    rc = lib.dxrv_default_view(b.ctypes.data, None if ps is None else ps.ctypes.data, width, height,
                               m.ctypes.data, eye.ctypes.data, light.ctypes.data)
    if rc != L.OK:
        raise L.DxrvError(rc, "dxrv_default_view: invalid arguments")
    return m.reshape(4, 4), eye, light


def save_image(path, rgba, comp=3):
    """dxrv_save_image: the reference's SaveImage (PNG screenshot) for an (H, W, 4) uint8 image, e.g. Voxelizer.render_view's."""
    img = np.ascontiguousarray(rgba, dtype=np.uint8)
    if img.ndim != 3 or img.shape[2] != 4:
        raise ValueError("rgba must be (height, width, 4) uint8")
    rc = L.lib().dxrv_save_image(str(path).encode(), img.ctypes.data, img.shape[1], img.shape[0], img.shape[1] * 4, comp)
    if rc != L.OK:
        raise L.DxrvError(rc, "dxrv_save_image: cannot write %s" % path)


class Voxelizer:
    """One context = one GPU + one stream.  Not thread-safe (same as the C ABI)."""

    MODE_SHADER = L.MODE_SHADER
    MODE_PARITY = L.MODE_PARITY

    def __init__(self, device=0):
        self._lib = L.lib()
        h = ctypes.c_void_p()
        rc = self._lib.dxrv_create(ctypes.byref(h), int(device))
        if rc != L.OK:
            raise L.DxrvError(rc, self._lib.dxrv_last_error(None).decode())
        self._h = h
        self.device = int(device)
        self._shape = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.dxrv_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != L.OK:
            raise L.DxrvError(rc, self._lib.dxrv_last_error(self._h).decode())

    # -- Voxelizer::Init's device half: createVB/createIB + buildAccelerationStructures ----------
    def build_bvh(self, mesh, bound=None):
        b = None if bound is None else np.ascontiguousarray(bound, dtype=np.float32)
        self._mesh = mesh  # keep host arrays alive
        self._check(self._lib.dxrv_build_bvh(self._h, mesh.vertex_bytes.ctypes.data, mesh.num_vertices, mesh.stride,
                                             mesh.indices.ctypes.data, mesh.indices.size,
                                             None if b is None else b.ctypes.data))

    def build_bvh_host_ptr(self, vptr, num_verts, stride, iptr, num_indices, bound=None):
        b = None if bound is None else np.ascontiguousarray(bound, dtype=np.float32)
        self._check(self._lib.dxrv_build_bvh(self._h, vptr, num_verts, stride, iptr, num_indices,
                                             None if b is None else b.ctypes.data))

    def build_bvh_device(self, d_vertices, num_verts, stride, d_indices, num_indices, bound=None):
        b = None if bound is None else np.ascontiguousarray(bound, dtype=np.float32)
        self._check(self._lib.dxrv_build_bvh_device(self._h, d_vertices, num_verts, stride, d_indices, num_indices,
                                                    None if b is None else b.ctypes.data))

    def bound(self):
        out = np.zeros(4, np.float32)
        self._check(self._lib.dxrv_get_bound(self._h, out.ctypes.data))
        return out

    # -- Voxelizer::voxelize ------------------------------------------------------------------------
    def voxelize(self, N, mode=L.MODE_SHADER, z0=0, z1=None, texels=False):
        """mode defaults to MODE_SHADER, the reference's function; MODE_PARITY is the opt-in fast path."""
        z1 = N if z1 is None else z1
        self._check(self._lib.dxrv_voxelize(self._h, N, mode | (L.EMIT_TEXELS if texels else 0), z0, z1))
        self._shape = (z1 - z0, N, (N + 31) // 32)
        self._N = N

    def set_read_back(self, transport):
        """dxrv_set_read_back: L.READ_BACK_AUTO / _DENSE (pipelined PCIe copy) / _SPARSE (compact blob + host expansion)."""
        self._check(self._lib.dxrv_set_read_back(self._h, transport))

    def voxelize_to_host(self, N, mode, z0, z1, ptr, nbytes, chunks=8):
        """dxrv_voxelize_to_host: voxelize + read-back of the dense bit grid into host memory at ptr (transport: set_read_back)."""
        self._check(self._lib.dxrv_voxelize_to_host(self._h, N, mode, z0, z1, ptr, nbytes, chunks))
        self._shape = (z1 - z0, N, (N + 31) // 32)
        self._N = N

    def voxelize_mesh_to_host(self, vptr, num_verts, stride, iptr, num_indices, N, mode, z0, z1, ptr, nbytes, bound=None, chunks=8):
        """dxrv_voxelize_mesh_to_host: dxrv_build_bvh (host arrays at vptr / iptr) + dxrv_voxelize_to_host as one call."""
        b = None if bound is None else np.ascontiguousarray(bound, dtype=np.float32)
        self._check(self._lib.dxrv_voxelize_mesh_to_host(self._h, vptr, num_verts, stride, iptr, num_indices, None if b is None else b.ctypes.data,
                                                         N, mode, z0, z1, ptr, nbytes, chunks))
        self._shape = (z1 - z0, N, (N + 31) // 32)
        self._N = N

    def fetch_bits(self, out=None):
        """uint32[(z1-z0), N, P] in the DXRV_FORMAT_BITS layout."""
        if out is None:
            out = np.empty(self._shape, np.uint32)
        self._check(self._lib.dxrv_fetch_grid(self._h, out.ctypes.data, out.nbytes, L.FORMAT_BITS))
        return out

    def fetch_sparse_into(self, ptr, capacity):
        """dxrv_fetch_grid_sparse into host memory at ptr; returns the bytes written."""
        n = ctypes.c_size_t()
        self._check(self._lib.dxrv_fetch_grid_sparse(self._h, ptr, capacity, ctypes.byref(n)))
        return int(n.value)

    def fetch_sparse(self):
        """The slab as a DXRV_FORMAT_SPARSE_BRICKS blob (numpy uint8)."""
        cap = 64 + self._shape[0] * self._shape[1] * self._shape[2] * 4 + (1 << 20) + self._shape[0] * self._shape[1] * self._shape[2] // 16
        buf = np.empty(cap, np.uint8)
        return buf[: self.fetch_sparse_into(buf.ctypes.data, cap)].copy()

    def fetch_into(self, ptr, nbytes, fmt=L.FORMAT_BITS):
        self._check(self._lib.dxrv_fetch_grid(self._h, ptr, nbytes, fmt))

    def fetch_u8(self):
        out = np.empty((self._shape[0], self._N, self._N), np.uint8)
        self._check(self._lib.dxrv_fetch_grid(self._h, out.ctypes.data, out.nbytes, L.FORMAT_U8))
        return out

    def fetch_texels(self):
        out = np.empty((self._shape[0], self._N, self._N), np.uint32)
        self._check(self._lib.dxrv_fetch_grid(self._h, out.ctypes.data, out.nbytes, L.FORMAT_R10G10B10A2))
        return out

    def grid_device(self):
        p, n = ctypes.c_void_p(), ctypes.c_size_t()
        self._check(self._lib.dxrv_grid_device(self._h, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def set_grid_target(self, d_ptr, nbytes):
        self._check(self._lib.dxrv_set_grid_target(self._h, d_ptr, nbytes))

    def build_mips(self):
        """Occupancy pyramid (OR of 2x2x2 children); returns the number of levels incl. level 0."""
        n = ctypes.c_uint32()
        self._check(self._lib.dxrv_build_mips(self._h, ctypes.byref(n)))
        return int(n.value)

    def fetch_mip(self, level):
        layers, N = self._shape[0] >> level, self._N >> level
        out = np.empty((layers, N, (N + 31) // 32), np.uint32)
        self._check(self._lib.dxrv_fetch_mip(self._h, level, out.ctypes.data, out.nbytes))
        return out

    def render_view(self, width, height, screen_to_local, eye, light):
        """RGBA8 image [height, width, 4] of the reference's viewer pass over the last (full) grid."""
        m = np.ascontiguousarray(screen_to_local, dtype=np.float32).reshape(16)
        e = np.ascontiguousarray(eye, dtype=np.float32)
        l = np.ascontiguousarray(light, dtype=np.float32)
        out = np.empty((height, width, 4), np.uint8)
        self._check(self._lib.dxrv_render_view(self._h, width, height, m.ctypes.data, e.ctypes.data, l.ctypes.data,
                                               out.ctypes.data, out.nbytes))
        return out

    def count_inside(self):
        c = ctypes.c_uint64()
        self._check(self._lib.dxrv_count_inside(self._h, ctypes.byref(c)))
        return int(c.value)

    def info(self, what):
        v = ctypes.c_uint64()
        self._check(self._lib.dxrv_get_info(self._h, what, ctypes.byref(v)))
        return int(v.value)

    def set_profiling(self, enable=True):
        self._check(self._lib.dxrv_set_profiling(self._h, 1 if enable else 0))

    def set_stream(self, cuda_stream):
        self._check(self._lib.dxrv_set_stream(self._h, cuda_stream))

    def synchronize(self):
        self._check(self._lib.dxrv_synchronize(self._h))

    def debug_read(self, what, dtype, count):
        out = np.empty(count, dtype)
        self._check(self._lib.dxrv_debug_read(self._h, what, out.ctypes.data, out.nbytes))
        return out

    def debug_sort_pairs(self, keys, values):
        k = np.ascontiguousarray(keys, dtype=np.uint32).copy()
        v = np.ascontiguousarray(values, dtype=np.uint32).copy()
        self._check(self._lib.dxrv_debug_sort_pairs(self._h, k.ctypes.data, v.ctypes.data, k.size))
        return k, v

    # -- multi-GPU (include/dxrv.h "multi-GPU" section; NCCL inside the library) -----------------------
    @staticmethod
    def comm_unique_id():
        buf = (ctypes.c_ubyte * 128)()
        rc = L.lib().dxrv_comm_get_unique_id(buf)
        if rc != L.OK:
            raise L.DxrvError(rc, L.lib().dxrv_last_error(None).decode())
        return bytes(buf)

    def comm_init(self, unique_id, rank, world):
        buf = (ctypes.c_ubyte * 128).from_buffer_copy(unique_id)
        self._check(self._lib.dxrv_comm_init(self._h, buf, rank, world))

    def comm_destroy(self):
        self._check(self._lib.dxrv_comm_destroy(self._h))

    def bcast_u32(self, values, root=0):
        a = np.ascontiguousarray(values, dtype=np.uint32).copy()
        self._check(self._lib.dxrv_bcast_u32(self._h, a.ctypes.data, a.size, root))
        return a

    def bcast_mesh(self, mesh, num_verts, stride, num_indices, root=0):
        """mesh: the host Mesh on the root, None elsewhere; the three sizes on every rank."""
        v = mesh.vertex_bytes.ctypes.data if mesh is not None else None
        i = mesh.indices.ctypes.data if mesh is not None else None
        self._mesh = mesh
        self._check(self._lib.dxrv_bcast_mesh(self._h, v, num_verts, stride, i, num_indices, root))

    def build_bvh_replicated(self, bound=None):
        b = None if bound is None else np.ascontiguousarray(bound, dtype=np.float32)
        self._check(self._lib.dxrv_build_bvh_replicated(self._h, None if b is None else b.ctypes.data))

    def gather_grid(self, root=-1, slabs=None):
        """slabs: [(z0, z1)] of every rank when the caller knows them (no internal exchange: usable inside a group)."""
        if slabs is None:
            self._check(self._lib.dxrv_gather_grid(self._h, root))
        else:
            t = np.ascontiguousarray(slabs, dtype=np.uint32).reshape(-1)
            self._check(self._lib.dxrv_gather_grid_slabs(self._h, root, t.ctypes.data))

    def full_grid_device(self):
        p, n = ctypes.c_void_p(), ctypes.c_size_t()
        self._check(self._lib.dxrv_full_grid_device(self._h, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def fetch_full_grid(self, N):
        out = np.empty((N, N, (N + 31) // 32), np.uint32)
        self._check(self._lib.dxrv_fetch_full_grid(self._h, out.ctypes.data, out.nbytes))
        return out

    def share_grid_target(self, owner, N, z0, z1):
        self._check(self._lib.dxrv_share_grid_target(self._h, owner._h, N, z0, z1))

    def ipc_export_grid(self, full_bytes):
        handle = (ctypes.c_ubyte * 64)()
        p = ctypes.c_void_p()
        self._check(self._lib.dxrv_ipc_export_grid(self._h, full_bytes, handle, ctypes.byref(p)))
        return bytes(handle), p.value

    def ipc_open(self, handle):
        buf = (ctypes.c_ubyte * 64).from_buffer_copy(handle)
        p = ctypes.c_void_p()
        self._check(self._lib.dxrv_ipc_open(self._h, buf, ctypes.byref(p)))
        return p.value

    def ipc_close(self, d_ptr):
        self._check(self._lib.dxrv_ipc_close(self._h, d_ptr))


def voxelize_obj_batch(voxelizers, paths, N, mode=L.MODE_SHADER, out=None, out_ptr=None, loader_threads=0, fetch=True):
    """dxrv_voxelize_obj_batch: OBJ files -> one N^3 bit grid each, as a pipeline inside the library (loader threads parse,
    every Voxelizer of `voxelizers` -- distinct contexts, typically 4 per GPU -- takes meshes s, s + len(voxelizers), ...).
    The grids land in `out` (uint32[len(paths), N, N, ceil(N/32)], allocated here when None) or at the raw address
    `out_ptr` (e.g. pinned memory); fetch=False voxelizes only.  Returns (out or None, triangles per mesh)."""
    lib = L.lib()
    n = len(paths)
    P = (N + 31) // 32
    grid_bytes = N * N * P * 4
    if fetch and out_ptr is None:
        if out is None:
            out = np.empty((n, N, N, P), np.uint32)
        if out.nbytes != n * grid_bytes or not out.flags["C_CONTIGUOUS"]:
            raise ValueError("out must be a C-contiguous array of len(paths) * N * N * ceil(N/32) * 4 bytes")
        out_ptr = out.ctypes.data
    handles = (ctypes.c_void_p * len(voxelizers))(*[v._h for v in voxelizers])
    cpaths = (ctypes.c_char_p * max(1, n))(*[str(p).encode() for p in paths])
    tris = np.zeros(max(1, n), np.uint32)
    rc = lib.dxrv_voxelize_obj_batch(handles, len(voxelizers), cpaths, n, N, mode, out_ptr if fetch else None,
                                     grid_bytes if fetch else 0, loader_threads, tris.ctypes.data)
    if rc != L.OK:
        raise L.DxrvError(rc, lib.dxrv_last_error(None).decode())
    for v in voxelizers:                       # every context holds the last grid it voxelized
        v._shape, v._N = (N, N, P), N
    return (out if fetch else None), tris[:n]


def sparse_decode(blob):
    """DXRV_FORMAT_SPARSE_BRICKS blob -> dense uint32[(z1-z0), N, P] (dxrv_sparse_decode, host code)."""
    b = np.ascontiguousarray(blob, dtype=np.uint8)
    h = b[:64].view(np.uint32)
    N, z0, z1, P = int(h[2]), int(h[3]), int(h[4]), int(h[5])
    out = np.empty((z1 - z0, N, P), np.uint32)
    rc = L.lib().dxrv_sparse_decode(b.ctypes.data, b.size, out.ctypes.data, out.nbytes)
    if rc != L.OK:
        raise L.DxrvError(rc, "dxrv_sparse_decode: malformed blob")
    return out


def sparse_encode(bits, N, z0=0):
    """dense uint32[layers, N, P] slab starting at layer z0 -> DXRV_FORMAT_SPARSE_BRICKS blob (dxrv_sparse_encode, host code;
    the same bytes the device encoder writes)."""
    g = np.ascontiguousarray(bits, dtype=np.uint32)
    layers = g.shape[0]
    n = ctypes.c_size_t()
    lib = L.lib()
    rc = lib.dxrv_sparse_encode(g.ctypes.data, g.nbytes, N, z0, z0 + layers, None, 0, ctypes.byref(n))   # size query
    if n.value == 0:
        raise L.DxrvError(rc, "dxrv_sparse_encode: invalid arguments")
    blob = np.empty(n.value, np.uint8)
    rc = lib.dxrv_sparse_encode(g.ctypes.data, g.nbytes, N, z0, z0 + layers, blob.ctypes.data, blob.size, ctypes.byref(n))
    if rc != L.OK:
        raise L.DxrvError(rc, "dxrv_sparse_encode failed")
    return blob


def unpack_bits(bits, N):
    """uint32[..., P] -> uint8[..., N] occupancy (bit x&31 of word x>>5)."""
    b = np.unpackbits(np.ascontiguousarray(bits).view(np.uint8), axis=-1, bitorder="little")
    return b[..., :N]
