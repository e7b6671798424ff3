"""ctypes binding of libdxrv.so (the C ABI declared in include/dxrv.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C dxrvoxelizer_b200/csrc``.
There is no Python or CPU fallback: if the shared library is missing, loading fails loudly.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdxrv.so")

OK = 0
ERR_INVALID_ARG, ERR_CUDA, ERR_NO_BVH, ERR_NO_GRID, ERR_IO, ERR_OOM, ERR_UNSUPPORTED = -1, -2, -3, -4, -5, -6, -7
MODE_SHADER, MODE_PARITY, EMIT_TEXELS = 0, 1, 0x100
FORMAT_BITS, FORMAT_U8, FORMAT_R10G10B10A2 = 0, 1, 2
READ_BACK_AUTO, READ_BACK_DENSE, READ_BACK_SPARSE = 0, 1, 2
INFO_NUM_TRIANGLES, INFO_NUM_NODES, INFO_KERNEL_LAUNCHES, INFO_CROSSINGS, INFO_SM_COUNT = 0, 1, 2, 3, 4
INFO_LAST_WALK_NS, INFO_LAST_FILL_NS, INFO_LAST_BUILD_NS, INFO_LAST_SORT_NS, INFO_LAST_D2H_BYTES = 5, 6, 7, 8, 9
DBG_MORTON_SORTED, DBG_PRIM_SORTED, DBG_NODES, DBG_TRIS, DBG_ROOT_BOX, DBG_BINS_STATE = 0, 1, 2, 3, 5, 6

_c = ctypes
_vp, _u32, _u64, _sz, _int = _c.c_void_p, _c.c_uint32, _c.c_uint64, _c.c_size_t, _c.c_int

# name -> (restype, argtypes); every symbol include/dxrv.h declares
SIGNATURES = {
    "dxrv_create": (_int, [_c.POINTER(_vp), _int]),
    "dxrv_destroy": (None, [_vp]),
    "dxrv_last_error": (_c.c_char_p, [_vp]),
    "dxrv_set_stream": (_int, [_vp, _vp]),
    "dxrv_synchronize": (_int, [_vp]),
    "dxrv_build_bvh": (_int, [_vp, _vp, _u32, _u32, _vp, _u32, _vp]),
    "dxrv_build_bvh_device": (_int, [_vp, _vp, _u32, _u32, _vp, _u32, _vp]),
    "dxrv_get_bound": (_int, [_vp, _vp]),
    "dxrv_voxelize": (_int, [_vp, _u32, _u32, _u32, _u32]),
    "dxrv_fetch_grid": (_int, [_vp, _vp, _sz, _u32]),
    "dxrv_fetch_grid_sparse": (_int, [_vp, _vp, _sz, _c.POINTER(_sz)]),
    "dxrv_sparse_decode": (_int, [_vp, _sz, _vp, _sz]),
    "dxrv_sparse_encode": (_int, [_vp, _sz, _u32, _u32, _u32, _vp, _sz, _c.POINTER(_sz)]),
    "dxrv_voxelize_to_host": (_int, [_vp, _u32, _u32, _u32, _u32, _vp, _sz, _u32]),
    "dxrv_voxelize_mesh_to_host": (_int, [_vp, _vp, _u32, _u32, _vp, _u32, _vp, _u32, _u32, _u32, _u32, _vp, _sz, _u32]),
    "dxrv_set_read_back": (_int, [_vp, _u32]),
    "dxrv_grid_device": (_int, [_vp, _c.POINTER(_vp), _c.POINTER(_sz)]),
    "dxrv_set_grid_target": (_int, [_vp, _vp, _sz]),
    "dxrv_count_inside": (_int, [_vp, _c.POINTER(_u64)]),
    "dxrv_default_view": (_int, [_vp, _vp, _u32, _u32, _vp, _vp, _vp]),
    "dxrv_render_view": (_int, [_vp, _u32, _u32, _vp, _vp, _vp, _vp, _sz]),
    "dxrv_save_image": (_int, [_c.c_char_p, _vp, _u32, _u32, _u32, _u32]),
    "dxrv_build_mips": (_int, [_vp, _c.POINTER(_u32)]),
    "dxrv_fetch_mip": (_int, [_vp, _u32, _vp, _sz]),
    "dxrv_get_info": (_int, [_vp, _u32, _c.POINTER(_u64)]),
    "dxrv_set_profiling": (_int, [_vp, _int]),
    "dxrv_debug_read": (_int, [_vp, _u32, _vp, _sz]),
    "dxrv_debug_sort_pairs": (_int, [_vp, _vp, _vp, _u32]),
    "dxrv_obj_load": (_int, [_c.c_char_p, _c.POINTER(_vp)]),
    "dxrv_obj_parse": (_int, [_c.c_char_p, _sz, _c.POINTER(_vp)]),
    "dxrv_obj_free": (None, [_vp]),
    "dxrv_obj_num_vertices": (_u32, [_vp]),
    "dxrv_obj_num_indices": (_u32, [_vp]),
    "dxrv_obj_vertex_stride": (_u32, [_vp]),
    "dxrv_obj_vertices": (_vp, [_vp]),
    "dxrv_obj_indices": (_vp, [_vp]),
    "dxrv_obj_aabb": (None, [_vp, _vp]),
    "dxrv_obj_bound": (None, [_vp, _vp]),
    "dxrv_voxelize_obj_batch": (_int, [_c.POINTER(_vp), _u32, _c.POINTER(_c.c_char_p), _u32, _u32, _u32, _vp, _sz, _u32, _vp]),
    "dxrv_host_alloc": (_vp, [_sz]),
    "dxrv_host_free": (None, [_vp]),
    "dxrv_ipc_export_grid": (_int, [_vp, _sz, _vp, _c.POINTER(_vp)]),
    "dxrv_ipc_open": (_int, [_vp, _vp, _c.POINTER(_vp)]),
    "dxrv_ipc_close": (_int, [_vp, _vp]),
    "dxrv_comm_get_unique_id": (_int, [_vp]),
    "dxrv_comm_init": (_int, [_vp, _vp, _int, _int]),
    "dxrv_comm_init_all": (_int, [_c.POINTER(_vp), _int]),
    "dxrv_comm_destroy": (_int, [_vp]),
    "dxrv_group_begin": (_int, []),
    "dxrv_group_end": (_int, []),
    "dxrv_bcast_u32": (_int, [_vp, _vp, _u32, _int]),
    "dxrv_bcast_mesh": (_int, [_vp, _vp, _u32, _u32, _vp, _u32, _int]),
    "dxrv_build_bvh_replicated": (_int, [_vp, _vp]),
    "dxrv_gather_grid": (_int, [_vp, _int]),
    "dxrv_gather_grid_slabs": (_int, [_vp, _int, _vp]),
    "dxrv_full_grid_device": (_int, [_vp, _c.POINTER(_vp), _c.POINTER(_sz)]),
    "dxrv_fetch_full_grid": (_int, [_vp, _vp, _sz]),
    "dxrv_share_grid_target": (_int, [_vp, _vp, _u32, _u32, _u32]),
}

_lib = None


def lib():
    """Load libdxrv.so once.  Raises if it has not been built -- never falls back to anything."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or make -C dxrvoxelizer_b200/csrc).  dxrvoxelizer_b200 has no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the ABI and the binding drift apart
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class DxrvError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("dxrv error %d: %s" % (code, message))
        self.code = code
