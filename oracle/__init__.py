"""CPU oracle bindings -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this package.  The product package (dxrvoxelizer_b200) never does.  See dxrv_oracle.h for
the contract ("Spec H"), the reference file:line map and the "parity unpinned" note.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

MODE_SHADER = 0
MODE_PARITY = 1
TIER_BRUTE = 0
TIER_ACCEL = 1


def build(quiet=True):
    """Compile libdxrv_oracle.so (and oracle/_ref when /root/reference is present)."""
    subprocess.run(["make", "-C", _HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libdxrv_oracle.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        L.oracle_bound.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p]
        L.oracle_bound.restype = None
        L.oracle_voxelize.argtypes = [
            ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32,
            ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
            ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]
        L.oracle_voxelize.restype = ctypes.c_int
        L.oracle_max_threads.restype = ctypes.c_int
        L.oracle_render_view.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.oracle_render_view.restype = ctypes.c_int
        _LIB = L
    return _LIB


def max_threads():
    return int(_lib().oracle_max_threads())


def bound(vertices, stride=None):
    """{cx,cy,cz,w} as Voxelizer::Init derives it (Content/Voxelizer.cpp:52-57)."""
    vb = np.ascontiguousarray(vertices)
    if stride is None:
        stride = vb.shape[-1] * vb.itemsize if vb.ndim == 2 else 24
    raw = vb.view(np.uint8).reshape(-1)
    out = np.zeros(4, np.float32)
    _lib().oracle_bound(raw.ctypes.data, raw.size // stride, stride, out.ctypes.data)
    return out


def voxelize(vertices, indices, N, mode, z0=0, z1=None, tier=TIER_ACCEL, threads=0, bound=None,
             texels=False, stride=None):
    """Returns dict(bits=uint32[(z1-z0), N, P], texels=uint32[(z1-z0),N,N] or None,
    crossings=int, odd_columns=int)."""
    vb = np.ascontiguousarray(vertices)
    if stride is None:
        stride = vb.shape[-1] * vb.itemsize if vb.ndim == 2 else 24
    raw = vb.view(np.uint8).reshape(-1)
    ib = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
    if z1 is None:
        z1 = N
    P = (N + 31) // 32
    bits = np.empty((z1 - z0, N, P), np.uint32)
    tex = np.empty((z1 - z0, N, N), np.uint32) if texels else None
    b = None if bound is None else np.ascontiguousarray(bound, dtype=np.float32)
    cr, odd = ctypes.c_uint64(0), ctypes.c_uint64(0)
    rc = _lib().oracle_voxelize(raw.ctypes.data, raw.size // stride, stride, ib.ctypes.data, ib.size,
                                None if b is None else b.ctypes.data, N, mode, z0, z1, tier, threads,
                                bits.ctypes.data, None if tex is None else tex.ctypes.data,
                                ctypes.byref(cr), ctypes.byref(odd))
    if rc != 0:
        raise ValueError("oracle_voxelize: invalid arguments")
    return {"bits": bits, "texels": tex, "crossings": int(cr.value), "odd_columns": int(odd.value)}


def render_view(bits, N, width, height, screen_to_local, eye, light, threads=0):
    """RGBA8 image [height, width, 4] of the viewer pass (PSRayCast.hlsl) over a full bit grid."""
    b = np.ascontiguousarray(bits, dtype=np.uint32)
    m = np.ascontiguousarray(screen_to_local, dtype=np.float32).reshape(16)
    e = np.ascontiguousarray(eye, dtype=np.float32)
    l = np.ascontiguousarray(light, dtype=np.float32)
    out = np.empty((height, width, 4), np.uint8)
    rc = _lib().oracle_render_view(b.ctypes.data, N, width, height, m.ctypes.data, e.ctypes.data, l.ctypes.data,
                                   out.ctypes.data, threads)
    if rc != 0:
        raise ValueError("oracle_render_view: invalid arguments")
    return out


# ---- the reference's own ObjLoader (compiled from /root/reference into oracle/_ref) -------------

def ref_loader_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_objloader.so"))


def ref_load_obj(path):
    """Run the UNMODIFIED reference ObjLoader::Import(path, true, true).  Returns
    (vertex_bytes uint8[nv*stride], indices uint32[ni], stride, aabb float32[6])."""
    global _REF
    if _REF is None:
        L = ctypes.CDLL(os.path.join(_HERE, "_ref", "libref_objloader.so"))
        L.ref_obj_import.restype = ctypes.c_void_p
        L.ref_obj_import.argtypes = [ctypes.c_char_p]
        for f in ("vertices", "indices"):
            getattr(L, "ref_obj_" + f).restype = ctypes.c_void_p
            getattr(L, "ref_obj_" + f).argtypes = [ctypes.c_void_p]
        for f in ("num_vertices", "num_indices", "stride"):
            getattr(L, "ref_obj_" + f).restype = ctypes.c_uint32
            getattr(L, "ref_obj_" + f).argtypes = [ctypes.c_void_p]
        L.ref_obj_aabb.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.ref_obj_free.argtypes = [ctypes.c_void_p]
        _REF = L
    L = _REF
    h = L.ref_obj_import(os.fsencode(path))
    if not h:
        raise IOError("reference ObjLoader::Import failed for %s" % path)
    try:
        nv, ni, st = L.ref_obj_num_vertices(h), L.ref_obj_num_indices(h), L.ref_obj_stride(h)
        vb = np.ctypeslib.as_array(ctypes.cast(L.ref_obj_vertices(h), ctypes.POINTER(ctypes.c_uint8)), (nv * st,)).copy()
        ib = np.ctypeslib.as_array(ctypes.cast(L.ref_obj_indices(h), ctypes.POINTER(ctypes.c_uint32)), (ni,)).copy()
        aabb = np.zeros(6, np.float32)
        L.ref_obj_aabb(h, aabb.ctypes.data)
    finally:
        L.ref_obj_free(h)
    return vb, ib, st, aabb
