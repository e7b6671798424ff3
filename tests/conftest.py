import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import ctypes
        from dxrvoxelizer_b200 import _lib as L
        h = ctypes.c_void_p()
        rc = L.lib().dxrv_create(ctypes.byref(h), 0)
        if rc == 0:
            L.lib().dxrv_destroy(h)
        return rc == 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly, not skip: a silent skip would hide a
    # missing CUDA extension.  Without `-m gpu` the GPU tests are simply deselected by the driver.
    pass


@pytest.fixture(scope="session")
def meshes_mod():
    from dxrvoxelizer_b200 import meshes
    return meshes


@pytest.fixture(scope="session")
def assets():
    """name -> Mesh for the reference's three shipped OBJ fixtures (product loader)."""
    import dxrvoxelizer_b200 as d
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = d.load_obj(d.asset_path(name))
        return cache[name]
    return get


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def vox():
    """One GPU context shared by the gpu tests; creation fails loudly without a device."""
    import dxrvoxelizer_b200 as d
    v = d.Voxelizer(0)
    yield v
    v.close()


def popcount(a):
    return int(np.unpackbits(np.ascontiguousarray(a).view(np.uint8)).sum())
