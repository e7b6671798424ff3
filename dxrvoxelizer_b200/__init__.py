"""dxrvoxelizer_b200 -- B200-native solid voxelizer behind the StarsX/DXRVoxelizer voxelization
surface.  The product is libdxrv.so (hand-written sm_100a CUDA behind the C ABI of include/dxrv.h);
this package is the Python face of that ABI plus the host-side helpers the tests and bench use.
"""
from ._lib import (MODE_PARITY, MODE_SHADER, FORMAT_BITS, FORMAT_R10G10B10A2, FORMAT_U8, DxrvError,  # noqa: F401
                   LIB_PATH)
from .voxelizer import Mesh, Voxelizer, default_view, load_obj, sparse_decode, unpack_bits, voxelize_obj_batch, save_image, sparse_encode  # noqa: F401
from .assets import asset_path  # noqa: F401
from . import sharding  # noqa: F401,E402
