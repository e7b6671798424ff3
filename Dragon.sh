#!/bin/sh
# Bin/Dragon.bat of the reference: start DXRVoxelizer.exe -mesh Assets/dragon.obj
HERE=$(dirname "$0")
MESH=$(python -c "import sys; sys.path.insert(0, '$HERE'); import dxrvoxelizer_b200 as d; print(d.asset_path('dragon.obj'))")
exec "$HERE/dxrvoxelizer_b200/dxrvoxelizer" -mesh "$MESH" "$@"
