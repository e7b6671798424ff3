#!/bin/sh
# Bin/TuringBowl.bat of the reference: start DXRVoxelizer.exe -mesh Assets/TuringBowl.obj 0.0 2.8 0.0 0.03
HERE=$(dirname "$0")
MESH=$(python -c "import sys; sys.path.insert(0, '$HERE'); import dxrvoxelizer_b200 as d; print(d.asset_path('TuringBowl.obj'))")
exec "$HERE/dxrvoxelizer_b200/dxrvoxelizer" -mesh "$MESH" 0.0 2.8 0.0 0.03 "$@"
