#!/bin/bash
# A/B on the GPU box: run "$@" with build_ab/libdxrv_base.so, then with the current library.
L=dxrvoxelizer_b200/libdxrv.so
cp $L /tmp/new.so
cp build_ab/libdxrv_base.so $L; echo "--- base"; "$@"
cp /tmp/new.so $L; echo "--- new"; "$@"
