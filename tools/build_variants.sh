#!/bin/bash
# build_ab/libdxrv_<name>.so for A/B runs on the GPU box (tools/abn.sh): tools/build_variants.sh name "-DFLAG=..." [name flags ...]
cd "$(dirname "$0")/../dxrvoxelizer_b200/csrc" || exit 1
mkdir -p ../../build_ab
while [ $# -ge 2 ]; do
    name=$1; flags=$2; shift 2
    make -s -j8 BUILD=build_v_$name LIB=../../build_ab/libdxrv_$name.so EXTRA="$flags" ../../build_ab/libdxrv_$name.so 2>&1 | grep -i "error" 
    ls -la ../../build_ab/libdxrv_$name.so | awk '{print $5, $9}'
done
