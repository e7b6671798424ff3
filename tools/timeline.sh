#!/bin/bash
# CTA timeline of the fill kernel (development aid).
#   here:       tools/timeline.sh build      -> build_ab/libdxrv_tl.so, compiled with -DDXRV_TIMELINE
#   GPU box:    tools/timeline.sh [N] [asset] -> swaps that library in, prints the timeline, swaps back
cd "$(dirname "$0")/.."
L=dxrvoxelizer_b200/libdxrv.so
if [ "$1" = build ]; then
    mkdir -p build_ab
    exec make -s -C dxrvoxelizer_b200/csrc BUILD=build_tl EXTRA=-DDXRV_TIMELINE LIB=../../build_ab/libdxrv_tl.so ../../build_ab/libdxrv_tl.so
fi
cp $L /tmp/libdxrv_keep.so; cp build_ab/libdxrv_tl.so $L
DXRV_NO_GRAPHS=1 DXRV_DBG_TIMELINE=1 python tools/prof_parity.py "${1:-1024}" 2 "${2:-dragon.obj}" ${3:-} ${4:-} | tail -${TL_LINES:-7}
cp /tmp/libdxrv_keep.so $L
