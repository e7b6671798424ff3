#!/bin/bash
# A/B/n on the GPU box: tools/abn.sh "lib1 lib2 ..." cmd...  runs cmd with each build_ab/<lib>.so swapped in
# (development aid; "cur" = the library as shipped).
cd "$(dirname "$0")/.."
L=dxrvoxelizer_b200/libdxrv.so
cp $L /tmp/libdxrv_cur.so
libs="$1"; shift
for v in $libs; do
    if [ "$v" = cur ]; then cp /tmp/libdxrv_cur.so $L; else cp build_ab/libdxrv_$v.so $L; fi
    echo "--- $v"; "$@"
done
cp /tmp/libdxrv_cur.so $L
