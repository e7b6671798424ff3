#!/bin/bash
# Second capture of round 2 (after the fused build / bit-31 masks / read-back transports): ncu launch list + full
# captures of the kernels that changed, sanitizer runs over the new code.  Run under gpurun; everything goes to gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 200 --csv --log-file $O/r02b_launches_bench_c3.csv python bench.py --steps 2 --warmup 1 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:"k_trace_fill_columns|k_build_fused|k_bin_columns|k_file_columns" -s 8 -c 4 -o $O/r02b_parity python bench.py --steps 2 --warmup 1 > /dev/null 2>&1
ls -la $O/r02b_*
export DXRV_NO_GRAPHS=1
( DXRV_FUSED_BUILD=1 timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_build.py tests/test_sparse.py tests/test_gpu_voxelize.py -m gpu -q -x \
    -k "fused and (cube or ico20480 or dragon or given_bound) or to_host or candidate_lists or round_trip and bunny" 2>&1 | tail -15 ) > $O/r02b_sanitizer_memcheck.txt 2>&1
( DXRV_FUSED_BUILD=1 timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_build.py tests/test_gpu_voxelize.py -m gpu -q -x \
    -k "fused and (cube or ico20480) and not True or candidate_lists" 2>&1 | tail -15 ) > $O/r02b_sanitizer_racecheck.txt 2>&1
tail -5 $O/r02b_sanitizer_memcheck.txt $O/r02b_sanitizer_racecheck.txt
