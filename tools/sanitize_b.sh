export DXRV_NO_GRAPHS=1
O=gpurun_out
( DXRV_FUSED_BUILD=1 timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_build.py tests/test_sparse.py tests/test_gpu_voxelize.py -m gpu -q \
    -k "fused and (cube or ico20480 or dragon or given_bound) or to_host or candidate_lists or round_trip and not 1024 or empty_and_full" 2>&1 | tail -12 ) > $O/r02b_sanitizer_memcheck.txt 2>&1
tail -6 $O/r02b_sanitizer_memcheck.txt
