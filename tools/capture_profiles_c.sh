#!/bin/bash
# Final capture of round 2 (after the exact tile sub-rectangle in the binning and the one-pass host grid): ncu launch list +
# full captures of the step's kernels, bench lines of every config, sanitizer run over the binning.  Run under gpurun;
# everything goes to gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
python bench.py > $O/r02c_bench_line.json 2> $O/r02c_bench.err
python bench.py --impl reference > $O/r02c_bench_line_reference_arm.json 2>> $O/r02c_bench.err
python bench.py --config c4 > $O/r02c_bench_line_c4.json 2>> $O/r02c_bench.err
python bench.py --config c5 > $O/r02c_bench_line_c5.json 2>> $O/r02c_bench.err
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/r02c_launches_bench_c3.csv python bench.py --steps 2 --warmup 1 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:"k_trace_fill_columns|k_build_fused|k_bin_columns|k_file_columns" -s 8 -c 4 -o $O/r02c_parity python bench.py --steps 2 --warmup 1 > /dev/null 2>&1
ls -la $O/r02c_*
export DXRV_NO_GRAPHS=1
( timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_voxelize.py tests/test_sparse.py -m gpu -q -x \
    -k "walked_and_binned or binned_candidates_overflow or candidate_lists or ragged or mesh_to_host and bunny or huge_triangles" 2>&1 | tail -12 ) > $O/r02c_sanitizer_memcheck.txt 2>&1
( timeout 300 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_voxelize.py -m gpu -q -x -k "binned_candidates_overflow or candidate_lists or ragged" 2>&1 | tail -12 ) > $O/r02c_sanitizer_racecheck.txt 2>&1
tail -4 $O/r02c_sanitizer_memcheck.txt $O/r02c_sanitizer_racecheck.txt
for f in $O/r02c_bench_line*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print(sys.argv[1], d.get("metric"), d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("ms_per_step"), (d.get("roofline") or {}).get("frac"), d.get("mismatched_voxels"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
