#!/bin/bash
# Round-end evidence on the GPU box (run under gpurun): ncu launch lists + `--set full` captures of the hot kernels.
# Everything goes to gpurun_out/; tools/ncu_summary.py turns the .ncu-rep files into the text summaries under profiles/.
set -u
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
# launch lists (cold-cache, serialised: shares of the step, not absolute times)
$NCU --metrics gpu__time_duration.sum -c 200 --csv --log-file $O/r02_launches_bench_c3.csv python bench.py --steps 2 --warmup 1 > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum -c 80 --csv --log-file $O/r02_launches_c4.csv python tools/prof_c4.py > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum -c 60 --csv --log-file $O/r02_launches_shader.csv python tools/shader_timing.py 512 > /dev/null 2>&1
# full captures
$NCU --set full --import-source on -k regex:"k_trace_fill_columns|k_bin_columns|k_file_columns" -s 6 -c 3 -o $O/r02_parity python bench.py --steps 2 --warmup 1 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:"k_trace_shader_bins" -c 1 -o $O/r02_shader_trace python tools/shader_timing.py 512 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:"k_bins_" -c 8 -o $O/r02_shader_bins_build python tools/shader_timing.py 512 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:"k_onesweep_pass_big|k_morton|k_leaf_setup|k_scatter_crossings" -s 7 -c 6 -o $O/r02_c4 python tools/prof_c4.py > /dev/null 2>&1
ls -la $O/*.ncu-rep $O/r02_launches_*.csv
