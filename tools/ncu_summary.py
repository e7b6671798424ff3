#!/usr/bin/env python3
"""Summarise an .ncu-rep: key metrics per kernel + hottest source lines.  usage: ncu_summary.py rep [kernel-substr] [top]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; sel = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); H, U = rows[0], rows[1]
want = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio',
        'launch__registers_per_thread','launch__shared_mem_per_block_dynamic','launch__grid_size','launch__block_size','launch__waves_per_multiprocessor',
        'launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio']
seen = set()
for r in rows[2:]:
    name = r[H.index('Kernel Name')]
    if sel not in name or name in seen: continue
    seen.add(name)
    print('===', name[:90])
    for w in want:
        if w in H: print('  %-80s %s %s' % (w, r[H.index(w)], U[H.index(w)]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"] + (["-k", "regex:" + sel] if sel else []), capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur = None; hdr = None; agg = collections.Counter(); samp = collections.Counter(); text = {}; fn = None; first = None
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name':
        fn = r[1]
        if first is None: first = fn
        continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None or fn != first: continue
    try: line = int(r[0])
    except: continue
    def num(col):
        try: return int(r[hdr.index(col)])
        except: return 0
    agg[(cur, line)] += num('Instructions Executed'); samp[(cur, line)] += num('# Samples'); text[(cur, line)] = r[1][:100]
tot = sum(agg.values()) or 1; tots = sum(samp.values()) or 1
print('--- hottest lines of', (first or '')[:80])
for k, v in sorted(agg.items(), key=lambda kv: -samp[kv[0]])[:top]:
    print('%5.1f%% samp %5.1f%% inst  %s:%d  %s' % (100 * samp[k] / tots, 100 * v / tot, k[0], k[1], text[k]))
