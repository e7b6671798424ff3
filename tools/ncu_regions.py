#!/usr/bin/env python3
"""Instruction / stall-sample share per source-line REGION of one kernel.  usage: ncu_regions.py rep kernel-substr file:lo-hi=name ..."""
import csv, subprocess, io, collections, sys
rep, sel = sys.argv[1], sys.argv[2]
regions = []
for a in sys.argv[3:]:
    rng, name = a.split('='); f, lh = rng.split(':'); lo, hi = lh.split('-'); regions.append((f, int(lo), int(hi), name))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "-k", "regex:" + sel], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur = None; hdr = None; fn = None; first = None
inst = collections.Counter(); samp = collections.Counter(); thr = collections.Counter()
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
    except ValueError: continue
    def num(col):
        try: return int(r[hdr.index(col)])
        except (ValueError, IndexError): return 0
    name = 'other:' + cur
    for f, lo, hi, n in regions:
        if cur == f and lo <= line <= hi: name = n; break
    inst[name] += num('Instructions Executed'); samp[name] += num('# Samples'); thr[name] += num('Thread Instructions Executed')
ti, ts = sum(inst.values()) or 1, sum(samp.values()) or 1
print('total warp instructions %d, samples %d' % (ti, ts))
for n, v in sorted(inst.items(), key=lambda kv: -samp[kv[0]]):
    print('%-28s %5.1f%% inst (%9d)  %5.1f%% samples   %4.1f thr/inst' % (n, 100 * v / ti, v, 100 * samp[n] / ts, thr[n] / max(v, 1)))
