#!/usr/bin/env python3
"""Per-phase breakdown of an ncu source-page CSV of lpcnet_sample_kernel: stall samples between named barriers,
with the (deferred-blocking) barrier wait attributed to the instruction right after each BAR.SYNC.
usage: tools/seg_profile.py <src.csv> <n_warps>"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
nw = int(sys.argv[2])
hdr = rows[1]; isrc = hdr.index('Source'); ismp = hdr.index('# Samples'); iex = hdr.index('Instructions Executed'); iw = hdr.index('L1 Wavefronts Shared')
data = rows[2:]
tot = sum(int(r[ismp]) for r in data)
print('total samples', tot, 'per warp', tot / nw)
segname = 'start'; acc = 0; accinst = 0; accw = 0; out = []; prev_bar = None
for r in data:
    s = r[isrc]; smp = int(r[ismp])
    if prev_bar is not None:
        out.append(('  WAIT after ' + prev_bar, smp, 0, 0)); prev_bar = None; continue
    acc += smp; accinst += int(r[iex]); accw += int(r[iw] or 0)
    if 'BAR.' in s or 'EXIT' in s:
        out.append((segname + ' -> ' + s.strip()[:30], acc, accinst, accw))
        segname = s.strip()[:30]; acc = 0; accinst = 0; accw = 0; prev_bar = (s.strip()[:34] if 'BAR.SYNC' in s else None)
for name, a, n, w in out:
    if a > 0.003 * tot:
        print(f"{100*a/tot:6.2f}% ({a/(tot/nw)*100:6.1f}% of one warp-time)  inst={n/1e6:8.1f}M smem_wf={w/1e6:7.1f}M  {name}")
