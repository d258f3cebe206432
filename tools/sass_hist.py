"""Opcode histogram (weighted by executed warp instructions) of one kernel from `ncu --page source --csv` (SASS view),
plus a coarse timeline: the kernel's SASS split into N chunks with executed instructions and stall samples."""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops, samp = Counter(), Counter()
seq = []
for r in rows[2:]:
    if r and r[0] == "Address":
        continue
    if len(r) < len(hdr):
        continue
    n = int(r[ix['Instructions Executed']]); s = int(r[ix['# Samples']])
    op = r[ix['Source']].split()
    op = [o for o in op if not o.startswith('@')]
    name = op[0].split('.')[0] if op else '?'
    ops[name] += n; samp[name] += s
    seq.append((n, s, r[ix['Source']].strip()))
tot = sum(ops.values()); ts = sum(samp.values())
print('total warp instructions', tot, 'samples', ts)
for k, v in ops.most_common(28):
    print(f"{k:10s} {v:>12d} {100*v/tot:5.1f}%   samples {100*samp[k]/max(ts,1):5.1f}%")
nchunk = int(sys.argv[2]) if len(sys.argv) > 2 else 24
step = (len(seq) + nchunk - 1) // nchunk
print('--- timeline (SASS order) ---')
for c in range(0, len(seq), step):
    part = seq[c:c + step]
    n = sum(p[0] for p in part); s = sum(p[1] for p in part)
    top = max(part, key=lambda p: p[1])
    print(f"[{c:5d}] inst {100*n/tot:5.1f}%  samples {100*s/max(ts,1):5.1f}%   hottest: {top[2][:70]}")
