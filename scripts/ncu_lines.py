"""Aggregate warp-stall samples per CUDA source line from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.
usage: python scripts/ncu_lines.py file.csv [launch_index] [top_n]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
launch = -1
first_file = None
fpath = None
hdr = None
agg = collections.Counter()
src = {}
stall = collections.defaultdict(collections.Counter)
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        if first_file is None:
            first_file = r[1]
        if r[1] == first_file:
            launch += 1
        fpath = r[1].split('/')[-1]
        continue
    if r[0] == 'Function Name':
        continue
    if r[0] == 'Line No':
        hdr = r
        continue
    if launch != which or hdr is None or r[0] == '':
        continue
    try:
        ln = int(r[0])
        s = int(r[hdr.index('# Samples')])
    except ValueError:
        continue
    agg[(fpath, ln)] += s
    src[(fpath, ln)] = r[1]
    for j, h in enumerate(hdr):
        if h.startswith('stall_') and 'Not Issued' not in h:
            try:
                stall[(fpath, ln)][h] += int(r[j])
            except ValueError:
                pass
tot = sum(agg.values())
print('total samples', tot)
for k, v in agg.most_common(topn):
    top = ', '.join(f'{a[6:]}:{b}' for a, b in stall[k].most_common(3))
    print(f'{v:6d} {100 * v / max(tot, 1):5.1f}% {k[0]}:{k[1]:4d} | {src[k][:100]} | {top}')
