"""Aggregate warp-stall samples per CUDA source line from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.
usage: python scripts/ncu_lines.py file.csv [launch_index] [top_n]   (launch index = position in the sequence of captured kernels)"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
# a launch = a maximal run of (File Path, Function Name) blocks with the same function name, restarting when the first file repeats
launch = -1
cur_fn = None
seen_files = set()
fpath = None
hdr = None
agg = collections.Counter()
inst = collections.Counter()
src = {}
stall = collections.defaultdict(collections.Counter)
name_of = {}
pending_file = None
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        pending_file = r[1]
        continue
    if r[0] == 'Function Name':
        if r[1] != cur_fn or pending_file in seen_files:
            launch += 1
            cur_fn = r[1]
            seen_files = set()
            name_of[launch] = r[1]
        seen_files.add(pending_file)
        fpath = pending_file.split('/')[-1]
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
    try:
        inst[(fpath, ln)] += int(r[hdr.index('Instructions Executed')])
    except ValueError:
        pass
    src[(fpath, ln)] = r[1]
    for j, h in enumerate(hdr):
        if h.startswith('stall_') and 'Not Issued' not in h:
            try:
                stall[(fpath, ln)][h] += int(r[j])
            except ValueError:
                pass
tot = sum(agg.values())
ti = sum(inst.values())
print('launch', which, name_of.get(which, '?')[:80], '| total samples', tot, '| instructions', ti)
for k, v in agg.most_common(topn):
    top = ', '.join(f'{a[6:]}:{b}' for a, b in stall[k].most_common(3))
    print(f'{v:6d} {100 * v / max(tot, 1):5.1f}% smp {100 * inst[k] / max(ti, 1):5.1f}% inst {k[0]}:{k[1]:4d} | {src[k][:88]} | {top}')
