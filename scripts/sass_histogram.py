"""Per-kernel SASS opcode histogram of libtaxo_sm100.so -> profiles/r2_sass_opcodes.csv (evidence that the GEMMs are tcgen05 + TMA +
TMEM and that the staged / star kernels use bulk copies): cuobjdump -sass, no GPU needed.

    python scripts/sass_histogram.py
"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "taxoexpan_b200", "libtaxo_sm100.so")
OUT = os.path.join(ROOT, "profiles", "r2_sass_opcodes.csv")
WATCH = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UBLKPF", "SYNCS", "ACQBULK", "PREEXIT", "LDGSTS", "HMMA", "FFMA", "LDG", "STG",
         "LDS", "STS", "SHFL", "ATOMG", "ATOMS", "RED", "MEMBAR", "BAR", "F2FP", "HADD2", "FMNMX", "IMAD"]

txt = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
kern = None
counts = collections.defaultdict(collections.Counter)
total = collections.Counter()
for ln in txt.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("void ", "")
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
    if m and kern:
        op = m.group(1)
        total[kern] += 1
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                counts[kern][w] += 1
                break
        else:
            base = op.split(".")[0]
            if base in WATCH:
                counts[kern][base] += 1
os.makedirs(os.path.dirname(OUT), exist_ok=True)
with open(OUT, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "sass_instructions"] + WATCH)
    for k in sorted(total, key=lambda k: -total[k]):
        w.writerow([k[:140], total[k]] + [counts[k][x] for x in WATCH])
    w.writerow(["TOTAL", sum(total.values())] + [sum(counts[k][x] for k in total) for x in WATCH])
tot = {x: sum(counts[k][x] for k in total) for x in WATCH}
print("kernels:", len(total), "| tcgen05/TMA/TMEM opcodes:", {x: tot[x] for x in ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "UBLKCP", "SYNCS", "ACQBULK", "PREEXIT") if tot[x]})
print("written", OUT)
