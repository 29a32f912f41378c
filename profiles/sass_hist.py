"""Opcode histogram + per-window share of executed warp instructions from
`ncu -i X.ncu-rep --page source --csv --print-source sass > X.csv`."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
si, ei, st = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
data = []
for r in rows[2:]:
    try:
        data.append((r[si], int(r[ei]), int(r[st])))
    except (ValueError, IndexError):
        pass
tot = sum(n for _, n, _ in data)
stot = sum(s for _, _, s in data) or 1
print("total warp instructions", tot, "SASS lines", len(data))
hist = collections.Counter()
for src, n, _ in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", src)
    hist[m.group(2) if m else src[:10]] += n
for op, n in hist.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print(f"{op:14s} {n:12d} {100 * n / tot:5.1f}%")
W = 100
for i in range(0, len(data), W):
    w = sum(n for _, n, _ in data[i:i + W])
    s = sum(s for _, _, s in data[i:i + W])
    print(f"line {i:5d}: instr {100 * w / tot:5.1f}%  stall samples {100 * s / stot:5.1f}%   {data[i][0].strip()[:50]}")
