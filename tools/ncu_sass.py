#!/usr/bin/env python
"""Per-opcode instruction histogram + hottest SASS lines of an ncu report (source page).
    python tools/ncu_sass.py rep.ncu-rep [topN]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ia = hdr.index("Instructions Executed"); isrc = hdr.index("Source"); ist = hdr.index("Warp Stall Sampling (All Samples)")
ops = collections.Counter(); stall = collections.Counter(); tot = 0; lines = []
for r in rows[2:]:
    if len(r) <= ia: continue
    try: n = int(r[ia]); s = int(r[ist])
    except ValueError: continue
    src = r[isrc].strip(); toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    op = op.split(".")[0]
    ops[op] += n; stall[op] += s; tot += n; lines.append((s, n, src))
print(f"total warp-instructions {tot}")
for op, n in ops.most_common(topn):
    print(f"{op:12s} {n:12d} {100*n/tot:5.1f}%   stall-samples {stall[op]}")
print("--- hottest lines by stall samples")
for s, n, src in sorted(lines, reverse=True)[:topn]:
    print(f"{s:6d} {n:10d}  {src[:110]}")
