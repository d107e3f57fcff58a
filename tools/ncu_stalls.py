#!/usr/bin/env python
"""Top stalled SASS instructions of one kernel of an ncu report (--page source), with the dominant stall reason and the
instructions just before each -- the view that exposed the load -> convert serialisations of round 2.

    python tools/ncu_stalls.py report.ncu-rep [kernel-name-substring] [topN]
"""
import csv, io, subprocess, sys
rep = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else ""; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 14
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
launches, cur, hdr = [], None, None
for r in rows:
    if not r: continue
    if r[0] == "Kernel Name": cur = []; launches.append((r[1], cur)); continue
    if r[0] == "Address": hdr = r; continue
    if cur is not None and r[0].startswith("0x"): cur.append(dict(zip(hdr, r)))
num = lambda v: int(v) if v and v.lstrip("-").isdigit() else 0
skeys = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
for name, ins in launches:
    if pat not in name: continue
    tot = sum(num(d["# Samples"]) for d in ins)
    print(name[:100], len(ins), "SASS instructions,", tot, "samples")
    agg = {k: sum(num(d[k]) for d in ins) for k in skeys}
    print("  stalls:", ", ".join(f"{k[6:]}:{100*v//max(tot,1)}%" for k, v in sorted(agg.items(), key=lambda t: -t[1])[:7]))
    top = sorted(enumerate(ins), key=lambda t: -num(t[1]["# Samples"]))[:topn]
    for i, d in top:
        why = max(skeys, key=lambda k: num(d[k]))
        prev = " <- ".join(ins[j]["Source"].split(";")[0].strip()[:38] for j in range(i - 1, max(i - 3, -1), -1))
        print(f"  {i:5d} {100*num(d['# Samples'])/max(tot,1):5.1f}%  {why[6:]:14s} {d['Source'].split(';')[0].strip()[:52]:52s} | {prev}")
    break
