#!/usr/bin/env python
"""Per-source-line instruction / stall-sample totals of an ncu report (needs -lineinfo + --import-source on).

    python tools/ncu_lines.py report.ncu-rep [--top 40] [--file block_fwd.cu]
"""
import csv, io, subprocess, sys, argparse
ap = argparse.ArgumentParser(); ap.add_argument("rep"); ap.add_argument("--top", type=int, default=40)
ap.add_argument("--file", default=""); ap.add_argument("--launch", type=int, default=0)
a = ap.parse_args()
out = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# sections start with a "File Path" row; take rows with a numeric line number
cur, hdr, data, nlaunch = None, None, {}, -1
for r in rows:
    if not r: continue
    if r[0] == "File Path":
        cur = r[1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        if cur and (not a.file or a.file in cur): nlaunch += 0
        continue
    if hdr and r[0].isdigit() and cur and (not a.file or a.file in cur):
        d = dict(zip(hdr, r))
        key = (cur.split("/")[-1], int(r[0]))
        e = data.setdefault(key, [0, 0, r[1][:90]])
        num = lambda v: int(v) if v and v.lstrip("-").isdigit() else 0
        e[0] += num(d["Instructions Executed"])
        e[1] += num(d["# Samples"])
tot_i = sum(e[0] for e in data.values()); tot_s = sum(e[1] for e in data.values())
print(f"total warp instr {tot_i}  samples {tot_s}")
for key, e in sorted(data.items(), key=lambda kv: -kv[1][0])[:a.top]:
    print(f"{key[0]}:{key[1]:4d} inst {e[0]:10d} {100*e[0]/max(tot_i,1):5.1f}%  samp {100*e[1]/max(tot_s,1):5.1f}%  {e[2]}")
