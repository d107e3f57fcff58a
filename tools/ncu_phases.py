#!/usr/bin/env python
"""Splits a kernel's SASS (ncu --page source, SASS view) into phases delimited by BAR.SYNC and prints
per-phase executed warp instructions, stall samples and the dominant opcodes.

    python tools/ncu_phases.py report.ncu-rep [--launch 0]
"""
import csv, io, subprocess, sys, argparse, collections
ap = argparse.ArgumentParser(); ap.add_argument("rep"); ap.add_argument("--launch", type=int, default=0)
a = ap.parse_args()
out = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
launches, cur, hdr = [], None, None
for r in rows:
    if not r: continue
    if r[0] == "Kernel Name":
        cur = []; launches.append((r[1], cur)); continue
    if r[0] == "Address":
        hdr = r; continue
    if cur is not None and r[0].startswith("0x"):
        cur.append(dict(zip(hdr, r)))
name, ins = launches[a.launch]
print(name, len(ins), "SASS instructions")
num = lambda v: int(v) if v and v.lstrip("-").isdigit() else 0
phases, cur = [], []
for d in ins:
    cur.append(d)
    if "BAR.SYNC" in d["Source"]:
        phases.append(cur); cur = []
phases.append(cur)
ti = sum(num(d["Instructions Executed"]) for d in ins); ts = sum(num(d["# Samples"]) for d in ins)
print(f"total warp-instr {ti}  samples {ts}")
stall_keys = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
for i, ph in enumerate(phases):
    pi = sum(num(d["Instructions Executed"]) for d in ph); ps = sum(num(d["# Samples"]) for d in ph)
    ops = collections.Counter()
    for d in ph:
        op = d["Source"].split()[0] if not d["Source"].strip().startswith("@") else d["Source"].split()[1]
        ops[op.split(".")[0]] += num(d["Instructions Executed"])
    st = collections.Counter()
    for d in ph:
        for k in stall_keys: st[k[6:]] += num(d[k])
    top = ", ".join(f"{k}:{v*100//max(pi,1)}%" for k, v in ops.most_common(8))
    stt = ", ".join(f"{k}:{v*100//max(ps,1)}%" for k, v in st.most_common(5))
    print(f"phase {i:2d}: sass {len(ph):5d}  inst {pi:9d} ({100*pi/ti:4.1f}%)  samples {ps:6d} ({100*ps/max(ts,1):4.1f}%)\n     ops: {top}\n     stalls: {stt}")
