#!/usr/bin/env python
"""SASS evidence per kernel (run here, no GPU): for every kernel of libfastvim_b200.so whose name matches a pattern, the
count of the Blackwell-specific mnemonics and a short excerpt around the first occurrence of each.

    python tools/sass_excerpt.py > profiles/r02_sass_excerpts.txt
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "fastvim_b200", "build")
TARGETS = [   # (object file, kernel-name regex, mnemonics of interest)
    ("gemm_tc.cu.o", r"gemm_tc_kernelILi256E", ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "SYNCS"]),
    ("gemm_tc2.cu.o", r"gemm_tc2_kernelILi256ELb1ELb1ELb1E", ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR"]),
    ("gemm_out_norm.cu.o", r"gemm_out_norm_kernelILi192ELb1E", ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "ATOMG", "LD.E.STRONG.GPU", "LDG.E.STRONG.GPU", "ACQBULK", "UTCBAR"]),
    ("block_cluster.cu.o", r"block_cluster_kernelILi12ELb1ELb1ELb0E", ["HMMA", "LDGSTS", "UCGABAR_ARV", "UCGABAR_WAIT", "MUFU.TANH", "MUFU.EX2", "FFMA2", "CCTL.E.PF2"]),
    ("block_fwd.cu.o", r"block_fwd_kernelILi14ELb1ELb1ELi12E", ["HMMA", "LDGSTS", "MUFU.TANH", "MUFU.EX2", "FFMA2", "ACQBULK", "ST.E.STRONG.GPU", "STG.E.STRONG.GPU"]),
    ("gate_bwd_v.cu.o", r"gate_bwd_v_kernelILb1ELi4E", ["MUFU.TANH", "FFMA2", "SHFL", "LDG.E.64", "STG.E.64"]),
    ("peer.cu.o", r"peer_copy_kernel", ["RED.E.ADD", "LDG.E.128", "MEMBAR", "NANOSLEEP"]),
    ("gate.cu.o", r"gate_fwd_kernel.*bf16", ["UBLKCP", "SYNCS"]),
]


def sass(obj, pat):
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, obj)], capture_output=True, text=True).stdout
    blocks = re.split(r"\n\s*Function : ", out)
    for b in blocks[1:]:
        name = b.split("\n", 1)[0].strip()
        if re.search(pat, name):
            return name, [ln.strip() for ln in b.split("\n") if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln)]
    return None, []


for obj, pat, mnems in TARGETS:
    name, lines = sass(obj, pat)
    if not name:
        print(f"== {obj}: no kernel matching {pat}\n")
        continue
    print(f"== {obj} :: {name}   ({len(lines)} SASS instructions)")
    for m in mnems:
        hits = [i for i, ln in enumerate(lines) if m in ln]
        print(f"   {m:14s} x{len(hits)}")
    for m in mnems:
        hits = [i for i, ln in enumerate(lines) if m in ln]
        if hits:
            i = hits[0]
            print(f"   -- first {m}:")
            for ln in lines[max(0, i - 1):i + 2]:
                print("      " + re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", ln)[:150])
    print()
