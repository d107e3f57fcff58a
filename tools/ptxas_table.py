"""Per-kernel resource table from ``nvcc -Xptxas -v`` logs (registers, spill bytes, static shared memory, barriers):
``python tools/ptxas_table.py <log dir> > profiles/r02_ptxas_resources.txt``.  Runs here (no GPU).  Logs:
``nvcc -c fastvim_b200/csrc/X.cu <build.py's NVCC_FLAGS> -Xptxas -v 2> X.cu.log``."""
import os
import re
import subprocess
import sys

d = sys.argv[1]
rows = []
for f in sorted(os.listdir(d)):
    if not f.endswith(".log"):
        continue
    txt = open(os.path.join(d, f)).read().splitlines()
    i = 0
    while i < len(txt):
        m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", txt[i])
        if not m:
            i += 1
            continue
        name = m.group(1)
        spill = regs = smem = bars = None
        for j in range(i + 1, min(i + 6, len(txt))):
            s = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", txt[j])
            if s:
                spill = (int(s.group(1)), int(s.group(2)), int(s.group(3)))
            r = re.search(r"Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes smem)?", txt[j])
            if r:
                regs, bars, smem = int(r.group(1)), int(r.group(2) or 0), int(r.group(3) or 0)
                break
        rows.append((f[:-4], name, regs, spill, smem, bars))
        i += 1
names = subprocess.run(["c++filt"], input="\n".join(r[1] for r in rows), capture_output=True, text=True).stdout.splitlines()
print(f"# {len(rows)} kernel instantiations, sm_100a, build.py flags; spills = (stack frame, spill stores, spill loads) bytes")
nsp = [r for r in rows if r[3] and (r[3][1] or r[3][2])]
print(f"# instantiations with register spills: {len(nsp)}")
print(f"{'file':22s} {'regs':>4s} {'spill':>14s} {'ssmem':>6s} {'bar':>3s}  kernel")
for (f, _, regs, spill, smem, bars), n in zip(rows, names):
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", n)
    print(f"{f:22s} {regs:4d} {str(spill):>14s} {smem:6d} {bars:3d}  {n[:110]}")
