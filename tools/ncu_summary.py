#!/usr/bin/env python
"""Compact per-kernel summary of an ncu report (run here, no GPU needed):

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--json out.json]

Prints, per profiled launch: duration, DRAM bytes, throughput percentages, occupancy limits,
instruction count, pipe utilisation and the top warp-stall reasons."""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "dur"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_registers", "lim_regs"),
    ("launch__occupancy_limit_shared_mem", "lim_smem"),
    ("launch__occupancy_limit_warps", "lim_warps"),
    ("launch__waves_per_multiprocessor", "waves"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_wf%"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank_conf"),
    ("sm__cycles_elapsed.max", "cycles"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall_cols = [(h, i) for h, i in idx.items() if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[idx["Kernel Name"]].split("(")[0][-60:]}
        for k, short in KEYS:
            if k in idx:
                d[short] = f"{r[idx[k]]} {units[idx[k]]}".strip()
        st = []
        for h, i in stall_cols:
            try:
                st.append((float(r[i]), h.split("stalled_")[1]))
            except ValueError:
                pass
        tot = sum(v for v, _ in st) or 1.0
        d["stalls"] = ", ".join(f"{n}:{100 * v / tot:.0f}%" for v, n in sorted(st, reverse=True)[:7])
        res.append(d)
        print(json.dumps(d))
    if "--json" in sys.argv:
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
