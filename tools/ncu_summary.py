#!/usr/bin/env python3
"""Summarise an .ncu-rep (read on the CPU box): one block of key metrics per captured launch.
Usage: tools/ncu_summary.py gpurun_out/X.ncu-rep [regex-of-extra-metrics]"""
import csv
import io
import re
import subprocess
import sys

KEYS = [
    r"^gpu__time_duration\.sum$", r"^dram__bytes_read\.sum$", r"^dram__bytes_write\.sum$",
    r"^launch__registers_per_thread$", r"^launch__grid_size$", r"^launch__block_size$", r"^launch__occupancy_limit",
    r"^launch__waves_per_multiprocessor$", r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$",
    r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^dram__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^sm__inst_executed_pipe_fp64.*", r"^sm__pipe_fp64_cycles_active.*", r"^smsp__inst_executed\.sum$",
    r"^sm__inst_executed\.sum$", r"^smsp__issue_active\.avg\.pct_of_peak_sustained_active$",
    r"^sm__inst_executed_pipe_(fma|fmaheavy|fmalite|alu|lsu|xu|fp64|uniform|cbu|adu)\.sum$",
    r"^sm__inst_executed_pipe_.*pct_of_peak_sustained_active$",
    r"^l1tex__t_sector_hit_rate\.pct$", r"^lts__t_sector_hit_rate\.pct$", r"^sm__cycles_elapsed\.avg$",
    r"^sm__cycles_active\.avg$", r"^smsp__cycles_active\.avg$", r"^smsp__warp_issue_stalled.*_per_warp_active\.pct$",
    r"^smsp__average_warp.*issue_stalled.*", r"^smsp__thread_inst_executed_per_inst_executed\.ratio$",
    r"^sm__sass_thread_inst_executed_op_d(fma|add|mul)_pred_on\.sum$", r"^smsp__sass_thread_inst_executed_op_d.*",
    r"^l1tex__t_bytes_pipe_lsu_mem_global_op_ld\.sum$", r"^lts__t_bytes\.sum$", r"^smsp__pcsamp_warps_issue_stalled.*",
]


def main():
    rep = sys.argv[1]
    extra = [sys.argv[2]] if len(sys.argv) > 2 else []
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    pats = [re.compile(k) for k in KEYS + extra]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"=== {d.get('Kernel Name', '?')[:110]}  grid={d.get('Grid Size')} block={d.get('Block Size')}")
        for i, h in enumerate(hdr):
            short = h.split(".", 2)[-1] if h.count(".") >= 2 and h.split(".")[1].startswith("Triage") else h
            if any(p.search(short) for p in pats) and r[i] not in ("", "n/a"):
                print(f"  {short} = {r[i]} {units[i]}")


if __name__ == "__main__":
    main()
