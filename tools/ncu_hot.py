#!/usr/bin/env python3
"""Hot SASS instructions of one kernel in an .ncu-rep (needs -lineinfo / --import-source on):
tools/ncu_hot.py X.ncu-rep <kernel-regex> [launch-index] [min-percent]"""
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
idx = sys.argv[3] if len(sys.argv) > 3 else "1"
minpct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.6
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f"::regex:{rx}:{idx}"],
                     capture_output=True, text=True).stdout
rows = [r for r in csv.reader(out.splitlines())]
hdr = next(r for r in rows if "Address" in r and "Source" in r)
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[rows.index(hdr) + 1:] if len(r) == len(hdr) and r[0].startswith("0x")]
first = data[0][0]
if [r[0] for r in data].count(first) > 1:      # several launches matched: keep the first
    data = data[:[r[0] for r in data].index(first, 1)]
tot = sum(int(r[ix["# Samples"]]) for r in data)
ninst = sum(int(r[ix["Instructions Executed"]]) for r in data)
print(f"total samples {tot}, SASS lines {len(data)}, warp instructions executed {ninst}")
agg = {}
for r in data:
    for h in hdr:
        if h.startswith("stall_") and "Not" not in h and r[ix[h]] not in ("", "0"):
            agg[h] = agg.get(h, 0) + int(r[ix[h]])
print("stall totals:", sorted(agg.items(), key=lambda x: -x[1])[:8])
ops = {}
for r in data:
    op = r[ix["Source"]].split()[0] if r[ix["Source"]].split() else "?"
    if op.startswith("@"):
        op = r[ix["Source"]].split()[1]
    op = op.split(".")[0]
    ops[op] = ops.get(op, 0) + int(r[ix["Instructions Executed"]])
print("opcode mix:", [(k, round(100 * v / ninst, 1)) for k, v in sorted(ops.items(), key=lambda x: -x[1])[:14]])
for k, r in enumerate(data):
    n = int(r[ix["# Samples"]])
    if n >= tot * minpct / 100:
        st = {h: int(r[ix[h]]) for h in hdr if h.startswith("stall_") and "Not" not in h and r[ix[h]] not in ("", "0")}
        top = sorted(st.items(), key=lambda x: -x[1])[:3]
        print(f"{k:4d} {n:5d} {100 * n / tot:4.1f}%  {r[ix['Source']].strip()[:64]:64s} {top}")
if len(sys.argv) > 5:                          # full per-SASS-line table: index, address, samples, warp instructions, thread instructions
    tix = ix.get("Thread Instructions Executed")
    with open(sys.argv[5], "w") as fh:
        fh.write("idx,address,samples,inst,thread_inst,sass\n")
        for k, r in enumerate(data):
            fh.write(f"{k},{r[0]},{r[ix['# Samples']]},{r[ix['Instructions Executed']]},{r[tix] if tix is not None else ''},\"{r[ix['Source']].strip()}\"\n")
