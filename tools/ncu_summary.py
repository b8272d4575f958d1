#!/usr/bin/env python
"""Text summary of an .ncu-rep (one profiled launch): key raw metrics, stall reasons, opcode mix,
executed code footprint and the hottest source lines.  Usage: ncu_summary.py report.ncu-rep [n_lines]"""
import csv
import io
import re
import subprocess
import sys


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    raw = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units, vals = raw[0], raw[1], raw[2]
    want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
            "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
    for w in want:
        if w in hdr:
            print(f"{w} = {vals[hdr.index(w)]} {units[hdr.index(w)]}")
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
            v = float(vals[i] or 0)
            if v >= 0.05:
                print(f"stall {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} = {v:.2f} warps per issue-active cycle")
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "cuda,sass"))))
    cur, h2, agg, ops, seen = None, None, {}, {}, set()
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) == 2:
            continue
        if r[0] == "Line No":
            h2 = r
            continue
        ie = num(r[h2.index("Instructions Executed")])
        if r[2] not in ("-", ""):
            if r[2] in seen:
                continue
            seen.add(r[2])
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[3])
            if m and ie:
                op = m.group(2).split(".")[0]
                ops[op] = ops.get(op, 0) + ie
            continue
        a = agg.setdefault((cur, num(r[0])), [0, 0, 0, r[1]])
        a[0] += num(r[h2.index("# Samples")])
        a[1] += ie
        a[2] += num(r[h2.index("Thread Instructions Executed")])
    executed = sum(1 for _ in ops)  # distinct opcodes (not footprint)
    tot = sum(a[0] for a in agg.values()) or 1
    toti = sum(a[1] for a in agg.values()) or 1
    so = sum(ops.values()) or 1
    print("opcode mix (executed):", ", ".join(f"{k} {v / so * 100:.1f}%" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:18]))
    fp = sum(v for k, v in ops.items() if k in ("DMUL", "DADD", "DFMA", "DSETP", "DMNMX"))
    print(f"FP64 share of executed instructions: {fp / so * 100:.1f}%")
    print(f"hottest source lines (of {toti} warp instructions):")
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        eff = a[2] / a[1] / 32 if a[1] else 0
        print(f"  samples {a[0] / tot * 100:5.1f}%  inst {a[1] / toti * 100:5.1f}%  lanes {eff:4.2f}  {f}:{l}  {a[3][:90]}")


if __name__ == "__main__":
    main()
