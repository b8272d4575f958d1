#!/usr/bin/env python
"""Per-kernel hottest source lines of an .ncu-rep holding several profiled launches (--import-source on).
Usage: ncu_lines.py report.ncu-rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
kernels = {}; cur_file = None; cur_fn = None; hdr = None
for r in csv.reader(io.StringIO(out)):
    if len(r) == 2 and r[0] in ("File Path", "File Name"): cur_file = r[1].split("/")[-1]; continue
    if len(r) == 2 and r[0] == "Function Name": cur_fn = r[1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if not hdr or len(r) != len(hdr) or r[2] not in ("-", ""): continue
    try: ie = int(r[hdr.index("Instructions Executed")]); sm = int(r[hdr.index("# Samples")]); te = int(r[hdr.index("Thread Instructions Executed")])
    except ValueError: continue
    if ie: kernels.setdefault(cur_fn, []).append((ie, sm, te, cur_file, r[0], r[1].strip()[:100]))
for fn, rows in kernels.items():
    tot = sum(a[0] for a in rows); tots = sum(a[1] for a in rows) or 1
    print("=====", fn[:110], "warp-inst", tot)
    byfile = {}
    for a in rows: byfile[a[3]] = byfile.get(a[3], 0) + a[0]
    print("   by file:", ", ".join(f"{k} {v / tot * 100:.1f}%" for k, v in sorted(byfile.items(), key=lambda kv: -kv[1])))
    for a in sorted(rows, reverse=True)[:top]:
        print(f"  {a[0] / tot * 100:5.1f}% inst {a[1] / tots * 100:5.1f}% smp lanes {a[2] / a[0]:4.1f}  {a[3]}:{a[4]}  {a[5]}")
