#!/usr/bin/env python
"""Per-kernel totals of the TIMED frames of `bench.py --steps K --warmup W` from an ncu launch list
(`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file list.csv python bench.py ...`).

    python tools/launch_summary.py list.csv [--warmup 3] [--steps 2]

A frame attempt starts at a k_raygen (round 1: k_camera_entity) that follows as many k_final as there were k_raygen before it
(the pipelines of a split frame interleave: two k_raygen, ..., two k_final); attempts whose k_final ran for real (> 20 us, i.e.
not an overflowed attempt that returned early) are complete.  --chunks-per-frame N > 0 selects the older reading (a sequence per
k_raygen, N sequences per frame) for the lists of round 1 and early round 2."""
import argparse
import csv
import re
from collections import OrderedDict

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--chunks-per-frame", type=int, default=0)
args = ap.parse_args()
rows = list(csv.reader(open(args.csv, errors="replace")))
hdr = next(r for r in rows if "Kernel Name" in r)
launches = []
for r in rows:
    if len(r) == len(hdr) and r is not hdr and r[hdr.index("Metric Name")] == "gpu__time_duration.sum":
        name = r[hdr.index("Kernel Name")]
        m = re.search(r"(k_[a-z_0-9]+(<[^>]*>)?)", name)
        val = float(r[hdr.index("Metric Value")].replace(",", ""))
        unit = r[hdr.index("Metric Unit")]
        ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        launches.append((m.group(1) if m else name[:40], ns))
seqs, cur, prev = [], None, ""
open_raygens = 0
for name, ns in launches:
    if args.chunks_per_frame > 0:
        # round 1 lists: a chunk starts at k_camera_entity; round 2: at k_raygen (the camera lookup moved into it)
        start = name.startswith("k_camera_entity") or (name.startswith("k_raygen") and not prev.startswith("k_camera_entity"))
    else:
        start = name.startswith("k_raygen") and open_raygens == 0
    if start:
        cur = []
        seqs.append(cur)
    if name.startswith("k_raygen"):
        open_raygens += 1
    elif name.startswith("k_final"):
        open_raygens = max(0, open_raygens - 1)
    prev = name
    if cur is not None and name.startswith("k_"):
        cur.append((name, ns))
complete = [s for s in seqs if any(n.startswith("k_final") and ns > 20e3 for n, ns in s)]
c = max(1, args.chunks_per_frame)
timed = complete[args.warmup * c:(args.warmup + args.steps) * c]
tot = OrderedDict()
for s in timed:
    for name, ns in s:
        t = tot.setdefault(name, [0, 0.0])
        t[0] += 1
        t[1] += ns
total = sum(v[1] for v in tot.values()) or 1.0
n_l = sum(v[0] for v in tot.values())
print(f"# per-kernel totals of the {args.steps} timed frames ({n_l} launches) of `python bench.py --steps {args.steps} --warmup {args.warmup} "
      f"--no-cpu-baseline` under")
print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares, not absolutes)")
for name, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:28s} n={n:4d} total={ns / 1e6:9.3f} ms share={ns / total * 100:5.1f}%")
print(f"# all: {total / 1e6:.3f} ms for {args.steps} frames; {len(seqs)} chunk sequences in the list, {len(complete)} complete")
