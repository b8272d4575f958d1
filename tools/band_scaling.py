#!/usr/bin/env python
"""Frame time of rank 0's share of a band-split frame on ONE GPU, for several world sizes: separates the per-frame fixed
cost (launch boundaries, kernel tails) from the per-pixel work.   python tools/band_scaling.py [--scene 3d_room]"""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import euclider_b200 as eb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="3d_room")
ap.add_argument("--width", type=int, default=3840)
ap.add_argument("--height", type=int, default=2160)
ap.add_argument("--worlds", default="1,2,4,8,16,32")
args = ap.parse_args()
out = torch.empty((args.height, args.width, 3), dtype=torch.uint8, device="cuda")
rows = []
for world in [int(v) for v in args.worlds.split(",")]:
    env = eb.load_reference_scene(args.scene)
    best, st = 1e9, None
    for i in range(18):
        st = env.render_device(out.data_ptr(), (args.width, args.height), 0.0, band_rows=16 if world > 1 else 0, band_rank=0,
                               band_world=world)
        if i >= 9:
            best = min(best, st["ms_total"])
    rows.append((world, st["pixels"], st["segments"], best, st["launches"], st["graph_replays"]))
    print(f"world {world:3d} pixels {st['pixels']:9d} segments {st['segments']:9d} ms {best:7.3f} launches {st['launches']} graph {st['graph_replays']}", flush=True)
    env.close()
# least squares ms = F + c * segments
import numpy as np
x = np.array([r[2] for r in rows], dtype=float)
y = np.array([r[3] for r in rows], dtype=float)
A = np.stack([np.ones_like(x), x], axis=1)
(f, c), *_ = np.linalg.lstsq(A, y, rcond=None)
print(f"fit: ms = {f:.3f} + {c * 1e6:.4f} per Msegment   (fixed cost per frame {f * 1e3:.0f} us)")
