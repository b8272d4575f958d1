#!/usr/bin/env python
"""Renders a few device-resident frames of one scene: the command profiled under ncu."""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import euclider_b200 as eb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="3d_room")
ap.add_argument("--width", type=int, default=3840)
ap.add_argument("--height", type=int, default=2160)
ap.add_argument("--frames", type=int, default=2)
ap.add_argument("--pipeline", default="wavefront")
ap.add_argument("--band-world", type=int, default=1, help="render only the bands of rank 0 of this many ranks (16-row bands)")
args = ap.parse_args()
env = eb.load_reference_scene(args.scene)
env.pipeline = eb.EUCL_PIPELINE_MEGAKERNEL if args.pipeline == "megakernel" else eb.EUCL_PIPELINE_WAVEFRONT
out = torch.empty((args.height, args.width, 3), dtype=torch.uint8, device="cuda")
for i in range(args.frames):
    st = env.render_device(out.data_ptr(), (args.width, args.height), 0.0, profile=(i == args.frames - 1),
                           band_rows=16 if args.band_world > 1 else 0, band_rank=0, band_world=args.band_world)
print({k: st[k] for k in ("segments", "launches", "ms_total", "ms_intersect", "ms_shade", "ms_resolve", "level_counts")})
