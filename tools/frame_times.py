#!/usr/bin/env python
"""Device-resident frame times of the config scenes under different tuning switches (environment
variables are read by the library at every render call).

    python tools/frame_times.py [--scenes 3d_room,4d_room] [--set EUCL_BIN_RAYS=0,1] [--frames 6]"""
import argparse
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import euclider_b200 as eb  # noqa: E402

SIZES = {"3d_fresnel": (1920, 1080), "3d_room": (3840, 2160), "3d_hallways": (3840, 2160), "4d_frame": (3840, 2160),
         "4d_cylinders": (3840, 2160), "4d_room": (7680, 4320)}
ap = argparse.ArgumentParser()
ap.add_argument("--scenes", default=",".join(SIZES))
ap.add_argument("--set", default="", help="VAR=v1,v2,... : one run per value (empty value = unset)")
ap.add_argument("--frames", type=int, default=6)
args = ap.parse_args()
var, values = None, [None]
if args.set:
    var, vs = args.set.split("=", 1)
    values = [v if v != "" else None for v in vs.split(",")]
for scene in args.scenes.split(","):
    w, h = SIZES[scene]
    out = torch.empty((h, w, 3), dtype=torch.uint8, device="cuda")
    for v in values:
        if var:
            if v is None:
                os.environ.pop(var, None)
            else:
                os.environ[var] = v
        env = eb.load_reference_scene(scene)  # a fresh scene per setting (arena, auto-tuning state)
        best, last = 1e9, None
        for i in range(args.frames + 4):
            st = env.render_device(out.data_ptr(), (w, h), 0.0, profile=(i == args.frames + 3))
            if i >= 4 and i < args.frames + 3:
                best = min(best, st["ms_total"])
            last = st
        print(f"{scene:13s} {var or ''}={v}  best {best:8.3f} ms | profiled: total {last['ms_total']:.2f} intersect {last['ms_intersect']:.2f} "
              f"shade {last['ms_shade']:.2f} resolve {last['ms_resolve']:.2f} | segments {last['segments']} retries {last['retries']}", flush=True)
        env.close()
