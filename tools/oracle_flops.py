#!/usr/bin/env python
"""F_oracle: counted floating-point operations per ray segment of the CPU oracle (counting build,
oracle/Makefile: liboracle_count.so) next to the algorithmic bound F_scene of SURVEY.md 8(d).

    python tools/oracle_flops.py [--width 384 --height 216]

CPU only.  Vector arithmetic is counted exactly, scalar tails of the intersectors by constants, one libm
call as one operation; colour arithmetic is not counted (see FLOPS in oracle/oracle.cc)."""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import euclider_b200 as eb  # noqa: E402
import oracle_api  # noqa: E402
from bench import scene_flops_per_segment  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--width", type=int, default=384)
ap.add_argument("--height", type=int, default=216)
args = ap.parse_args()
print("| scene | F_scene (bound) | F_oracle (counted) | ratio | segments / pixel |")
print("|---|---|---|---|---|")
for scene in ("3d_fresnel", "3d_room", "3d_hallways", "4d_frame", "4d_cylinders", "4d_room"):
    env = eb.load_reference_scene(scene)
    _, _, st = oracle_api.render(env, args.width, args.height, time=0.0, variant="count")
    f_scene = scene_flops_per_segment(env)
    f_oracle = st["flops"] / max(st["segments"], 1)
    print(f"| {scene} | {f_scene} | {f_oracle:.0f} | {f_oracle / f_scene:.2f} | {st['segments'] / (args.width * args.height):.2f} |")
