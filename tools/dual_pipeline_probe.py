#!/usr/bin/env python
"""Does running rank 0's share of a band-split frame as TWO independent pipelines (two scenes, two streams, two host
threads, each on half of the bands) hide the per-stage latency of the level loop?  Wall clock over K frames, one GPU.
    python tools/dual_pipeline_probe.py [--scene 3d_room] [--worlds 1,8]"""
import argparse
import sys
import threading
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import euclider_b200 as eb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="3d_room")
ap.add_argument("--width", type=int, default=3840)
ap.add_argument("--height", type=int, default=2160)
ap.add_argument("--worlds", default="1,2,4,8")
ap.add_argument("--frames", type=int, default=30)
ap.add_argument("--split", type=int, default=2)
args = ap.parse_args()
out = torch.empty((args.height, args.width, 3), dtype=torch.uint8, device="cuda")
size = (args.width, args.height)


def run(env, rank, world, frames):
    st = None
    for _ in range(frames):
        st = env.render_device(out.data_ptr(), size, 0.0, band_rows=16 if world > 1 else 0, band_rank=rank, band_world=world)
    return st


for world in [int(v) for v in args.worlds.split(",")]:
    # (a) one pipeline
    env = eb.load_reference_scene(args.scene)
    run(env, 0, world, 12)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st = run(env, 0, world, args.frames)
    torch.cuda.synchronize()
    one = (time.perf_counter() - t0) / args.frames * 1e3
    seg_one = st["segments"]
    env.close()
    # (b) `split` pipelines on interleaved halves of the same bands
    k = args.split
    envs = [eb.load_reference_scene(args.scene) for _ in range(k)]
    ranks = [j * world for j in range(k)]
    for e, r in zip(envs, ranks):
        run(e, r, world * k, 12)
    torch.cuda.synchronize()
    segs = [0] * k

    def body(j):
        segs[j] = run(envs[j], ranks[j], world * k, args.frames)["segments"]

    threads = [threading.Thread(target=body, args=(j,)) for j in range(k)]
    t0 = time.perf_counter()
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    torch.cuda.synchronize()
    two = (time.perf_counter() - t0) / args.frames * 1e3
    for e in envs:
        e.close()
    print(f"world {world:2d}: one pipeline {one:7.3f} ms/frame ({seg_one} segments) | {k} pipelines {two:7.3f} ms/frame "
          f"({sum(segs)} segments)", flush=True)
