#!/usr/bin/env python
"""Renders every reference scene on the GPU (both pipelines) and compares with the CPU oracle."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import euclider_b200 as eb  # noqa: E402
import oracle_api  # noqa: E402

SCENES = ["3d_fresnel", "3d_room", "3d_hallways", "4d_frame", "4d_cylinders", "4d_room", "3d_frame", "3d_fresnel_2",
          "3d_photo", "4d_fresnel"]


def compare(a, b):
    diff = np.abs(a.astype(np.int16) - b.astype(np.int16)).max(axis=-1)
    return float((diff == 0).mean()), float((diff <= 1).mean()), int(diff.max())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=320)
    ap.add_argument("--height", type=int, default=180)
    ap.add_argument("--scenes", nargs="*", default=SCENES)
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "gpu_check.json"))
    args = ap.parse_args()
    results = {}
    for name in args.scenes:
        env = eb.load_reference_scene(name)
        t0 = time.time()
        ref_rgb, ref_hit, ref_stats = oracle_api.render(env, args.width, args.height, time=1.234, variant="det")
        t_cpu = time.time() - t0
        gl_rgb, gl_hit, gl_stats = oracle_api.render(env, args.width, args.height, time=1.234, variant="glibc")
        entry = {"oracle_s": t_cpu, "oracle_stats": ref_stats, "glibc_stats": gl_stats}
        for pipe_name, pipe in (("wavefront", eb.EUCL_PIPELINE_WAVEFRONT), ("megakernel", eb.EUCL_PIPELINE_MEGAKERNEL)):
            env.pipeline = pipe
            img = env.render((args.width, args.height), time=1.234, want_hit_ids=True)
            img = env.render((args.width, args.height), time=1.234, want_hit_ids=True)
            exact, within1, maxdiff = compare(img.data, ref_rgb)
            hit_same = float((img.hit_ids == ref_hit).mean())
            g_exact, g_within1, g_maxdiff = compare(img.data, gl_rgb)
            entry[pipe_name] = {"exact": exact, "within1": within1, "maxdiff": maxdiff, "hit_same": hit_same,
                                "glibc_exact": g_exact, "glibc_within1": g_within1, "glibc_maxdiff": g_maxdiff,
                                "glibc_hit_same": float((img.hit_ids == gl_hit).mean()),
                                "levels_match": img.stats["level_counts"] == ref_stats["level_counts"],
                                "segments": img.stats["segments"], "level_counts": img.stats["level_counts"],
                                "ms_total": img.stats["ms_total"], "retries": img.stats["retries"],
                                "segments_match": img.stats["segments"] == ref_stats["segments"]}
            print(name, pipe_name, json.dumps(entry[pipe_name]), flush=True)
        results[name] = entry
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps(results, indent=1))


if __name__ == "__main__":
    main()
