#!/usr/bin/env python
"""Headline benchmark: Mrays/s (ray segments per second) and frame time on a reference scene.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scene 3d_room]
                    [--width 3840 --height 2160] [--pipeline wavefront|megakernel]

A "step" is one fixed-pose headless frame (Environment::render, src/universe/mod.rs:300-357).
 * value    : whole-job ray segments / second with the frame left in HBM (eucl_render_device)
 * e2e      : the same through the reference-facing call with a HOST output buffer
              (Environment.render -> eucl_render; frame copied device->host every step)
 * roofline : FP64 issue roofline of the dominant kernel; the peak is
              measured live by a DADD/DMUL/DFMA microbenchmark
 * cpu_baseline : the CPU oracle (restatement of the reference, oracle/) on a bounded sample
`--impl reference` times the CPU oracle alone with all host threads (the Rust reference cannot be
built here: no rustc/cargo).
N > 1: one process per GPU (torchrun); interleaved row bands, one gather to rank 0, no other
collective -- weak/strong: the frame is fixed, so scaling is "strong".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

# algorithmic lower-bound flops per ray segment: every primitive of every surfaced
# entity is tested once per segment; Sphere 7D+4, (half)plane 4D, Cylinder 15D+2
PRIM_FLOPS = {1: lambda d: 7 * d + 4, 2: lambda d: 4 * d, 3: lambda d: 4 * d, 4: lambda d: 15 * d + 2}

DEFAULT_CONFIGS = {
    "3d_fresnel": (1920, 1080), "3d_room": (3840, 2160), "3d_hallways": (3840, 2160), "4d_frame": (3840, 2160),
    "4d_cylinders": (3840, 2160), "4d_room": (7680, 4320),
}


def scene_flops_per_segment(env) -> int:
    flat = env.flat
    total = 0
    for e in range(flat.n_entities):
        ent = flat.entities[e]
        if ent.surface < 0:
            continue
        for n in range(ent.node_first, ent.node_root + 1):
            node = flat.nodes[n]
            if node.op == 0:
                kind = flat.prims[node.prim].kind
                if kind in PRIM_FLOPS:
                    total += PRIM_FLOPS[kind](flat.dim)
    return total


class ClockSampler:
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_sample(env, width, height, t, n_blocks=9, rows_per_block=8, threads=None):
    """Times the CPU oracle on `n_blocks` row blocks spread over the frame; returns (segments, seconds, desc)."""
    import oracle_api

    threads = threads or os.cpu_count() or 1
    rows_per_block = min(rows_per_block, height)
    n_blocks = max(1, min(n_blocks, height // rows_per_block))
    starts = [int((k + 0.5) * height / n_blocks - rows_per_block / 2) for k in range(n_blocks)]
    segs, secs = 0, 0.0
    for r0 in starts:
        r0 = max(0, min(height - rows_per_block, r0))
        t0 = time.perf_counter()
        _, _, st = oracle_api.render(env, width, height, time=t, threads=threads, rows=(r0, r0 + rows_per_block))
        secs += time.perf_counter() - t0
        segs += st["segments"]
    return segs, secs, f"{n_blocks} blocks x {rows_per_block} rows of the {width}x{height} frame, spread evenly", threads


def run_reference(args, width, height):
    """--impl reference: the CPU restatement of the reference's path, all host threads."""
    import euclider_b200 as eb

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    env = eb.load_reference_scene(args.scene)
    if args.max_depth:
        env.camera.max_depth = args.max_depth
    threads = os.cpu_count() or 1
    blocks = 9
    for _ in range(args.warmup):
        oracle_sample(env, width, height, args.time, n_blocks=2, rows_per_block=2)
    seg_total, sec_total, desc = 0, 0.0, ""
    for _ in range(args.steps):
        segs, secs, desc, threads = oracle_sample(env, width, height, args.time, n_blocks=blocks, rows_per_block=args.ref_rows)
        seg_total += segs
        sec_total += secs
    value = seg_total / sec_total / 1e6
    # frame time extrapolated from the sampled rows
    sample_pixels = blocks * args.ref_rows * width
    ms_frame = sec_total / args.steps * 1e3 * (width * height) / sample_pixels
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_frame, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.scene} {width}x{height} max_depth {env.camera.max_depth} time {args.time}",
                   "note": "ms_per_step extrapolated from the sampled rows to the full frame"},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_RESULT_OUT = sys.stdout


def emit(obj) -> None:
    """The result lines go to the process's ORIGINAL stdout (see main)."""
    print(json.dumps(obj), file=_RESULT_OUT, flush=True)


def main():
    # stdout carries the JSON result lines and nothing else: libraries that write to file descriptor 1 on their own
    # (NCCL prints "NCCL version ..." there) are sent to stderr instead
    global _RESULT_OUT
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scene", default="3d_room")
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--time", type=float, default=0.0)
    ap.add_argument("--max-depth", type=int, default=0)
    ap.add_argument("--pipeline", default="wavefront", choices=["wavefront", "megakernel"])
    ap.add_argument("--band-rows", type=int, default=16)
    ap.add_argument("--ref-rows", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="N > 1: ranks store their rows into GPU 0's frame through CUDA IPC (peer) or NCCL gather")
    ap.add_argument("--check", action="store_true", help="N > 1: compare the gathered frame with a single-GPU render")
    args = ap.parse_args()
    width, height = DEFAULT_CONFIGS.get(args.scene, (3840, 2160))
    width, height = args.width or width, args.height or height

    if args.impl == "reference":
        run_reference(args, width, height)
        return

    import torch
    import torch.distributed as dist

    import euclider_b200 as eb
    from euclider_b200 import bands

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    env = eb.load_reference_scene(args.scene)
    if args.max_depth:
        env.camera.max_depth = args.max_depth
    env.pipeline = eb.EUCL_PIPELINE_MEGAKERNEL if args.pipeline == "megakernel" else eb.EUCL_PIPELINE_WAVEFRONT
    stream = torch.cuda.current_stream()
    env.set_stream(stream.cuda_stream, device=local_rank)

    band_rows = args.band_rows if world > 1 else 0
    opts = eb.EuclRenderOpts(width=width, height=height, band_rows=band_rows, band_rank=rank, band_world=world)
    my_rows = int(eb.lib().eucl_band_rows_for_rank(opts)) if world > 1 else height
    frame_bytes = height * width * 3
    peer = world > 1 and args.gather == "peer"
    frame_ptr, frame = 0, None
    if world == 1:
        frame = torch.empty((height, width, 3), dtype=torch.uint8, device=device)
        frame_ptr = frame.data_ptr()
    elif peer:
        # One frame buffer on GPU 0, mapped into every rank through CUDA IPC: each rank's pack kernel
        # stores its rows straight into it over NVLink, so the gather IS the last kernel of the frame.
        import ctypes as C

        handle = torch.zeros(eb._capi.EUCL_IPC_HANDLE_BYTES, dtype=torch.uint8, device=device)
        ptr = C.c_void_p()
        if rank == 0:
            eb._capi.check(eb.lib().eucl_device_malloc(local_rank, frame_bytes, C.byref(ptr)))
            buf = (C.c_uint8 * eb._capi.EUCL_IPC_HANDLE_BYTES)()
            eb._capi.check(eb.lib().eucl_ipc_export(ptr, buf))
            handle.copy_(torch.tensor(list(buf), dtype=torch.uint8))
        dist.broadcast(handle, src=0)
        if rank != 0:
            buf = (C.c_uint8 * eb._capi.EUCL_IPC_HANDLE_BYTES)(*handle.cpu().tolist())
            eb._capi.check(eb.lib().eucl_ipc_open(buf, local_rank, C.byref(ptr)))
        frame_ptr = ptr.value
        if rank == 0:  # torch view of the raw allocation (for the copy to the host)
            class _Raw:
                __cuda_array_interface__ = {"shape": (height, width, 3), "typestr": "|u1", "data": (frame_ptr, False),
                                            "version": 2}
            frame = torch.as_tensor(_Raw(), device=device)
    else:
        # NCCL gather of compact row blocks, scattered to frame rows on rank 0
        d_rows = torch.empty((max(my_rows, 1), width, 3), dtype=torch.uint8, device=device)
        frame = torch.empty((height, width, 3), dtype=torch.uint8, device=device) if rank == 0 else None
        rows_of = []
        for r in range(world):
            idx = bands.local_rows(height, band_rows, r, world)
            rows_of.append(torch.tensor(idx, dtype=torch.long, device=device))
        max_rows = max(len(x) for x in rows_of)
        send = torch.zeros((max_rows, width, 3), dtype=torch.uint8, device=device)
        recv = [torch.empty_like(send) for _ in range(world)] if rank == 0 else None
    host = torch.empty((height, width, 3), dtype=torch.uint8).pin_memory() if rank == 0 else None

    def step(profile=False):
        if world == 1 or peer:
            st = env.render_device(frame_ptr, (width, height), args.time, device=local_rank, band_rows=band_rows,
                                   band_rank=rank, band_world=world, compact_rows=False, profile=profile)
            if peer:
                dist.barrier()  # every rank's rows have landed in GPU 0's frame (render_device is synchronous)
            return st
        st = env.render_device(d_rows.data_ptr(), (width, height), args.time, device=local_rank, band_rows=band_rows,
                               band_rank=rank, band_world=world, compact_rows=True, profile=profile)
        send[:my_rows].copy_(d_rows[:my_rows])
        dist.gather(send, recv, dst=0)
        if rank == 0:
            for r in range(world):
                frame.index_copy_(0, rows_of[r], recv[r][:len(rows_of[r])])
        return st

    def frame_to_host():
        """rank 0: device frame -> pinned host buffer (the e2e leg)."""
        host.copy_(frame, non_blocking=False)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    segs, launches, retries = 0, 0, 0
    ev0.record(stream)
    for _ in range(args.steps):
        st = step()
        segs += st["segments"]
        launches += st["launches"]
        retries += st["retries"]
    ev1.record(stream)
    sync_all()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms, float(segs), float(launches)], dtype=torch.float64, device=device)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, segs, launches = float(tmax[0]), int(tsum[1]), int(tsum[2])
    ms_per_step = ms / args.steps
    value = segs / (ms * 1e-3) / 1e6

    if rank == 0 and world == 1:
        # profile pass: per-kernel-family device times (extra events; not part of the timed region)
        prof = step(profile=True)
        # end-to-end through the reference-facing call with a pinned HOST frame buffer
        host_np = host.numpy()
        for _ in range(3):
            env.render((width, height), args.time, device=local_rank, out=host_np)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_segs = 0
        for _ in range(args.steps):
            img = env.render((width, height), args.time, device=local_rank, out=host_np)
            e2e_segs += img.stats["segments"]
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        e2e = {"value": e2e_segs / e2e_s / 1e6, "unit": "Mrays/s", "ms_per_step": e2e_s / args.steps * 1e3,
               "h2d_bytes_per_step": 256, "d2h_bytes_per_step": width * height * 3}
    elif rank == 0:
        prof = None
        e2e = None  # N > 1: measured below (bands -> GPU 0 -> pinned host)
    if world > 1:
        # e2e at N GPUs: bands -> gather -> host copy on rank 0, timed by wall clock between barriers
        sync_all()
        t0 = time.perf_counter()
        e2e_segs = 0
        for _ in range(args.steps):
            st = step()
            e2e_segs += st["segments"]
            if rank == 0:
                frame_to_host()
        sync_all()
        e2e_s = time.perf_counter() - t0
        tt = torch.tensor([e2e_s, float(e2e_segs)], dtype=torch.float64, device=device)
        tm = tt.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ts = tt.clone()
        dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        if rank == 0:
            e2e = {"value": float(ts[1]) / float(tm[0]) / 1e6, "unit": "Mrays/s", "ms_per_step": float(tm[0]) / args.steps * 1e3,
                   "h2d_bytes_per_step": 256 * world, "d2h_bytes_per_step": width * height * 3}

    if rank == 0:
        import ctypes as C

        f_scene = scene_flops_per_segment(env)
        dadd, dmul, dfma = C.c_double(), C.c_double(), C.c_double()
        eb.lib().eucl_fp64_peak(local_rank, C.byref(dadd), C.byref(dmul), C.byref(dfma))
        peak = max(dadd.value, dmul.value)  # non-fused FP64 instruction rate, T op/s (-fmad=false build)
        seg_per_frame = segs / args.steps / world if world > 1 else segs / args.steps
        roofline = {"bound": "fp64_issue", "unit": "Tflop/s", "peak": peak, "peak_source": "measured live (DADD/DMUL microbenchmark)",
                    "peak_dfma_tops": dfma.value, "traffic": None, "flops_per_segment": f_scene}
        if args.scene == "3d_room" and (width, height) == (3840, 2160):
            # dram__bytes_read.sum + dram__bytes_write.sum of ONE k_intersect launch (level 8 of a whole-frame chunk,
            # 4.15 M rays) from the committed ncu --set full capture (profiles/r1d_ncu_k_intersect_level8.txt).
            # Algorithmic bytes of that launch: none beyond its queue records (ray 52 B in, hit 64 B out per ray =
            # 482 MB); the frame itself is 3 B/pixel, written by k_final.
            roofline["traffic"] = 353.219328e6 + 294.449664e6
            roofline["traffic_note"] = "one level-8 k_intersect launch, 4.15 M rays, ncu --set full capture (profiles/r1d_*)"
        if prof is not None and prof["ms_intersect"] > 0:
            achieved = prof["segments"] * f_scene / (prof["ms_intersect"] * 1e-3) / 1e12
            roofline.update({"kernel": "k_intersect (all levels of one frame)", "achieved": achieved, "frac": achieved / peak,
                             "kernel_ms_per_frame": prof["ms_intersect"],
                             "family_ms": {k: prof[k] for k in ("ms_raygen", "ms_intersect", "ms_shade", "ms_resolve", "ms_total")}})
        else:
            achieved = value * 1e6 * f_scene / 1e12
            roofline.update({"kernel": "whole frame", "achieved": achieved, "frac": achieved / (peak * world)})
        cpu = None
        if not args.no_cpu_baseline and world == 1:  # reported at N = 1 only
            try:
                c_segs, c_secs, desc, threads = oracle_sample(env, width, height, args.time, rows_per_block=args.ref_rows)
                cpu = {"value": c_segs / c_secs / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": desc}
            except Exception as exc:  # the oracle is test infrastructure; its absence must not hide the GPU number
                cpu = {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "port", "sample": f"unavailable: {exc}"}
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.scene} {width}x{height} max_depth {env.camera.max_depth} time {args.time}",
                       "pipeline": args.pipeline, "segments_per_frame": seg_per_frame * (world if world > 1 else 1),
                       "parallelism": (f"row bands of {band_rows} x {world} ranks, "
                                       + ("peer stores into GPU 0's frame (CUDA IPC over NVLink)" if peer else "NCCL gather to rank 0"))
                       if world > 1 else "single GPU",
                       "l2": "node arena (GBs per chunk) is far larger than the 126 MB L2; no explicit flush"},
            "fps": 1e3 / ms_per_step, "e2e": e2e, "gpu_launches": launches, "retries": retries, "clocks": clocks,
            "roofline": roofline, "cpu_baseline": cpu,
        }
        emit(line)
    if world > 1 and args.check:
        step()
        sync_all()
        if rank == 0:
            frame_to_host()
            env.set_stream(0, device=local_rank)
            whole = env.render((width, height), args.time, device=local_rank)
            same = bool(np.array_equal(whole.data, host.numpy()))
            emit({"check": "gathered frame == single-GPU frame", "equal": same})
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
