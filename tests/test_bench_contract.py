"""The CPU-runnable parts of the bench contract: the reference arm prints exactly one JSON line on stdout with
the agreed keys; the committed launch list summarises to the kernel shares quoted in profiles/README.md."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_one_json_line(built_lib, oracle):
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--ref-rows", "1", "--scene", "3d_fresnel"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mrays/s" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("3d_fresnel 1920x1080")


def test_reference_arm_renders_whole_frames_by_default(built_lib, oracle):
    """Without --ref-rows the reference arm renders every row of every timed frame: ms_per_step is measured, not
    extrapolated (small frame here so that the CPU suite stays short)."""
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1",
                        "--scene", "3d_fresnel", "--width", "160", "--height", "90"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.strip()][0])
    assert d["steps"] == 3 and d["cpu_baseline"]["sample"] == "whole 160x90 frames"
    assert "measured" in d["config"]["note"] and d["ms_per_step"] > 0
    # segments per frame are deterministic: value * time reproduces an integer multiple of one frame's count
    segs = d["value"] * 1e6 * d["ms_per_step"] * 1e-3
    assert abs(segs - round(segs)) < 1e-3 * max(1.0, segs)


def test_reference_arm_is_silent_on_other_ranks(built_lib, oracle):
    """Under torchrun only rank 0 runs the reference arm; the other ranks exit 0 without work."""
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                       text=True, timeout=120, env={**__import__("os").environ, "RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_committed_launch_list_matches_its_summary():
    csv = ROOT / "profiles" / "r1d_launches_bench_3d_room_4k.csv"
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "launch_summary.py"), str(csv), "--warmup", "3", "--steps", "2",
                        "--chunks-per-frame", "1"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout == (ROOT / "profiles" / "r1d_launches_bench_3d_room_4k_summary.txt").read_text()
    shares = {}
    for ln in r.stdout.splitlines():
        if "share=" in ln:
            family = ln.split("<")[0].split()[0]
            shares[family] = shares.get(family, 0.0) + float(ln.split("share=")[1].rstrip("%"))
    assert 40 < shares["k_intersect"] < 48 and 42 < shares["k_shade"] < 50 and shares["k_resolve"] + shares["k_final"] < 12


def test_split_launch_list_matches_its_summary():
    """The launch list of the final build: two pipelines per frame interleave in it (two k_raygen ... two k_final)."""
    csv = ROOT / "profiles" / "r2_launches_bench_3d_room_4k_split.csv"
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "launch_summary.py"), str(csv), "--warmup", "8", "--steps", "2"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    committed = (ROOT / "profiles" / "r2_launches_bench_3d_room_4k_split_summary.txt").read_text().splitlines()
    assert r.stdout.splitlines() == [ln for ln in committed if not ln.startswith("# EUCL_GRAPH=0")]
    counts, shares = {}, {}
    for ln in r.stdout.splitlines():
        if "share=" in ln:
            family = ln.split("<")[0].split()[0]
            counts[family] = counts.get(family, 0) + int(ln.split("n=")[1].split()[0])
            shares[family] = shares.get(family, 0.0) + float(ln.split("share=")[1].rstrip("%"))
    assert counts["k_raygen"] == 4 and counts["k_final"] == 4  # 2 frames x 2 pipelines
    assert 40 < shares["k_intersect"] < 50 and 42 < shares["k_shade"] < 52 and shares["k_resolve"] + shares["k_final"] < 8
