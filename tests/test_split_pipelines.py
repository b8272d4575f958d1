"""A rank's share of a frame rendered as TWO concurrent pipelines (api_device.cu: render_split) is the same picture,
the same hit map and the same node counts as one pipeline -- bit for bit, for whole frames, band-split frames
(frame-row and compact layouts), ragged heights, both precisions, and through both entry points."""
from pathlib import Path

import numpy as np
import pytest

import euclider_b200 as eb

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(params=[2, 3])
def split(monkeypatch, request):
    monkeypatch.setenv("EUCL_SPLIT", str(request.param))
    monkeypatch.setenv("EUCL_SPLIT_MIN_PIXELS", "0")


def one_pipeline(env, size, **kw):
    import os
    saved = os.environ.get("EUCL_SPLIT")
    os.environ["EUCL_SPLIT"] = "1"
    try:
        return env.render(size, **kw)
    finally:
        if saved is None:
            del os.environ["EUCL_SPLIT"]
        else:
            os.environ["EUCL_SPLIT"] = saved


@pytest.mark.parametrize("name,size", [("3d_room", (160, 90)), ("4d_room", (96, 67)), ("3d_hallways", (64, 33)),
                                       ("4d_frame", (50, 32)), ("3d_fresnel", (40, 31))])
def test_split_frame_equals_oracle(built_lib, oracle, split, name, size):
    env = eb.load_reference_scene(name)
    w, h = size
    img = env.render(size, time=1.234, want_hit_ids=True)
    rgb, hit, stats = oracle.render(env, w, h, time=1.234, variant="det")
    assert np.array_equal(img.hit_ids, hit) and np.array_equal(img.data, rgb)
    assert img.stats["level_counts"] == stats["level_counts"] and img.stats["segments"] == stats["segments"]
    assert img.stats["pixels"] == w * h
    if h >= 32:  # two bands of 16 rows at least: really split
        assert img.stats["launches"] > one_pipeline(env, size).stats["launches"]


def test_split_repeated_frames_replay_graphs(built_lib, split):
    env = eb.load_reference_scene("3d_room")
    first = env.render((128, 80), want_hit_ids=True)
    st = None
    for _ in range(8):
        again = env.render((128, 80), want_hit_ids=True)
        st = again.stats
        assert np.array_equal(again.data, first.data) and np.array_equal(again.hit_ids, first.hit_ids)
    assert st["graph_replays"] == 1 and st["retries"] == 0
    whole = one_pipeline(env, (128, 80), want_hit_ids=True)
    assert np.array_equal(whole.data, first.data) and np.array_equal(whole.hit_ids, first.hit_ids)
    assert whole.stats["segments"] == first.stats["segments"]


@pytest.mark.parametrize("world,band", [(2, 8), (3, 4), (4, 16)])
def test_split_bands_reassemble(built_lib, split, world, band):
    """Host entry point, band-split: every rank renders its bands as two pipelines into a compact device buffer and
    scatters them into the one frame."""
    env = eb.load_reference_scene("4d_room")
    w, h = 72, 131  # ragged: the last band is partial
    whole = one_pipeline(env, (w, h), want_hit_ids=True)
    out = np.zeros((h, w, 3), np.uint8)
    segs = 0
    for rank in range(world):
        part = env.render((w, h), band_rows=band, band_rank=rank, band_world=world, out=out)
        segs += part.stats["segments"]
    assert np.array_equal(out, whole.data) and segs == whole.stats["segments"]


def test_split_device_entry_point_both_layouts(built_lib, split):
    import torch

    env = eb.load_reference_scene("3d_hallways")
    w, h, world, band = 96, 83, 2, 8
    whole = torch.from_numpy(one_pipeline(env, (w, h)).data).cuda()
    frame = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
    for rank in range(world):
        env.render_device(frame.data_ptr(), (w, h), band_rows=band, band_rank=rank, band_world=world)
    assert torch.equal(frame, whole)
    for rank in range(world):  # compact: the rank's bands one after the other
        rows = [r for r in range(h) if (r // band) % world == rank]
        compact = torch.zeros((len(rows), w, 3), dtype=torch.uint8, device="cuda")
        st = env.render_device(compact.data_ptr(), (w, h), band_rows=band, band_rank=rank, band_world=world, compact_rows=True)
        assert st["pixels"] == len(rows) * w
        assert torch.equal(compact, whole[rows])


def test_split_low_precision(built_lib, oracle, split):
    env = eb.load_reference_scene("3d_room")
    env.precision = "f32"
    img = env.render((120, 70), time=0.5, want_hit_ids=True)
    rgb, hit, stats = oracle.render(env, 120, 70, time=0.5, variant="f32")
    assert np.array_equal(img.hit_ids, hit) and np.array_equal(img.data, rgb)
    assert img.stats["segments"] == stats["segments"]


def test_split_memory_is_reported_for_both_arenas(built_lib, split):
    env = eb.load_reference_scene("3d_room")
    for _ in range(3):  # (the frame after a learning frame re-allocates its arena to fit)
        env.render((256, 160))
    both = env.memory()
    env2 = eb.load_reference_scene("3d_room")
    for _ in range(3):
        one_pipeline(env2, (256, 160))
    one = env2.memory()
    assert both["node_capacity"] > 0 and one["node_capacity"] > 0
    assert both["arena_bytes"] < 1.3 * one["arena_bytes"] + (1 << 20)  # two half arenas, not two whole ones


def test_split_profile_mode_and_retries(built_lib, oracle, split, monkeypatch):
    monkeypatch.setenv("EUCL_ARENA_FACTOR_X10", "11")
    env = eb.load_reference_scene("3d_fresnel_2")
    img = env.render((160, 96), want_hit_ids=True, profile=True)
    rgb, hit, stats = oracle.render(env, 160, 96)
    assert img.stats["retries"] > 0
    assert np.array_equal(img.data, rgb) and np.array_equal(img.hit_ids, hit)
    assert img.stats["level_counts"] == stats["level_counts"]
