"""CUDA path vs the CPU oracle, through the C ABI (eucl_render / eucl_render_device).

Bars: bit-exact RGB8, primary hit-entity maps and per-level node counts against the `det` oracle
(same deterministic libm, include/eucl_detmath.h); against the `glibc` oracle (the reference
platform's libm) the north-star tolerance: >= 99.5 % of pixels within 1/255 per channel and
identical hit-entity maps."""
from pathlib import Path

import numpy as np
import pytest

import euclider_b200 as eb

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
REF_SCENES = ["3d_fresnel", "3d_room", "3d_hallways", "4d_frame", "4d_cylinders", "4d_room", "3d_frame", "3d_fresnel_2",
              "3d_photo", "4d_fresnel"]
OWN_SCENES = ["csg_mix_3d", "blend_4d", "no_void_3d"]
PIPELINES = {"wavefront": eb.EUCL_PIPELINE_WAVEFRONT, "megakernel": eb.EUCL_PIPELINE_MEGAKERNEL}


def load(name):
    if name in OWN_SCENES:
        return eb.Parser.default(resource_root=ROOT).parse_file(ROOT / "tests" / "scenes" / f"{name}.json")
    return eb.load_reference_scene(name)


def assert_exact(img, ref_rgb, ref_hit, ref_stats):
    assert np.array_equal(img.hit_ids, ref_hit)
    assert img.stats["level_counts"] == ref_stats["level_counts"]
    assert img.stats["segments"] == ref_stats["segments"]
    assert np.array_equal(img.data, ref_rgb)


def within_one(a, b):
    return float((np.abs(a.astype(np.int16) - b.astype(np.int16)).max(axis=-1) <= 1).mean())


@pytest.mark.parametrize("pipeline", sorted(PIPELINES))
@pytest.mark.parametrize("name", REF_SCENES + OWN_SCENES)
def test_scene_bit_exact(built_lib, oracle, name, pipeline):
    env = load(name)
    env.pipeline = PIPELINES[pipeline]
    w, h, t = 160, 90, 1.234
    img = env.render((w, h), time=t, want_hit_ids=True)
    assert_exact(img, *oracle.render(env, w, h, time=t, variant="det"))
    gl_rgb, gl_hit, _ = oracle.render(env, w, h, time=t, variant="glibc")
    assert np.array_equal(img.hit_ids, gl_hit)
    assert within_one(img.data, gl_rgb) >= 0.995


@pytest.mark.parametrize("path", sorted((ROOT / "tests" / "golden").glob("*.npz")), ids=lambda p: p.stem)
def test_against_committed_golden(built_lib, path):
    g = np.load(path)
    env = load(path.stem.rsplit("_", 1)[0])
    img = env.render((int(g["width"]), int(g["height"])), time=float(g["time"]), want_hit_ids=True)
    assert np.array_equal(img.data, g["rgb_det"])
    assert np.array_equal(img.hit_ids.astype(np.int8), g["hit_det"])
    assert img.stats["level_counts"] == g["levels_det"].tolist()


@pytest.mark.parametrize("size", [(1, 1), (33, 17), (241, 135), (64, 3)])
def test_odd_and_tiny_frames(built_lib, oracle, size):
    """Odd sizes hit rel = 0 columns/rows exactly (d3/entity/camera.rs:170-173)."""
    env = load("3d_room")
    img = env.render(size, time=0.0, want_hit_ids=True)
    assert_exact(img, *oracle.render(env, size[0], size[1], time=0.0))


@pytest.mark.parametrize("depth", [0, 1, 2, 16])
def test_max_depth_is_a_render_parameter(built_lib, oracle, depth):
    env = load("4d_room")
    env.camera.max_depth = depth
    img = env.render((128, 72), want_hit_ids=True)
    assert_exact(img, *oracle.render(env, 128, 72))
    assert img.stats["levels"] == depth + 1


def test_moved_camera_and_resolution_divisor(built_lib, oracle):
    env = load("4d_cylinders")
    env.camera.location[0], env.camera.location[1], env.camera.location[2], env.camera.location[3] = -7.0, 0.5, 0.8, 0.6
    img = env.render((512, 256), time=0.5, context=eb.SimulationContext(resolution=4), want_hit_ids=True)
    assert (img.width, img.height) == (128, 64)
    assert_exact(img, *oracle.render(env, 128, 64, time=0.5))
    assert (img.hit_ids >= 0).any()


def test_time_only_feeds_perlin_truncated_to_ms(built_lib, oracle):
    env = load("3d_room")
    a = env.render((96, 54), time=1.2341)
    b = env.render((96, 54), time=1.2349)  # same millisecond
    c = env.render((96, 54), time=1.2351)
    assert np.array_equal(a.data, b.data) and not np.array_equal(a.data, c.data)


def test_bands_reassemble_to_the_whole_frame(built_lib):
    """Interleaved row bands rendered rank by rank into one buffer == the single-call frame."""
    env = load("3d_hallways")
    w, h = 120, 67
    whole = env.render((w, h), want_hit_ids=True)
    out = np.zeros((h, w, 3), np.uint8)
    segs = 0
    for rank in range(3):
        part = env.render((w, h), band_rows=8, band_rank=rank, band_world=3, out=out)
        segs += part.stats["segments"]
    assert np.array_equal(out, whole.data) and segs == whole.stats["segments"]


def test_tiny_arena_forces_retries_but_not_errors(built_lib, oracle, monkeypatch):
    monkeypatch.setenv("EUCL_ARENA_FACTOR_X10", "11")
    monkeypatch.setenv("EUCL_CHUNK_PIXELS", "4096")
    env = load("3d_fresnel_2")  # the ray tree grows with depth: level 10 alone exceeds the pixel count
    img = env.render((160, 90), want_hit_ids=True)
    assert img.stats["retries"] > 0
    assert_exact(img, *oracle.render(env, 160, 90))


def test_full_size_rows_match_the_oracle(built_lib, oracle):
    """BASELINE config 3d_room at 3840x2160: the oracle renders sampled rows of the full-size frame."""
    env = load("3d_room")
    w, h = 3840, 2160
    img = env.render((w, h), want_hit_ids=True)
    for r0 in (0, 700, 1079, 1080, 1500, 2159):
        rgb, hit, _ = oracle.render(env, w, h, rows=(r0, r0 + 1))
        assert np.array_equal(img.data[r0:r0 + 1], rgb) and np.array_equal(img.hit_ids[r0:r0 + 1], hit)
    again = env.render((w, h))
    assert np.array_equal(again.data, img.data)  # idempotent
    # device memory of the whole-frame arena: 172 B per ray-tree node (ray, hit, node record, colour) plus index lists that
    # hold one level each; the second frame has given back what the learning frame over-allocated
    mem = env.memory()
    assert img.stats["retries"] >= 1 and again.stats["retries"] == 0
    assert mem["node_capacity"] >= img.stats["nodes"] and mem["node_capacity"] <= 1.1 * img.stats["nodes"]
    assert mem["arena_bytes"] + mem["list_bytes"] < 11 * 2**30
    print(f"3d_room 4K arena: {mem['arena_bytes'] / 2**30:.2f} GiB nodes + {mem['list_bytes'] / 2**30:.2f} GiB lists, "
          f"{mem['node_capacity']} nodes for {img.stats['nodes']} used")
    env.pipeline = eb.EUCL_PIPELINE_MEGAKERNEL
    mega = env.render((w, h))
    assert np.array_equal(mega.data, img.data) and mega.stats["level_counts"] == img.stats["level_counts"]


@pytest.mark.parametrize("name,size,depth", [("3d_fresnel", (1920, 1080), 10), ("3d_hallways", (3840, 2160), 10),
                                             ("4d_frame", (3840, 2160), 10), ("4d_cylinders", (3840, 2160), 10),
                                             ("4d_room", (7680, 4320), 10), ("4d_room", (7680, 4320), 16)])
def test_baseline_configs_at_full_size(built_lib, oracle, name, size, depth):
    """Every BASELINE.json config at its full resolution: sampled rows against the oracle (bit-exact),
    idempotence, and the segment count of the sampled rows."""
    env = load(name)
    env.camera.max_depth = depth
    w, h = size
    img = env.render((w, h), want_hit_ids=True)
    for r0 in (0, h // 3, h // 2 - 1, h // 2, h - 1):
        rgb, hit, _ = oracle.render(env, w, h, rows=(r0, r0 + 1))
        assert np.array_equal(img.data[r0:r0 + 1], rgb) and np.array_equal(img.hit_ids[r0:r0 + 1], hit)
    assert img.stats["pixels"] == w * h and img.stats["level_counts"][0] == w * h
    again = env.render((w, h))
    assert np.array_equal(again.data, img.data)


def test_device_buffer_entry_point(built_lib):
    import torch

    env = load("3d_fresnel")
    w, h = 200, 100
    host = env.render((w, h))
    d = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
    env.set_stream(torch.cuda.current_stream().cuda_stream)
    st = env.render_device(d.data_ptr(), (w, h))
    torch.cuda.synchronize()
    assert np.array_equal(d.cpu().numpy(), host.data) and st["segments"] == host.stats["segments"]


def test_errors_are_statuses_not_aborts(built_lib):
    env = load("3d_fresnel")
    for depth in (100, 64, 2**32 - 1):  # EUCL_MAX_LEVELS = 64 levels = max_depth <= 63; UINT32_MAX must not wrap
        env.camera.max_depth = depth
        with pytest.raises(eb.EuclError) as err:
            env.render((8, 8))
        assert err.value.status == -1


@pytest.mark.parametrize("name", ["3d_room", "3d_hallways"])
def test_turned_cameras_3d(built_lib, oracle, name):
    """Poses reached through the rotation entry points (yaw, pitch, roll: csrc/host/camera.cc): oblique
    camera frames through ray generation, both grouping modes of the wavefront (the first frames of a scene
    alternate between them while it tunes itself)."""
    env = load(name)
    env.rotate_yaw(0.4)
    env.rotate_pitch(-0.25)
    env.rotate_roll(0.3)
    ref = oracle.render(env, 128, 72, time=0.5, variant="det")
    for _ in range(4):
        assert_exact(env.render((128, 72), time=0.5, want_hit_ids=True), *ref)


def test_turned_camera_4d(built_lib, oracle):
    env = load("4d_room")
    env.rotate_plane4(0, 3, 0.35)  # forward towards ana: the W axis comes into view
    env.rotate_plane4(1, 2, -0.2)
    ref = oracle.render(env, 128, 72, time=0.0, variant="det")
    assert_exact(env.render((128, 72), time=0.0, want_hit_ids=True), *ref)


def test_ray_grouping_modes_are_identical(built_lib, oracle, monkeypatch):
    """EUCL_BIN_RAYS only reorders the rays of a level: same picture, same counts, either way."""
    env = load("3d_room")
    ref = oracle.render(env, 160, 90, time=0.75, variant="det")
    for mode in ("0", "1"):
        monkeypatch.setenv("EUCL_BIN_RAYS", mode)
        assert_exact(load("3d_room").render((160, 90), time=0.75, want_hit_ids=True), *ref)


def test_chunks_shrink_to_the_memory_budget(built_lib, oracle, monkeypatch):
    """A frame whose node arena does not fit the budget is rendered in smaller chunks of rows, not refused
    (EUCL_ARENA_MAX_MB stands in for a full device); the picture and the counts do not depend on the split."""
    env = load("3d_room")
    ref = oracle.render(env, 160, 90, time=0.25, variant="det")
    whole = env.render((160, 90), time=0.25, want_hit_ids=True)
    assert_exact(whole, *ref)
    monkeypatch.setenv("EUCL_ARENA_MAX_MB", "3")
    split = load("3d_room").render((160, 90), time=0.25, want_hit_ids=True)
    assert_exact(split, *ref)
    assert split.stats["launches"] > 2 * whole.stats["launches"]
    monkeypatch.setenv("EUCL_ARENA_MAX_MB", "1")  # not even one row fits: a status, not a crash
    with pytest.raises(eb.EuclError) as err:
        load("3d_room").render((3840, 90), time=0.25)
    assert err.value.status == -31


def test_repeated_frames_replay_as_a_cuda_graph(built_lib, oracle, monkeypatch):
    """A chunk whose launch parameters repeat is captured once and replayed as one CUDA graph launch
    (api_device.cu: render_impl); a new pose or time goes back to direct launches.  Same picture either way."""
    import torch

    env = load("3d_room")
    w, h = 128, 72
    ref = oracle.render(env, w, h, time=0.25, variant="det")
    d = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
    replays = []
    for _ in range(8):  # the first frames of a scene also settle its arena size and its ray-grouping mode
        st = env.render_device(d.data_ptr(), (w, h), 0.25)
        replays.append(st["graph_replays"])
        assert st["segments"] == ref[2]["segments"] and st["launches"] > 20
        assert np.array_equal(d.cpu().numpy(), ref[0])
    assert replays[0] == 0 and replays[-1] == 1 and replays[-2] == 1
    st = env.render_device(d.data_ptr(), (w, h), 0.75)  # another time: other parameters, direct launches again
    assert st["graph_replays"] == 0
    assert np.array_equal(d.cpu().numpy(), oracle.render(env, w, h, time=0.75, variant="det")[0])
    monkeypatch.setenv("EUCL_GRAPH", "0")
    for _ in range(3):
        st = env.render_device(d.data_ptr(), (w, h), 0.75)
        assert st["graph_replays"] == 0
