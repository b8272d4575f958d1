"""Host-side mirror of the reference interface for the trace-loop path.

Reference (Rust): ``scene::Parser::default().parse::<Box<Environment>>(json)`` (src/scene.rs:564,
1466-1478) builds an ``Environment``; ``Environment::render(dimensions, time, threads, context)``
(src/universe/mod.rs:300-357) renders one frame into a ``RawImage2d<u8>`` (RGB8, row 0 = bottom).
The same names and argument meanings are kept here; the work is done by libeuclider_b200.so
(C ABI in include/euclider_b200.h) on a B200.  There is no CPU rendering path in this package.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, Optional, Sequence, Tuple, Union

import numpy as np

from . import _capi
from ._capi import (EuclCamera, EuclError, EuclFlatScene, EuclRenderOpts, EuclStats, EUCL_PIPELINE_MEGAKERNEL,
                    EUCL_PIPELINE_WAVEFRONT, check, lib)

ParserError = EuclError  # the reference's `ParserError` (src/scene.rs:524-552); see EuclError.status

# `resources/universe_dim.jpg` (background of 3d_room / 3d_hallways) is absent from the reference
# checkout (.MISSING_LARGE_BLOBS); the declared substitute is used by the library and the oracle alike.
DEFAULT_TEXTURE_SUBSTITUTES = {"universe_dim.jpg": "universe_bright.jpg", "moon.jpg": "universe_bright.jpg"}


@dataclass
class SimulationContext:
    """The fields of the reference's SimulationContext that `render` reads (src/simulation.rs:167-186)."""
    resolution: int = 1  # the reference defaults to 8 (window / 8); headless renders use 1
    debugging: bool = False


@dataclass
class RawImage2d:
    """glium's RawImage2d<u8> as returned by Environment::render: RGB8, row 0 = bottom."""
    data: np.ndarray  # uint8, shape (height, width, 3)
    width: int
    height: int
    format: str = "U8U8U8"
    hit_ids: Optional[np.ndarray] = None  # int32 (height, width): primary hit entity, -1 background, -2 checkerboard
    stats: Optional[dict] = None

    def to_top_down(self) -> np.ndarray:
        return self.data[::-1]

    def save(self, path: Union[str, os.PathLike]) -> None:
        """Writes the frame top-down: `.ppm` through the library (eucl_write_ppm), anything else via Pillow."""
        path = str(path)
        if path.lower().endswith(".ppm"):
            check(lib().eucl_write_ppm(path.encode(), self.width, self.height, np.ascontiguousarray(self.data).ctypes.data))
        else:
            from PIL import Image

            Image.fromarray(np.ascontiguousarray(self.to_top_down())).save(path)


def _decode_image(path: Path) -> Tuple[int, int, bytes]:
    """Decodes to RGBA8, row 0 = top -- the layout `image::DynamicImage::get_pixel` exposes
    (alpha = 255 for RGB / L images, grey replicated)."""
    from PIL import Image

    with Image.open(path) as im:
        rgba = im.convert("RGBA")
        return rgba.width, rgba.height, rgba.tobytes()


def stats_to_dict(st: EuclStats) -> dict:
    levels = int(st.levels)
    return {
        "pixels": int(st.pixels), "segments": int(st.segments), "nodes": int(st.nodes),
        "level_counts": [int(st.level_counts[i]) for i in range(levels)], "levels": levels,
        "retries": int(st.retries), "launches": int(st.launches), "ray_grouping": int(st.ray_grouping), "ms_total": float(st.ms_total),
        "ms_raygen": float(st.ms_raygen), "ms_intersect": float(st.ms_intersect), "ms_shade": float(st.ms_shade),
        "ms_resolve": float(st.ms_resolve), "graph_replays": int(st.graph_replays),
    }


class Environment:
    """A parsed universe (Universe3 / Universe4).  Mirrors `trait Environment` (src/universe/mod.rs:289-360)."""

    def __init__(self, parsed_handle: int, texture_paths: Sequence[str]):
        self._parsed = C.c_void_p(parsed_handle)
        self._scenes: Dict[Tuple[int, str], C.c_void_p] = {}
        self.texture_paths = list(texture_paths)
        flat = lib().eucl_parsed_flat(self._parsed)
        self.camera: EuclCamera = flat.contents.camera.copy()  # pose the JSON constructs; mutable by the caller
        self.pipeline = EUCL_PIPELINE_WAVEFRONT
        # the reference's scalar type F: "f64" (default build) or "f32" (cargo feature `low_precision`, src/main.rs:46-49)
        self.precision = "f64"

    # -- reference API ---------------------------------------------------------------------------
    def max_depth(self) -> int:
        return int(self.camera.max_depth)

    def render(self, dimensions: Tuple[int, int], time: float = 0.0, threads: int = 0,
               context: Optional[SimulationContext] = None, *, device: int = 0, want_hit_ids: bool = False,
               band_rows: int = 0, band_rank: int = 0, band_world: int = 1, out: Optional[np.ndarray] = None,
               profile: bool = False) -> RawImage2d:
        """Environment::render.  `time` is seconds since start (the reference passes a Duration);
        `threads` is accepted for signature parity and ignored (the GPU schedules the pixels)."""
        del threads
        context = context or SimulationContext()
        width, height = int(dimensions[0]) // context.resolution, int(dimensions[1]) // context.resolution
        opts = EuclRenderOpts(width=width, height=height, time_seconds=float(time), band_rows=band_rows,
                              band_rank=band_rank, band_world=band_world, pipeline=self.pipeline, compact_rows=0,
                              want_hit_ids=int(want_hit_ids), profile=int(profile))
        if out is None:
            out = np.zeros((height, width, 3), dtype=np.uint8)
        assert out.dtype == np.uint8 and out.size == height * width * 3 and out.flags["C_CONTIGUOUS"]
        hit = np.full((height, width), -3, dtype=np.int32) if want_hit_ids else None
        stats = EuclStats()
        check(lib().eucl_render(self._device_scene(device), C.byref(self.camera), C.byref(opts), out.ctypes.data,
                                hit.ctypes.data if hit is not None else None, C.byref(stats)))
        return RawImage2d(out, width, height, hit_ids=hit, stats=stats_to_dict(stats))

    # -- device-resident variant (inputs and outputs stay in HBM) --------------------------------
    def render_device(self, d_out_rgb8: int, dimensions: Tuple[int, int], time: float = 0.0, *, device: int = 0,
                      d_out_hit_ids: int = 0, band_rows: int = 0, band_rank: int = 0, band_world: int = 1,
                      compact_rows: bool = False, profile: bool = False) -> dict:
        width, height = int(dimensions[0]), int(dimensions[1])
        opts = EuclRenderOpts(width=width, height=height, time_seconds=float(time), band_rows=band_rows,
                              band_rank=band_rank, band_world=band_world, pipeline=self.pipeline,
                              compact_rows=int(compact_rows), want_hit_ids=int(bool(d_out_hit_ids)),
                              profile=int(profile))
        stats = EuclStats()
        check(lib().eucl_render_device(self._device_scene(device), C.byref(self.camera), C.byref(opts),
                                       C.c_void_p(d_out_rgb8), C.c_void_p(d_out_hit_ids) if d_out_hit_ids else None,
                                       C.byref(stats)))
        return stats_to_dict(stats)

    def trace_path_unknown(self, location: Sequence[float], direction: Sequence[float], distance: float, *,
                           device: int = 0) -> Optional[Tuple[list, list]]:
        """Universe::trace_path_unknown (src/universe/mod.rs:273-286): the point reached by travelling
        `distance` along `direction` through the universe (voids stretch or shrink the step) and the
        direction of travel there; None when `location` lies in no entity.  The reference's cameras
        move with this call (d3/entity/camera.rs:223-243)."""
        dim = self.dim
        loc = (C.c_double * dim)(*[float(v) for v in location[:dim]])
        dirv = (C.c_double * dim)(*[float(v) for v in direction[:dim]])
        out_l, out_d = (C.c_double * dim)(), (C.c_double * dim)()
        status = lib().eucl_trace_path(self._device_scene(device), loc, dirv, float(distance), out_l, out_d)
        if status == 1:
            return None
        check(status)
        return list(out_l), list(out_d)

    def move_camera(self, direction: Sequence[float], distance: float, *, device: int = 0) -> bool:
        """Translates `self.camera.location` like the translation part of the reference's camera
        `update` (d3/entity/camera.rs:223-243): normalised direction, trace_path_unknown, new location.
        (The reference additionally turns the view when the void bends the path; it labels that code
        "not tested", and it is not reproduced here.)  Returns False when the camera is in no entity."""
        dim = self.dim
        norm = sum(float(v) * float(v) for v in direction[:dim]) ** 0.5
        if norm == 0.0:
            return True
        unit = [float(v) / norm for v in direction[:dim]]
        res = self.trace_path_unknown([self.camera.location[k] for k in range(dim)], unit, distance * norm, device=device)
        if res is None:
            return False
        for k in range(dim):
            self.camera.location[k] = res[0][k]
        return True

    # The view-turning part of the reference's camera `update` (d3/entity/camera.rs:94-145,303-350;
    # d4/entity/camera.rs:68-130).  The reference derives the angles from mouse deltas and held keys;
    # here the caller passes them (radians).  `kind` names the reference camera type being mimicked.
    def rotate_yaw(self, angle: float, kind: str = "PitchYawCamera3") -> None:
        check(lib().eucl_camera_rotate_yaw(C.byref(self.camera), float(angle), 0 if kind == "PitchYawCamera3" else 1))

    def rotate_pitch(self, angle: float, kind: str = "PitchYawCamera3") -> None:
        check(lib().eucl_camera_rotate_pitch(C.byref(self.camera), float(angle), 1 if kind == "PitchYawCamera3" else 0))

    def rotate_roll(self, angle: float) -> None:
        check(lib().eucl_camera_rotate_roll(C.byref(self.camera), float(angle)))

    def rotate_plane4(self, axis_a: int, axis_b: int, angle: float) -> None:
        """FreeCamera4: rotation in the plane of two camera axes (0 forward, 1 left, 2 up, 3 ana)."""
        check(lib().eucl_camera_rotate_plane4(C.byref(self.camera), int(axis_a), int(axis_b), float(angle)))

    # -- plumbing ----------------------------------------------------------------------------------
    @property
    def flat(self) -> EuclFlatScene:
        return lib().eucl_parsed_flat(self._parsed).contents

    @property
    def dim(self) -> int:
        return int(self.flat.dim)

    def _device_scene(self, device: int) -> C.c_void_p:
        key = (device, self.precision)
        if key not in self._scenes:
            if self.precision not in ("f64", "f32"):
                raise ValueError("precision must be 'f64' or 'f32'")
            handle = C.c_void_p()
            check(lib().eucl_scene_create_precision(lib().eucl_parsed_flat(self._parsed), device,
                                                    _capi.EUCL_PRECISION_F32 if self.precision == "f32" else _capi.EUCL_PRECISION_F64,
                                                    C.byref(handle)))
            self._scenes[key] = handle
        return self._scenes[key]

    def set_texture(self, slot: int, width: int, height: int, rgba8: bytes) -> None:
        """Fills texture slot `slot` with decoded RGBA8 pixels (row 0 = top); for callers that decode
        images themselves.  Must happen before the first render."""
        assert len(rgba8) == width * height * 4
        check(lib().eucl_parsed_set_texture(self._parsed, slot, width, height, rgba8))

    def memory(self, device: int = 0) -> dict:
        """Device memory this environment's scene holds for frames: node arena, index lists, node capacity."""
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        check(lib().eucl_scene_memory(self._device_scene(device), C.byref(a), C.byref(b), C.byref(c)))
        return {"arena_bytes": a.value, "list_bytes": b.value, "node_capacity": c.value}

    def set_stream(self, cuda_stream: int, device: int = 0) -> None:
        """Runs this environment's kernels on `cuda_stream` (e.g. torch.cuda.current_stream().cuda_stream)."""
        check(lib().eucl_scene_set_stream(self._device_scene(device), C.c_void_p(cuda_stream)))

    def close(self) -> None:
        for handle in self._scenes.values():
            lib().eucl_scene_destroy(handle)
        self._scenes.clear()
        if self._parsed:
            lib().eucl_parsed_destroy(self._parsed)
            self._parsed = C.c_void_p()

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass


@dataclass
class Parser:
    """scene::Parser (src/scene.rs:554-1478).  `resource_root` plays the role of the reference's
    working directory for the `./resources/...` texture paths (src/scene.rs:1050-1072)."""
    resource_root: Union[str, os.PathLike, None] = None
    texture_substitutes: Dict[str, str] = field(default_factory=lambda: dict(DEFAULT_TEXTURE_SUBSTITUTES))

    @classmethod
    def default(cls, resource_root: Union[str, os.PathLike, None] = None) -> "Parser":
        return cls(resource_root=resource_root)

    def _resolve(self, texture_path: str) -> Path:
        root = Path(self.resource_root) if self.resource_root is not None else Path.cwd()
        p = root / texture_path
        if not p.exists() and p.name in self.texture_substitutes:
            p = p.with_name(self.texture_substitutes[p.name])
        if not p.exists():
            raise FileNotFoundError(f"texture `{texture_path}` not found under {root}")
        return p

    def parse(self, json_text: str, *, load_textures: bool = True) -> Environment:
        handle = C.c_void_p()
        check(lib().eucl_scene_parse(json_text.encode("utf-8"), C.byref(handle)))
        n = lib().eucl_parsed_texture_count(handle)
        paths = [lib().eucl_parsed_texture_path(handle, k).decode("utf-8") for k in range(n)]
        if load_textures:
            try:
                for slot, tex in enumerate(paths):
                    w, h, data = _decode_image(self._resolve(tex))
                    check(lib().eucl_parsed_set_texture(handle, slot, w, h, data))
            except Exception:
                lib().eucl_parsed_destroy(handle)
                raise
        return Environment(handle.value, paths)

    def parse_file(self, path: Union[str, os.PathLike], **kw) -> Environment:
        return self.parse(Path(path).read_text(), **kw)
