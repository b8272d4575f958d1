"""ctypes binding of oracle/liboracle.so -- the CPU restatement used as the parity checker.

Test infrastructure: imported only by tests/, __graft_entry__.smoke() and bench.py's CPU legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

from euclider_b200._capi import EUCL_MAX_LEVELS, EuclCamera, EuclFlatScene, EuclPrim

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
LIB_PATHS = {"glibc": ORACLE_DIR / "liboracle.so", "det": ORACLE_DIR / "liboracle_det.so",
             "count": ORACLE_DIR / "liboracle_count.so",  # "count": det + flop counters (tools/oracle_flops.py)
             "f32": ORACLE_DIR / "liboracle_f32.so"}  # the reference's `low_precision` feature (type F = f32), det libm
_libs = {}

dptr = C.POINTER(C.c_double)


def build(target: str = "all") -> None:
    subprocess.run(["make", "-C", str(ORACLE_DIR), target], check=True, capture_output=True)


def lib(variant: str = "det") -> C.CDLL:
    """variant "det": transcendental functions from include/eucl_detmath.h (bit-comparable with the
    CUDA path); "glibc": the host libm (what the Rust reference would link on this platform); "f32": the `det` build
    with `type F = f32` (renderer and camera path only: the unit-test hooks probe the f64 builds)."""
    if variant not in _libs:
        path = LIB_PATHS[variant]
        deps = [ORACLE_DIR / "oracle.cc", ROOT / "include" / "euclider_b200.h", ROOT / "include" / "eucl_detmath.h"]
        if not path.exists() or path.stat().st_mtime < max(d.stat().st_mtime for d in deps):
            build("liboracle_count.so" if variant == "count" else "all")
        h = C.CDLL(str(path))
        h.oracle_render.restype = C.c_int
        h.oracle_render.argtypes = [C.POINTER(EuclFlatScene), C.POINTER(EuclCamera), C.c_uint32, C.c_uint32, C.c_double,
                                    C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        if variant != "f32":
            h.oracle_entity_intersections.restype = C.c_int
            h.oracle_entity_intersections.argtypes = [C.POINTER(EuclFlatScene), C.c_int, dptr, dptr, C.c_int, dptr]
            h.oracle_prim_intersect.restype = C.c_int
            h.oracle_prim_intersect.argtypes = [C.c_int, C.POINTER(EuclPrim), dptr, dptr, dptr]
            h.oracle_prim_inside.restype = C.c_int
            h.oracle_prim_inside.argtypes = [C.c_int, C.POINTER(EuclPrim), dptr]
            h.oracle_entity_inside.restype = C.c_int
            h.oracle_entity_inside.argtypes = [C.POINTER(EuclFlatScene), C.c_int, dptr]
            h.oracle_material_at.restype = C.c_int
            h.oracle_material_at.argtypes = [C.POINTER(EuclFlatScene), dptr]
            h.oracle_angle_between.restype = C.c_double
            h.oracle_angle_between.argtypes = [C.c_int, dptr, dptr]
            h.oracle_angle_between_f32.restype = C.c_float
            h.oracle_angle_between_f32.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float)]
            h.oracle_combine_palette_color.restype = None
            h.oracle_combine_palette_color.argtypes = [dptr, dptr, C.c_double, dptr]
            h.oracle_combine_palette_color_f32.restype = None
            h.oracle_combine_palette_color_f32.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float,
                                                           C.POINTER(C.c_float)]
            h.oracle_remainder_i.restype = C.c_int64
            h.oracle_remainder_i.argtypes = [C.c_int64, C.c_int64]
            h.oracle_remainder_f.restype = C.c_double
            h.oracle_remainder_f.argtypes = [C.c_double, C.c_double]
            h.oracle_blend.restype = None
            h.oracle_blend.argtypes = [C.c_int, C.c_double, dptr, dptr, dptr]
            h.oracle_to_pixel.restype = None
            h.oracle_to_pixel.argtypes = [dptr, C.POINTER(C.c_uint8)]
            h.oracle_perlin4.restype = C.c_double
            h.oracle_perlin4.argtypes = [C.POINTER(C.c_uint8), dptr]
            h.oracle_hsv_to_rgb.restype = None
            h.oracle_hsv_to_rgb.argtypes = [C.c_double, C.c_double, C.c_double, dptr]
            h.oracle_general_rotation.restype = None
            h.oracle_general_rotation.argtypes = [C.c_int, dptr, dptr, C.c_double, dptr, dptr]
            h.oracle_surface_probe.restype = None
            h.oracle_surface_probe.argtypes = [C.POINTER(EuclFlatScene), C.c_int, dptr, dptr, C.c_int, dptr, dptr, dptr]
            h.oracle_mapped_color.restype = None
            h.oracle_mapped_color.argtypes = [C.POINTER(EuclFlatScene), C.c_int, dptr, dptr]
            h.oracle_ray_vector.restype = None
            h.oracle_ray_vector.argtypes = [C.POINTER(EuclFlatScene), C.POINTER(EuclCamera), C.c_int, C.c_int, C.c_int,
                                            C.c_int, dptr]
            h.oracle_detmath_unary.restype = None
            h.oracle_detmath_unary.argtypes = [C.c_int, dptr, dptr, C.c_int]
            h.oracle_detmath_atan2.restype = None
            h.oracle_detmath_atan2.argtypes = [dptr, dptr, dptr, C.c_int]
        h.oracle_trace_path.restype = C.c_int
        h.oracle_trace_path.argtypes = [C.POINTER(EuclFlatScene), dptr, dptr, C.c_double, dptr, dptr]
        h.oracle_uses_detmath.restype = C.c_int
        h.oracle_real_bytes.restype = C.c_int
        _libs[variant] = h
    return _libs[variant]


def darr(values):
    values = list(values)
    return (C.c_double * len(values))(*values)


def render(env, width: int, height: int, time: float = 0.0, threads: int | None = None, rows=None, camera=None,
           variant: str = "det"):
    """Oracle frame for an euclider_b200.Environment: (rgb uint8 [rows,w,3], hit int32 [rows,w], stats dict)."""
    threads = threads or os.cpu_count() or 1
    r0, r1 = rows if rows is not None else (0, height)
    rgb = np.zeros((r1 - r0, width, 3), dtype=np.uint8)
    hit = np.zeros((r1 - r0, width), dtype=np.int32)
    stats = np.zeros(8 + EUCL_MAX_LEVELS, dtype=np.uint64)
    cam = camera if camera is not None else env.camera
    flat = env.flat
    rc = lib(variant).oracle_render(C.byref(flat), C.byref(cam), width, height, float(time), r0, r1, threads,
                             rgb.ctypes.data, hit.ctypes.data, stats.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"oracle_render failed: {rc}")
    levels = int(cam.max_depth) + 1
    return rgb, hit, {
        "segments": int(stats[0]), "nodes": int(stats[1]), "nan_channel": int(stats[2]), "bad_texcoord": int(stats[3]),
        "no_material": int(stats[4]), "csg_runaway": int(stats[5]), "flops": int(stats[6]),
        "level_counts": [int(v) for v in stats[8:8 + levels]],
    }


def trace_path(env, location, direction, distance: float, variant: str = "det"):
    """Universe::trace_path_unknown on the oracle: (location, direction) or None."""
    dim = env.dim
    out_l, out_d = (C.c_double * dim)(), (C.c_double * dim)()
    flat = env.flat
    rc = lib(variant).oracle_trace_path(C.byref(flat), darr(location), darr(direction), float(distance), out_l, out_d)
    if rc == 1:
        return None
    if rc != 0:
        raise RuntimeError("oracle_trace_path: runaway")
    return list(out_l), list(out_d)
