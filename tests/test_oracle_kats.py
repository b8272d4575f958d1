"""The oracle against the reference's OWN known-answer tests for this path.

Sources (relative to the reference's src/): entity/shape.rs:1040-1149 (four primitive
intersection KATs in 2-D), util.rs:1007-1037 (angle_between), util.rs:947-958
(combine_palette_color), util.rs:960-969 (remainder).  These are the only tests the reference has
on the trace-loop path; both oracle builds (glibc libm / deterministic libm) must satisfy them.
"""
import ctypes as C
import math

import numpy as np
import pytest

from euclider_b200._capi import EuclPrim, PRIM_CYLINDER, PRIM_HALFSPACE, PRIM_HYPERPLANE, PRIM_SPHERE

VARIANTS = ["det", "glibc"]


def ulps_apart(a: float, b: float, f32: bool = False) -> int:
    dt, it = (np.float32, np.int32) if f32 else (np.float64, np.int64)
    ia, ib = np.array(a, dt).view(it).item(), np.array(b, dt).view(it).item()
    return abs(ia - ib)


def prim(kind, v0=(0, 0), v1=(0, 0), s0=0.0, s1=0.0):
    p = EuclPrim()
    p.kind = kind
    for k, v in enumerate(v0):
        p.v0[k] = v
    for k, v in enumerate(v1):
        p.v1[k] = v
    p.s0, p.s1 = s0, s1
    return p


def intersect(oracle, variant, dim, p, loc, direction):
    out = (C.c_double * (2 * (1 + 2 * dim)))()
    n = oracle.lib(variant).oracle_prim_intersect(dim, C.byref(p), oracle.darr(loc), oracle.darr(direction), out)
    vals = list(out)
    w = 1 + 2 * dim
    return [dict(distance=vals[i * w], location=vals[i * w + 1:i * w + 1 + dim], normal=vals[i * w + 1 + dim:(i + 1) * w])
            for i in range(n)]


@pytest.mark.parametrize("variant", VARIANTS)
def test_intersect_sphere_linear(oracle, variant):  # shape.rs:1048-1073
    hits = intersect(oracle, variant, 2, prim(PRIM_SPHERE, v0=(2, 0), s0=1.0), (0, 0), (1, 0))
    assert len(hits) == 2
    assert hits[0]["location"] == [1.0, 0.0] and hits[0]["normal"] == [-1.0, 0.0]
    assert ulps_apart(hits[0]["distance"], 1.0) <= 2
    assert hits[1]["location"] == [3.0, 0.0] and hits[1]["normal"] == [1.0, 0.0]
    assert ulps_apart(hits[1]["distance"], 3.0) <= 2


@pytest.mark.parametrize("variant", VARIANTS)
def test_intersect_plane_linear(oracle, variant):  # shape.rs:1075-1096
    # Hyperplane::new_with_point(normal (-1, 0), point (1, 0)): constant = -(n . p) = 1
    hits = intersect(oracle, variant, 2, prim(PRIM_HYPERPLANE, v0=(-1, 0), s0=1.0), (0, 0), (1, 0))
    assert len(hits) == 1
    assert hits[0]["location"] == [1.0, 0.0] and hits[0]["normal"] == [-1.0, 0.0]
    assert ulps_apart(hits[0]["distance"], 1.0) <= 2


@pytest.mark.parametrize("variant", VARIANTS)
def test_intersect_halfspace_linear(oracle, variant):  # shape.rs:1098-1121
    # HalfSpace::new_with_point(plane, inside (2, 0)): signum = sign(n . p + c) = sign(-2 + 1) = -1
    hits = intersect(oracle, variant, 2, prim(PRIM_HALFSPACE, v0=(-1, 0), s0=1.0, s1=-1.0), (0, 0), (1, 0))
    assert len(hits) == 1
    assert hits[0]["location"] == [1.0, 0.0] and hits[0]["normal"] == [-1.0, 0.0]
    assert ulps_apart(hits[0]["distance"], 1.0) <= 2


@pytest.mark.parametrize("variant", VARIANTS)
def test_intersect_cylinder_linear(oracle, variant):  # shape.rs:1123-1148
    hits = intersect(oracle, variant, 2, prim(PRIM_CYLINDER, v0=(2, 0), v1=(0, 1), s0=1.0), (0, 0), (1, 0))
    assert len(hits) == 2
    assert hits[0]["location"] == [1.0, 0.0] and hits[0]["normal"] == [-1.0, 0.0]
    assert ulps_apart(hits[0]["distance"], 1.0) <= 2
    assert hits[1]["location"] == [3.0, 0.0] and hits[1]["normal"] == [1.0, 0.0]
    assert ulps_apart(hits[1]["distance"], 3.0) <= 2


@pytest.mark.parametrize("dim", [3, 4])
@pytest.mark.parametrize("variant", VARIANTS)
def test_primitive_kats_zero_padded(oracle, variant, dim):
    """The same four KATs restated in the dimensions the renderer uses (zero padded)."""
    pad = lambda v: tuple(v) + (0,) * (dim - len(v))
    for p in (prim(PRIM_SPHERE, v0=pad((2, 0)), s0=1.0), prim(PRIM_CYLINDER, v0=pad((2, 0)), v1=pad((0, 1)), s0=1.0)):
        hits = intersect(oracle, variant, dim, p, pad((0, 0)), pad((1, 0)))
        assert [h["distance"] for h in hits] == [1.0, 3.0]
        assert hits[0]["normal"] == list(pad((-1.0, 0.0))) and hits[1]["normal"] == list(pad((1.0, 0.0)))
    for p in (prim(PRIM_HYPERPLANE, v0=pad((-1, 0)), s0=1.0), prim(PRIM_HALFSPACE, v0=pad((-1, 0)), s0=1.0, s1=-1.0)):
        hits = intersect(oracle, variant, dim, p, pad((0, 0)), pad((1, 0)))
        assert len(hits) == 1 and hits[0]["distance"] == 1.0 and hits[0]["normal"] == list(pad((-1.0, 0.0)))


def test_angle_between_f32(oracle):  # util.rs:1007-1037 (the reference runs this test in f32)
    f = oracle.lib("glibc").oracle_angle_between_f32
    f3 = lambda v: (C.c_float * 3)(*v)
    cases = [((1, 0, 0), (0, 1, 0), math.pi / 2), ((1, 0, 0), (1, 1, 0), math.pi / 4),
             ((1, 0, 0), (-1, 1, 0), 3 * math.pi / 4), ((1, 0, 0), (-1, 0, 0), math.pi)]
    for a, b, want in cases:
        assert ulps_apart(f(f3(a), f3(b)), np.float32(want), f32=True) <= 2


@pytest.mark.parametrize("variant", VARIANTS)
def test_angle_between_f64(oracle, variant):
    f = oracle.lib(variant).oracle_angle_between
    for a, b, want in [((1, 0, 0), (0, 1, 0), math.pi / 2), ((1, 0, 0), (1, 1, 0), math.pi / 4),
                       ((1, 0, 0), (-1, 1, 0), 3 * math.pi / 4), ((1, 0, 0), (-1, 0, 0), math.pi)]:
        assert ulps_apart(f(3, oracle.darr(a), oracle.darr(b)), want) <= 2
    # NaN -> 0 (util.rs:717-721): zero vector
    assert f(3, oracle.darr((0, 0, 0)), oracle.darr((1, 0, 0))) == 0.0


def test_combine_palette_color(oracle):  # util.rs:947-958, f32 like the reference's test
    f = oracle.lib("glibc").oracle_combine_palette_color_f32
    a, b, out = (C.c_float * 4)(1.0, 0.5, 0.0, 1.0), (C.c_float * 4)(0.0, 1.0, 0.5, 0.5), (C.c_float * 4)()
    f(a, b, C.c_float(1.0 / 3.0), out)
    third = np.float32(1.0) / np.float32(3.0)
    want = [third, np.float32(0.5) * third + np.float32(1.0) * (np.float32(1.0) - third),
            np.float32(0.5) * (np.float32(1.0) - third), third + np.float32(0.5) * (np.float32(1.0) - third)]
    for got, w in zip(out, want):
        assert ulps_apart(got, w, f32=True) <= 2
    # f64 path used by the renderer: edge ratios return the inputs unchanged (util.rs:268-272)
    g = oracle.lib("det").oracle_combine_palette_color
    o = (C.c_double * 4)()
    g(oracle.darr((1, .5, 0, 1)), oracle.darr((0, 1, .5, .5)), 0.0, o)
    assert list(o) == [0, 1, .5, .5]
    g(oracle.darr((1, .5, 0, 1)), oracle.darr((0, 1, .5, .5)), 1.0, o)
    assert list(o) == [1, .5, 0, 1]


def test_remainder(oracle):  # util.rs:960-969
    f = oracle.lib("det").oracle_remainder_i
    assert [f(a, 3) for a in range(-3, 4)] == [0, 1, 2, 0, 1, 2, 0]
    g = oracle.lib("det").oracle_remainder_f
    assert g(-0.5, 4.0) == 3.5 and g(4.0, 4.0) == 0.0 and g(5.25, 4.0) == 1.25
