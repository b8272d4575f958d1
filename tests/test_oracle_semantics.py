"""Behaviour the reference's own tests do NOT pin but its code defines:
CSG stream quirks, NaN flow, Fresnel / Snell, quantisation.  Run against the oracle on the CPU;
the GPU suite then holds the CUDA path to the oracle bit for bit."""
import ctypes as C
import json
import math

import numpy as np
import pytest

import euclider_b200 as eb

P3 = lambda x, y, z: {"Point3::new": [x, y, z]}
V3 = lambda x, y, z: {"Vector3::new": [x, y, z]}
OP = lambda name: {"SetOperation": [name]}
SPH = lambda c, r: {"Sphere3::new": [P3(*c), r]}
HS = lambda n, p, inside: {"HalfSpace3::new_with_point": [{"Hyperplane3::new_with_point": [V3(*n), P3(*p)]}, P3(*inside)]}
OF = lambda shapes, op: {"ComposableShape3::of": [shapes, OP(op)]}


def scene_with(shapes, surface=None):
    surf = surface or {"ComposableSurface3": {
        "reflection_ratio": {"reflection_ratio_fresnel_3": [1.5, 1.0]},
        "reflection_direction": {"reflection_direction_specular_3": []},
        "threshold_direction": {"threshold_direction_snell_3": [1.5]},
        "surface_color": {"surface_color_uniform_3": [{"Rgba::new": [0, 0, 0, 0]}]}}}
    ents = [{"Entity3Impl::new": [s, {"Vacuum3::new": []}, surf]} for s in shapes] + [{"Void3::new_with_vacuum": []}]
    text = json.dumps({"Universe3": {"camera": {"PitchYawCamera3": []}, "entities": ents,
                                     "background": {"MappedTextureImpl3::new": [{"uv_sphere_3": [P3(0, 0, 0)]},
                                                                                {"texture_image_linear": ["./t.png"]}]}}})
    env = eb.Parser.default().parse(text, load_textures=False)
    env.set_texture(0, 2, 2, bytes([255, 0, 0, 255, 0, 255, 0, 255, 0, 0, 255, 255, 255, 255, 255, 255]))
    return env


def stream(oracle, env, entity, loc, direction, n=16):
    out = (C.c_double * (7 * n))()
    flat = env.flat
    k = oracle.lib("det").oracle_entity_intersections(C.byref(flat), entity, oracle.darr(loc), oracle.darr(direction), n, out)
    v = list(out)
    return [(v[i * 7], tuple(v[i * 7 + 4:i * 7 + 7])) for i in range(k)]


def test_union_two_overlapping_spheres(oracle, built_lib):  # shape.rs:212-264
    env = scene_with([OF([SPH((4, 0, 0), 1), SPH((5, 0, 0), 1)], "Union")])
    s = stream(oracle, env, 0, (0, 0, 0), (1, 0, 0))
    assert [t for t, _ in s] == [3.0, 6.0]  # inner crossings (4 and 5) are inside the other sphere
    assert s[0][1] == (-1.0, 0.0, 0.0) and s[1][1] == (1.0, 0.0, 0.0)


def test_intersection_lens(oracle, built_lib):  # shape.rs:291-340
    env = scene_with([OF([SPH((4, 0, 0), 1), SPH((5, 0, 0), 1)], "Intersection")])
    assert [t for t, _ in stream(oracle, env, 0, (0, 0, 0), (1, 0, 0))] == [4.0, 5.0]


def test_complement_flips_b_normals(oracle, built_lib):  # shape.rs:365-409
    env = scene_with([OF([SPH((5, 0, 0), 2), SPH((5, 0, 0), 1)], "Complement")])
    s = stream(oracle, env, 0, (0, 0, 0), (1, 0, 0), n=4)
    # a enters at 3; b's hits 4 and 6 are inside a and come back with negated normals; then only `a`
    # (exit at 7) is left and the iterator yields it WITHOUT advancing (shape.rs:392): it repeats
    assert [t for t, _ in s] == [3.0, 4.0, 6.0, 7.0]
    assert s[1][1] == (1.0, 0.0, 0.0) and s[2][1] == (-1.0, 0.0, 0.0)
    assert [t for t, _ in stream(oracle, env, 0, (0, 0, 0), (1, 0, 0), n=7)] == [3.0, 4.0, 6.0, 7.0, 7.0, 7.0, 7.0]


def test_symmetric_difference_flips_inside_hits(oracle, built_lib):  # shape.rs:436-496
    env = scene_with([OF([SPH((4, 0, 0), 1), SPH((5, 0, 0), 1)], "SymmetricDifference")])
    s = stream(oracle, env, 0, (0, 0, 0), (1, 0, 0))
    assert [t for t, _ in s] == [3.0, 4.0, 5.0, 6.0]
    assert [n[0] for _, n in s] == [-1.0, 1.0, -1.0, 1.0]  # 4 (b's entry, inside a) and 5 (a's exit, inside b) flipped


def test_union_early_none(oracle, built_lib):
    """Union with `b` exhausted returns None as soon as an `a` hit lies inside b -- even though a
    later `a` hit would be a boundary (shape.rs:243-250)."""
    # b = half-space x >= 3.5 hit at 3.5; a = sphere [3, 5]
    env = scene_with([OF([SPH((4, 0, 0), 1), HS((1, 0, 0), (3.5, 0, 0), (10, 0, 0))], "Union")])
    s = stream(oracle, env, 0, (0, 0, 0), (1, 0, 0))
    assert [t for t, _ in s] == [3.0]  # a(3) emitted, b(3.5) inside a skipped, then a(5) inside b -> None


def test_ties_pick_b(oracle, built_lib):  # `a.distance < b.distance` strict
    env = scene_with([OF([HS((1, 0, 0), (3, 0, 0), (10, 0, 0)), HS((2, 0, 0), (3, 0, 0), (10, 0, 0))], "SymmetricDifference")])
    s = stream(oracle, env, 0, (0, 0, 0), (1, 0, 0))
    assert [t for t, _ in s] == [3.0, 3.0]
    # b first: b's normal is n * -signum = (2,0,0) * -1, and the hit lies on a's boundary (inside, signum(+0) = 1) -> flipped
    assert s[0][1] == (2.0, 0.0, 0.0)


def test_plane_parallel_ray_passes_nan(oracle, built_lib):
    """t = -x/0 = -inf/NaN: `t < 0` rejects -inf but a NaN passes (shape.rs:792)."""
    env = scene_with([HS((0, 0, 1), (0, 0, 0), (0, 0, 1))])
    s = stream(oracle, env, 0, (0, 0, 0), (1, 0, 0))  # origin ON the plane, ray parallel: 0/0
    assert len(s) == 1 and math.isnan(s[0][0])
    assert stream(oracle, env, 0, (0, 0, 1), (1, 0, 0)) == []  # -1/0 = -inf < 0


def test_cylinder_second_normal_uses_first_axis_point(oracle, built_lib):  # shape.rs:999,1017
    env = scene_with([{"Cylinder3::new": [P3(5, 0, 0), V3(0, 0, 1), 1]}])
    s = stream(oracle, env, 0, (0, 0, 0), (1, 0, 0.5))
    (t1, n1), (t2, n2) = s
    assert (t1, t2) == (4.0, 6.0) and n1 == (-1.0, 0.0, 0.0)
    # second hit p2 = (6,0,3); axis point of the FIRST hit q = (5,0,2): normal = normalize((1,0,1))
    assert n2 == pytest.approx((math.sqrt(0.5), 0.0, math.sqrt(0.5)), abs=1e-15)


def test_material_at_first_entity_in_list_order(oracle, built_lib):
    env = scene_with([SPH((0, 0, 0), 2), SPH((0, 0, 0), 1)])
    f = oracle.lib("det").oracle_material_at
    flat = env.flat
    assert f(C.byref(flat), oracle.darr((0.5, 0, 0))) == 0  # both contain it: list order wins
    assert f(C.byref(flat), oracle.darr((5, 0, 0))) == 2    # only the Void


def probe(oracle, env, direction, normal_closer, exiting):
    r, refl, thr = C.c_double(), (C.c_double * 3)(), (C.c_double * 3)()
    flat = env.flat
    oracle.lib("det").oracle_surface_probe(C.byref(flat), 0, oracle.darr(direction), oracle.darr(normal_closer), exiting,
                                          C.byref(r), refl, thr)
    return r.value, list(refl), list(thr)


def test_fresnel_and_snell(oracle, built_lib):  # surface.rs:214-288
    env = scene_with([SPH((9, 0, 0), 1)])
    # head-on: R = ((n1 - n2) / (n1 + n2))^2 = 0.04 for 1 -> 1.5
    ratio, refl, thr = probe(oracle, env, (1, 0, 0), (-1, 0, 0), 0)
    assert ratio == pytest.approx(0.04, abs=1e-15) and refl == [-1.0, 0.0, 0.0]
    assert all(math.isnan(c) for c in thr)  # ray parallel to the normal: Gram-Schmidt degenerates (util.rs:631-666)
    # 45 degrees entering: Snell sin(t2) = sin(45) / 1.5
    d = (math.sqrt(0.5), math.sqrt(0.5), 0.0)
    ratio, refl, thr = probe(oracle, env, d, (-1, 0, 0), 0)
    t2 = math.asin(math.sin(math.pi / 4) / 1.5)
    assert thr == pytest.approx((math.cos(t2), math.sin(t2), 0.0), abs=1e-12)
    assert refl == pytest.approx((-d[0], d[1], 0.0), abs=1e-15)
    # total internal reflection when exiting at 60 degrees: asin(1.5 * sin 60) is NaN -> ratio 1
    d = (0.5, math.sqrt(0.75), 0.0)
    ratio, _, thr = probe(oracle, env, d, (-1, 0, 0), 1)
    assert ratio == 1.0 and all(math.isnan(c) for c in thr)


def test_general_rotation_nan_for_z_normals(oracle):
    """Any 3-D rotation plane containing +-z as first vector degenerates: e2 is in the span (App. A.5)."""
    out = (C.c_double * 3)()
    f = oracle.lib("det").oracle_general_rotation
    f(3, oracle.darr((0, 0, 1)), oracle.darr((0.6, 0, -0.8)), 0.1, oracle.darr((0.6, 0, -0.8)), out)
    assert all(math.isnan(c) for c in out)
    f(3, oracle.darr((0, 1, 0)), oracle.darr((0.6, -0.8, 0)), 0.25, oracle.darr((0.6, -0.8, 0)), out)
    got = np.array(list(out))
    assert np.linalg.norm(got) == pytest.approx(1.0, abs=1e-12)
    assert math.acos(float(np.dot(got, (0.6, -0.8, 0)))) == pytest.approx(0.25, abs=1e-12)


def test_to_pixel_truncates(oracle):
    out = (C.c_uint8 * 4)()
    oracle.lib("det").oracle_to_pixel(oracle.darr((0.999, 1.0, -0.5, 254.9999 / 255.0)), out)
    assert list(out) == [254, 255, 0, 254]
    oracle.lib("det").oracle_to_pixel(oracle.darr((float("nan"), 2.0, 0.5, 0.0)), out)
    assert list(out) == [0, 255, 127, 0]  # NaN -> 0 is this build's definition (the reference panics)


def test_blends_used_by_the_benchmark_scenes(oracle):
    out = (C.c_double * 4)()
    src, dst = (1.0, 1.0, 1.0, 0.0), (0.0, 0.0, 1.0, 0.25)
    oracle.lib("det").oracle_blend(eb._capi.BLEND_DARKEN, 0.0, oracle.darr(src), oracle.darr(dst), out)
    assert list(out) == [0.0, 0.0, 1.0, 0.25]  # a fully transparent source leaves the destination
    src, dst = (1.0, 0.0, 0.0, 1.0), (0.0, 1.0, 0.0, 1.0)
    oracle.lib("det").oracle_blend(eb._capi.BLEND_DIFFERENCE, 0.0, oracle.darr(src), oracle.darr(dst), out)
    assert list(out) == [1.0, 1.0, 0.0, 1.0]
    oracle.lib("det").oracle_blend(eb._capi.BLEND_OVER, 0.0, oracle.darr((1, 0, 0, 0.5)), oracle.darr((0, 0, 1, 1.0)), out)
    assert list(out) == [0.5, 0.0, 0.5, 1.0]


def test_perlin_is_deterministic_and_bounded(oracle, built_lib):
    env = scene_with([SPH((9, 0, 0), 1)])
    perm = bytes(env.flat.perlin_perm)
    assert sorted(perm) == list(range(256))  # a permutation (noise 0.4.1 PermutationTable, seed 0)
    f = oracle.lib("det").oracle_perlin4
    buf = (C.c_uint8 * 256)(*perm)
    rng = np.random.default_rng(3)
    vals = [f(buf, oracle.darr(p)) for p in rng.uniform(-20, 20, (2000, 4))]
    assert max(abs(v) for v in vals) <= 1.0 and len(set(vals)) > 1900
    assert f(buf, oracle.darr((1.0, 2.0, 3.0, 4.0))) == 0.0  # lattice points are zeros of gradient noise


def test_texture_wrap_and_orientation(oracle, built_lib):
    env = scene_with([SPH((9, 0, 0), 1)])  # 2x2 texture: (R, G / B, W), row 0 = top
    out = (C.c_double * 4)()
    flat = env.flat
    f = oracle.lib("det").oracle_mapped_color
    f(C.byref(flat), 0, oracle.darr((-1, 1e-9, 0)), out)  # u -> 1 (wraps to column 0 / 1 blend), v = 0.5
    assert 0.0 <= out[0] <= 1.0 and out[3] == 1.0
    # u = 0.5 (x between the two columns), v = 0.5 - asin(z)/pi: above the equator the TOP row (R, G)
    # dominates, below it the bottom row (B, W) -> the blue channel tells the orientation
    f(C.byref(flat), 0, oracle.darr((1, 0, 0.5)), out)
    up_blue = out[2]
    f(C.byref(flat), 0, oracle.darr((1, 0, -0.5)), out)
    down_blue = out[2]
    v_up = 0.5 - math.asin(0.5 / math.sqrt(1.25)) / math.pi
    assert up_blue == pytest.approx(2 * v_up - 0.5, abs=1e-12) and down_blue == pytest.approx(1 - (2 * v_up - 0.5), abs=1e-12)
    assert up_blue < 0.5 < down_blue
