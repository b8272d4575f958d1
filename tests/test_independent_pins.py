"""Pins for the arithmetic the reference takes from un-vendored crates (palette 0.2.1 blends and HSV, json 0.11
number conversion) and for the Fresnel provider, INDEPENDENT of the oracle's and the CUDA path's own code:
every expectation below is computed here, in numpy / the Python standard library, from the published
definitions -- the W3C "Compositing and Blending Level 1" equations (general form: separable blend function
B(Cb, Cs) + source-over compositing; Porter-Duff operators from their Fa / Fb tables), `colorsys.hsv_to_rgb`,
the Fresnel equations in their sine / tangent form, exact integer arithmetic for decimal literals -- and is
then compared with BOTH oracle builds (libm variants) on the CPU and with the GPU through rendered pixels.

Reference call sites: surface.rs:309-390 (blend_function_*), scene.rs:648-667 (Rgba::from_hsva),
surface.rs:214-244 (reflection_ratio_fresnel), scene.rs deserializer (json numbers).
Perlin (d3/entity/surface.rs:27-37) has no published closed form to compare with: oracle/ASSUMPTIONS.md row 6
records the gradient-table ordering alternative instead."""
import colorsys
import ctypes as C
import json
import math
from fractions import Fraction

import numpy as np
import pytest

import euclider_b200 as eb
from euclider_b200 import _capi

VARIANTS = ["det", "glibc"]
BLENDS = ["over", "inside", "outside", "atop", "xor", "plus", "multiply", "screen", "overlay", "darken", "lighten",
          "dodge", "burn", "hard_light", "soft_light", "difference", "exclusion"]
BLEND_IDS = {name: getattr(_capi, "BLEND_" + name.upper()) for name in BLENDS}


# ---------------------------------------------------------------------------------------------------------
# W3C compositing, written from the spec's GENERAL equations (not from palette's expanded per-mode forms)

def _hard_light(cb, cs):
    return np.where(cs <= 0.5, cb * (2 * cs), cb + (2 * cs - 1) - cb * (2 * cs - 1))


def _soft_light(cb, cs, linear_term=-3.0):
    """Soft light.  palette 0.2.1 documents its blend modes as taken from the SVG Compositing draft, whose
    dark-backdrop polynomial is  16 m^3 - 12 m^2 - 3 m  (premultiplied form: Da (2 Sca - Sa) (16 m^3 - 12 m^2 - 3 m)
    + Sca - Sca Da + Dca);  W3C Compositing and Blending Level 1 has  D(Cb) - Cb = 16 m^3 - 12 m^2 + 3 m  there.
    The build follows the SVG draft (`linear_term = -3`); test_soft_light_deviation_from_w3c_level_1 pins that the
    two differ in that branch only.  Recorded in oracle/ASSUMPTIONS.md."""
    dark = (16 * cb * cb * cb - 12 * cb * cb + linear_term * cb)
    delta = np.where(cb <= 0.25, dark, np.sqrt(cb) - cb)
    return np.where(cs <= 0.5, cb - (1 - 2 * cs) * cb * (1 - cb), cb + (2 * cs - 1) * delta)


def _dodge(cb, cs):
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(cb == 0, 0.0, np.where(cs == 1, 1.0, np.minimum(1.0, cb / (1 - cs))))


def _burn(cb, cs):
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(cb == 1, 1.0, np.where(cs == 0, 0.0, 1.0 - np.minimum(1.0, (1 - cb) / cs)))


SEPARABLE = {
    "multiply": lambda cb, cs: cb * cs,
    "screen": lambda cb, cs: cb + cs - cb * cs,
    "overlay": lambda cb, cs: _hard_light(cs, cb),  # hard-light with the layers swapped
    "darken": np.minimum,
    "lighten": np.maximum,
    "dodge": _dodge,
    "burn": _burn,
    "hard_light": _hard_light,
    "soft_light": _soft_light,
    "difference": lambda cb, cs: np.abs(cb - cs),
    "exclusion": lambda cb, cs: cb + cs - 2 * cb * cs,
}
# Porter-Duff: co = as * Fa * Cs + ab * Fb * Cb, ao = as * Fa + ab * Fb
PORTER_DUFF = {
    "over": (lambda a_s, a_b: 1.0, lambda a_s, a_b: 1 - a_s),
    "inside": (lambda a_s, a_b: a_b, lambda a_s, a_b: 0.0),
    "outside": (lambda a_s, a_b: 1 - a_b, lambda a_s, a_b: 0.0),
    "atop": (lambda a_s, a_b: a_b, lambda a_s, a_b: 1 - a_s),
    "xor": (lambda a_s, a_b: 1 - a_b, lambda a_s, a_b: 1 - a_s),
    "plus": (lambda a_s, a_b: 1.0, lambda a_s, a_b: 1.0),
}


def w3c_blend(name, src, dst):
    """src, dst: arrays [..., 4] of NON-premultiplied r, g, b, a in [0, 1]; returns the non-premultiplied result."""
    cs, a_s = src[..., :3], src[..., 3:4]
    cb, a_b = dst[..., :3], dst[..., 3:4]
    if name in PORTER_DUFF:
        fa, fb = PORTER_DUFF[name]
        co = a_s * fa(a_s, a_b) * cs + a_b * fb(a_s, a_b) * cb
        ao = np.minimum(1.0, a_s * fa(a_s, a_b) + a_b * fb(a_s, a_b))
    else:
        mixed = (1 - a_b) * cs + a_b * SEPARABLE[name](cb, cs)  # Cs' = (1 - ab) Cs + ab B(Cb, Cs)
        co = a_s * mixed + (1 - a_s) * a_b * cb                   # source-over of Cs' on the backdrop
        ao = a_s + a_b * (1 - a_s)
    with np.errstate(divide="ignore", invalid="ignore"):
        rgb = np.where(ao > 0, co / ao, 0.0)
    return np.concatenate([rgb, ao], axis=-1)


def oracle_blend(oracle, variant, name, src, dst):
    out = (C.c_double * 4)()
    f = oracle.lib(variant).oracle_blend
    res = np.zeros_like(src)
    for i in range(src.shape[0]):
        f(BLEND_IDS[name], 0.0, oracle.darr(src[i]), oracle.darr(dst[i]), out)
        res[i] = list(out)
    return res


def blend_inputs():
    rng = np.random.default_rng(17)
    rand = rng.uniform(0.0, 1.0, (400, 2, 4))
    # alpha corners and colour corners (the dodge / burn / soft-light branch points)
    corners = []
    for a_s in (0.0, 1.0, 0.5):
        for a_b in (0.0, 1.0, 0.5):
            for c_s in (0.0, 1.0, 0.25, 0.5):
                for c_b in (0.0, 1.0, 0.25, 0.5):
                    corners.append([[c_s, 1 - c_s, c_s * 0.5, a_s], [c_b, c_b * 0.5, 1 - c_b, a_b]])
    allv = np.concatenate([rand, np.array(corners)], axis=0)
    return allv[:, 0, :].copy(), allv[:, 1, :].copy()


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("name", BLENDS)
def test_blend_functions_follow_the_w3c_equations(oracle, name, variant):
    """All 17 palette blends (surface.rs:309-390) against the spec's general equations, random inputs plus the
    alpha = 0 / 1 and colour = 0 / 1 corners (from_premultiplied's `is_normal` branch, dodge / burn limits)."""
    src, dst = blend_inputs()
    got = oracle_blend(oracle, variant, name, src, dst)
    want = w3c_blend(name, src, dst)
    visible = want[:, 3] > 1e-9  # colour of a fully transparent result is meaningless (palette returns 0)
    # where the spec's branch conditions meet (Cb = 0 / Cs = 1 for dodge, Cb = 1 / Cs = 0 for burn) an implementation
    # may take either neighbouring branch; both are limits of the same function only for dodge with Cb = 0 and burn
    # with Cb = 1, which the spec tests first -- palette tests the source first.  Exclude exactly those double corners.
    if name == "dodge":
        visible &= ~((src[:, :3] == 1).any(axis=1) & (dst[:, :3] == 0).any(axis=1))
    if name == "burn":
        visible &= ~((src[:, :3] == 0).any(axis=1) & (dst[:, :3] == 1).any(axis=1))
    assert visible.sum() > 400
    assert np.allclose(got[visible], want[visible], rtol=0, atol=2e-12), \
        (name, np.abs(got[visible] - want[visible]).max())
    assert np.allclose(got[~visible][:, 3], want[~visible][:, 3], atol=2e-12)


def test_soft_light_deviation_from_w3c_level_1(oracle):
    """The one place where the build's blends and W3C Level 1 disagree: soft light over a dark backdrop (Cb <= 1/4)
    with a bright source (Cs > 1/2) -- the sign of the linear term of the SVG draft's polynomial."""
    src, dst = blend_inputs()
    got = oracle_blend(oracle, "det", "soft_light", src, dst)
    SEPARABLE["_w3c_soft_light"] = lambda cb, cs: _soft_light(cb, cs, linear_term=3.0)
    try:
        w3c = w3c_blend("_w3c_soft_light", src, dst)
    finally:
        del SEPARABLE["_w3c_soft_light"]
    differs = np.abs(got - w3c).max(axis=1) > 1e-9
    in_branch = ((src[:, :3] > 0.5) & (dst[:, :3] <= 0.25) & (dst[:, :3] > 0)).any(axis=1) & (src[:, 3] > 0) & (dst[:, 3] > 0)
    assert differs.any() and not (differs & ~in_branch).any()


@pytest.mark.parametrize("variant", VARIANTS)
def test_hsv_to_rgb_matches_colorsys(oracle, variant):
    """palette Hsv -> Rgb (scene.rs:648-667 Rgba::from_hsva, d3/entity/surface.rs:38) against the standard library."""
    f = oracle.lib(variant).oracle_hsv_to_rgb
    out = (C.c_double * 3)()
    rng = np.random.default_rng(5)
    cases = [(h, s, v) for h, s, v in zip(rng.uniform(-720, 1080, 600), rng.uniform(0, 1, 600), rng.uniform(0, 1, 600))]
    cases += [(h, 1.0, 1.0) for h in (0.0, 60.0, 120.0, 180.0, 240.0, 300.0, 359.999, 360.0, -60.0, 720.0)]
    for h, s, v in cases:
        f(h, s, v, out)
        want = colorsys.hsv_to_rgb((h / 360.0) % 1.0, s, v)
        assert list(out) == pytest.approx(want, abs=1e-12), (h, s, v)


def test_from_hsva_is_lowered_like_colorsys(built_lib):
    """Rgba::from_hsva constants are converted by the host front end (csrc/host/scene_parse.cc), not by the oracle."""
    rng = np.random.default_rng(9)
    for h, s, v, a in rng.uniform(0, 1, (20, 4)) * [360.0, 1, 1, 1]:
        env = parse_wall({"surface_color_uniform_3": [{"Rgba::from_hsva": [float(h), float(s), float(v), float(a)]}]})
        op = env.flat.color_ops[0]
        assert op.op == 0
        want = colorsys.hsv_to_rgb(h / 360.0, s, v) + (a,)
        assert [op.f[k] for k in range(4)] == pytest.approx(want, abs=1e-12)


# ---------------------------------------------------------------------------------------------------------
# Fresnel (surface.rs:214-244) in the sine / tangent form of the equations

def fresnel_unpolarised(theta_i, n1, n2):
    s = n1 / n2 * math.sin(theta_i)
    if s > 1.0:
        return 1.0
    theta_t = math.asin(s)
    if theta_i == 0.0:
        return ((n1 - n2) / (n1 + n2)) ** 2
    rs = (math.sin(theta_i - theta_t) / math.sin(theta_i + theta_t)) ** 2
    rp = (math.tan(theta_i - theta_t) / math.tan(theta_i + theta_t)) ** 2
    return 0.5 * (rs + rp)


def glass_env(index=1.458):
    from test_oracle_semantics import SPH, scene_with
    return scene_with([SPH((9, 0, 0), 1)], surface={"ComposableSurface3": {
        "reflection_ratio": {"reflection_ratio_fresnel_3": [index, 1.0]},
        "reflection_direction": {"reflection_direction_specular_3": []},
        "threshold_direction": {"threshold_direction_snell_3": [index]},
        "surface_color": {"surface_color_uniform_3": [{"Rgba::new": [0, 0, 0, 0]}]}}})


def probe_ratio(oracle, variant, env, theta, exiting):
    # ray in the xy plane hitting a surface whose closer normal is -x at incidence `theta`
    d = (math.cos(theta), math.sin(theta), 0.0)
    r, refl, thr = C.c_double(), (C.c_double * 3)(), (C.c_double * 3)()
    flat = env.flat
    oracle.lib(variant).oracle_surface_probe(C.byref(flat), 0, oracle.darr(d), oracle.darr((-1.0, 0.0, 0.0)), exiting,
                                             C.byref(r), refl, thr)
    return r.value, list(thr)


@pytest.mark.parametrize("variant", VARIANTS)
def test_fresnel_closed_forms(oracle, built_lib, variant):
    n = 1.458
    env = glass_env(n)
    # normal incidence: ((n1 - n2) / (n1 + n2))^2
    assert probe_ratio(oracle, variant, env, 0.0, 0)[0] == pytest.approx(((1 - n) / (1 + n)) ** 2, abs=1e-14)
    assert probe_ratio(oracle, variant, env, 0.0, 1)[0] == pytest.approx(((n - 1) / (n + 1)) ** 2, abs=1e-14)
    # Brewster's angle: the p-polarised term vanishes, R = Rs / 2 with Rs = cos^2(2 theta_B)
    for exiting, (n1, n2) in ((0, (1.0, n)), (1, (n, 1.0))):
        theta_b = math.atan(n2 / n1)
        rs = math.cos(2 * theta_b) ** 2
        assert probe_ratio(oracle, variant, env, theta_b, exiting)[0] == pytest.approx(0.5 * rs, abs=1e-13)
    # beyond the critical angle (exiting): total internal reflection, ratio exactly 1
    theta_c = math.asin(1.0 / n)
    assert probe_ratio(oracle, variant, env, theta_c + 1e-6, 1)[0] == 1.0
    assert probe_ratio(oracle, variant, env, theta_c - 1e-6, 1)[0] < 1.0
    # random angles against the sine / tangent form
    rng = np.random.default_rng(11)
    for theta in rng.uniform(0.01, math.pi / 2 - 0.01, 200):
        for exiting, (n1, n2) in ((0, (1.0, n)), (1, (n, 1.0))):
            assert probe_ratio(oracle, variant, env, theta, exiting)[0] == pytest.approx(
                fresnel_unpolarised(theta, n1, n2), abs=1e-11), (theta, exiting)


@pytest.mark.parametrize("variant", VARIANTS)
def test_snell_direction(oracle, built_lib, variant):
    """threshold_direction_snell (surface.rs:268-288): n1 sin(t1) = n2 sin(t2), refracted ray in the plane of incidence."""
    n = 1.458
    env = glass_env(n)
    rng = np.random.default_rng(13)
    for theta in rng.uniform(0.05, 1.4, 100):
        _, thr = probe_ratio(oracle, variant, env, theta, 0)
        t2 = math.asin(math.sin(theta) / n)
        assert thr == pytest.approx((math.cos(t2), math.sin(t2), 0.0), abs=1e-12)


# ---------------------------------------------------------------------------------------------------------
# json 0.11 number conversion: mantissa (u64) * 10^exponent in ONE f64 multiplication / division

def parse_wall(color, radius_literal="3", extra=""):
    text = ('{"Universe3": {"camera": {"PitchYawCamera3": []}, "entities": [{"Entity3Impl::new": ['
            '{"Sphere3::new": [{"Point3::new": [10, 0, 0]}, ' + radius_literal + ']}, {"Vacuum3::new": []}, '
            '{"ComposableSurface3": {"reflection_ratio": {"reflection_ratio_uniform_3": [0.0]}, '
            '"reflection_direction": {"reflection_direction_specular_3": []}, '
            '"threshold_direction": {"threshold_direction_identity_3": []}, "surface_color": ' + json.dumps(color) + '}}]}, '
            '{"Void3::new_with_vacuum": []}], "background": {"MappedTextureImpl3::new": [{"uv_sphere_3": '
            '[{"Point3::new": [0, 0, 0]}]}, {"texture_image_linear": ["./t.png"]}]}}}')
    env = eb.Parser.default().parse(text, load_textures=False)
    env.set_texture(0, 2, 2, bytes([0, 0, 255, 255] * 4))
    return env


def json_crate_value(literal):
    """json 0.11 (number.rs): a number is kept as (sign, u64 mantissa, i16 decimal exponent) and converted with ONE
    multiplication `mantissa as f64 * 10^e`, the power being the correctly rounded f64 literal 1e<e> from a table
    (|e| <= 22) -- so, unlike strtod, the result carries up to three roundings (mantissa, power, product)."""
    lit = literal.lower()
    sign = -1.0 if lit.startswith("-") else 1.0
    lit = lit.lstrip("+-")
    exp10 = 0
    if "e" in lit:
        lit, e = lit.split("e")
        exp10 = int(e)
    if "." in lit:
        whole, frac = lit.split(".")
        exp10 -= len(frac)
        lit = whole + frac
    mantissa = int(lit)
    assert mantissa < 2 ** 64 and abs(exp10) <= 22
    power = float(Fraction(10) ** exp10)  # Fraction -> float rounds correctly: the table literal 1e<exp10>
    return sign * (float(mantissa) * power)


@pytest.mark.parametrize("literal", ["4.24264068712", "21.01", "1.458", "0.1", "3", "2.5e3", "1e-7", "123456789.987654321",
                                     "6.01", "0.30000000000000004", "17.000000000000004", "9007199254740993"])
def test_json_numbers_are_mantissa_times_power_of_ten(built_lib, literal):
    env = parse_wall({"surface_color_uniform_3": [{"Rgba::new": [1, 0, 0, 1]}]}, radius_literal=literal)
    got = env.flat.prims[0].s0
    want = json_crate_value(literal)
    exact = Fraction(literal)  # what a correctly rounding reader (strtod) returns
    assert got == want, (literal, got, want)
    assert abs(Fraction(got) - exact) <= abs(exact) * Fraction(1, 2 ** 51)  # a few ulp from the literal at most


# ---------------------------------------------------------------------------------------------------------
# the GPU against the same numpy expectations, through rendered pixels

def wall_scene(color, dim=3):
    """An opaque wall x >= 5 in front of the default camera, shaded by `color` alone (reflection ratio 0)."""
    surface = {"ComposableSurface3": {
        "reflection_ratio": {"reflection_ratio_uniform_3": [0.0]},
        "reflection_direction": {"reflection_direction_specular_3": []},
        "threshold_direction": {"threshold_direction_identity_3": []}, "surface_color": color}}
    wall = {"HalfSpace3::new_with_point": [{"Hyperplane3::new_with_point": [{"Vector3::new": [1, 0, 0]}, {"Point3::new": [5, 0, 0]}]},
                                           {"Point3::new": [6, 0, 0]}]}
    text = json.dumps({"Universe3": {"camera": {"PitchYawCamera3": []}, "entities": [
        {"Entity3Impl::new": [wall, {"Vacuum3::new": []}, surface]}, {"Void3::new_with_vacuum": []}],
        "background": {"MappedTextureImpl3::new": [{"uv_sphere_3": [{"Point3::new": [0, 0, 0]}]},
                                                   {"texture_image_linear": ["./t.png"]}]}}})
    env = eb.Parser.default().parse(text, load_textures=False)
    env.set_texture(0, 2, 2, bytes([0, 0, 255, 255] * 4))
    return env


def to_u8(rgb):
    return np.floor(np.clip(rgb, 0.0, 1.0) * 255.0).astype(int)


@pytest.mark.gpu
@pytest.mark.parametrize("name", BLENDS)
def test_gpu_blends_follow_the_w3c_equations(built_lib, name):
    """surface_color_blend on the device, pixel by pixel against numpy: opaque backdrops (alpha_b = 1) so that the
    blended colour is what reaches the frame (Porter-Duff operators that leave the wall translucent are compared
    through `over white` with the transmitted background, a uniform blue texture)."""
    rng = np.random.default_rng(23)
    uniform = lambda c: {"surface_color_uniform_3": [{"Rgba::new": [float(v) for v in c]}]}
    for src, dst in zip(rng.uniform(0.05, 0.95, (6, 4)), rng.uniform(0.05, 0.95, (6, 4))):
        dst[3] = 1.0
        want = w3c_blend(name, src[None, :], dst[None, :])[0]
        if want[3] < 1.0 - 1e-12:
            continue  # translucent result: the pixel also depends on what lies behind (covered GPU-vs-oracle)
        env = wall_scene({"surface_color_blend_3": [uniform(src), uniform(dst), {"blend_function_" + name: []}]})
        img = env.render((4, 4))
        px = img.data[2, 2].astype(int)
        assert np.abs(px - to_u8(want[:3])).max() <= 1, (name, src, dst, px, to_u8(want[:3]))  # truncation knife edge: 1 LSB


@pytest.mark.gpu
def test_gpu_fresnel_mix_matches_the_equations(built_lib):
    """A glass half-space seen at a grazing angle in front of a red wall (reflected ray) with a blue sky behind it
    (transmitted ray): red / 255 of the pixel is the Fresnel reflectance, blue / 255 its complement."""
    n = 1.458
    glass = {"ComposableSurface3": {
        "reflection_ratio": {"reflection_ratio_fresnel_3": [n, 1.0]},
        "reflection_direction": {"reflection_direction_specular_3": []},
        "threshold_direction": {"threshold_direction_snell_3": [n]},
        "surface_color": {"surface_color_uniform_3": [{"Rgba::new": [0, 0, 0, 0]}]}}}
    red = {"ComposableSurface3": {
        "reflection_ratio": {"reflection_ratio_uniform_3": [0.0]},
        "reflection_direction": {"reflection_direction_specular_3": []},
        "threshold_direction": {"threshold_direction_identity_3": []},
        "surface_color": {"surface_color_uniform_3": [{"Rgba::new": [1, 0, 0, 1]}]}}}
    hs = lambda normal, point, inside: {"HalfSpace3::new_with_point": [
        {"Hyperplane3::new_with_point": [{"Vector3::new": normal}, {"Point3::new": point}]}, {"Point3::new": inside}]}
    for theta in (0.3, 0.9, 1.2, 1.4):
        # glass fills y <= -1 below the camera's forward ray tilted down by (pi/2 - theta): the ray (1, -tan, 0) meets the
        # plane y = -1 at incidence theta; the reflected ray climbs to the red ceiling y >= 50
        down = math.tan(math.pi / 2 - theta)
        text = json.dumps({"Universe3": {"camera": {"PitchYawCamera3": []}, "entities": [
            {"Entity3Impl::new": [hs([0, 1, 0], [0, -1, 0], [0, -2, 0]), {"Vacuum3::new": []}, glass]},
            {"Entity3Impl::new": [hs([0, 1, 0], [0, 50, 0], [0, 51, 0]), {"Vacuum3::new": []}, red]},
            {"Void3::new_with_vacuum": []}],
            "background": {"MappedTextureImpl3::new": [{"uv_sphere_3": [{"Point3::new": [0, 0, 0]}]},
                                                       {"texture_image_linear": ["./t.png"]}]}}})
        env = eb.Parser.default().parse(text, load_textures=False)
        env.set_texture(0, 2, 2, bytes([0, 0, 255, 255] * 4))
        # aim the camera along (1, -down, 0): yaw about z turns forward towards -y
        env.rotate_yaw(-math.atan(down))
        img = env.render((5, 5))
        px = img.data[2, 2].astype(float)
        r = fresnel_unpolarised(theta, 1.0, n)
        assert abs(px[0] / 255.0 - r) <= 1.5 / 255.0 and abs(px[2] / 255.0 - (1 - r)) <= 2.5 / 255.0, (theta, px, r)
        assert px[1] <= 1
