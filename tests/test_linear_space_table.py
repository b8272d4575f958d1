"""LinearSpace in table form (scene_dev.cuh: LinRow; api_device.cu: lower_lin_row): component expressions that are
sums of `v`, `v * c`, `c * v`, `v / c`, `c / v` terms run as a per-void operation table on the device, everything else
through the RPN interpreter (material.rs:70-163 evaluates meval expressions per ray transition).  Both must equal the
oracle -- which walks the expression programs recursively -- bit for bit, and each other."""
import json

import numpy as np
import pytest

import euclider_b200 as eb

pytestmark = pytest.mark.gpu

P3 = lambda x, y, z: {"Point3::new": [x, y, z]}
V3 = lambda x, y, z: {"Vector3::new": [x, y, z]}

# (forward, inverse) per component; the oracle and the device only need them to be evaluated identically
EXPRESSIONS = {
    "stretch": [("x * 4", "x / 4"), ("y", "y"), ("z", "z")],
    "const_first": [("4 * x", "0.25 * x"), ("y / 2", "2 * y"), ("z", "z")],
    "sums": [("x * 2 + y", "x / 2 - y * 0.5"), ("y - z * 0.25", "y + z / 4"), ("z + x * 0.125 - y / 8", "z")],
    "reciprocal": [("x", "x"), ("y", "y"), ("3 / z", "3 / z")],
    "fallback_nonlinear": [("x * sqrt(abs(x) + 1)", "x / 2"), ("y", "y"), ("sin(z) + z", "z")],
    "fallback_parenthesised": [("(x + y) * 2", "x / 2"), ("y", "-y"), ("z", "z")],
}


def hallway(exprs):
    """A void box with the LinearSpace under test between the camera and a lit wall, plus a mirror sphere inside it so
    that rays enter, leave and re-enter the void."""
    surface = lambda color, ratio=0.0: {"ComposableSurface3": {
        "reflection_ratio": {"reflection_ratio_uniform_3": [ratio]},
        "reflection_direction": {"reflection_direction_specular_3": []},
        "threshold_direction": {"threshold_direction_identity_3": []}, "surface_color": color}}
    lit = lambda rgba: {"surface_color_blend_3": [
        {"surface_color_illumination_global_3": [{"Rgba::new": [1, 1, 1, 0]}, {"Rgba::new": [0, 0, 0, 0.6]}]},
        {"surface_color_uniform_3": [{"Rgba::new": rgba}]}, {"blend_function_darken": []}]}
    space = {"LinearSpace3": ["xyz", [{"ComponentTransformation3": [[
        {"ComponentTransformationExpr": [f, i]} for f, i in exprs]]}]]}
    wall = lambda n, p, inside: {"HalfSpace3::new_with_point": [{"Hyperplane3::new_with_point": [V3(*n), P3(*p)]}, P3(*inside)]}
    return json.dumps({"Universe3": {"camera": {"PitchYawCamera3": []}, "entities": [
        {"Entity3Impl::new": [{"Sphere3::new": [P3(9, 1, 0.5), 1.5]}, {"Vacuum3::new": []},
                              surface({"surface_color_uniform_3": [{"Rgba::new": [0.9, 0.9, 0.2, 1]}]}, 0.6)]},
        {"Entity3Impl::new": [{"HalfSpace3::cuboid": [P3(8, 0, 0), V3(8, 9, 7)]}, space,
                              surface({"surface_color_uniform_3": [{"Rgba::new": [0, 0, 0, 0]}]})]},
        {"Entity3Impl::new": [wall((1, 0, 0), (20, 0, 0), (21, 0, 0)), {"Vacuum3::new": []}, surface(lit([0.2, 0.4, 1, 1]))]},
        {"Entity3Impl::new": [wall((0, 0, 1), (0, 0, -5), (0, 0, -6)), {"Vacuum3::new": []}, surface(lit([1, 0.3, 0.2, 1]))]},
        {"Void3::new_with_vacuum": []}],
        "background": {"MappedTextureImpl3::new": [{"uv_sphere_3": [P3(0, 0, 0)]}, {"texture_image_linear": ["./t.png"]}]}}})


def load(exprs):
    env = eb.Parser.default().parse(hallway(exprs), load_textures=False)
    env.set_texture(0, 2, 2, bytes([255, 0, 0, 255, 0, 255, 0, 255, 0, 0, 255, 255, 255, 255, 255, 255]))
    return env


@pytest.mark.parametrize("name", sorted(EXPRESSIONS))
def test_table_rows_equal_the_interpreter_and_the_oracle(built_lib, oracle, monkeypatch, name):
    w, h = 144, 81
    env = load(EXPRESSIONS[name])
    ref_rgb, ref_hit, ref_stats = oracle.render(env, w, h, variant="det")
    img = env.render((w, h), want_hit_ids=True)
    assert np.array_equal(img.hit_ids, ref_hit) and img.stats["level_counts"] == ref_stats["level_counts"]
    assert np.array_equal(img.data, ref_rgb)
    assert ref_stats["level_counts"][2] > 0  # rays do cross the void
    monkeypatch.setenv("EUCL_LIN_TABLE", "0")  # read at eucl_scene_create: every component through the RPN interpreter
    env_rpn = load(EXPRESSIONS[name])
    img_rpn = env_rpn.render((w, h), want_hit_ids=True)
    assert np.array_equal(img_rpn.data, img.data) and np.array_equal(img_rpn.hit_ids, img.hit_ids)
    # the camera path through the void uses the same transformations (Universe::trace_path, mod.rs:186-227)
    loc, direction = [0.0, 0.2, 0.1], [1.0, 0.05, 0.02]
    assert env.trace_path_unknown(loc, direction, 9.0) == env_rpn.trace_path_unknown(loc, direction, 9.0) == oracle.trace_path(env, loc, direction, 9.0)
