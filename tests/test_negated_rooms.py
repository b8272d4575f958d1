"""Complement(VoidShape, X) -- a room described by its interior, like 4d_room's walls -- is lowered to X's program with
normals flipped and membership inverted (api_device.cu: ENT_NEGATED) instead of running the general Complement merge:
VoidShape yields no hits and contains every point, so ComplementIterator only ever takes its `b`-only branch
(shape.rs:394-408) and is_point_inside is `true && !X` (shape.rs:596).  With X a chain of half-spaces the room becomes
a root plane chain (first-item shortcut, light intersect kernel).  Bit-exact against the oracle, which evaluates the
original Complement, and against the general evaluator (EUCL_NEGATED_ROOMS=0)."""
import json

import numpy as np
import pytest

import euclider_b200 as eb

pytestmark = pytest.mark.gpu

P3 = lambda x, y, z: {"Point3::new": [x, y, z]}
V3 = lambda x, y, z: {"Vector3::new": [x, y, z]}
SPH = lambda c, r: {"Sphere3::new": [P3(*c), r]}
BOX = lambda c, d: {"HalfSpace3::cuboid": [P3(*c), V3(*d)]}
OF = lambda shapes, op: {"ComposableShape3::of": [shapes, {"SetOperation": [op]}]}
EVERYTHING_BUT = lambda x: OF([{"VoidShape3": []}, x], "Complement")

SHAPES = {
    "box_room": EVERYTHING_BUT(BOX((3, 0, 0), (30, 16, 9))),                     # root plane chain after lowering
    "sphere_room": EVERYTHING_BUT(SPH((2, 0, 0), 14)),                            # a primitive after lowering
    "lens_room": EVERYTHING_BUT(OF([SPH((0, 0, 0), 12), SPH((9, 0, 0), 12)], "Intersection")),  # general program
    "two_boxes_room": EVERYTHING_BUT(OF([BOX((0, 0, 0), (14, 14, 8)), BOX((9, 2, 0), (14, 9, 6))], "Union")),
}


def scene(room, camera_x=0.0):
    surface = lambda color, ratio: {"ComposableSurface3": {
        "reflection_ratio": {"reflection_ratio_uniform_3": [ratio]},
        "reflection_direction": {"reflection_direction_specular_3": []},
        "threshold_direction": {"threshold_direction_identity_3": []}, "surface_color": color}}
    lit = lambda rgba: {"surface_color_blend_3": [
        {"surface_color_illumination_global_3": [{"Rgba::new": [1, 1, 1, 0]}, {"Rgba::new": [0, 0, 0, 0.6]}]},
        {"surface_color_illumination_directional_3": [V3(0.3, 0.2, -1), {"Rgba::new": rgba}, {"Rgba::new": [0.1, 0.1, 0.1, 1]}]},
        {"blend_function_darken": []}]}
    glass = {"ComposableSurface3": {
        "reflection_ratio": {"reflection_ratio_fresnel_3": [1.458, 1.0]},
        "reflection_direction": {"reflection_direction_specular_3": []},
        "threshold_direction": {"threshold_direction_snell_3": [1.458]},
        "surface_color": {"surface_color_uniform_3": [{"Rgba::new": [0, 0, 0, 0]}]}}}
    return json.dumps({"Universe3": {"camera": {"PitchYawCamera3::new_with_location": [P3(camera_x, 0.3, 0.2)]}, "entities": [
        {"Entity3Impl::new": [SPH((6, 1, 0), 1.5), {"Vacuum3::new": []}, glass]},
        {"Entity3Impl::new": [SPH((5, -3, 1), 1.0), {"Vacuum3::new": []}, surface(lit([0.9, 0.9, 0.2, 1]), 0.5)]},
        {"Entity3Impl::new": [room, {"Vacuum3::new": []}, surface(lit([0.3, 0.5, 1, 1]), 0.2)]},
        {"Void3::new_with_vacuum": []}],
        "background": {"MappedTextureImpl3::new": [{"uv_sphere_3": [P3(0, 0, 0)]}, {"texture_image_linear": ["./t.png"]}]}}})


def load(text):
    env = eb.Parser.default().parse(text, load_textures=False)
    env.set_texture(0, 2, 2, bytes([255, 0, 0, 255, 0, 255, 0, 255, 0, 0, 255, 255, 255, 255, 255, 255]))
    return env


@pytest.mark.parametrize("camera_x", [0.0, -40.0], ids=["inside", "outside"])
@pytest.mark.parametrize("name", sorted(SHAPES))
def test_everything_but_x_rooms(built_lib, oracle, monkeypatch, name, camera_x):
    """camera inside the room (in the Void, looking at the walls from within) and outside it (inside the wall material:
    material_at must say so)."""
    w, h = 144, 81
    env = load(scene(SHAPES[name], camera_x))
    ref_rgb, ref_hit, ref_stats = oracle.render(env, w, h, variant="det")
    for pipeline in (eb.EUCL_PIPELINE_WAVEFRONT, eb.EUCL_PIPELINE_MEGAKERNEL):
        env.pipeline = pipeline
        for _ in range(6 if pipeline == eb.EUCL_PIPELINE_WAVEFRONT else 1):
            img = env.render((w, h), want_hit_ids=True)
            assert np.array_equal(img.hit_ids, ref_hit) and img.stats["level_counts"] == ref_stats["level_counts"]
            assert np.array_equal(img.data, ref_rgb)
    monkeypatch.setenv("EUCL_NEGATED_ROOMS", "0")  # read at eucl_scene_create: the general Complement merge
    general = load(scene(SHAPES[name], camera_x)).render((w, h), want_hit_ids=True)
    assert np.array_equal(general.data, ref_rgb) and np.array_equal(general.hit_ids, ref_hit)
    env.precision = "f32"
    f32_rgb, f32_hit, _ = oracle.render(env, w, h, variant="f32")
    img = env.render((w, h), want_hit_ids=True)
    assert np.array_equal(img.data, f32_rgb) and np.array_equal(img.hit_ids, f32_hit)
