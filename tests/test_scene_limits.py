"""eucl_scene_create validates a flat scene BEFORE it touches a device (so this runs without a GPU): programs
that exceed a fixed device evaluation stack are rejected with EUCL_ERR_SCENE_LIMIT, out-of-range table
indices and opcodes with EUCL_ERR_INVALID_ARGUMENT -- never silently corrupted thread-local memory.

The reference accepts blends and expressions of any depth (surface.rs:295-307 recurses through closures,
material.rs:99-110 through meval's tree); the device evaluators use fixed stacks (scene_dev.cuh:
kExprStackMax = 16 values, kColorStackMax = 8 colours)."""
import ctypes as C
import json

import pytest

import euclider_b200 as eb
from euclider_b200 import _capi
from euclider_b200._capi import lib

from pathlib import Path

EUCL_ERR_INVALID_ARGUMENT, EUCL_ERR_SCENE_LIMIT, EUCL_ERR_NO_DEVICE = -1, -21, -32
PARSER = eb.Parser.default(resource_root=Path(__file__).resolve().parent.parent)


def uniform(r, g, b, a):
    return {"surface_color_uniform_3": [{"Rgba::new": [r, g, b, a]}]}


def nested_blend(depth):
    """`depth` blends nested through `destination`: the colour program needs depth + 1 stack slots."""
    color = uniform(0.1, 0.2, 0.3, 1.0)
    for k in range(depth):
        color = {"surface_color_blend_3": [uniform(0.5, 0.5 / (k + 1), 0.25, 0.5), color, {"blend_function_darken": []}]}
    return color


def scene(color, expression="x * 2", inverse="x / 2"):
    material = {"LinearSpace3": {"legend": "xyz", "transformations": [{"ComponentTransformation3": {"expressions": [
        {"ComponentTransformationExpr": {"expression": expression, "inverse_expression": inverse}},
        {"ComponentTransformationExpr": {"expression": "y", "inverse_expression": "y"}},
        {"ComponentTransformationExpr": {"expression": "z", "inverse_expression": "z"}}]}}]}}
    surface = {"ComposableSurface3": {
        "reflection_ratio": {"reflection_ratio_uniform_3": [0.0]},
        "reflection_direction": {"reflection_direction_specular_3": []},
        "threshold_direction": {"threshold_direction_identity_3": []},
        "surface_color": color}}
    return json.dumps({"Universe3": {
        "camera": {"PitchYawCamera3": []},
        "entities": [{"Entity3Impl::new": [{"Sphere3::new": [{"Point3::new": [10, 0, 0]}, 3]}, material, surface]},
                     {"Void3::new_with_vacuum": []}],
        "background": {"MappedTextureImpl3::new": [{"uv_sphere_3": [{"Point3::new": [0, 0, 0]}]},
                                                   {"texture_image_linear": ["./tests/scenes/checker_rgba.png"]}]}}})


def create_status(env, flat=None):
    handle = C.c_void_p()
    st = lib().eucl_scene_create(C.byref(flat) if flat is not None else lib().eucl_parsed_flat(env._parsed), 0, C.byref(handle))
    if st == 0:
        lib().eucl_scene_destroy(handle)
    return st, lib().eucl_last_error().decode()


def accepted(status):
    # a valid scene gets past validation: created on a GPU box, EUCL_ERR_NO_DEVICE here
    return status in (0, EUCL_ERR_NO_DEVICE)


def right_nested_sum(operands):
    expr = "x"
    for _ in range(operands - 1):
        expr = f"x + ({expr})"
    return expr


def test_blend_nesting_up_to_the_device_stack_is_accepted(built_lib):
    env = PARSER.parse(scene(nested_blend(7)))  # 8 slots
    assert accepted(create_status(env)[0])


def test_deeper_blend_nesting_is_a_scene_limit(built_lib):
    env = PARSER.parse(scene(nested_blend(8)))  # 9 slots: used to overflow `Rgba stack[8]` silently
    status, msg = create_status(env)
    assert status == EUCL_ERR_SCENE_LIMIT and "colour program" in msg


def test_left_nested_expressions_need_two_slots(built_lib):
    expr = " + ".join(["x"] * 24)  # ((x + x) + x) ...: stack depth 2 whatever the length
    env = PARSER.parse(scene(uniform(1, 0, 0, 1), expr, "x / 24"))
    assert accepted(create_status(env)[0])


def test_right_nested_expression_up_to_the_device_stack_is_accepted(built_lib):
    env = PARSER.parse(scene(uniform(1, 0, 0, 1), right_nested_sum(16), "x / 16"))
    assert accepted(create_status(env)[0])


def test_deeper_right_nested_expression_is_a_scene_limit(built_lib):
    env = PARSER.parse(scene(uniform(1, 0, 0, 1), right_nested_sum(17), "x / 17"))  # used to overflow `double st[16]`
    status, msg = create_status(env)
    assert status == EUCL_ERR_SCENE_LIMIT and "expression" in msg


def _copy_table(flat, name, ctype, count):
    arr = (ctype * max(count, 1))()
    src = getattr(flat, name)
    for i in range(count):
        arr[i] = src[i]
    return arr


MUTATIONS = {
    "entity.material": ("entities", _capi.EuclEntity, "n_entities", lambda t: setattr(t[0], "material", 99)),
    "entity.surface": ("entities", _capi.EuclEntity, "n_entities", lambda t: setattr(t[0], "surface", 7)),
    "entity.node_root": ("entities", _capi.EuclEntity, "n_entities", lambda t: setattr(t[0], "node_root", 1000)),
    "node.prim": ("nodes", _capi.EuclNode, "n_nodes", lambda t: setattr(t[0], "prim", -4)),
    "prim.kind": ("prims", _capi.EuclPrim, "n_prims", lambda t: setattr(t[0], "kind", 9)),
    "material.transform_first": ("materials", _capi.EuclMaterial, "n_materials", lambda t: setattr(t[0], "transform_first", 5)),
    "material.kind": ("materials", _capi.EuclMaterial, "n_materials", lambda t: setattr(t[0], "kind", 3)),
    "transform.fwd_len": ("transforms", _capi.EuclTransform, "n_transforms", lambda t: t[0].fwd_len.__setitem__(0, 1000)),
    "transform.inv_first": ("transforms", _capi.EuclTransform, "n_transforms", lambda t: t[0].inv_first.__setitem__(1, -2)),
    "expr.op": ("expr_ops", _capi.EuclExprOp, "n_expr_ops", lambda t: setattr(t[0], "op", 42)),
    "expr.var": ("expr_ops", _capi.EuclExprOp, "n_expr_ops", lambda t: (setattr(t[0], "op", 1), setattr(t[0], "arg", 3))),
    "expr.underflow": ("expr_ops", _capi.EuclExprOp, "n_expr_ops", lambda t: setattr(t[0], "op", 2)),
    "surface.color_len": ("surfaces", _capi.EuclSurface, "n_surfaces", lambda t: setattr(t[0], "color_len", 1000)),
    "surface.ratio_op": ("surfaces", _capi.EuclSurface, "n_surfaces", lambda t: setattr(t[0], "ratio_op", 2)),
    "surface.thr_op": ("surfaces", _capi.EuclSurface, "n_surfaces", lambda t: setattr(t[0], "thr_op", -1)),
    "color.op": ("color_ops", _capi.EuclColorOp, "n_color_ops", lambda t: setattr(t[0], "op", 6)),
    "color.blend_fn": ("color_ops", _capi.EuclColorOp, "n_color_ops", lambda t: setattr(t[2], "i0", 18)),
    "color.underflow": ("color_ops", _capi.EuclColorOp, "n_color_ops", lambda t: setattr(t[0], "op", 5)),
    "color.texture": ("color_ops", _capi.EuclColorOp, "n_color_ops", lambda t: (setattr(t[0], "op", 4), setattr(t[0], "i0", 3))),
}


@pytest.mark.parametrize("what", sorted(MUTATIONS))
def test_out_of_range_indices_and_opcodes_are_invalid_arguments(built_lib, what):
    """Caller-built flat scenes come in through the public C ABI: every index the device dereferences is checked."""
    env = PARSER.parse(scene(nested_blend(1)))
    base = env.flat
    assert accepted(create_status(env)[0])
    table, ctype, count_field, mutate = MUTATIONS[what]
    flat = _capi.EuclFlatScene()
    C.memmove(C.byref(flat), C.byref(base), C.sizeof(flat))
    arr = _copy_table(base, table, ctype, getattr(base, count_field))
    mutate(arr)
    setattr(flat, table, C.cast(arr, C.POINTER(ctype)))
    status, msg = create_status(env, flat)
    assert status == EUCL_ERR_INVALID_ARGUMENT, (what, status, msg)


def test_background_index_is_checked(built_lib):
    env = PARSER.parse(scene(uniform(1, 0, 0, 1)))
    flat = _capi.EuclFlatScene()
    C.memmove(C.byref(flat), C.byref(env.flat), C.sizeof(flat))
    flat.background = 5
    assert create_status(env, flat)[0] == EUCL_ERR_INVALID_ARGUMENT
