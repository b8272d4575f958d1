"""Random scenes in the reference's JSON vocabulary: nested CSG of every SetOperation over every
primitive, boxes, capped cylinders, mirrors, glass, LinearSpace voids.  Used to stress the device CSG
evaluator (macro chains, plane chains, bound culling) against the oracle far beyond the shipped scenes."""
import json

import numpy as np


def _vec(d, xs, kind):
    return {f"{kind}{d}::new": [float(v) for v in xs[:d]]}


def random_scene(seed: int, dim: int, rooms: bool = True) -> str:
    rng = np.random.default_rng(seed)
    d = dim
    P = lambda xs: _vec(d, xs, "Point")
    V = lambda xs: _vec(d, xs, "Vector")

    def rpoint(scale=4.0, x0=8.0):
        p = rng.uniform(-scale, scale, 4)
        p[0] += x0
        return p

    def leaf():
        kind = rng.integers(0, 6)
        c = rpoint()
        if kind == 0:
            return {f"Sphere{d}::new": [P(c), float(rng.uniform(0.8, 3.0))]}
        if kind == 1:
            return {f"Cylinder{d}::new": [P(c), V(rng.normal(size=4)), float(rng.uniform(0.5, 2.0))]}
        if kind == 2:
            n = rng.normal(size=4)
            return {f"HalfSpace{d}::new_with_point": [{f"Hyperplane{d}::new_with_point": [V(n), P(c)]}, P(c + rng.normal(size=4))]}
        if kind == 3:
            dims = rng.uniform(1.0, 5.0, 4)
            return {"HalfSpace3::cuboid": [P(c), V(dims)]} if d == 3 else {"HalfSpace4::hypercuboid": [P(c), V(dims)]}
        if kind == 4 and d == 4:
            return {"Cylinder4::new_with_height": [P(c), V(rng.normal(size=4)), float(rng.uniform(0.5, 2.0)), float(rng.uniform(1, 5))]}
        if kind == 4:
            return {"Cylinder3::new_with_height": [P(c), V(rng.normal(size=4)), float(rng.uniform(0.5, 2.0)), float(rng.uniform(1, 5))]}
        return {f"Hyperplane{d}::new_with_point": [V(rng.normal(size=4)), P(c)]}

    def shape(depth, top=True):
        if depth == 0 or rng.random() < 0.35:
            return leaf()
        # a Complement whose `b` runs dry repeats `a` forever (shape.rs:392): nested under another
        # operation the reference never returns, so (like the shipped scenes) it mostly sits at the root
        ops = ["Union", "Intersection", "SymmetricDifference"] + (["Complement"] * 2 if top or rng.random() < 0.1 else [])
        op = ops[rng.integers(0, len(ops))]
        n = int(rng.integers(2, 5)) if op != "Complement" else 2
        return {f"ComposableShape{d}::of": [[shape(depth - 1, False) for _ in range(n)], {"SetOperation": [op]}]}

    def rgba():
        return {"Rgba::new": [float(v) for v in rng.uniform(0, 1, 4)]}

    def color(depth=2):
        k = rng.integers(0, 4 if depth else 3)
        if k == 0:
            return {f"surface_color_uniform_{d}": [rgba()]}
        if k == 1:
            return {f"surface_color_illumination_global_{d}": [rgba(), rgba()]}
        if k == 2:
            return {f"surface_color_illumination_directional_{d}": [V(rng.normal(size=4)), rgba(), rgba()]}
        fn = ["darken", "difference", "over", "multiply", "screen", "lighten", "xor", "exclusion"][rng.integers(0, 8)]
        return {f"surface_color_blend_{d}": [color(depth - 1), color(depth - 1), {f"blend_function_{fn}": []}]}

    def surface():
        if rng.random() < 0.4:
            n = float(rng.uniform(1.1, 1.8))
            return {f"ComposableSurface{d}": {
                "reflection_ratio": {f"reflection_ratio_fresnel_{d}": [n, 1.0]},
                "reflection_direction": {f"reflection_direction_specular_{d}": []},
                "threshold_direction": {f"threshold_direction_snell_{d}": [n]},
                "surface_color": {f"surface_color_uniform_{d}": [{"Rgba::new": [0, 0, 0, float(rng.uniform(0, 0.3))]}]}}}
        return {f"ComposableSurface{d}": {
            "reflection_ratio": {f"reflection_ratio_uniform_{d}": [float(rng.choice([0.0, 0.0, 0.3, 0.7, 1.0]))]},
            "reflection_direction": {f"reflection_direction_specular_{d}": []},
            "threshold_direction": {f"threshold_direction_identity_{d}": []},
            "surface_color": color()}}

    def material():
        if d == 3 and rng.random() < 0.25:
            k = float(rng.choice([2.0, 0.5, 3.0]))
            exprs = [{"ComponentTransformationExpr": [f"x * {k}", f"x / {k}"]}, {"ComponentTransformationExpr": ["y", "y"]},
                     {"ComponentTransformationExpr": ["z + 0", "z - 0"]}]
            return {"LinearSpace3": ["xyz", [{"ComponentTransformation3": [exprs]}]]}
        return {f"Vacuum{d}::new": []}

    entities = [{f"Entity{d}Impl::new": [shape(int(rng.integers(1, 4))), material(), surface()]} for _ in range(int(rng.integers(2, 7)))]
    if rooms:
        # every third scene is enclosed by a room described by its interior, Complement(VoidShape, X), like 4d_room's walls:
        # X a big box (a chain of half-spaces), a sphere, or a small CSG program around the camera
        rng2 = np.random.default_rng(seed + 7_000_000)
        if rng2.random() < 1.0 / 3.0:
            centre = np.concatenate([[6.0], rng2.uniform(-1, 1, 3)])
            kind = rng2.integers(0, 3)
            if kind == 0:
                dims = rng2.uniform(30.0, 44.0, 4)
                x = {"HalfSpace3::cuboid": [P(centre), V(dims)]} if d == 3 else {"HalfSpace4::hypercuboid": [P(centre), V(dims)]}
            elif kind == 1:
                x = {f"Sphere{d}::new": [P(centre), float(rng2.uniform(18.0, 26.0))]}
            else:
                x = {f"ComposableShape{d}::of": [[{f"Sphere{d}::new": [P(centre), 24.0]},
                                                 {f"Sphere{d}::new": [P(centre + np.array([9.0, 0, 0, 0])), 24.0]}],
                                                {"SetOperation": ["Intersection"]}]}
            room = {f"ComposableShape{d}::of": [[{f"VoidShape{d}": []}, x], {"SetOperation": ["Complement"]}]}
            entities.append({f"Entity{d}Impl::new": [room, {f"Vacuum{d}::new": []}, surface()]})
    entities.append({f"Void{d}::new_with_vacuum": []})
    uv = {"uv_sphere_3": [{"Point3::new": [0, 0, 0]}]}
    if d == 4:
        uv = {"uv_derank_4": [uv]}
    cam = {"FreeCamera3::new_with_location": [P(rng.uniform(-1, 1, 4))]} if d == 3 else {"FreeCamera4::new_with_location": [P(rng.uniform(-1, 1, 4))]}
    return json.dumps({f"Universe{d}": {"camera": cam, "entities": entities,
                                         "background": {f"MappedTextureImpl{d}::new": [uv, {"texture_image_linear": ["./tests/scenes/checker_rgba.png"]}]}}})
