#!/usr/bin/env python
"""Writes this repo's OWN test scenes (in the reference's JSON vocabulary) and a small procedural
texture.  They exercise constructors the six benchmark scenes do not: every SetOperation, nested
CSG, Hyperplane entities, capped cylinders, every blend function, nearest-neighbour and surface
textures, LinearSpace with two chained transformations, Perlin with a time offset."""
import json
from pathlib import Path

import numpy as np
from PIL import Image

HERE = Path(__file__).resolve().parent


def P3(x, y, z): return {"Point3::new": [x, y, z]}
def V3(x, y, z): return {"Vector3::new": [x, y, z]}
def P4(x, y, z, w): return {"Point4::new": [x, y, z, w]}
def V4(x, y, z, w): return {"Vector4::new": [x, y, z, w]}
def rgba(r, g, b, a): return {"Rgba::new": [r, g, b, a]}
def op(name): return {"SetOperation": [name]}


def surface(d, ratio, thr, color):
    return {f"ComposableSurface{d}": {"reflection_ratio": ratio, "reflection_direction": {f"reflection_direction_specular_{d}": []},
                                     "threshold_direction": thr, "surface_color": color}}


def uniform(d, c): return {f"surface_color_uniform_{d}": [c]}
def illum(d, light, dark): return {f"surface_color_illumination_global_{d}": [light, dark]}
def blend(d, src, dst, fn): return {f"surface_color_blend_{d}": [src, dst, fn]}
def fresnel(d, a=1.458, b=1.0): return {f"reflection_ratio_fresnel_{d}": [a, b]}
def ratio(d, r): return {f"reflection_ratio_uniform_{d}": [r]}
def snell(d, n=1.458): return {f"threshold_direction_snell_{d}": [n]}
def ident(d): return {f"threshold_direction_identity_{d}": []}
def entity(d, shape, material, surf): return {f"Entity{d}Impl::new": [shape, material, surf]}
def vac(d): return {f"Vacuum{d}::new": []}


def tex(d, path, filt="texture_image_linear"):
    uv = {"uv_sphere_3": [P3(0, 0, 0)]}
    if d == 4:
        uv = {"uv_derank_4": [uv]}
    return {f"MappedTextureImpl{d}::new": [uv, {filt: [path]}]}


def texture_png():
    y, x = np.mgrid[0:16, 0:32]
    img = np.zeros((16, 32, 4), np.uint8)
    img[..., 0] = (x * 8) % 256
    img[..., 1] = (y * 16) % 256
    img[..., 2] = ((x ^ y) & 1) * 255
    img[..., 3] = 255 - (x % 4) * 40
    Image.fromarray(img, "RGBA").save(HERE / "checker_rgba.png")


def csg_mix_3d():
    d = 3
    grey = illum(d, rgba(1, 1, 1, 1), rgba(0.1, 0.1, 0.1, 1))
    glass = surface(d, fresnel(d), snell(d), uniform(d, rgba(0, 0, 0, 0)))
    ents = [
        # symmetric difference of two spheres, half transparent mirror
        entity(d, {"ComposableShape3::of": [[{"Sphere3::new": [P3(9, 3, 0), 1.5]}, {"Sphere3::new": [P3(9, 4.5, 0), 1.5]}],
                                            op("SymmetricDifference")]}, vac(d),
               surface(d, ratio(d, 0.3), ident(d), blend(d, grey, uniform(d, rgba(1, 0.5, 0, 0.6)), {"blend_function_multiply": []}))),
        # capped cylinder minus a sphere (Complement of an Intersection chain)
        entity(d, {"ComposableShape3::of": [[{"Cylinder3::new_with_height": [P3(8, -3, 0), V3(0.2, 0.1, 1), 1.2, 3]},
                                             {"Sphere3::new": [P3(8, -3, 1.2), 0.9]}], op("Complement")]}, vac(d), glass),
        # union of an intersection (lens) and a cuboid, surface texture
        entity(d, {"ComposableShape3::of": [[
            {"ComposableShape3::of": [[{"Sphere3::new": [P3(12, 0, 1.2), 1.5]}, {"Sphere3::new": [P3(12, 0, -0.2), 1.5]}], op("Intersection")]},
            {"HalfSpace3::cuboid": [P3(12, 0, -2), V3(2, 2, 1)]}], op("Union")]}, vac(d),
               surface(d, ratio(d, 0.0), ident(d), {"surface_color_texture_3": [tex(3, "./tests/scenes/checker_rgba.png", "texture_image_nearest_neighbor")]})),
        # a bare hyperplane with a perlin surface
        entity(d, {"Hyperplane3::new_with_point": [V3(0, 0, 1), P3(0, 0, -3)]}, vac(d),
               surface(d, ratio(d, 0.1), ident(d), blend(d, {"surface_color_perlin_hue_seed_3": [7, 2.0, 0.5]},
                                                        uniform(d, rgba(0.2, 0.2, 0.2, 1)), {"blend_function_ratio": [0.5]}))),
        # a void that stretches x then squeezes y: two chained transformations
        {"Entity3Impl::new_with_surface": [{"HalfSpace3::cuboid": [P3(6, 0, 0), V3(1, 4, 4)]},
                                           {"LinearSpace3": ["xyz", [
                                               {"ComponentTransformation3": [[{"ComponentTransformationExpr": ["x * 2", "x / 2"]},
                                                                              {"ComponentTransformationExpr": ["y", "y"]},
                                                                              {"ComponentTransformationExpr": ["z", "z"]}]]},
                                               {"ComponentTransformation3": [[{"ComponentTransformationExpr": ["x", "x"]},
                                                                              {"ComponentTransformationExpr": ["y / 2 + z * 0", "y * 2"]},
                                                                              {"ComponentTransformationExpr": ["z", "z"]}]]}]]},
                                           surface(d, ratio(d, 0.0), ident(d), uniform(d, rgba(0, 0, 0, 0)))]},
        {"Void3::new_with_vacuum": []},
    ]
    return {"Universe3": {"camera": {"FreeCamera3::new_with_location": [P3(-1, 0.2, 0.1)]}, "entities": ents,
                          "background": tex(3, "./tests/scenes/checker_rgba.png")}}


BLENDS = ["over", "inside", "outside", "atop", "xor", "plus", "multiply", "screen", "overlay", "darken", "lighten", "dodge",
          "burn", "hard_light", "soft_light", "difference", "exclusion"]


def blend_4d():
    d = 4
    ents = []
    for i, name in enumerate(BLENDS):
        y, z = (i % 6 - 2.5) * 2.2, (i // 6 - 1) * 2.4
        col = blend(d, illum(d, rgba(0.9, 0.8, 0.3, 0.9), rgba(0.1, 0.2, 0.6, 0.4)),
                    {"surface_color_illumination_directional_4": [V4(0.3, -0.2, -1, 0.1), {"Rgba::from_hsva": [i * 21.0, 0.8, 0.9, 0.7]},
                                                                   {"Rgba::new_u8": [20, 40, 60, 200]}]},
                    {f"blend_function_{name}": []})
        ents.append(entity(d, {"Sphere4::new": [P4(10, y, z, 0.1 * i), 1.0]}, vac(d), surface(d, ratio(d, 0.15), ident(d), col)))
    ents.append(entity(d, {"HalfSpace4::hypercuboid": {"center": P4(14, 0, 0, 0), "dimensions": V4(1, 16, 9, 4)}}, vac(d),
                       surface(d, fresnel(d), snell(d, 1.3), uniform(d, rgba(0.1, 0.3, 0.1, 0.2)))))
    ents.append({"Void4::new_with_vacuum": []})
    return {"Universe4": {"camera": {"FreeCamera4::new_with_location": [P4(0, 0, 0, 0.3)]}, "entities": ents,
                          "background": tex(4, "./tests/scenes/checker_rgba.png")}}


def no_void_3d():
    """No Void entity: the camera is in no entity -> the 8-pixel checkerboard (mod.rs:387-395)."""
    d = 3
    return {"Universe3": {"camera": {"PitchYawCamera3": []},
                          "entities": [entity(d, {"Sphere3::new": [P3(10, 0, 0), 3]}, vac(d),
                                              surface(d, ratio(d, 0.0), ident(d), uniform(d, rgba(1, 0, 0, 1))))],
                          "background": tex(3, "./tests/scenes/checker_rgba.png")}}


if __name__ == "__main__":
    texture_png()
    for name, scene in (("csg_mix_3d", csg_mix_3d()), ("blend_4d", blend_4d()), ("no_void_3d", no_void_3d())):
        (HERE / f"{name}.json").write_text(json.dumps(scene, indent=1))
        print("wrote", name)
