"""Camera rotations (euclider_b200/csrc/host/camera.cc) against an independent numpy restatement: rotation
matrices by Rodrigues' formula instead of the library's quaternion form, 4-D plane rotations as explicit
matrices.  nalgebra itself is not available (parity unpinned, oracle/ASSUMPTIONS.md), so the tolerance is a
few ulps of accumulated rounding rather than bit equality.  Host arithmetic only: no GPU needed."""
import ctypes as C

import numpy as np
import pytest

import euclider_b200 as eb
from euclider_b200._capi import EuclCamera

TOL = 1e-13


def cam3(forward=(1, 0, 0), up=(0, 0, 1)):
    c = EuclCamera()
    c.dim = 3
    for k in range(3):
        c.forward[k], c.up[k] = float(forward[k]), float(up[k])
    return c


def vec(c, name, n=3):
    return np.array([getattr(c, name)[k] for k in range(n)])


def rodrigues(axis, angle, v):
    axis = np.asarray(axis, float)
    axis = axis / np.linalg.norm(axis)
    v = np.asarray(v, float)
    return v * np.cos(angle) + np.cross(axis, v) * np.sin(angle) + axis * np.dot(axis, v) * (1 - np.cos(angle))


def unit(v):
    return v / np.linalg.norm(v)


@pytest.mark.parametrize("angle", [0.0, 0.3, -1.1, 2.9, 7.0])
def test_pitch_yaw_camera_yaw_turns_forward_and_up_about_z(built_lib, angle):
    f0, u0 = unit(np.array([0.8, 0.1, 0.3])), None
    u0 = unit(np.cross(np.cross(f0, [0, 0, 1]), f0))
    c = cam3(f0, u0)
    assert built_lib.eucl_camera_rotate_yaw(C.byref(c), angle, 0) == 0
    assert np.allclose(vec(c, "forward"), unit(rodrigues([0, 0, 1], angle, f0)), atol=TOL)
    assert np.allclose(vec(c, "up"), unit(rodrigues([0, 0, 1], angle, u0)), atol=TOL)


@pytest.mark.parametrize("angle", [0.25, -0.6])
def test_free_camera_yaw_and_roll_turn_about_the_cameras_own_axes(built_lib, angle):
    f0 = unit(np.array([0.5, -0.4, 0.2]))
    u0 = unit(np.cross(np.cross(f0, [0.1, 0.2, 1.0]), f0))
    c = cam3(f0, u0)
    assert built_lib.eucl_camera_rotate_yaw(C.byref(c), angle, 1) == 0
    f1 = unit(rodrigues(u0, angle, f0))
    assert np.allclose(vec(c, "forward"), f1, atol=TOL) and np.allclose(vec(c, "up"), u0, atol=TOL)
    assert built_lib.eucl_camera_rotate_roll(C.byref(c), angle) == 0
    assert np.allclose(vec(c, "up"), unit(rodrigues(f1, angle, u0)), atol=TOL)
    assert abs(np.dot(vec(c, "up"), vec(c, "forward"))) < 1e-12


def test_pitch_turns_about_forward_cross_up_and_rebuilds_up(built_lib):
    c = cam3()
    assert built_lib.eucl_camera_rotate_pitch(C.byref(c), 0.4, 0) == 0
    axis_h = np.cross([1, 0, 0], [0, 0, 1])
    f1 = unit(rodrigues(axis_h, 0.4, [1, 0, 0]))
    assert np.allclose(vec(c, "forward"), f1, atol=TOL)
    assert np.allclose(vec(c, "up"), unit(np.cross(unit(axis_h), f1)), atol=TOL)


def test_pitch_snaps_at_the_poles_for_the_pitch_yaw_camera(built_lib):
    """rotate_pitch_static with snap (d3/entity/camera.rs:119-131): the angle to +z is smaller than the step
    -> forward becomes exactly +z; the free camera (snap off) turns past the pole instead."""
    f0 = unit(np.array([0.2, 0.0, 1.0]))
    u0 = unit(np.cross(np.cross(f0, [0, 0, 1]), f0))
    c = cam3(f0, u0)
    assert built_lib.eucl_camera_rotate_pitch(C.byref(c), 0.5, 1) == 0
    assert list(vec(c, "forward")) == [0.0, 0.0, 1.0]
    assert abs(np.linalg.norm(vec(c, "up")) - 1) < 1e-15 and abs(vec(c, "up")[2]) < 1e-15
    c = cam3(f0, u0)
    assert built_lib.eucl_camera_rotate_pitch(C.byref(c), -3.0, 1) == 0  # towards -z, past the pole
    assert list(vec(c, "forward")) == [-0.0, -0.0, -1.0]
    c = cam3(f0, u0)
    assert built_lib.eucl_camera_rotate_pitch(C.byref(c), 0.5, 0) == 0
    assert vec(c, "forward")[2] < 1.0


def cam4():
    c = EuclCamera()
    c.dim = 4
    rng = np.random.default_rng(5)
    q, _ = np.linalg.qr(rng.normal(size=(4, 4)))
    for k in range(4):
        c.forward[k], c.left[k], c.up[k] = q[k, 0], q[k, 1], q[k, 2]
    return c, q


def ana_of(f, l, u):
    """Generalised cross product: cofactors of the first row of | e ; f ; l ; u |."""
    m = np.array([f, l, u])
    return np.array([(-1) ** k * np.linalg.det(np.delete(m, k, axis=1)) for k in range(4)])


@pytest.mark.parametrize("axes", [(0, 1), (0, 3), (2, 3), (1, 2)])
def test_plane_rotation_of_the_4d_camera(built_lib, axes):
    c, q = cam4()
    angle = 0.37
    f0, l0, u0 = q[:, 0].copy(), q[:, 1].copy(), q[:, 2].copy()
    a0 = ana_of(f0, l0, u0)
    frame = np.stack([f0, l0, u0, a0], axis=1)
    rot = np.eye(4)
    i, j = sorted(axes)
    rot[i, i] = rot[j, j] = np.cos(angle)
    rot[i, j], rot[j, i] = -np.sin(angle), np.sin(angle)
    expect = [frame @ (rot @ (frame.T @ v)) for v in (f0, l0, u0)]
    assert built_lib.eucl_camera_rotate_plane4(C.byref(c), axes[0], axes[1], angle) == 0
    got = [vec(c, n, 4) for n in ("forward", "left", "up")]
    for g, e in zip(got, expect):  # reorthonormalize_4 of an already orthonormal frame changes rounding only
        assert np.allclose(g, e, atol=1e-12)
    gram = np.array([[np.dot(x, y) for y in got] for x in got])
    assert np.allclose(gram, np.eye(3), atol=1e-14)


def test_rotation_entry_points_check_the_camera_dimension(built_lib):
    c, _ = cam4()
    assert built_lib.eucl_camera_rotate_yaw(C.byref(c), 0.1, 0) == -1
    assert built_lib.eucl_camera_rotate_plane4(C.byref(cam3()), 0, 1, 0.1) == -1
    assert built_lib.eucl_camera_rotate_plane4(C.byref(c), 2, 2, 0.1) == -1


def test_environment_mirror(built_lib):
    env = eb.Parser.default().parse(MINIMAL, load_textures=False)
    env.rotate_yaw(np.pi / 2)
    assert np.allclose([env.camera.forward[k] for k in range(3)], [0, 1, 0], atol=1e-15)
    env.rotate_pitch(0.2, kind="FreeCamera3")
    env.rotate_roll(0.1)
    f = np.array([env.camera.forward[k] for k in range(3)])
    u = np.array([env.camera.up[k] for k in range(3)])
    assert abs(np.dot(f, u)) < 1e-14 and abs(np.linalg.norm(f) - 1) < 1e-15


MINIMAL = """{"Universe3": {"camera": {"PitchYawCamera3": []},
 "entities": [{"Void3::new_with_vacuum": []}],
 "background": {"MappedTextureImpl3::new": [{"uv_sphere_3": [{"Point3::new": [0, 0, 0]}]},
                                             {"texture_image_linear": ["./none.png"]}]}}}"""
