// Camera rotations: the part of the reference's per-frame `Camera::update` that turns the view before
// the hot path runs.  Pure host arithmetic on an EuclCamera pose;
// the translation part (`trace_path_unknown` through voids) runs on the device (eucl_trace_path).
//
// Reference: PitchYawCamera3::{rotate_yaw_static, rotate_pitch_static} (src/universe/d3/entity/
// camera.rs:110-136), FreeCamera3::{rotate_yaw_static, rotate_roll_static} (:329-337),
// FreeCamera4::update_rotation (src/universe/d4/entity/camera.rs:68-130), util::find_orthonormal_4 /
// reorthonormalize_4 (src/util.rs:301-322).  Mouse / keyboard decoding (which angle, which axes) is the
// caller's business, as it is glium's in the reference.
//
// nalgebra 0.8.2's UnitQuaternion is not vendored with the reference; its `new(axis * angle)` and
// `rotate` are restated from the published algorithm (oracle/ASSUMPTIONS.md): parity unpinned.
#include <cmath>
#include <string>

#include "error.h"
#include "euclider_b200.h"

namespace {

struct V3 {
    double x, y, z;
};
V3 load3(const double* p) { return V3{p[0], p[1], p[2]}; }
void store3(double* p, const V3& v) {
    p[0] = v.x;
    p[1] = v.y;
    p[2] = v.z;
}
V3 operator+(const V3& a, const V3& b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
V3 operator*(const V3& a, double s) { return V3{a.x * s, a.y * s, a.z * s}; }
V3 operator-(const V3& a) { return V3{-a.x, -a.y, -a.z}; }
double dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
V3 cross(const V3& a, const V3& b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
double norm(const V3& a) { return std::sqrt(dot(a, a)); }
V3 normalize(const V3& a) {
    const double n = norm(a);
    return V3{a.x / n, a.y / n, a.z / n};
}
// util.rs:712-722: acos(dot / (|a| |b|))
double angle_between(const V3& a, const V3& b) { return std::acos(dot(a, b) / (norm(a) * norm(b))); }

// UnitQuaternion::new(axisangle): rotation by |axisangle| about axisangle (identity for the zero vector)
struct Quat {
    double w;
    V3 v;
};
Quat quat_from_axis_angle(const V3& axisangle) {
    const double sq = dot(axisangle, axisangle);
    if (sq == 0.0) return Quat{1.0, V3{0.0, 0.0, 0.0}};
    const double ang = std::sqrt(sq);
    const double s = std::sin(ang / 2.0), c = std::cos(ang / 2.0);
    const double s_ang = s / ang;
    return Quat{c, axisangle * s_ang};
}
// UnitQuaternion * Vector3: v + 2 w (q x v) + q x (2 (q x v))
V3 rotate(const Quat& q, const V3& v) {
    V3 t = cross(q.v, v);
    t = t * 2.0;
    return (t * q.w + cross(q.v, t)) + v;
}

bool pose3(const EuclCamera* cam, const char* who) {
    if (!cam || cam->dim != 3) {
        eucl::fail(EUCL_ERR_INVALID_ARGUMENT, std::string(who) + ": needs a 3-D camera");
        return false;
    }
    return true;
}

// 4-D generalised cross product: the vector orthogonal to a, b, c given by the formal determinant
// | e_x e_y e_z e_w ; a ; b ; c | expanded along the first row (util.rs:301-307)
void find_orthonormal_4(const double* a, const double* b, const double* c, double* out) {
    auto det3 = [](double m00, double m01, double m02, double m10, double m11, double m12, double m20, double m21, double m22) {
        return m00 * (m11 * m22 - m12 * m21) - m01 * (m10 * m22 - m12 * m20) + m02 * (m10 * m21 - m11 * m20);
    };
    out[0] = det3(a[1], a[2], a[3], b[1], b[2], b[3], c[1], c[2], c[3]);
    out[1] = -det3(a[0], a[2], a[3], b[0], b[2], b[3], c[0], c[2], c[3]);
    out[2] = det3(a[0], a[1], a[3], b[0], b[1], b[3], c[0], c[1], c[3]);
    out[3] = -det3(a[0], a[1], a[2], b[0], b[1], b[2], c[0], c[1], c[2]);
}
void normalize4(double* v) {
    const double n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3]);
    for (int k = 0; k < 4; ++k) v[k] /= n;
}

} // namespace

extern "C" {

// PitchYawCamera3 (about_up = 0): both forward and up turn about +z.  FreeCamera3 (about_up = 1): forward
// turns about the camera's own up vector.
int eucl_camera_rotate_yaw(EuclCamera* cam, double angle, int about_up) {
    if (!pose3(cam, "eucl_camera_rotate_yaw")) return EUCL_ERR_INVALID_ARGUMENT;
    V3 forward = load3(cam->forward), up = load3(cam->up);
    if (about_up) {
        const Quat q = quat_from_axis_angle(up * angle);
        forward = normalize(rotate(q, forward));
    } else {
        const Quat q = quat_from_axis_angle(V3{0.0, 0.0, 1.0} * angle);
        forward = normalize(rotate(q, forward));
        up = normalize(rotate(q, up));
    }
    store3(cam->forward, forward);
    store3(cam->up, up);
    return EUCL_OK;
}

// rotate_pitch_static: turn about the horizontal axis forward x up; with `snap` (PitchYawCamera3) the view stops
// at straight up / straight down instead of flipping over.
int eucl_camera_rotate_pitch(EuclCamera* cam, double angle, int snap) {
    if (!pose3(cam, "eucl_camera_rotate_pitch")) return EUCL_ERR_INVALID_ARGUMENT;
    V3 forward = load3(cam->forward), up = load3(cam->up);
    const V3 axis_h = normalize(cross(forward, up));
    const V3 z{0.0, 0.0, 1.0};
    const double pi = 3.14159265358979323846264338327950288;
    bool snapped = false;
    if (snap) {
        const double result_angle = angle_between(forward, z);
        if (result_angle < angle) {
            forward = z;
            snapped = true;
        } else if (pi - result_angle < -angle) {
            forward = -z;
            snapped = true;
        }
    }
    if (!snapped) {
        const Quat q = quat_from_axis_angle(axis_h * angle);
        forward = normalize(rotate(q, forward));
    }
    up = normalize(cross(axis_h, forward));
    store3(cam->forward, forward);
    store3(cam->up, up);
    return EUCL_OK;
}

// FreeCamera3::rotate_roll_static: up turns about forward
int eucl_camera_rotate_roll(EuclCamera* cam, double angle) {
    if (!pose3(cam, "eucl_camera_rotate_roll")) return EUCL_ERR_INVALID_ARGUMENT;
    const V3 forward = load3(cam->forward);
    const Quat q = quat_from_axis_angle(forward * angle);
    store3(cam->up, normalize(rotate(q, load3(cam->up))));
    return EUCL_OK;
}

// FreeCamera4::update_rotation: rotation by `angle` in the plane of two of the camera's OWN axes
// (0 forward, 1 left, 2 up, 3 ana = the vector orthogonal to the other three), applied to forward,
// left and up, followed by reorthonormalize_4.
int eucl_camera_rotate_plane4(EuclCamera* cam, int axis_a, int axis_b, double angle) {
    if (!cam || cam->dim != 4) return eucl::fail(EUCL_ERR_INVALID_ARGUMENT, "eucl_camera_rotate_plane4: needs a 4-D camera");
    if (axis_a < 0 || axis_a > 3 || axis_b < 0 || axis_b > 3 || axis_a == axis_b)
        return eucl::fail(EUCL_ERR_INVALID_ARGUMENT, "eucl_camera_rotate_plane4: two different axes in 0..3");
    bool chosen[4] = {false, false, false, false};
    chosen[axis_a] = chosen[axis_b] = true;
    double rot[4][4];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            rot[r][c] = r == c ? 1.0 : 0.0;
            if (chosen[r] && chosen[c]) rot[r][c] = r == c ? std::cos(angle) : (r < c ? -std::sin(angle) : std::sin(angle));
        }
    double ana[4];
    find_orthonormal_4(cam->forward, cam->left, cam->up, ana);
    // normalize_matrix: columns forward, left, up, ana
    double m[4][4];
    for (int r = 0; r < 4; ++r) {
        m[r][0] = cam->forward[r];
        m[r][1] = cam->left[r];
        m[r][2] = cam->up[r];
        m[r][3] = ana[r];
    }
    auto apply = [&](double* v) { // m * (rot * (m^T * v)), each product accumulated from zero like nalgebra
        double a[4], b[4], c[4];
        for (int i = 0; i < 4; ++i) {
            double acc = 0.0;
            for (int k = 0; k < 4; ++k) acc = acc + m[k][i] * v[k];
            a[i] = acc;
        }
        for (int i = 0; i < 4; ++i) {
            double acc = 0.0;
            for (int k = 0; k < 4; ++k) acc = acc + rot[i][k] * a[k];
            b[i] = acc;
        }
        for (int i = 0; i < 4; ++i) {
            double acc = 0.0;
            for (int k = 0; k < 4; ++k) acc = acc + m[i][k] * b[k];
            c[i] = acc;
        }
        for (int i = 0; i < 4; ++i) v[i] = c[i];
    };
    apply(cam->forward);
    apply(cam->left);
    apply(cam->up);
    find_orthonormal_4(cam->forward, cam->left, cam->up, ana); // to_ana() of the rotated frame
    // reorthonormalize_4(forward, left, up, ana): util.rs:309-322
    double t[4];
    find_orthonormal_4(cam->up, cam->left, ana, t);
    normalize4(t);
    for (int k = 0; k < 4; ++k) cam->forward[k] = t[k];
    find_orthonormal_4(cam->forward, cam->up, ana, t);
    normalize4(t);
    for (int k = 0; k < 4; ++k) cam->left[k] = t[k];
    find_orthonormal_4(cam->forward, ana, cam->left, t);
    normalize4(t);
    for (int k = 0; k < 4; ++k) cam->up[k] = t[k];
    return EUCL_OK;
}

} // extern "C"
