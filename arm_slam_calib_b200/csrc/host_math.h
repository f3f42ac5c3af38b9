// Small fixed-size fp64 rigid-transform helpers for the host side (generator, packer).
// The container has no Eigen/Boost, so this is all we need of them.
#pragma once
#include <cmath>
#include <cstring>

namespace asc {

struct Xf {              // rigid transform, row-major R (3x3) and t (3)
    double R[9];
    double t[3];
};

inline Xf xf_identity() {
    Xf x;
    std::memset(&x, 0, sizeof(x));
    x.R[0] = x.R[4] = x.R[8] = 1.0;
    return x;
}

inline Xf xf_from12(const double* m) {   // [R | t] row-major 3x4
    Xf x;
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) x.R[3 * r + c] = m[4 * r + c];
        x.t[r] = m[4 * r + 3];
    }
    return x;
}

inline void xf_to12(const Xf& x, double* m) {
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) m[4 * r + c] = x.R[3 * r + c];
        m[4 * r + 3] = x.t[r];
    }
}

inline Xf xf_mul(const Xf& a, const Xf& b) {
    Xf o;
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c)
            o.R[3 * r + c] = a.R[3 * r] * b.R[c] + a.R[3 * r + 1] * b.R[3 + c] + a.R[3 * r + 2] * b.R[6 + c];
        o.t[r] = a.R[3 * r] * b.t[0] + a.R[3 * r + 1] * b.t[1] + a.R[3 * r + 2] * b.t[2] + a.t[r];
    }
    return o;
}

inline Xf xf_inv(const Xf& a) {
    Xf o;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) o.R[3 * r + c] = a.R[3 * c + r];
    for (int r = 0; r < 3; ++r)
        o.t[r] = -(o.R[3 * r] * a.t[0] + o.R[3 * r + 1] * a.t[1] + o.R[3 * r + 2] * a.t[2]);
    return o;
}

// Rodrigues rotation about a unit axis (Eigen::AngleAxisd::toRotationMatrix equivalent)
inline Xf xf_rot(const double* axis, double angle) {
    const double s = std::sin(angle), c = std::cos(angle), v = 1.0 - c;
    const double x = axis[0], y = axis[1], z = axis[2];
    Xf o = xf_identity();
    o.R[0] = c + x * x * v;     o.R[1] = x * y * v - z * s; o.R[2] = x * z * v + y * s;
    o.R[3] = y * x * v + z * s; o.R[4] = c + y * y * v;     o.R[5] = y * z * v - x * s;
    o.R[6] = z * x * v - y * s; o.R[7] = z * y * v + x * s; o.R[8] = c + z * z * v;
    return o;
}

// p_local = T^-1 * p_world
inline void xf_transform_to(const Xf& T, const double* pw, double* pl) {
    const double d[3] = {pw[0] - T.t[0], pw[1] - T.t[1], pw[2] - T.t[2]};
    for (int c = 0; c < 3; ++c) pl[c] = T.R[c] * d[0] + T.R[3 + c] * d[1] + T.R[6 + c] * d[2];
}

}  // namespace asc
