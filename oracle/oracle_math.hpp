// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product: only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may build, load or call anything under oracle/.
//
// CPU restatement (C++17, f64, -ffp-contract=off) of the math layer the reference hot path uses:
//   /root/reference/src/vec3.rs, src/ray.rs, src/interval.rs and the parts of glam 0.29.2
//   (Cargo.lock:372-374; NOT vendored under /root/reference — restated from its published scalar
//   f64 implementation) that those files call.
//
// PARITY UNPINNED at the third-party boundary: the reference ships no tests/golden vectors and cannot
// be compiled here (no Rust toolchain), so glam/rand semantics below are restated from the published
// algorithms.  The only reference-produced pins are demo/*.png (see tests/golden/README.md).
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>

namespace orc {

constexpr double PI = 3.14159265358979323846264338327950288;  // std::f64::consts::PI
constexpr double INF = std::numeric_limits<double>::infinity();

// Rust f64::min/max ignore a NaN operand (same as C fmin/fmax).
inline double fmin_(double a, double b) { return std::fmin(a, b); }
inline double fmax_(double a, double b) { return std::fmax(a, b); }
// Rust f64::clamp: NaN stays NaN.
inline double clamp_(double x, double lo, double hi) {
    if (x < lo) x = lo;
    if (x > hi) x = hi;
    return x;
}
// Rust f64::signum: +1 for +0.0 and positives, -1 for -0.0 and negatives, NaN for NaN.
inline double signum_(double x) { return std::isnan(x) ? x : std::copysign(1.0, x); }
// f64::powi lowered by LLVM to multiplications (exponentiation by squaring).
inline double powi2(double x) { return x * x; }
inline double powi5(double x) { double x2 = x * x; double x4 = x2 * x2; return x4 * x; }
inline double to_radians(double deg) { return deg * (PI / 180.0); }  // f64::to_radians

struct Vec3 {
    double x = 0, y = 0, z = 0;
    Vec3() = default;
    Vec3(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
    static Vec3 splat(double v) { return {v, v, v}; }
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator*(Vec3 a, Vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline Vec3 operator/(Vec3 a, Vec3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline Vec3 operator*(Vec3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline Vec3 operator*(double s, Vec3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline Vec3 operator/(Vec3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
inline Vec3 operator-(double s, Vec3 a) { return {s - a.x, s - a.y, s - a.z}; }  // f64 - DVec3
inline Vec3 operator-(Vec3 a) { return {-a.x, -a.y, -a.z}; }
inline Vec3& operator+=(Vec3& a, Vec3 b) { a = a + b; return a; }
inline Vec3& operator*=(Vec3& a, Vec3 b) { a = a * b; return a; }
inline Vec3& operator*=(Vec3& a, double s) { a = a * s; return a; }
inline Vec3& operator/=(Vec3& a, double s) { a = a / s; return a; }
inline bool operator==(Vec3 a, Vec3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

// glam DVec3::dot: (x*x) + (y*y) + (z*z), left to right.
inline double dot(Vec3 a, Vec3 b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }
// glam DVec3::cross.
inline Vec3 cross(Vec3 a, Vec3 b) {
    return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}
inline double length_squared(Vec3 a) { return dot(a, a); }
inline double length(Vec3 a) { return std::sqrt(dot(a, a)); }
// glam normalize: self * length_recip(), length_recip = length().recip().
inline Vec3 normalize(Vec3 a) { return a * (1.0 / length(a)); }
inline Vec3 vmin(Vec3 a, Vec3 b) { return {fmin_(a.x, b.x), fmin_(a.y, b.y), fmin_(a.z, b.z)}; }
inline Vec3 vmax(Vec3 a, Vec3 b) { return {fmax_(a.x, b.x), fmax_(a.y, b.y), fmax_(a.z, b.z)}; }
inline double max_element(Vec3 a) { return fmax_(a.x, fmax_(a.y, a.z)); }
inline double min_element(Vec3 a) { return fmin_(a.x, fmin_(a.y, a.z)); }
inline Vec3 recip(Vec3 a) { return {1.0 / a.x, 1.0 / a.y, 1.0 / a.z}; }
// glam reflect: self - 2.0 * self.dot(normal) * normal
inline Vec3 reflect(Vec3 v, Vec3 n) { return v - (2.0 * dot(v, n)) * n; }
// glam refract.
inline Vec3 refract(Vec3 v, Vec3 n, double eta) {
    double n_dot_i = dot(n, v);
    double k = 1.0 - eta * eta * (1.0 - n_dot_i * n_dot_i);
    if (k >= 0.0) return eta * v - (eta * n_dot_i + std::sqrt(k)) * n;
    return Vec3(0, 0, 0);
}
// glam DVec3::lerp: self * (1 - s) + rhs * s ; FloatExt::lerp for f64: self + (rhs - self) * t
inline Vec3 lerp(Vec3 a, Vec3 b, double s) { return a * (1.0 - s) + b * s; }
inline double lerp(double a, double b, double t) { return a + (b - a) * t; }
// vec3.rs:36-44
inline double luminance(Vec3 c) { return 0.2126 * c.x + 0.7152 * c.y + 0.0722 * c.z; }

struct Quat { double x, y, z, w; };
inline Quat quat_normalize(Quat q) {
    double len = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    double r = 1.0 / len;
    return {q.x * r, q.y * r, q.z * r, q.w * r};
}
inline Quat quat_inverse(Quat q) { return {-q.x, -q.y, -q.z, q.w}; }  // conjugate
// glam DQuat::mul_vec3 (scalar path).
inline Vec3 quat_mul_vec3(Quat q, Vec3 v) {
    double w = q.w;
    Vec3 b(q.x, q.y, q.z);
    double b2 = dot(b, b);
    return v * (w * w - b2) + b * (dot(v, b) * 2.0) + cross(b, v) * (w * 2.0);
}
// glam DQuat::from_axis_angle
inline Quat quat_from_axis_angle(Vec3 axis, double angle) {
    double s = std::sin(angle * 0.5), c = std::cos(angle * 0.5);
    Vec3 v = axis * s;
    return {v.x, v.y, v.z, c};
}
// vec3.rs:23-29
inline Quat get_rotation_to_z(Vec3 n) {
    if (n.z < -0.99999) return {1.0, 0.0, 0.0, 0.0};
    return quat_normalize({n.y, -n.x, 0.0, 1.0 + n.z});
}

// Column-major 4x4 like glam::DMat4: c[col][row].
struct Mat4 {
    double c[4][4];
};
inline Mat4 mat4_from_rotation_translation(Quat r, Vec3 t) {
    double x = r.x, y = r.y, z = r.z, w = r.w;
    double x2 = x + x, y2 = y + y, z2 = z + z;
    double xx = x * x2, xy = x * y2, xz = x * z2;
    double yy = y * y2, yz = y * z2, zz = z * z2;
    double wx = w * x2, wy = w * y2, wz = w * z2;
    Mat4 m;
    m.c[0][0] = 1.0 - (yy + zz); m.c[0][1] = xy + wz; m.c[0][2] = xz - wy; m.c[0][3] = 0.0;
    m.c[1][0] = xy - wz; m.c[1][1] = 1.0 - (xx + zz); m.c[1][2] = yz + wx; m.c[1][3] = 0.0;
    m.c[2][0] = xz + wy; m.c[2][1] = yz - wx; m.c[2][2] = 1.0 - (xx + yy); m.c[2][3] = 0.0;
    m.c[3][0] = t.x; m.c[3][1] = t.y; m.c[3][2] = t.z; m.c[3][3] = 1.0;
    return m;
}
inline Mat4 mat4_from_quat(Quat r) { return mat4_from_rotation_translation(r, Vec3(0, 0, 0)); }
inline Mat4 mat4_transpose(const Mat4& m) {
    Mat4 t;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) t.c[i][j] = m.c[j][i];
    return t;
}
// glam DMat4::inverse (scalar cofactor form).
inline Mat4 mat4_inverse(const Mat4& m) {
    double m00 = m.c[0][0], m01 = m.c[0][1], m02 = m.c[0][2], m03 = m.c[0][3];
    double m10 = m.c[1][0], m11 = m.c[1][1], m12 = m.c[1][2], m13 = m.c[1][3];
    double m20 = m.c[2][0], m21 = m.c[2][1], m22 = m.c[2][2], m23 = m.c[2][3];
    double m30 = m.c[3][0], m31 = m.c[3][1], m32 = m.c[3][2], m33 = m.c[3][3];
    double coef00 = m22 * m33 - m32 * m23, coef02 = m12 * m33 - m32 * m13, coef03 = m12 * m23 - m22 * m13;
    double coef04 = m21 * m33 - m31 * m23, coef06 = m11 * m33 - m31 * m13, coef07 = m11 * m23 - m21 * m13;
    double coef08 = m21 * m32 - m31 * m22, coef10 = m11 * m32 - m31 * m12, coef11 = m11 * m22 - m21 * m12;
    double coef12 = m20 * m33 - m30 * m23, coef14 = m10 * m33 - m30 * m13, coef15 = m10 * m23 - m20 * m13;
    double coef16 = m20 * m32 - m30 * m22, coef18 = m10 * m32 - m30 * m12, coef19 = m10 * m22 - m20 * m12;
    double coef20 = m20 * m31 - m30 * m21, coef22 = m10 * m31 - m30 * m11, coef23 = m10 * m21 - m20 * m11;
    double fac0[4] = {coef00, coef00, coef02, coef03}, fac1[4] = {coef04, coef04, coef06, coef07};
    double fac2[4] = {coef08, coef08, coef10, coef11}, fac3[4] = {coef12, coef12, coef14, coef15};
    double fac4[4] = {coef16, coef16, coef18, coef19}, fac5[4] = {coef20, coef20, coef22, coef23};
    double vec0[4] = {m10, m00, m00, m00}, vec1[4] = {m11, m01, m01, m01};
    double vec2[4] = {m12, m02, m02, m02}, vec3[4] = {m13, m03, m03, m03};
    const double sign_a[4] = {1.0, -1.0, 1.0, -1.0}, sign_b[4] = {-1.0, 1.0, -1.0, 1.0};
    Mat4 inv;
    for (int i = 0; i < 4; i++) {
        double inv0 = (vec1[i] * fac0[i] - vec2[i] * fac1[i]) + vec3[i] * fac2[i];
        double inv1 = (vec0[i] * fac0[i] - vec2[i] * fac3[i]) + vec3[i] * fac4[i];
        double inv2 = (vec0[i] * fac1[i] - vec1[i] * fac3[i]) + vec3[i] * fac5[i];
        double inv3 = (vec0[i] * fac2[i] - vec1[i] * fac4[i]) + vec2[i] * fac5[i];
        inv.c[0][i] = inv0 * sign_a[i];
        inv.c[1][i] = inv1 * sign_b[i];
        inv.c[2][i] = inv2 * sign_a[i];
        inv.c[3][i] = inv3 * sign_b[i];
    }
    double d0 = m.c[0][0] * inv.c[0][0], d1 = m.c[0][1] * inv.c[1][0];
    double d2 = m.c[0][2] * inv.c[2][0], d3 = m.c[0][3] * inv.c[3][0];
    double det = d0 + d1 + d2 + d3;
    double rcp = 1.0 / det;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) inv.c[i][j] = inv.c[i][j] * rcp;
    return inv;
}
// glam DMat4::transform_point3 / transform_vector3.
inline Vec3 transform_point3(const Mat4& m, Vec3 p) {
    double r[3];
    for (int i = 0; i < 3; i++) {
        double res = m.c[0][i] * p.x;
        res = m.c[1][i] * p.y + res;
        res = m.c[2][i] * p.z + res;
        res = m.c[3][i] + res;
        r[i] = res;
    }
    return {r[0], r[1], r[2]};
}
inline Vec3 transform_vector3(const Mat4& m, Vec3 v) {
    double r[3];
    for (int i = 0; i < 3; i++) {
        double res = m.c[0][i] * v.x;
        res = m.c[1][i] * v.y + res;
        res = m.c[2][i] * v.z + res;
        r[i] = res;
    }
    return {r[0], r[1], r[2]};
}

// src/interval.rs
struct Interval {
    double min, max;
    bool contains(double x) const { return min <= x && x <= max; }   // interval.rs:26-28
    bool surrounds(double x) const { return min < x && x < max; }    // interval.rs:30-32
};

// src/ray.rs — Ray::new normalises the direction (ray.rs:23-29).
struct Ray {
    Vec3 origin, direction;
    double time;
    static Ray make(Vec3 o, Vec3 d, double time) { return Ray{o, normalize(d), time}; }
    Vec3 at(double t) const { return origin + direction * t; }
};

// ---- counter-based RNG shared by the oracle and the device (DESIGN.md "RNG contract") ----
// Philox4x32-10, key = (seed_lo, seed_hi), counter = (draw/2, pixel, sample, 0).  Each block
// yields two 53-bit uniforms in [0,1) like rand 0.8.5's gen::<f64>() ((u64 >> 11) * 2^-53).
struct Philox {
    static inline void block(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                             uint32_t out[4]) {
        for (int r = 0; r < 10; r++) {
            uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
            uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
            uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    }
};
struct Rng {
    // mode A: explicit uniforms (unit tests feed the same numbers to oracle and device)
    const double* arr = nullptr; int arr_n = 0;
    // mode B: Philox stream of path (pixel, sample)
    uint64_t seed = 0; uint32_t pixel = 0, sample = 0;
    uint32_t used = 0;
    uint32_t cached_block = 0xFFFFFFFFu; double cached[2];
    double next() {
        uint32_t k = used++;
        if (arr) return (int)k < arr_n ? arr[k] : 0.5;
        uint32_t b = k >> 1;
        if (b != cached_block) {
            uint32_t o[4];
            Philox::block((uint32_t)seed, (uint32_t)(seed >> 32), b, pixel, sample, 0u, o);
            cached[0] = (double)((((uint64_t)o[0] << 32) | o[1]) >> 11) * (1.0 / 9007199254740992.0);
            cached[1] = (double)((((uint64_t)o[2] << 32) | o[3]) >> 11) * (1.0 / 9007199254740992.0);
            cached_block = b;
        }
        return cached[k & 1];
    }
};

// Keyed uniforms for decisions taken INSIDE an intersection test (the constant-density medium of include/pt_b200.h —
// ours, the reference's volume.rs is a stub): they must not depend on traversal order, so they are addressed by
// (path, bounce, stream) instead of being drawn sequentially: uniform #0 of philox(key = seed, counter = (bounce, pixel,
// sample, 1 + stream)).  The integrator sets the key before each intersect_all; ray batches use (0, ray index, 0, 0).
struct PathKey { uint64_t seed = 0; uint32_t pixel = 0, sample = 0, bounce = 0; };
inline thread_local PathKey g_path_key;
inline double keyed_uniform(uint32_t stream) {
    uint32_t o[4];
    Philox::block((uint32_t)g_path_key.seed, (uint32_t)(g_path_key.seed >> 32), g_path_key.bounce, g_path_key.pixel, g_path_key.sample, 1u + stream, o);
    return (double)((((uint64_t)o[0] << 32) | o[1]) >> 11) * (1.0 / 9007199254740992.0);
}

}  // namespace orc
