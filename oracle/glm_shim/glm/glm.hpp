// TEST INFRASTRUCTURE — stand-in for g-truc/glm @ 0af55ccecd98d4e5a8d1fad7de25ba429d60e863
// (pinned by the reference in hab/DEPS:66-71, NOT vendored under /root/reference).
//
// Only what the reference's software-raster path touches is provided.  The
// functions that carry arithmetic (mat4 inverse / determinant / rotate and
// mat4*vec4) restate glm's published formulas with the same operation order,
// so float results match a real glm build:
//   * inverse      — glm/detail/func_matrix.inl  compute_inverse<4,4>
//   * determinant  — glm/detail/func_matrix.inl  compute_determinant<4,4>
//   * rotate       — glm/ext/matrix_transform.inl rotate()
//   * mat4 * vec4  — glm/detail/type_mat4x4.inl  operator*(mat, col_type)
// Call sites in the reference: src/geometry/matrix.cc:35,51,55,160,270-274,463.
// Everything else is a thin alias of <cmath>.
#pragma once

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>

namespace glm {

template <typename T>
struct tvec2 {
  union { T x, r, s; };
  union { T y, g, t; };
  constexpr tvec2() : x(T(0)), y(T(0)) {}
  constexpr tvec2(T a, T b) : x(a), y(b) {}
  constexpr explicit tvec2(T a) : x(a), y(a) {}
  T& operator[](int i) { return (&x)[i]; }
  const T& operator[](int i) const { return (&x)[i]; }
};

template <typename T>
struct tvec3 {
  union { T x, r, s; };
  union { T y, g, t; };
  union { T z, b, p; };
  constexpr tvec3() : x(T(0)), y(T(0)), z(T(0)) {}
  constexpr tvec3(T a, T b_, T c) : x(a), y(b_), z(c) {}
  constexpr explicit tvec3(T a) : x(a), y(a), z(a) {}
  T& operator[](int i) { return (&x)[i]; }
  const T& operator[](int i) const { return (&x)[i]; }
};

template <typename T>
struct tvec4 {
  union { T x, r, s; };
  union { T y, g, t; };
  union { T z, b, p; };
  union { T w, a, q; };
  constexpr tvec4() : x(T(0)), y(T(0)), z(T(0)), w(T(0)) {}
  constexpr tvec4(T a_, T b_, T c, T d) : x(a_), y(b_), z(c), w(d) {}
  constexpr explicit tvec4(T a_) : x(a_), y(a_), z(a_), w(a_) {}
  T& operator[](int i) { return (&x)[i]; }
  const T& operator[](int i) const { return (&x)[i]; }
  tvec4& operator+=(const tvec4& o) {
    x += o.x; y += o.y; z += o.z; w += o.w;
    return *this;
  }
  tvec4& operator-=(const tvec4& o) {
    x -= o.x; y -= o.y; z -= o.z; w -= o.w;
    return *this;
  }
  template <typename U>
  tvec4& operator*=(U s) {
    x *= s; y *= s; z *= s; w *= s;
    return *this;
  }
  template <typename U>
  tvec4& operator/=(U s) {
    x /= s; y /= s; z /= s; w /= s;
    return *this;
  }
};

template <typename T>
inline tvec3<T> operator*(const tvec3<T>& a, T s) { return {a.x * s, a.y * s, a.z * s}; }
template <typename T>
inline tvec3<T> operator*(T s, const tvec3<T>& a) { return {s * a.x, s * a.y, s * a.z}; }

template <typename T>
inline tvec4<T> operator+(const tvec4<T>& a, const tvec4<T>& b) {
  return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w};
}
template <typename T>
inline tvec4<T> operator-(const tvec4<T>& a, const tvec4<T>& b) {
  return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w};
}
template <typename T>
inline tvec4<T> operator*(const tvec4<T>& a, const tvec4<T>& b) {
  return {a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w};
}
template <typename T>
inline tvec4<T> operator*(const tvec4<T>& a, T s) {
  return {a.x * s, a.y * s, a.z * s, a.w * s};
}
template <typename T>
inline tvec4<T> operator*(T s, const tvec4<T>& a) {
  return {s * a.x, s * a.y, s * a.z, s * a.w};
}
template <typename T>
inline tvec4<T> operator/(const tvec4<T>& a, T s) {
  return {a.x / s, a.y / s, a.z / s, a.w / s};
}

typedef tvec2<float> vec2;
typedef tvec3<float> vec3;
typedef tvec4<float> vec4;
typedef tvec4<double> dvec4;
typedef tvec2<uint32_t> uvec2;
typedef tvec2<int32_t> ivec2;
typedef tvec3<int32_t> ivec3;
typedef tvec4<int32_t> ivec4;
typedef tvec4<uint64_t> u64vec4;
typedef tvec2<int32_t> i32vec2;

struct mat4 {
  vec4 c[4];
  mat4() : mat4(1.f) {}
  explicit mat4(float d) {
    c[0] = vec4(d, 0, 0, 0);
    c[1] = vec4(0, d, 0, 0);
    c[2] = vec4(0, 0, d, 0);
    c[3] = vec4(0, 0, 0, d);
  }
  mat4(const vec4& a, const vec4& b, const vec4& cc, const vec4& d) {
    c[0] = a; c[1] = b; c[2] = cc; c[3] = d;
  }
  vec4& operator[](int i) { return c[i]; }
  const vec4& operator[](int i) const { return c[i]; }
};

// glm/detail/type_mat4x4.inl — operator*(mat<4,4>, col_type): (m0*v0 + m1*v1) + (m2*v2 + m3*v3)
inline vec4 operator*(const mat4& m, const vec4& v) {
  vec4 const Mov0(v[0]);
  vec4 const Mov1(v[1]);
  vec4 const Mul0 = m[0] * Mov0;
  vec4 const Mul1 = m[1] * Mov1;
  vec4 const Add0 = Mul0 + Mul1;
  vec4 const Mov2(v[2]);
  vec4 const Mov3(v[3]);
  vec4 const Mul2 = m[2] * Mov2;
  vec4 const Mul3 = m[3] * Mov3;
  vec4 const Add1 = Mul2 + Mul3;
  vec4 const Add2 = Add0 + Add1;
  return Add2;
}

// glm/detail/type_mat4x4.inl — operator*(mat<4,4>, mat<4,4>)
inline mat4 operator*(const mat4& m1, const mat4& m2) {
  mat4 r(0.f);
  for (int i = 0; i < 4; i++) {
    r[i] = m1[0] * m2[i][0] + m1[1] * m2[i][1] + m1[2] * m2[i][2] + m1[3] * m2[i][3];
  }
  return r;
}

inline mat4 operator*(const mat4& m, float s) {
  return mat4(m[0] * s, m[1] * s, m[2] * s, m[3] * s);
}

// glm/detail/func_matrix.inl — compute_determinant<4,4>
inline float determinant(const mat4& m) {
  float SubFactor00 = m[2][2] * m[3][3] - m[3][2] * m[2][3];
  float SubFactor01 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
  float SubFactor02 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
  float SubFactor03 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
  float SubFactor04 = m[2][0] * m[3][2] - m[3][0] * m[2][2];
  float SubFactor05 = m[2][0] * m[3][1] - m[3][0] * m[2][1];

  vec4 DetCof(+(m[1][1] * SubFactor00 - m[1][2] * SubFactor01 + m[1][3] * SubFactor02),
              -(m[1][0] * SubFactor00 - m[1][2] * SubFactor03 + m[1][3] * SubFactor04),
              +(m[1][0] * SubFactor01 - m[1][1] * SubFactor03 + m[1][3] * SubFactor05),
              -(m[1][0] * SubFactor02 - m[1][1] * SubFactor04 + m[1][2] * SubFactor05));

  return m[0][0] * DetCof[0] + m[0][1] * DetCof[1] + m[0][2] * DetCof[2] + m[0][3] * DetCof[3];
}

// glm/detail/func_matrix.inl — compute_inverse<4,4>
inline mat4 inverse(const mat4& m) {
  float Coef00 = m[2][2] * m[3][3] - m[3][2] * m[2][3];
  float Coef02 = m[1][2] * m[3][3] - m[3][2] * m[1][3];
  float Coef03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];

  float Coef04 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
  float Coef06 = m[1][1] * m[3][3] - m[3][1] * m[1][3];
  float Coef07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];

  float Coef08 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
  float Coef10 = m[1][1] * m[3][2] - m[3][1] * m[1][2];
  float Coef11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];

  float Coef12 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
  float Coef14 = m[1][0] * m[3][3] - m[3][0] * m[1][3];
  float Coef15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];

  float Coef16 = m[2][0] * m[3][2] - m[3][0] * m[2][2];
  float Coef18 = m[1][0] * m[3][2] - m[3][0] * m[1][2];
  float Coef19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];

  float Coef20 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
  float Coef22 = m[1][0] * m[3][1] - m[3][0] * m[1][1];
  float Coef23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];

  vec4 Fac0(Coef00, Coef00, Coef02, Coef03);
  vec4 Fac1(Coef04, Coef04, Coef06, Coef07);
  vec4 Fac2(Coef08, Coef08, Coef10, Coef11);
  vec4 Fac3(Coef12, Coef12, Coef14, Coef15);
  vec4 Fac4(Coef16, Coef16, Coef18, Coef19);
  vec4 Fac5(Coef20, Coef20, Coef22, Coef23);

  vec4 Vec0(m[1][0], m[0][0], m[0][0], m[0][0]);
  vec4 Vec1(m[1][1], m[0][1], m[0][1], m[0][1]);
  vec4 Vec2(m[1][2], m[0][2], m[0][2], m[0][2]);
  vec4 Vec3(m[1][3], m[0][3], m[0][3], m[0][3]);

  vec4 Inv0(Vec1 * Fac0 - Vec2 * Fac1 + Vec3 * Fac2);
  vec4 Inv1(Vec0 * Fac0 - Vec2 * Fac3 + Vec3 * Fac4);
  vec4 Inv2(Vec0 * Fac1 - Vec1 * Fac3 + Vec3 * Fac5);
  vec4 Inv3(Vec0 * Fac2 - Vec1 * Fac4 + Vec2 * Fac5);

  vec4 SignA(+1, -1, +1, -1);
  vec4 SignB(-1, +1, -1, +1);
  mat4 Inverse(Inv0 * SignA, Inv1 * SignB, Inv2 * SignA, Inv3 * SignB);

  vec4 Row0(Inverse[0][0], Inverse[1][0], Inverse[2][0], Inverse[3][0]);

  vec4 Dot0(m[0] * Row0);
  float Dot1 = (Dot0.x + Dot0.y) + (Dot0.z + Dot0.w);

  float OneOverDeterminant = 1.f / Dot1;

  return Inverse * OneOverDeterminant;
}

template <typename T>
inline T dot(const tvec3<T>& a, const tvec3<T>& b) {
  return a.x * b.x + a.y * b.y + a.z * b.z;
}
template <typename T>
inline T dot(const tvec4<T>& a, const tvec4<T>& b) {
  // glm compute_dot<vec4>: tmp = a*b; (tmp.x + tmp.y) + (tmp.z + tmp.w)
  tvec4<T> tmp = a * b;
  return (tmp.x + tmp.y) + (tmp.z + tmp.w);
}
template <typename T>
inline tvec3<T> normalize(const tvec3<T>& v) {
  return v * (T(1) / std::sqrt(dot(v, v)));
}
template <typename T>
inline tvec4<T> normalize(const tvec4<T>& v) {
  return v * (T(1) / std::sqrt(dot(v, v)));
}

// glm/ext/matrix_transform.inl — rotate(m, angle, v)
inline mat4 rotate(const mat4& m, float angle, const vec3& v) {
  float const a = angle;
  float const c = std::cos(a);
  float const s = std::sin(a);

  vec3 axis(normalize(v));
  vec3 temp((1.f - c) * axis);

  mat4 Rotate(0.f);
  Rotate[0][0] = c + temp[0] * axis[0];
  Rotate[0][1] = temp[0] * axis[1] + s * axis[2];
  Rotate[0][2] = temp[0] * axis[2] - s * axis[1];

  Rotate[1][0] = temp[1] * axis[0] - s * axis[2];
  Rotate[1][1] = c + temp[1] * axis[1];
  Rotate[1][2] = temp[1] * axis[2] + s * axis[0];

  Rotate[2][0] = temp[2] * axis[0] + s * axis[1];
  Rotate[2][1] = temp[2] * axis[1] - s * axis[0];
  Rotate[2][2] = c + temp[2] * axis[2];

  mat4 Result(0.f);
  Result[0] = m[0] * Rotate[0][0] + m[1] * Rotate[0][1] + m[2] * Rotate[0][2];
  Result[1] = m[0] * Rotate[1][0] + m[1] * Rotate[1][1] + m[2] * Rotate[1][2];
  Result[2] = m[0] * Rotate[2][0] + m[1] * Rotate[2][1] + m[2] * Rotate[2][2];
  Result[3] = m[3];
  return Result;
}

template <typename T>
constexpr T epsilon() { return std::numeric_limits<T>::epsilon(); }
template <typename T>
constexpr T pi() { return static_cast<T>(3.14159265358979323846264338327950288); }

// glm/gtx/matrix_query.inl — isIdentity
inline bool isIdentity(const mat4& m, float eps) {
  bool result = true;
  for (int i = 0; result && i < 4; ++i) {
    for (int j = 0; result && j < i; ++j) result = std::abs(m[i][j]) <= eps;
    if (result) result = std::abs(m[i][i] - 1.f) <= eps;
    for (int j = i + 1; result && j < 4; ++j) result = std::abs(m[i][j]) <= eps;
  }
  return result;
}

// camera.cc only (never reached by the raster path): right-handed, [-1,1] depth.
inline mat4 lookAt(const vec3& eye, const vec3& center, const vec3& up) {
  auto sub = [](const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); };
  auto cross = [](const vec3& a, const vec3& b) {
    return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
  };
  vec3 const f(normalize(sub(center, eye)));
  vec3 const s(normalize(cross(f, up)));
  vec3 const u(cross(s, f));
  mat4 R(1.f);
  R[0][0] = s.x; R[1][0] = s.y; R[2][0] = s.z;
  R[0][1] = u.x; R[1][1] = u.y; R[2][1] = u.z;
  R[0][2] = -f.x; R[1][2] = -f.y; R[2][2] = -f.z;
  R[3][0] = -dot(s, eye); R[3][1] = -dot(u, eye); R[3][2] = dot(f, eye);
  return R;
}
inline mat4 perspective(float fovy, float aspect, float zNear, float zFar) {
  float const tanHalfFovy = std::tan(fovy / 2.f);
  mat4 R(0.f);
  R[0][0] = 1.f / (aspect * tanHalfFovy);
  R[1][1] = 1.f / (tanHalfFovy);
  R[2][2] = -(zFar + zNear) / (zFar - zNear);
  R[2][3] = -1.f;
  R[3][2] = -(2.f * zFar * zNear) / (zFar - zNear);
  return R;
}

// --- scalar aliases of <cmath> -------------------------------------------
template <typename T>
inline T clamp(T x, T lo, T hi) { return std::min(std::max(x, lo), hi); }
template <typename T>
inline T floor(T x) { return std::floor(x); }
template <typename T>
inline T ceil(T x) { return std::ceil(x); }
template <typename T>
inline T fract(T x) { return x - std::floor(x); }
template <typename T>
inline T mod(T x, T y) { return x - y * std::floor(x / y); }
template <typename T>
inline T abs(T x) { return std::abs(x); }
template <typename T>
inline T sqrt(T x) { return std::sqrt(x); }
template <typename T>
inline T sin(T x) { return std::sin(x); }
template <typename T>
inline T cos(T x) { return std::cos(x); }
template <typename T>
inline T min(T a, T b) { return (b < a) ? b : a; }
template <typename T>
inline T max(T a, T b) { return (a < b) ? b : a; }
template <typename T>
inline bool isinf(T x) { return std::isinf(x); }
template <typename T>
inline bool isnan(T x) { return std::isnan(x); }
template <typename T>
constexpr T radians(T deg) { return deg * static_cast<T>(0.01745329251994329576923690768489); }
template <typename T>
constexpr T degrees(T rad) { return rad * static_cast<T>(57.295779513082320876798154814105); }

}  // namespace glm
