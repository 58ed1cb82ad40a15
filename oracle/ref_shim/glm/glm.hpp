/*
 * Minimal stand-in for the subset of glm 0.9.9.8 (external/CMakeLists.txt:22-26, not vendored in the
 * reference tree, no network here) that the reference's geometry.h / material.h / bvh.h /
 * bvh_builder.{h,cpp} use. Written for oracle/ref_layout.cpp, which compiles those reference
 * files UNMODIFIED from /root/reference to check struct layouts and constructors against the
 * C ABI. Test infrastructure only; float32, every operation separately rounded.
 */
#pragma once
#include <cmath>

namespace glm
{
struct vec3
{
    float x, y, z;
    vec3() : x(0), y(0), z(0) {}
    explicit vec3(float s) : x(s), y(s), z(s) {}
    vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    template <typename V4, typename = decltype(V4().w)>
    explicit vec3(const V4& v) : x(v.x), y(v.y), z(v.z) {}
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
};
struct vec4
{
    float x, y, z, w;
    vec4() : x(0), y(0), z(0), w(0) {}
    vec4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
    template <typename T>
    vec4(T x_, int y_, int z_, int w_) : x((float)x_), y((float)y_), z((float)z_), w((float)w_) {}
    vec4(const vec3& v, float w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
    vec4(const vec3& v, int w_) : x(v.x), y(v.y), z(v.z), w((float)w_) {}
};
inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, const vec3& a) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 min(const vec3& a, const vec3& b) { return vec3(std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z)); }
inline vec3 max(const vec3& a, const vec3& b) { return vec3(std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z)); }
inline vec3 cross(const vec3& a, const vec3& b)
{
    return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
inline float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 normalize(const vec3& v) { return v * (1.0f / std::sqrt(dot(v, v))); }
} /* namespace glm */
