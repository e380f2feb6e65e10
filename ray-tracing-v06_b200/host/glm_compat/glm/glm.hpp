// glm_compat/glm/glm.hpp — a minimal stand-in for the parts of glm the reference's host-side API
// uses in its signatures (glm::vec2/3/4 and a handful of free functions).  It exists so that
// the host mirror compiles where the real glm (vendored in the reference under
// Libraries/include/glm) is not on the include path; put the real glm first on the include
// path and this header is never seen.  Written from scratch; host arithmetic only.
#pragma once
#include <cmath>

namespace glm {

struct vec2 {
	float x, y;
	constexpr vec2() : x(0), y(0) {}
	constexpr explicit vec2(float s) : x(s), y(s) {}
	constexpr vec2(float a, float b) : x(a), y(b) {}
	float& operator[](int i) { return (&x)[i]; }
	const float& operator[](int i) const { return (&x)[i]; }
};

struct vec3 {
	float x, y, z;
	constexpr vec3() : x(0), y(0), z(0) {}
	constexpr explicit vec3(float s) : x(s), y(s), z(s) {}
	template <typename A, typename B, typename C>
	constexpr vec3(A a, B b, C c) : x(static_cast<float>(a)), y(static_cast<float>(b)), z(static_cast<float>(c)) {}
	float& operator[](int i) { return (&x)[i]; }
	const float& operator[](int i) const { return (&x)[i]; }
	vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
	vec3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
};

struct vec4 {
	float x, y, z, w;
	constexpr vec4() : x(0), y(0), z(0), w(0) {}
	constexpr explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
	constexpr vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
	constexpr vec4(const vec3& v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
	float& operator[](int i) { return (&x)[i]; }
	const float& operator[](int i) const { return (&x)[i]; }
};

inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, const vec3& a) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline bool operator==(const vec3& a, const vec3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

inline float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(const vec3& a, const vec3& b) { return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
inline float length(const vec3& a) { return std::sqrt(dot(a, a)); }
inline vec3 normalize(const vec3& a) { return a * (1.0f / std::sqrt(dot(a, a))); }
inline vec3 min(const vec3& a, const vec3& b) { return vec3(b.x < a.x ? b.x : a.x, b.y < a.y ? b.y : a.y, b.z < a.z ? b.z : a.z); }
inline vec3 max(const vec3& a, const vec3& b) { return vec3(a.x < b.x ? b.x : a.x, a.y < b.y ? b.y : a.y, a.z < b.z ? b.z : a.z); }
inline vec3 abs(const vec3& a) { return vec3(std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)); }
inline float radians(float deg) { return deg * 0.01745329251994329576923690768489f; }
inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline vec3 mix(const vec3& a, const vec3& b, float t) { return a * (1.0f - t) + b * t; }

}  // namespace glm
