// TEST INFRASTRUCTURE ONLY.  Minimal stand-in for the subset of GLM that
// actorshq/dataset/native/ray_sampler.cu and actorshq/toolbox/native/occupancy_grid_generation.cu use
// (vec2/vec3/vec4/mat3/mat4, min, max, normalize), so the
// UNMODIFIED reference source can be compiled into oracle/_ref on a box without libglm-dev.
// Written from the GLM API documentation; semantics follow GLM's generic (non-SIMD) definitions:
// column-major mat3, normalize(v) = v * inversesqrt(dot(v,v)), inversesqrt(x) = 1/sqrt(x),
// min(a,b) = (b<a)?b:a, max(a,b) = (a<b)?b:a.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define GLM_FUNC __host__ __device__ inline

namespace glm {

struct vec2 {
  float x, y;
  vec2() = default;
  GLM_FUNC vec2(float a, float b) : x(a), y(b) {}
  GLM_FUNC float& operator[](int i) { return (&x)[i]; }
  GLM_FUNC const float& operator[](int i) const { return (&x)[i]; }
};

struct vec3 {
  float x, y, z;
  vec3() = default;
  GLM_FUNC vec3(float a, float b, float c) : x(a), y(b), z(c) {}
  GLM_FUNC explicit vec3(float s) : x(s), y(s), z(s) {}
  GLM_FUNC vec3(int a, int b, int c) : x(float(a)), y(float(b)), z(float(c)) {}
  GLM_FUNC float& operator[](int i) { return (&x)[i]; }
  GLM_FUNC const float& operator[](int i) const { return (&x)[i]; }
};

GLM_FUNC vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
GLM_FUNC vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
GLM_FUNC vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
GLM_FUNC vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
GLM_FUNC vec3 operator*(float s, const vec3& a) { return vec3(a.x * s, a.y * s, a.z * s); }
GLM_FUNC vec3 operator+(const vec3& a, float s) { return vec3(a.x + s, a.y + s, a.z + s); }
GLM_FUNC vec3 operator-(const vec3& a, float s) { return vec3(a.x - s, a.y - s, a.z - s); }
GLM_FUNC vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
GLM_FUNC vec3 operator/(float s, const vec3& a) { return vec3(s / a.x, s / a.y, s / a.z); }

struct mat3 {
  vec3 c[3];  // columns
  mat3() = default;
  GLM_FUNC vec3& operator[](int i) { return c[i]; }
  GLM_FUNC const vec3& operator[](int i) const { return c[i]; }
};
struct vec4 {
  float x, y, z, w;
  vec4() = default;
  GLM_FUNC vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
  GLM_FUNC vec4(const vec3& v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
  GLM_FUNC float& operator[](int i) { return (&x)[i]; }
  GLM_FUNC const float& operator[](int i) const { return (&x)[i]; }
};
GLM_FUNC vec4 operator+(const vec4& a, const vec4& b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
GLM_FUNC vec4 operator*(const vec4& a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }
struct mat4 {
  vec4 c[4];  // columns
  mat4() = default;
  GLM_FUNC vec4& operator[](int i) { return c[i]; }
  GLM_FUNC const vec4& operator[](int i) const { return c[i]; }
};
// GLM's generic mat4 * vec4 associates as (m0*x + m1*y) + (m2*z + m3*w)
GLM_FUNC vec4 operator*(const mat4& m, const vec4& v) { return (m[0] * v.x + m[1] * v.y) + (m[2] * v.z + m[3] * v.w); }

GLM_FUNC vec3 operator*(const mat3& m, const vec3& v) {
  return vec3(m[0][0] * v.x + m[1][0] * v.y + m[2][0] * v.z, m[0][1] * v.x + m[1][1] * v.y + m[2][1] * v.z,
              m[0][2] * v.x + m[1][2] * v.y + m[2][2] * v.z);
}

GLM_FUNC float min(float a, float b) { return (b < a) ? b : a; }
GLM_FUNC float max(float a, float b) { return (a < b) ? b : a; }
GLM_FUNC vec3 min(const vec3& a, const vec3& b) { return vec3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
GLM_FUNC vec3 max(const vec3& a, const vec3& b) { return vec3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
GLM_FUNC float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
GLM_FUNC float inversesqrt(float x) { return 1.0f / sqrtf(x); }
GLM_FUNC vec3 normalize(const vec3& v) { return v * inversesqrt(dot(v, v)); }

}  // namespace glm
