// Minimal stand-in for the three GLM headers the reference's DOGM core includes (glm/vec4.hpp, glm/mat4x4.hpp,
// glm/glm.hpp).  GLM itself is not vendored by the reference (README.md:42-46: `apt install libglm-dev`) and is
// absent from this image.  HARNESS CODE for oracle/_ref only: it lets the reference's own .cu files compile
// unmodified.  Semantics follow GLM 0.9.9: trivially-copyable PODs, column-major mat4x4, and
// mat4x4 * vec4 evaluated as (m0*v0 + m1*v1) + (m2*v2 + m3*v3)  (glm/detail/type_mat4x4.inl, operator*).
#pragma once
// GLM's own setup header pulls these in; the reference relies on that (dogm_types.h uses assert/malloc)
#include <cassert>
#include <cmath>
#include <cstdlib>
#ifdef __CUDACC__
#define GLM_SHIM_FN __host__ __device__ inline
#else
#define GLM_SHIM_FN inline
#endif

namespace glm
{
struct vec2 // used by the demo's simulator and evaluator only (host code)
{
    float x, y;
    vec2() = default;
    template <typename A, typename B>
    GLM_SHIM_FN vec2(A a, B b) : x(static_cast<float>(a)), y(static_cast<float>(b))
    {
    }
    GLM_SHIM_FN float& operator[](int i) { return (&x)[i]; }
    GLM_SHIM_FN const float& operator[](int i) const { return (&x)[i]; }
    GLM_SHIM_FN vec2& operator+=(const vec2& o)
    {
        x += o.x;
        y += o.y;
        return *this;
    }
    GLM_SHIM_FN vec2& operator-=(const vec2& o)
    {
        x -= o.x;
        y -= o.y;
        return *this;
    }
};
GLM_SHIM_FN vec2 operator+(const vec2& a, const vec2& b) { return vec2(a.x + b.x, a.y + b.y); }
GLM_SHIM_FN vec2 operator-(const vec2& a, const vec2& b) { return vec2(a.x - b.x, a.y - b.y); }
GLM_SHIM_FN vec2 operator*(const vec2& a, float s) { return vec2(a.x * s, a.y * s); }
GLM_SHIM_FN vec2 operator*(float s, const vec2& a) { return vec2(s * a.x, s * a.y); }
GLM_SHIM_FN bool operator==(const vec2& a, const vec2& b) { return a.x == b.x && a.y == b.y; }

struct vec4
{
    float x, y, z, w;
    vec4() = default;
    GLM_SHIM_FN vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    GLM_SHIM_FN float& operator[](int i) { return (&x)[i]; }
    GLM_SHIM_FN const float& operator[](int i) const { return (&x)[i]; }
};

GLM_SHIM_FN vec4 operator+(const vec4& a, const vec4& b)
{
    return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
GLM_SHIM_FN vec4 operator*(const vec4& a, const vec4& b)
{
    return vec4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
}
GLM_SHIM_FN vec4 operator*(float s, const vec4& a)
{
    return vec4(s * a.x, s * a.y, s * a.z, s * a.w);
}
GLM_SHIM_FN bool operator==(const vec4& a, const vec4& b)
{
    return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w;
}
} // namespace glm
