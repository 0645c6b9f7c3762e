// see vec4.hpp: GLM stand-in, harness code for oracle/_ref only
#pragma once
#include "vec4.hpp"

namespace glm
{
struct mat4x4
{
    vec4 value[4]; // columns
    mat4x4() = default;
    GLM_SHIM_FN mat4x4(float x0, float y0, float z0, float w0, float x1, float y1, float z1, float w1, float x2, float y2,
                       float z2, float w2, float x3, float y3, float z3, float w3)
    {
        value[0] = vec4(x0, y0, z0, w0);
        value[1] = vec4(x1, y1, z1, w1);
        value[2] = vec4(x2, y2, z2, w2);
        value[3] = vec4(x3, y3, z3, w3);
    }
    GLM_SHIM_FN vec4& operator[](int i) { return value[i]; }
    GLM_SHIM_FN const vec4& operator[](int i) const { return value[i]; }
};

GLM_SHIM_FN vec4 operator*(const mat4x4& m, const vec4& v)
{
    const vec4 Mov0(v[0], v[0], v[0], v[0]);
    const vec4 Mov1(v[1], v[1], v[1], v[1]);
    const vec4 Mul0 = m[0] * Mov0;
    const vec4 Mul1 = m[1] * Mov1;
    const vec4 Add0 = Mul0 + Mul1;
    const vec4 Mov2(v[2], v[2], v[2], v[2]);
    const vec4 Mov3(v[3], v[3], v[3], v[3]);
    const vec4 Mul2 = m[2] * Mov2;
    const vec4 Mul3 = m[3] * Mov3;
    const vec4 Add1 = Mul2 + Mul3;
    return Add0 + Add1;
}

GLM_SHIM_FN mat4x4 transpose(const mat4x4& m)
{
    return mat4x4(m[0][0], m[1][0], m[2][0], m[3][0], m[0][1], m[1][1], m[2][1], m[3][1], m[0][2], m[1][2], m[2][2],
                  m[3][2], m[0][3], m[1][3], m[2][3], m[3][3]);
}
} // namespace glm
