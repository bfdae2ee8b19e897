// oracle/oracle_vec.h -- TEST INFRASTRUCTURE (CPU oracle), NOT PRODUCT CODE.
//
// Scalar restatement of the float3 helpers the reference's device programs use, with the
// reference's exact operation order (every function cites the line it follows).  Must be
// compiled with -ffp-contract=off so that no multiply-add is fused behind our back.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace ora {

struct V3 {
    float x, y, z;
};

static inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
static inline V3 v3(float s) { return V3{s, s, s}; }

// sutil/vec_math.h float3 operators: all component-wise, no reassociation.
static inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
static inline V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
static inline V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
// sutil/vec_math.h:479-481 component-wise divide
static inline V3 operator/(V3 a, V3 b) { return v3(a.x / b.x, a.y / b.y, a.z / b.z); }
// sutil/vec_math.h:483-487: float3 / float multiplies by the reciprocal
static inline V3 operator/(V3 a, float s)
{
    float inv = 1.0f / s;
    return a * inv;
}
static inline V3& operator+=(V3& a, V3 b) { a = a + b; return a; }
static inline V3& operator*=(V3& a, V3 b) { a = a * b; return a; }

// sutil/vec_math.h:527-530
static inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// sutil/vec_math.h:533-536
static inline V3 cross(V3 a, V3 b)
{
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// sutil/vec_math.h:539-542
static inline float length(V3 v) { return sqrtf(dot(v, v)); }
// sutil/vec_math.h:545-549
static inline V3 normalize(V3 v)
{
    float invLen = 1.0f / sqrtf(dot(v, v));
    return v * invLen;
}
// sutil/vec_math.h:558-561  i - 2.0f * n * dot(n,i)  == i - ((2.0f*n) * dot(n,i))
static inline V3 reflect(V3 i, V3 n) { return i - (2.0f * n) * dot(n, i); }
// sutil/vec_math.h:567-570
static inline V3 faceforward(V3 n, V3 i, V3 nref) { return n * copysignf(1.0f, dot(i, nref)); }
// sutil/vec_math.h:500-503
static inline V3 lerp(V3 a, V3 b, float t) { return a + t * (b - a); }
// sutil/vec_math.h:114-117
static inline float clampf(float f, float a, float b) { return fmaxf(a, fminf(f, b)); }
static inline V3 clamp(V3 v, float a, float b) { return v3(clampf(v.x, a, b), clampf(v.y, a, b), clampf(v.z, a, b)); }

static inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

}  // namespace ora
