// TEST SCAFFOLDING -- not product code.
//
// Minimal stand-in for g-truc/glm, used ONLY to compile the reference rasterizer
// (third_party/gaussian-splatting/submodules/diff-gaussian-rasterization) into
// oracle/_ref/ so it can serve as the GPU-side parity oracle and speed baseline.
// The reference lists glm as an un-vendored git submodule (DGR/.gitmodules:1-3; no
// pinned commit) and it is not installed in this image.  The reference uses exactly
// these symbols: vec3, vec4, mat3, dot, length, max, transpose, mat3*mat3,
// float*mat3 and the usual vector/scalar operators (call sites: forward.cu:21-153,
// backward.cu:20-396).  Semantics follow glm's published behaviour: column-major
// storage, mat3(a..i) fills column 0, then 1, then 2, m[c] is a column, m[c][r] an
// element, (A*B)[j][i] = A[0][i]*B[j][0] + A[1][i]*B[j][1] + A[2][i]*B[j][2]
// evaluated left to right, dot(a,b) = (a.x*b.x + a.y*b.y) + a.z*b.z.
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define GLMS_FN __host__ __device__ inline
#else
#define GLMS_FN inline
#endif

namespace glm {

struct vec3 {
	float x, y, z;
	GLMS_FN vec3() : x(0.f), y(0.f), z(0.f) {}
	template <typename A, typename B, typename C>
	GLMS_FN vec3(A a, B b, C c) : x(float(a)), y(float(b)), z(float(c)) {}
	GLMS_FN explicit vec3(float s) : x(s), y(s), z(s) {}
	GLMS_FN float& operator[](int i) { return (&x)[i]; }
	GLMS_FN const float& operator[](int i) const { return (&x)[i]; }
	GLMS_FN vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
	GLMS_FN vec3& operator+=(float s) { x += s; y += s; z += s; return *this; }
	GLMS_FN vec3& operator-=(const vec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
	GLMS_FN vec3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
};

struct vec4 {
	float x, y, z, w;
	GLMS_FN vec4() : x(0.f), y(0.f), z(0.f), w(0.f) {}
	template <typename A, typename B, typename C, typename D>
	GLMS_FN vec4(A a, B b, C c, D d) : x(float(a)), y(float(b)), z(float(c)), w(float(d)) {}
	GLMS_FN float& operator[](int i) { return (&x)[i]; }
	GLMS_FN const float& operator[](int i) const { return (&x)[i]; }
};

GLMS_FN vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
GLMS_FN vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
GLMS_FN vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
GLMS_FN vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
GLMS_FN vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
GLMS_FN vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
GLMS_FN vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }

GLMS_FN float dot(const vec3& a, const vec3& b)
{
	vec3 t = a * b;
	return t.x + t.y + t.z;
}
GLMS_FN float length(const vec3& a) { return sqrt(dot(a, a)); }
GLMS_FN vec3 max(const vec3& a, float s)
{
	return vec3(a.x < s ? s : a.x, a.y < s ? s : a.y, a.z < s ? s : a.z);
}

struct mat3 {
	vec3 c[3];
	GLMS_FN mat3() {}
	GLMS_FN explicit mat3(float d)
	{
		c[0] = vec3(d, 0.f, 0.f);
		c[1] = vec3(0.f, d, 0.f);
		c[2] = vec3(0.f, 0.f, d);
	}
	template <typename T0, typename T1, typename T2, typename T3, typename T4, typename T5, typename T6, typename T7, typename T8>
	GLMS_FN mat3(T0 x0, T1 y0, T2 z0, T3 x1, T4 y1, T5 z1, T6 x2, T7 y2, T8 z2)
	{
		c[0] = vec3(x0, y0, z0);
		c[1] = vec3(x1, y1, z1);
		c[2] = vec3(x2, y2, z2);
	}
	GLMS_FN vec3& operator[](int i) { return c[i]; }
	GLMS_FN const vec3& operator[](int i) const { return c[i]; }
};

GLMS_FN mat3 transpose(const mat3& m)
{
	return mat3(
		m[0][0], m[1][0], m[2][0],
		m[0][1], m[1][1], m[2][1],
		m[0][2], m[1][2], m[2][2]);
}

GLMS_FN mat3 operator*(const mat3& a, const mat3& b)
{
	mat3 r;
	for (int j = 0; j < 3; j++)
		for (int i = 0; i < 3; i++)
			r[j][i] = a[0][i] * b[j][0] + a[1][i] * b[j][1] + a[2][i] * b[j][2];
	return r;
}

GLMS_FN mat3 operator*(float s, const mat3& m)
{
	mat3 r;
	r[0] = s * m[0];
	r[1] = s * m[1];
	r[2] = s * m[2];
	return r;
}
GLMS_FN mat3 operator*(const mat3& m, float s) { return s * m; }

} // namespace glm
