// Minimal stand-in for the g-truc/glm header that the reference rasterizer
// includes (`#include <glm/glm.hpp>`; reference .gitmodules pins glm at commit
// 6f14f4792a0cde5d0cf2c910506724d61cb95834 but the submodule directory is EMPTY
// in /root/reference, so the reference does not compile as shipped).
//
// TEST INFRASTRUCTURE ONLY.  This file exists so that oracle/build_ref.py can
// compile the *unmodified* reference .cu files into oracle/_ref/ for GPU-vs-GPU
// parity checks and for the `bench.py --impl reference` arm.  Nothing in the
// product (reduced-3dgs_b200/) includes it.
//
// Only the symbols the reference uses are provided (forward.cu, backward.cu,
// rasterizer_impl.cu, reduced_3dgs.cu, reduced_3dgs/*.cu): vec1, vec3, vec4, bvec3,
// mat3 / mat4 (column-major, m[c][r]), dot, length, max, sign, pow, all,
// lessThanEqual, greaterThanEqual, transpose and the arithmetic operators.  The
// summation orders follow upstream GLM's generic (non-SIMD) code paths:
//   dot(a,b)      = (a.x*b.x + a.y*b.y) + a.z*b.z
//   (A*B)[c][r]   = (A[0][r]*B[c][0] + A[1][r]*B[c][1]) + A[2][r]*B[c][2]
//   M4*v          = (M[0]*v.x + M[1]*v.y) + (M[2]*v.z + M[3]*v.w)
//   v*M3          = (dot(M[0],v), dot(M[1],v), dot(M[2],v))   (row vector times matrix)
//   length(v)     = sqrt(dot(v,v))
//   sign(x)       = (0 < x) - (x < 0)
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define GLM_SHIM_FN __host__ __device__ inline
#else
#define GLM_SHIM_FN inline
#endif

namespace glm {

struct vec1 {
	float x;
	GLM_SHIM_FN vec1() : x(0) {}
	GLM_SHIM_FN explicit vec1(float s) : x(s) {}
};
struct vec4;

struct vec3 {
	float x, y, z;
	GLM_SHIM_FN vec3() : x(0), y(0), z(0) {}
	GLM_SHIM_FN explicit vec3(float s) : x(s), y(s), z(s) {}
	GLM_SHIM_FN explicit vec3(int s) : x(float(s)), y(float(s)), z(float(s)) {}
	GLM_SHIM_FN explicit vec3(const vec4& v);                    // drops w
	template <typename A, typename B, typename C>
	GLM_SHIM_FN vec3(A a, B b, C c) : x(float(a)), y(float(b)), z(float(c)) {}
	GLM_SHIM_FN float& operator[](int i) { return (&x)[i]; }
	GLM_SHIM_FN const float& operator[](int i) const { return (&x)[i]; }
	GLM_SHIM_FN vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
	GLM_SHIM_FN vec3& operator+=(float s) { x += s; y += s; z += s; return *this; }
	GLM_SHIM_FN vec3& operator-=(const vec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
	GLM_SHIM_FN vec3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
	GLM_SHIM_FN vec3& operator/=(float s) { x /= s; y /= s; z /= s; return *this; }
};

struct vec4 {
	float x, y, z, w;
	GLM_SHIM_FN vec4() : x(0), y(0), z(0), w(0) {}
	GLM_SHIM_FN explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
	template <typename A, typename B, typename C, typename D>
	GLM_SHIM_FN vec4(A a, B b, C c, D d) : x(float(a)), y(float(b)), z(float(c)), w(float(d)) {}
	GLM_SHIM_FN vec4(const vec3& v, const vec1& s) : x(v.x), y(v.y), z(v.z), w(s.x) {}
	GLM_SHIM_FN vec4(const vec3& v, float s) : x(v.x), y(v.y), z(v.z), w(s) {}
	GLM_SHIM_FN float& operator[](int i) { return (&x)[i]; }
	GLM_SHIM_FN const float& operator[](int i) const { return (&x)[i]; }
};

GLM_SHIM_FN vec3::vec3(const vec4& v) : x(v.x), y(v.y), z(v.z) {}
GLM_SHIM_FN vec4 operator+(const vec4& a, const vec4& b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
GLM_SHIM_FN vec4 operator*(const vec4& a, const vec4& b) { return vec4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
GLM_SHIM_FN vec4 operator*(const vec4& a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }

GLM_SHIM_FN vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
GLM_SHIM_FN vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
GLM_SHIM_FN vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
GLM_SHIM_FN vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
GLM_SHIM_FN vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
GLM_SHIM_FN vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
GLM_SHIM_FN vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
GLM_SHIM_FN vec3 operator+(const vec3& a, float s) { return vec3(a.x + s, a.y + s, a.z + s); }
GLM_SHIM_FN vec3 operator/(const vec3& a, const vec3& b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
GLM_SHIM_FN vec3 pow(const vec3& a, const vec3& e) { return vec3(powf(a.x, e.x), powf(a.y, e.y), powf(a.z, e.z)); }

struct bvec3 {
	bool x, y, z;
	GLM_SHIM_FN bvec3(bool a, bool b, bool c) : x(a), y(b), z(c) {}
};
GLM_SHIM_FN bvec3 lessThanEqual(const vec3& a, const vec3& b) { return bvec3(a.x <= b.x, a.y <= b.y, a.z <= b.z); }
GLM_SHIM_FN bvec3 greaterThanEqual(const vec3& a, const vec3& b) { return bvec3(a.x >= b.x, a.y >= b.y, a.z >= b.z); }
GLM_SHIM_FN bool all(const bvec3& v) { return v.x && v.y && v.z; }

GLM_SHIM_FN float dot(const vec3& a, const vec3& b)
{
	vec3 tmp(a * b);
	return tmp.x + tmp.y + tmp.z;
}
GLM_SHIM_FN float length(const vec3& v) { return sqrtf(dot(v, v)); }
GLM_SHIM_FN vec3 max(const vec3& a, float s)
{
	return vec3(a.x < s ? s : a.x, a.y < s ? s : a.y, a.z < s ? s : a.z);
}
GLM_SHIM_FN float sign1(float v) { return float(0.0f < v) - float(v < 0.0f); }
GLM_SHIM_FN vec3 sign(const vec3& v) { return vec3(sign1(v.x), sign1(v.y), sign1(v.z)); }

// Column-major 3x3: c[k] is column k, m[c][r].
struct mat3 {
	vec3 c[3];
	GLM_SHIM_FN mat3() { c[0] = vec3(1, 0, 0); c[1] = vec3(0, 1, 0); c[2] = vec3(0, 0, 1); }
	GLM_SHIM_FN explicit mat3(float s) { c[0] = vec3(s, 0, 0); c[1] = vec3(0, s, 0); c[2] = vec3(0, 0, s); }
	template <typename X0, typename Y0, typename Z0, typename X1, typename Y1, typename Z1,
		typename X2, typename Y2, typename Z2>
	GLM_SHIM_FN mat3(X0 x0, Y0 y0, Z0 z0, X1 x1, Y1 y1, Z1 z1, X2 x2, Y2 y2, Z2 z2)
	{
		c[0] = vec3(x0, y0, z0); c[1] = vec3(x1, y1, z1); c[2] = vec3(x2, y2, z2);
	}
	GLM_SHIM_FN vec3& operator[](int i) { return c[i]; }
	GLM_SHIM_FN const vec3& operator[](int i) const { return c[i]; }
};

GLM_SHIM_FN mat3 operator*(const mat3& m1, const mat3& m2)
{
	const float A00 = m1[0][0], A01 = m1[0][1], A02 = m1[0][2];
	const float A10 = m1[1][0], A11 = m1[1][1], A12 = m1[1][2];
	const float A20 = m1[2][0], A21 = m1[2][1], A22 = m1[2][2];
	const float B00 = m2[0][0], B01 = m2[0][1], B02 = m2[0][2];
	const float B10 = m2[1][0], B11 = m2[1][1], B12 = m2[1][2];
	const float B20 = m2[2][0], B21 = m2[2][1], B22 = m2[2][2];
	mat3 r(0.0f);
	r[0][0] = A00 * B00 + A10 * B01 + A20 * B02;
	r[0][1] = A01 * B00 + A11 * B01 + A21 * B02;
	r[0][2] = A02 * B00 + A12 * B01 + A22 * B02;
	r[1][0] = A00 * B10 + A10 * B11 + A20 * B12;
	r[1][1] = A01 * B10 + A11 * B11 + A21 * B12;
	r[1][2] = A02 * B10 + A12 * B11 + A22 * B12;
	r[2][0] = A00 * B20 + A10 * B21 + A20 * B22;
	r[2][1] = A01 * B20 + A11 * B21 + A21 * B22;
	r[2][2] = A02 * B20 + A12 * B21 + A22 * B22;
	return r;
}
GLM_SHIM_FN mat3 operator*(float s, const mat3& m)
{
	mat3 r(0.0f);
	r[0] = m[0] * s; r[1] = m[1] * s; r[2] = m[2] * s;
	return r;
}
GLM_SHIM_FN mat3 operator*(const mat3& m, float s) { return s * m; }
// row vector times matrix (type_mat3x3.inl: operator*(column_type const& v, mat const& m))
GLM_SHIM_FN vec3 operator*(const vec3& v, const mat3& m)
{
	return vec3(m[0][0] * v.x + m[0][1] * v.y + m[0][2] * v.z,
		m[1][0] * v.x + m[1][1] * v.y + m[1][2] * v.z,
		m[2][0] * v.x + m[2][1] * v.y + m[2][2] * v.z);
}

// Column-major 4x4, only what redundancy_score.cu needs: M * v.
struct mat4 {
	vec4 c[4];
	GLM_SHIM_FN vec4& operator[](int i) { return c[i]; }
	GLM_SHIM_FN const vec4& operator[](int i) const { return c[i]; }
};
GLM_SHIM_FN vec4 operator*(const mat4& m, const vec4& v)
{
	const vec4 Mul0 = m[0] * vec4(v.x), Mul1 = m[1] * vec4(v.y);
	const vec4 Add0 = Mul0 + Mul1;
	const vec4 Mul2 = m[2] * vec4(v.z), Mul3 = m[3] * vec4(v.w);
	const vec4 Add1 = Mul2 + Mul3;
	return Add0 + Add1;
}

GLM_SHIM_FN mat3 transpose(const mat3& m)
{
	return mat3(m[0][0], m[1][0], m[2][0],
		m[0][1], m[1][1], m[2][1],
		m[0][2], m[1][2], m[2][2]);
}

} // namespace glm
