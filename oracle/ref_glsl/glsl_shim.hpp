// glsl_shim.hpp -- TEST INFRASTRUCTURE.  The subset of GLSL 4.50 the reference's shaders use, as C++, so that the reference's OWN shader
// source (gpugi/shader/*.glsl, *.comp under /root/reference, read where it lies by oracle/ref_glsl/glsl2cpp.py and piped into g++ -- no copy is
// kept) runs on the CPU and pins the oracle's restatement of it (tests/test_ref_glsl.py).  Nothing in gpugi_b200/ includes, links or loads this.
//
// Arithmetic definition (the one DESIGN.md section 2 states for the oracle and the CUDA path): every GLSL float operation is one IEEE binary32
// operation, no contraction (compile with -ffp-contract=off); min / max drop a NaN operand (fminf / fmaxf, what NVIDIA's FMNMX does);
// dot(a, b) = (a.x b.x + a.y b.y) + a.z b.z; cross as written in the GLSL specification 8.5; length = sqrt(dot); normalize(v) = v * (1 / sqrt(dot(v, v)));
// clamp(x, a, b) = min(max(x, a), b); mix(x, y, a) = x (1 - a) + y a; transcendental functions are libm's float versions.
#pragma once

#include <math.h>   // the C++ wrapper: float overloads of sqrt / pow / sin / ... in the global namespace, as GLSL spells them
#include <cmath>
#include <stdlib.h>
#include <cstdint>
#include <cstring>

typedef unsigned int uint;

struct vec2; struct vec3; struct vec4;

// ---- swizzle proxies: live inside the vector's union, read / write the named components in place
template <int N, int A, int B> struct swz2
{
	float d[N];
	inline operator vec2() const;
	inline swz2& operator=(const vec2& v);
};
template <int N, int A, int B, int C> struct swz3
{
	float d[N];
	inline operator vec3() const;
	inline swz3& operator=(const vec3& v);
};

struct glsl_ordered2 { float a, b; }; // braced initialisation: evaluated left to right
struct vec2
{
	union { struct { float x, y; }; struct { float r, g; }; swz2<2, 0, 0> xx; swz2<2, 0, 1> xy; swz2<2, 1, 0> yx; swz2<2, 1, 1> yy; };
	vec2() : x(0.0f), y(0.0f) {}
	explicit vec2(float s) : x(s), y(s) {}
	vec2(float a, float b) : x(a), y(b) {}
	explicit vec2(const glsl_ordered2& o) : x(o.a), y(o.b) {}
	vec2(const vec2& o) : x(o.x), y(o.y) {}
	vec2& operator=(const vec2& o) { x = o.x; y = o.y; return *this; }
};
struct vec3
{
	union
	{
		struct { float x, y, z; }; struct { float r, g, b; };
		swz2<3, 0, 0> xx; swz2<3, 0, 1> xy; swz2<3, 0, 2> xz; swz2<3, 1, 2> yz; swz2<3, 1, 1> yy; swz2<3, 2, 2> zz;
		swz3<3, 0, 1, 2> xyz; swz3<3, 0, 1, 2> rgb;
	};
	vec3() : x(0.0f), y(0.0f), z(0.0f) {}
	explicit vec3(float s) : x(s), y(s), z(s) {}
	vec3(float a, float b, float c) : x(a), y(b), z(c) {}
	vec3(const vec2& a, float c) : x(a.x), y(a.y), z(c) {}
	vec3(float a, const vec2& b) : x(a), y(b.x), z(b.y) {}
	vec3(const vec3& o) : x(o.x), y(o.y), z(o.z) {}
	vec3& operator=(const vec3& o) { x = o.x; y = o.y; z = o.z; return *this; }
};
struct vec4
{
	union
	{
		struct { float x, y, z, w; }; struct { float r, g, b, a; };
		swz2<4, 0, 0> xx; swz2<4, 0, 1> xy; swz2<4, 0, 2> xz; swz2<4, 1, 2> yz; swz2<4, 2, 3> zw;
		swz3<4, 0, 1, 2> xyz; swz3<4, 0, 1, 2> rgb;
	};
	vec4() : x(0.0f), y(0.0f), z(0.0f), w(0.0f) {}
	explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
	vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
	vec4(const vec3& a, float d) : x(a.x), y(a.y), z(a.z), w(d) {}
	vec4(const vec2& a, float c, float d) : x(a.x), y(a.y), z(c), w(d) {}
	vec4(const vec2& a, const vec2& b) : x(a.x), y(a.y), z(b.x), w(b.y) {}
	vec4(const vec4& o) : x(o.x), y(o.y), z(o.z), w(o.w) {}
	vec4& operator=(const vec4& o) { x = o.x; y = o.y; z = o.z; w = o.w; return *this; }
};
template <int N, int A, int B> inline swz2<N, A, B>::operator vec2() const { return vec2(d[A], d[B]); }
template <int N, int A, int B> inline swz2<N, A, B>& swz2<N, A, B>::operator=(const vec2& v) { d[A] = v.x; d[B] = v.y; return *this; }
template <int N, int A, int B, int C> inline swz3<N, A, B, C>::operator vec3() const { return vec3(d[A], d[B], d[C]); }
template <int N, int A, int B, int C> inline swz3<N, A, B, C>& swz3<N, A, B, C>::operator=(const vec3& v) { d[A] = v.x; d[B] = v.y; d[C] = v.z; return *this; }

struct ivec2
{
	int x, y;
	ivec2() : x(0), y(0) {}
	ivec2(int a, int b) : x(a), y(b) {}
	explicit ivec2(int s) : x(s), y(s) {}
	explicit ivec2(const vec2& v) : x(int(v.x)), y(int(v.y)) {}
	operator vec2() const { return vec2(float(x), float(y)); } // GLSL's implicit ivec2 -> vec2
};
struct uvec2
{
	uint x, y;
	uvec2() : x(0u), y(0u) {}
	uvec2(uint a, uint b) : x(a), y(b) {}
};
inline ivec2 to_ivec2(const uvec2& v) { return ivec2(int(v.x), int(v.y)); }
struct ivec4
{
	int x, y, z, w;
	ivec4() : x(0), y(0), z(0), w(0) {}
	ivec4(int a, int b, int c, int d) : x(a), y(b), z(c), w(d) {}
};
struct bvec3 { bool x, y, z; };
struct mat4 { float m[16]; };

inline ivec2 operator+(ivec2 a, ivec2 b) { return ivec2(a.x + b.x, a.y + b.y); }
inline ivec2 operator-(ivec2 a, ivec2 b) { return ivec2(a.x - b.x, a.y - b.y); }
inline ivec2 operator*(ivec2 a, int s) { return ivec2(a.x * s, a.y * s); }
inline bool operator==(ivec2 a, ivec2 b) { return a.x == b.x && a.y == b.y; }

// ---- component-wise arithmetic (one IEEE operation per component and operator)
inline vec2 operator+(const vec2& a, const vec2& b) { return vec2(a.x + b.x, a.y + b.y); }
inline vec2 operator+(const vec2& a, float s) { return vec2(a.x + s, a.y + s); }
inline vec2 operator+(float s, const vec2& a) { return vec2(s + a.x, s + a.y); }
inline vec2& operator+=(vec2& a, const vec2& b) { a = a + b; return a; }
inline vec2& operator+=(vec2& a, float s) { a = a + s; return a; }
inline vec2 operator-(const vec2& a, const vec2& b) { return vec2(a.x - b.x, a.y - b.y); }
inline vec2 operator-(const vec2& a, float s) { return vec2(a.x - s, a.y - s); }
inline vec2 operator-(float s, const vec2& a) { return vec2(s - a.x, s - a.y); }
inline vec2& operator-=(vec2& a, const vec2& b) { a = a - b; return a; }
inline vec2& operator-=(vec2& a, float s) { a = a - s; return a; }
inline vec2 operator*(const vec2& a, const vec2& b) { return vec2(a.x * b.x, a.y * b.y); }
inline vec2 operator*(const vec2& a, float s) { return vec2(a.x * s, a.y * s); }
inline vec2 operator*(float s, const vec2& a) { return vec2(s * a.x, s * a.y); }
inline vec2& operator*=(vec2& a, const vec2& b) { a = a * b; return a; }
inline vec2& operator*=(vec2& a, float s) { a = a * s; return a; }
inline vec2 operator/(const vec2& a, const vec2& b) { return vec2(a.x / b.x, a.y / b.y); }
inline vec2 operator/(const vec2& a, float s) { return vec2(a.x / s, a.y / s); }
inline vec2 operator/(float s, const vec2& a) { return vec2(s / a.x, s / a.y); }
inline vec2& operator/=(vec2& a, const vec2& b) { a = a / b; return a; }
inline vec2& operator/=(vec2& a, float s) { a = a / s; return a; }
inline vec2 operator-(const vec2& a) { return vec2(-a.x, -a.y); }
inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator+(const vec3& a, float s) { return vec3(a.x + s, a.y + s, a.z + s); }
inline vec3 operator+(float s, const vec3& a) { return vec3(s + a.x, s + a.y, s + a.z); }
inline vec3& operator+=(vec3& a, const vec3& b) { a = a + b; return a; }
inline vec3& operator+=(vec3& a, float s) { a = a + s; return a; }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator-(const vec3& a, float s) { return vec3(a.x - s, a.y - s, a.z - s); }
inline vec3 operator-(float s, const vec3& a) { return vec3(s - a.x, s - a.y, s - a.z); }
inline vec3& operator-=(vec3& a, const vec3& b) { a = a - b; return a; }
inline vec3& operator-=(vec3& a, float s) { a = a - s; return a; }
inline vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3& operator*=(vec3& a, const vec3& b) { a = a * b; return a; }
inline vec3& operator*=(vec3& a, float s) { a = a * s; return a; }
inline vec3 operator/(const vec3& a, const vec3& b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator/(float s, const vec3& a) { return vec3(s / a.x, s / a.y, s / a.z); }
inline vec3& operator/=(vec3& a, const vec3& b) { a = a / b; return a; }
inline vec3& operator/=(vec3& a, float s) { a = a / s; return a; }
inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
inline vec4 operator+(const vec4& a, const vec4& b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline vec4 operator+(const vec4& a, float s) { return vec4(a.x + s, a.y + s, a.z + s, a.w + s); }
inline vec4 operator+(float s, const vec4& a) { return vec4(s + a.x, s + a.y, s + a.z, s + a.w); }
inline vec4& operator+=(vec4& a, const vec4& b) { a = a + b; return a; }
inline vec4& operator+=(vec4& a, float s) { a = a + s; return a; }
inline vec4 operator-(const vec4& a, const vec4& b) { return vec4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
inline vec4 operator-(const vec4& a, float s) { return vec4(a.x - s, a.y - s, a.z - s, a.w - s); }
inline vec4 operator-(float s, const vec4& a) { return vec4(s - a.x, s - a.y, s - a.z, s - a.w); }
inline vec4& operator-=(vec4& a, const vec4& b) { a = a - b; return a; }
inline vec4& operator-=(vec4& a, float s) { a = a - s; return a; }
inline vec4 operator*(const vec4& a, const vec4& b) { return vec4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
inline vec4 operator*(const vec4& a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline vec4 operator*(float s, const vec4& a) { return vec4(s * a.x, s * a.y, s * a.z, s * a.w); }
inline vec4& operator*=(vec4& a, const vec4& b) { a = a * b; return a; }
inline vec4& operator*=(vec4& a, float s) { a = a * s; return a; }
inline vec4 operator/(const vec4& a, const vec4& b) { return vec4(a.x / b.x, a.y / b.y, a.z / b.z, a.w / b.w); }
inline vec4 operator/(const vec4& a, float s) { return vec4(a.x / s, a.y / s, a.z / s, a.w / s); }
inline vec4 operator/(float s, const vec4& a) { return vec4(s / a.x, s / a.y, s / a.z, s / a.w); }
inline vec4& operator/=(vec4& a, const vec4& b) { a = a / b; return a; }
inline vec4& operator/=(vec4& a, float s) { a = a / s; return a; }
inline vec4 operator-(const vec4& a) { return vec4(-a.x, -a.y, -a.z, -a.w); }

// ---- scalar builtins
inline float min(float a, float b) { return fminf(a, b); }
inline float max(float a, float b) { return fmaxf(a, b); }
inline int min(int a, int b) { return b < a ? b : a; }
inline int max(int a, int b) { return a < b ? b : a; }
inline uint min(uint a, uint b) { return b < a ? b : a; }
inline uint max(uint a, uint b) { return a < b ? b : a; }
inline float clamp(float x, float a, float b) { return min(max(x, a), b); }
inline int clamp(int x, int a, int b) { return min(max(x, a), b); }
inline float sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
inline float step(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
inline float mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
inline float fract(float x) { return x - floorf(x); }
inline float inversesqrt(float x) { return 1.0f / sqrtf(x); }
inline float atan(float y, float x) { return atan2f(y, x); }
inline uint floatBitsToUint(float f) { uint u; std::memcpy(&u, &f, 4); return u; }
inline int floatBitsToInt(float f) { int i; std::memcpy(&i, &f, 4); return i; }
inline float uintBitsToFloat(uint u) { float f; std::memcpy(&f, &u, 4); return f; }
inline float intBitsToFloat(int i) { float f; std::memcpy(&f, &i, 4); return f; }

// ---- vector builtins
#define GLSL_MAP1(F) \
	inline vec2 F(const vec2& a) { return vec2(F(a.x), F(a.y)); } \
	inline vec3 F(const vec3& a) { return vec3(F(a.x), F(a.y), F(a.z)); } \
	inline vec4 F(const vec4& a) { return vec4(F(a.x), F(a.y), F(a.z), F(a.w)); }
inline float glsl_abs(float x) { return fabsf(x); }
inline vec2 abs(const vec2& a) { return vec2(fabsf(a.x), fabsf(a.y)); }
inline vec3 abs(const vec3& a) { return vec3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
inline vec4 abs(const vec4& a) { return vec4(fabsf(a.x), fabsf(a.y), fabsf(a.z), fabsf(a.w)); }
GLSL_MAP1(sign)
inline vec3 sqrt(const vec3& a) { return vec3(sqrtf(a.x), sqrtf(a.y), sqrtf(a.z)); }
inline vec3 ceil(const vec3& a) { return vec3(ceilf(a.x), ceilf(a.y), ceilf(a.z)); }
inline vec3 log2(const vec3& a) { return vec3(log2f(a.x), log2f(a.y), log2f(a.z)); }
inline vec2 floor(const vec2& a) { return vec2(floorf(a.x), floorf(a.y)); }
inline vec2 round(const vec2& a) { return vec2(roundf(a.x), roundf(a.y)); }
#undef GLSL_MAP1
#define GLSL_MAP2(F) \
	inline vec2 F(const vec2& a, const vec2& b) { return vec2(F(a.x, b.x), F(a.y, b.y)); } \
	inline vec3 F(const vec3& a, const vec3& b) { return vec3(F(a.x, b.x), F(a.y, b.y), F(a.z, b.z)); } \
	inline vec4 F(const vec4& a, const vec4& b) { return vec4(F(a.x, b.x), F(a.y, b.y), F(a.z, b.z), F(a.w, b.w)); } \
	inline vec2 F(const vec2& a, float b) { return vec2(F(a.x, b), F(a.y, b)); } \
	inline vec3 F(const vec3& a, float b) { return vec3(F(a.x, b), F(a.y, b), F(a.z, b)); } \
	inline vec4 F(const vec4& a, float b) { return vec4(F(a.x, b), F(a.y, b), F(a.z, b), F(a.w, b)); }
GLSL_MAP2(min)
GLSL_MAP2(max)
#undef GLSL_MAP2
inline vec2 clamp(const vec2& x, const vec2& a, const vec2& b) { return min(max(x, a), b); }
inline vec3 clamp(const vec3& x, const vec3& a, const vec3& b) { return min(max(x, a), b); }
inline vec4 clamp(const vec4& x, const vec4& a, const vec4& b) { return min(max(x, a), b); }
inline vec2 clamp(const vec2& x, float a, float b) { return min(max(x, a), b); }
inline vec3 clamp(const vec3& x, float a, float b) { return min(max(x, a), b); }
inline vec4 clamp(const vec4& x, float a, float b) { return min(max(x, a), b); }
inline vec3 mix(const vec3& x, const vec3& y, float a) { return x * (1.0f - a) + y * a; }
inline vec3 pow(const vec3& a, const vec3& b) { return vec3(powf(a.x, b.x), powf(a.y, b.y), powf(a.z, b.z)); }
inline vec3 exp(const vec3& a) { return vec3(expf(a.x), expf(a.y), expf(a.z)); }

inline float dot(const vec2& a, const vec2& b) { return a.x * b.x + a.y * b.y; }
inline float dot(const vec3& a, const vec3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float dot(const vec4& a, const vec4& b) { return ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w; }
inline vec3 cross(const vec3& a, const vec3& b) { return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
inline float length(const vec2& a) { return sqrtf(dot(a, a)); }
inline float length(const vec3& a) { return sqrtf(dot(a, a)); }
inline vec3 normalize(const vec3& a) { const float inv = 1.0f / sqrtf(dot(a, a)); return a * inv; }
inline vec2 normalize(const vec2& a) { const float inv = 1.0f / sqrtf(dot(a, a)); return a * inv; }
inline bvec3 lessThan(const vec3& a, const vec3& b) { bvec3 r = {a.x < b.x, a.y < b.y, a.z < b.z}; return r; }
inline bvec3 greaterThanEqual(const vec3& a, const vec3& b) { bvec3 r = {a.x >= b.x, a.y >= b.y, a.z >= b.z}; return r; }
inline bool all(const bvec3& b) { return b.x && b.y && b.z; }
inline bool any(const bvec3& b) { return b.x || b.y || b.z; }

inline uint packSnorm2x16(const vec2& v)
{
	const int a = int(roundf(clamp(v.x, -1.0f, 1.0f) * 32767.0f)), b = int(roundf(clamp(v.y, -1.0f, 1.0f) * 32767.0f));
	return (uint(a) & 0xffffu) | (uint(b) << 16);
}
inline vec2 unpackSnorm2x16(uint p)
{
	return vec2(clamp(float(int16_t(p & 0xffffu)) / 32767.0f, -1.0f, 1.0f), clamp(float(int16_t(p >> 16)) / 32767.0f, -1.0f, 1.0f));
}
inline uint packUnorm4x8(const vec4& v)
{
	uint r = 0;
	const float c[4] = {v.x, v.y, v.z, v.w};
	for (int i = 0; i < 4; ++i) r |= uint(roundf(clamp(c[i], 0.0f, 1.0f) * 255.0f)) << (8 * i);
	return r;
}
inline vec4 unpackUnorm4x8(uint p) { return vec4(float(p & 255u) / 255.0f, float((p >> 8) & 255u) / 255.0f, float((p >> 16) & 255u) / 255.0f, float(p >> 24) / 255.0f); }

// ---- resources
struct samplerBuffer   // GL_RGBA32F / GL_RGB32F texture buffer
{
	const float* data = nullptr; int comps = 4;
};
struct isamplerBuffer  // GL_RGBA32I
{
	const int* data = nullptr;
};
inline vec4 texelFetch(const samplerBuffer& b, int i)
{
	const float* p = b.data + size_t(b.comps) * size_t(i);
	return b.comps == 4 ? vec4(p[0], p[1], p[2], p[3]) : vec4(p[0], p[1], p[2], 1.0f);
}
inline ivec4 texelFetch(const isamplerBuffer& b, int i) { const int* p = b.data + 4 * size_t(i); return ivec4(p[0], p[1], p[2], p[3]); }

// bindless 2D textures: the handle's low word indexes this table (filled by the driver from GiSceneDesc::textures, texels already decoded to
// float RGBA).  Sampler state of scene.cpp:16-21 (mag LINEAR, REPEAT, level 0): GL 4.5 section 8.14.2, texel centres at (i + 0.5) / size.
struct GlslTexture { int w = 0, h = 0; const float* rgba = nullptr; };
extern GlslTexture* g_glslTextures;
struct sampler2D
{
	uint index;
	explicit sampler2D(const uvec2& handle) : index(handle.x) {}
};
inline int glsl_wrap(int i, int n) { const int m = i % n; return m < 0 ? m + n : m; }
inline vec4 glsl_texel(const GlslTexture& t, int x, int y) { const float* p = t.rgba + 4 * (size_t(y) * t.w + x); return vec4(p[0], p[1], p[2], p[3]); }
inline vec4 textureLod(const sampler2D& s, const vec2& uv, float)
{
	const GlslTexture& t = g_glslTextures[s.index];
	if (t.w == 1 && t.h == 1) return glsl_texel(t, 0, 0);
	const float fu = uv.x * float(t.w) - 0.5f, fv = uv.y * float(t.h) - 0.5f;
	const float iu = floorf(fu), iv = floorf(fv);
	const float a = fu - iu, b = fv - iv;
	const int i0 = glsl_wrap(int(iu), t.w), i1 = glsl_wrap(int(iu) + 1, t.w), j0 = glsl_wrap(int(iv), t.h), j1 = glsl_wrap(int(iv) + 1, t.h);
	const vec4 t00 = glsl_texel(t, i0, j0), t10 = glsl_texel(t, i1, j0), t01 = glsl_texel(t, i0, j1), t11 = glsl_texel(t, i1, j1);
	const float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
	return vec4(t00.x * w00 + t10.x * w10 + t01.x * w01 + t11.x * w11, t00.y * w00 + t10.y * w10 + t01.y * w01 + t11.y * w11,
	            t00.z * w00 + t10.z * w10 + t01.z * w01 + t11.z * w11, t00.w * w00 + t10.w * w10 + t01.w * w01 + t11.w * w11);
}
// environment cube map: the pinned scenes have none bound -- RendererSystem binds a black one (pathtracer.cpp:41-45 is only called by "envMap")
struct samplerCubeArray { vec4 constant; };
inline vec4 textureLod(const samplerCubeArray& s, const vec4&, float) { return s.constant; }

// RGBA32F image (the backbuffer): the driver points it at the caller's accumulator
struct image2D { float* rgba = nullptr; int w = 0, h = 0; };
// (OpenGL 4.5 section 8.26: an access outside the image loads zero / is discarded)
inline vec4 imageLoad(const image2D& im, const ivec2& p) { if (p.x < 0 || p.y < 0 || p.x >= im.w || p.y >= im.h) return vec4(0.0f); const float* q = im.rgba + 4 * (size_t(p.y) * im.w + p.x); return vec4(q[0], q[1], q[2], q[3]); }
inline void imageStore(const image2D& im, const ivec2& p, const vec4& v) { if (p.x < 0 || p.y < 0 || p.x >= im.w || p.y >= im.h) return; float* q = im.rgba + 4 * (size_t(p.y) * im.w + p.x); q[0] = v.x; q[1] = v.y; q[2] = v.z; q[3] = v.w; }

// the same memory seen as an R32F image four times as wide (screenreproject.glsl:7,43-46: WriteAtomic adds one channel at a time with
// GL_NV_shader_atomic_float); returns the old value like the GLSL built-in
inline float imageAtomicAdd(const image2D& im, const ivec2& p, float v)
{
	if (p.x < 0 || p.y < 0 || p.x >= 4 * im.w || p.y >= im.h) return 0.0f;
	float* q = im.rgba + size_t(p.y) * size_t(im.w) * 4 + size_t(p.x);
	float old;
#pragma omp atomic capture
	{ old = *q; *q += v; }
	return old;
}
struct uimage2D { uint* data = nullptr; int w = 0, h = 0; };   // LockTexture: declared by the shaders, only used in commented-out code
inline void memoryBarrierImage() {}
// shader storage counters (lightcache.glsl:27-37)
inline int atomicAdd(int& mem, int v)
{
	int old;
#pragma omp atomic capture
	{ old = mem; mem += v; }
	return old;
}

// per-invocation built-in (glsl2cpp.py rewrites gl_GlobalInvocationID.xy)
extern thread_local uvec2 gl_GlobalInvocationID_xy;
extern thread_local uint gl_GlobalInvocationID_x;
