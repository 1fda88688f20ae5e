// vec.hpp -- the ~100 lines of float3 math the host side needs (stands in for dependencies/epsilon).
// Every operation is a single IEEE binary32 op; compile with -ffp-contract=off.
#pragma once
#include <cmath>

namespace ei {

const float PI = 3.14159265358979323846f;

struct Vec2 { float x, y; Vec2() : x(0), y(0) {} Vec2(float x_, float y_) : x(x_), y(y_) {} };
struct IVec2 { int x, y; IVec2() : x(0), y(0) {} IVec2(int x_, int y_) : x(x_), y(y_) {} };
struct Vec4 { float x, y, z, w; Vec4() : x(0), y(0), z(0), w(0) {} Vec4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {} };

struct Vec3
{
	float x, y, z;
	Vec3() : x(0), y(0), z(0) {}
	explicit Vec3(float s) : x(s), y(s), z(s) {}
	Vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
	float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline Vec3 operator+(const Vec3& a, const Vec3& b) { return Vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vec3 operator-(const Vec3& a, const Vec3& b) { return Vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vec3 operator*(const Vec3& a, float s) { return Vec3(a.x * s, a.y * s, a.z * s); }
inline Vec3 operator*(float s, const Vec3& a) { return Vec3(s * a.x, s * a.y, s * a.z); }
inline Vec3 operator/(const Vec3& a, float s) { return Vec3(a.x / s, a.y / s, a.z / s); }
inline Vec3& operator+=(Vec3& a, const Vec3& b) { a = a + b; return a; }
inline Vec3& operator*=(Vec3& a, float s) { a = a * s; return a; }
inline bool operator==(const Vec3& a, const Vec3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline bool operator!=(const Vec3& a, const Vec3& b) { return !(a == b); }
inline float dot(const Vec3& a, const Vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross(const Vec3& a, const Vec3& b) { return Vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float len(const Vec3& a) { return std::sqrt(dot(a, a)); }
inline Vec3 normalize(const Vec3& a) { return a / len(a); }

struct Triangle { Vec3 v0, v1, v2; };
struct Box { Vec3 min, max; };
// ei::surface(Triangle), epsilon 3dtypes.inl:412-417
inline float surface(const Triangle& t) { return len(cross(t.v1 - t.v0, t.v2 - t.v0)) * 0.5f; }

enum class Types3D { BOX, OBOX };

} // namespace ei
