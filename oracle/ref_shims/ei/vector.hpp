// TEST INFRASTRUCTURE ONLY.  Stand-in for <ei/vector.hpp> so that the REFERENCE'S OWN gpugi/imageut/hdrimage.cpp compiles with g++
// (oracle/Makefile, target _ref/libref_hdrimage.so): the real epsilon header is MSVC-only.  Only what that file
// touches: plain aggregates with epsilon's member names and memory layout (x, y, z[, w] packed floats / ints).
#pragma once
#include <cstring>
#include <string>

namespace ei {
struct Vec3 { float x, y, z; Vec3() = default; Vec3(float _x, float _y, float _z) : x(_x), y(_y), z(_z) {} };
struct Vec4 { float x, y, z, w; };
struct IVec2 { int x, y; };
}
