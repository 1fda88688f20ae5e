// Force-included (-include) in front of every reference translation unit of oracle/_ref/libref_bim.so.  TEST INFRASTRUCTURE ONLY.
// The reference's dependencies/bim + dependencies/epsilon are written for MSVC; three of its habits are resolved from OUTSIDE the sources:
//   1. headers MSVC pulls in transitively (std::string in material.hpp, memcpy in bim_kdtree.cpp, uint8_t in bim_tangentspace.cpp);
//   2. `swap(member, std::vector<T>(n))` binds an rvalue to std::swap's non-const reference (bim_io.cpp:215,232, bim_chunk.cpp:250-252,
//      bim_hierarchy.cpp:223) -- an overload taking the temporary by rvalue reference does what MSVC's extension does;
//   3. (Makefile) ei/vector.hpp:183-198 guards subcol()/subrow() with a default template argument that is not value-dependent
//      (`sizeof(FROM)`), which MSVC evaluates lazily and GCC at class instantiation: the build pipes that one header through sed
//      (`sizeof(FROM)` -> `(FROM <= TO)`, same truth value) into a temporary include directory and deletes it afterwards.
//   4. bim/log.hpp:24-43 relies on MSVC's one-phase lookup: its generic buildMessageString calls std::to_string on string arguments because
//      the string overloads are declared after it; two pass-through std::to_string overloads give that call a meaning;
//   5. (Makefile) -std=c++14: with C++17's string_view, `std::string = json` (bim_io.cpp:459) is ambiguous in the vendored nlohmann json.
// No arithmetic, no algorithm and no file-format code of the reference is replaced.
#pragma once
#include <cstdint>
#include <cstring>
#include <limits>
#include <string>
#include <utility>
#include <vector>

namespace std {
template <class T, class A>
inline void swap(vector<T, A>& a, vector<T, A>&& b) { a.swap(b); }
inline string to_string(const char* s) { return string(s); }
inline string to_string(const string& s) { return s; }
}
