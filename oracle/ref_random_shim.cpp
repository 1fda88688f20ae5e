// TEST INFRASTRUCTURE ONLY.  C-linkage doors onto the REFERENCE'S OWN utilities/random.cpp, which oracle/Makefile compiles from
// /root/reference/gpugi/utilities/random.cpp where it lies (no reference source is copied into this repository) into
// oracle/_ref/libref_random.so.  Only tests/ and tests/golden/make_golden_rng.py load it: it pins the oracle's and the host layer's
// WangHash / xorshift32 / xorshift64* / XorshiftF against the real reference.
#include "random.hpp" // the reference's header, found through -I/root/reference/gpugi/utilities
#include <cstdint>

extern "C" __attribute__((visibility("default"))) std::uint32_t ref_wanghash(std::uint32_t x) { return WangHash(x); }
extern "C" __attribute__((visibility("default"))) std::uint32_t ref_xorshift32(std::uint32_t x) { return Xorshift(x); }
extern "C" __attribute__((visibility("default"))) std::uint64_t ref_xorshift64(std::uint64_t x) { return Xorshift(x); }
extern "C" __attribute__((visibility("default"))) float ref_xorshift_f(std::uint64_t* state) { return XorshiftF(*state); }
extern "C" __attribute__((visibility("default"))) double ref_xorshift_d(std::uint64_t* state) { return XorshiftD(*state); }
extern "C" __attribute__((visibility("default"))) float ref_xorshift_range(std::uint64_t* state, float lo, float hi) { return Xorshift(*state, lo, hi); }
