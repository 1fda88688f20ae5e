// Stand-in for dependencies/chinoise/include/cn/chinoise.hpp.  TEST INFRASTRUCTURE ONLY.
// The real umbrella header also includes fieldnoise.hpp, whose .inl calls `_x.subcol<0,N-1>()` on a dependent type without the
// `template` keyword (MSVC accepts, GCC cannot parse).  bim_hierarchy.cpp only uses cn::HaltonRng and cn::barycentric (SGGX fitting,
// off the hot path), which live in the two headers below -- included from the reference tree, unmodified.
#pragma once
#include <cn/rnd.hpp>
#include <cn/sampler.hpp>
