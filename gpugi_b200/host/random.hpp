// random.hpp -- host RNG of the reference (gpugi/utilities/random.hpp:1-14).
#pragma once
#include <cstdint>

std::uint32_t WangHash(std::uint32_t seed);
std::uint32_t Xorshift(std::uint32_t rndState);
std::uint64_t Xorshift(std::uint64_t rndState);
double XorshiftD(std::uint64_t& rndState);
float XorshiftF(std::uint64_t& rndState);
float Xorshift(std::uint64_t& rndState, float min, float max);
