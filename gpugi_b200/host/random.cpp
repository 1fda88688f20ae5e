// random.cpp -- integer-exact restatement of gpugi/utilities/random.cpp:5-49.
// WangHash feeds PerIteration.FrameSeed (renderersystem.cpp:101); XorshiftF drives the light sampler.
#include "random.hpp"

std::uint32_t WangHash(std::uint32_t s)
{
	s = (s ^ 61u) ^ (s >> 16);
	s *= 9u;
	s ^= s >> 4;
	s *= 0x27d4eb2du;
	s ^= s >> 15;
	return s;
}

std::uint32_t Xorshift(std::uint32_t s)
{
	s ^= s << 13;
	s ^= s >> 17;
	s ^= s << 5;
	return s;
}

// xorshift64* (Vigna), the multiplier of random.cpp:30
std::uint64_t Xorshift(std::uint64_t s)
{
	s ^= s >> 12;
	s ^= s << 25;
	s ^= s >> 27;
	return s * UINT64_C(2685821657736338717);
}

double XorshiftD(std::uint64_t& s)
{
	s = Xorshift(s);
	return static_cast<double>(s % 900719925474099) / 900719925474099.0;
}

float XorshiftF(std::uint64_t& s)
{
	s = Xorshift(s);
	return static_cast<float>(s % 8388593) / 8388593.0f;
}

float Xorshift(std::uint64_t& s, float min, float max)
{
	s = Xorshift(s);
	return (static_cast<float>(s % 8388593) / 8388593.0f) * (max - min) + min;
}
