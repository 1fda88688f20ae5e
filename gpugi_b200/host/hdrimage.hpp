// hdrimage.hpp -- PFM read/write of the reference (gpugi/imageut/hdrimage.hpp, hdrimage.cpp:6-60).
#pragma once
#include "vec.hpp"
#include <cstdint>
#include <memory>
#include <string>

/// Writes RGB little-endian PFM; every pixel is divided by _divisor (the iteration count).
bool WritePfm(const ei::Vec4* _data, const ei::IVec2& _size, const std::string& _filename, std::uint32_t _divisor = 1);
/// Returns nullptr if the file is missing or is not a "PF" image.
std::unique_ptr<ei::Vec3[]> LoadPfm(const std::string& _filename, ei::IVec2& _size);
