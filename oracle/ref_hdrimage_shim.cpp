// TEST INFRASTRUCTURE ONLY.  C-linkage doors onto the REFERENCE'S OWN imageut/hdrimage.cpp (WritePfm / LoadPfm), compiled where it lies
// against the stand-in headers of oracle/ref_shims (see oracle/Makefile).  Pins the PFM read-out format of SURVEY §8 a-18.
#include "hdrimage.hpp" // the reference's header, found through -I/root/reference/gpugi/imageut
#include <cstdint>

extern "C" __attribute__((visibility("default"))) int ref_write_pfm(const char* path, const float* rgba, int w, int h, std::uint32_t divisor)
{
	return WritePfm(reinterpret_cast<const ei::Vec4*>(rgba), ei::IVec2{w, h}, path, divisor) ? 0 : 1;
}

extern "C" __attribute__((visibility("default"))) int ref_load_pfm(const char* path, float* rgb, int capacity_pixels, int* w, int* h)
{
	ei::IVec2 size{0, 0};
	std::unique_ptr<ei::Vec3[]> img = LoadPfm(path, size);
	if (!img) return 1;
	*w = size.x; *h = size.y;
	if (size.x * size.y > capacity_pixels) return 2;
	std::memcpy(rgb, img.get(), sizeof(float) * 3 * size_t(size.x) * size_t(size.y));
	return 0;
}
