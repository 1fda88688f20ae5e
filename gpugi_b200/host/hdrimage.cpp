// hdrimage.cpp -- PFM I/O in the reference's exact byte layout (hdrimage.cpp:6-60): header
// "PF\n<w> <h>\n-1.000000\n", then w*h RGB float triples, row 0 first, little endian.
#include "hdrimage.hpp"
#include <cstring>
#include <fstream>

bool WritePfm(const ei::Vec4* _data, const ei::IVec2& _size, const std::string& _filename, std::uint32_t _divisor)
{
	std::ofstream file(_filename.c_str(), std::ios::binary);
	if (!file) return false;
	file << "PF\n" << _size.x << " " << _size.y << "\n-1.000000\n";
	const size_t count = size_t(_size.x) * size_t(_size.y);
	for (size_t i = 0; i < count; ++i)
	{
		float rgb[3] = {_data[i].x / _divisor, _data[i].y / _divisor, _data[i].z / _divisor};
		file.write(reinterpret_cast<const char*>(rgb), sizeof(rgb));
	}
	return bool(file);
}

std::unique_ptr<ei::Vec3[]> LoadPfm(const std::string& _filename, ei::IVec2& _size)
{
	std::ifstream file(_filename.c_str(), std::ios::binary);
	if (!file) return nullptr;
	char magic[4] = {0, 0, 0, 0};
	file.read(magic, 3);
	if (std::strcmp(magic, "PF\n") != 0) return nullptr;
	file >> _size.x >> _size.y;
	if (!file || _size.x <= 0 || _size.y <= 0) return nullptr;
	file.ignore(11); // "\n-1.000000\n"
	const size_t count = size_t(_size.x) * size_t(_size.y);
	std::unique_ptr<ei::Vec3[]> out(new ei::Vec3[count]);
	file.read(reinterpret_cast<char*>(out.get()), std::streamsize(count * sizeof(ei::Vec3)));
	if (!file) return nullptr;
	return out;
}
