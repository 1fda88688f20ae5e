// imageio.hpp -- texture file decoding for the host layer.  The reference loads every material texture and sky face through
// stb_image: `stbi_load(file, &w, &h, &n, 4)` -> 8-bit RGBA, first row = top row of the picture (dependencies/glhelper/glhelper/texture2d.cpp:31-52,
// called from gpugi/scene/scene.cpp:196 and renderersystem.cpp:173-186).  This is an independent decoder with the same contract for the formats
// scene files use: PNG (all colour types and bit depths, Adam7), JPEG (baseline / extended sequential Huffman, any sampling), TGA (true colour,
// grey, colour-mapped; raw or RLE), BMP (1-32 bpp uncompressed / bit fields), binary PNM, and the raw "GITX" container of the synthetic scenes.
#pragma once

#include <string>
#include <vector>

namespace gi {

/// Decodes `file` into width * height RGBA8 texels, row 0 = top.  Alpha is 255 where the file has none.  false + *error on failure
/// (missing file, unknown or unsupported format -- progressive JPEG, RLE-compressed BMP, 16-bit PNM -- or a corrupt stream).
bool LoadImageRGBA8(const std::string& file, int& width, int& height, std::vector<unsigned char>& rgba, std::string* error = nullptr);

/// The same from memory.
bool DecodeImageRGBA8(const unsigned char* data, size_t size, int& width, int& height, std::vector<unsigned char>& rgba, std::string* error = nullptr);

} // namespace gi
