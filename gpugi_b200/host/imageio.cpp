// imageio.cpp -- see imageio.hpp.  Written from the format specifications (PNG: RFC 2083 / W3C PNG; JPEG: ITU-T T.81 + JFIF; TGA: Truevision 2.0;
// BMP: the Windows DIB headers), not from stb_image.
#include "imageio.hpp"

#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <new>

namespace gi {
namespace {

typedef unsigned char u8;

struct Fail
{
	std::string* error;
	bool operator()(const std::string& text) const { if (error) *error = text; return false; }
};

struct Reader
{
	const u8* p; size_t n, pos;
	bool has(size_t k) const { return pos + k <= n; }
	uint32_t u8at(size_t i) const { return i < n ? p[i] : 0u; }
	uint32_t get8() { return pos < n ? p[pos++] : (pos++, 0u); }
	uint32_t le16() { uint32_t a = get8(); return a | (get8() << 8); }
	uint32_t le32() { uint32_t a = le16(); return a | (le16() << 16); }
	uint32_t be16() { uint32_t a = get8(); return (a << 8) | get8(); }
	uint32_t be32() { uint32_t a = be16(); return (a << 16) | be16(); }
};

// stb_image's own limits (STBI_MAX_DIMENSIONS, stbi__mad3sizes_valid): either side <= 2^24 and width * height * 4 must fit an int;
// a corrupt header cannot ask for more than 2 GB
bool sizeOk(long long w, long long h)
{
	return w > 0 && h > 0 && w <= (1 << 24) && h <= (1 << 24) && w * h * 4 <= 0x7fffffffLL;
}

// ------------------------------------------------------------------------------------------------ PNG
int paeth(int a, int b, int c)
{
	const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
	return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// one (sub)image of w x h pixels: scanlines with a leading filter byte -> unfiltered bytes, `stride` bytes per row
bool pngUnfilter(const u8* in, size_t inSize, int w, int h, int bitsPerPixel, std::vector<u8>& out, size_t& consumed)
{
	const size_t stride = (size_t(w) * bitsPerPixel + 7) / 8;
	const int bpp = std::max(1, bitsPerPixel / 8);
	if (inSize < (stride + 1) * size_t(h)) return false;
	out.assign(stride * size_t(h), 0);
	for (int y = 0; y < h; ++y)
	{
		const u8* src = in + (stride + 1) * size_t(y);
		u8* cur = out.data() + stride * size_t(y);
		const u8* up = y ? cur - stride : nullptr;
		const int filter = src[0];
		++src;
		for (size_t x = 0; x < stride; ++x)
		{
			const int a = x >= size_t(bpp) ? cur[x - bpp] : 0, b = up ? up[x] : 0, c = (up && x >= size_t(bpp)) ? up[x - bpp] : 0;
			int v = src[x];
			switch (filter)
			{
			case 0: break;
			case 1: v += a; break;
			case 2: v += b; break;
			case 3: v += (a + b) >> 1; break;
			case 4: v += paeth(a, b, c); break;
			default: return false;
			}
			cur[x] = u8(v);
		}
	}
	consumed = (stride + 1) * size_t(h);
	return true;
}

bool decodePng(Reader r, int& W, int& H, std::vector<u8>& rgba, const Fail& fail)
{
	r.pos = 8;
	int depth = 0, colorType = 0, interlace = 0;
	std::vector<u8> idat, palette, trns;
	bool haveHeader = false;
	for (;;)
	{
		if (!r.has(8)) return fail("PNG: truncated");
		const uint32_t len = r.be32(), type = r.be32();
		if (!r.has(size_t(len) + 4)) return fail("PNG: truncated chunk");
		const u8* d = r.p + r.pos;
		if (type == 0x49484452u) // IHDR
		{
			if (len != 13) return fail("PNG: bad IHDR");
			Reader q{d, len, 0};
			W = int(q.be32()); H = int(q.be32()); depth = int(q.get8()); colorType = int(q.get8());
			const int comp = int(q.get8()), flt = int(q.get8()); interlace = int(q.get8());
			if (!sizeOk(W, H) || comp || flt || interlace > 1) return fail("PNG: unsupported header");
			haveHeader = true;
		}
		else if (type == 0x504c5445u) palette.assign(d, d + len);  // PLTE
		else if (type == 0x74524e53u) trns.assign(d, d + len);     // tRNS
		else if (type == 0x49444154u) idat.insert(idat.end(), d, d + len); // IDAT
		else if (type == 0x49454e44u) break;                       // IEND
		r.pos += size_t(len) + 4;
	}
	if (!haveHeader) return fail("PNG: no IHDR");
	static const int channelsOf[7] = {1, 0, 3, 1, 2, 0, 4};
	if (colorType > 6 || !channelsOf[colorType]) return fail("PNG: bad colour type");
	const int channels = channelsOf[colorType];
	if (!(depth == 8 || depth == 16 || (depth < 8 && (depth == 1 || depth == 2 || depth == 4) && (colorType == 0 || colorType == 3)))) return fail("PNG: bad bit depth");
	if (colorType == 3 && (depth == 16 || palette.size() < 3)) return fail("PNG: bad palette image");
	const int bitsPerPixel = channels * depth;
	// inflate
	std::vector<u8> raw;
	{
		size_t cap = 0;
		if (!interlace) cap = ((size_t(W) * bitsPerPixel + 7) / 8 + 1) * size_t(H);
		else cap = ((size_t(W) * bitsPerPixel + 7) / 8 + 8) * size_t(H) + 64 * 7;
		// deflate expands at most 1032 : 1 -- a header that claims more than the IDAT bytes can hold is corrupt (no allocation for it)
		if (!interlace && cap / 1032 > idat.size() + 16) return fail("PNG: truncated image data");
		raw.resize(cap + 64);
		z_stream zs;
		std::memset(&zs, 0, sizeof(zs));
		if (inflateInit(&zs) != Z_OK) return fail("PNG: zlib init failed");
		zs.next_in = idat.data(); zs.avail_in = uInt(idat.size());
		zs.next_out = raw.data(); zs.avail_out = uInt(raw.size());
		const int st = inflate(&zs, Z_FINISH);
		const size_t got = zs.total_out;
		inflateEnd(&zs);
		if (st != Z_STREAM_END && st != Z_OK && st != Z_BUF_ERROR) return fail("PNG: corrupt zlib stream");
		raw.resize(got);
	}
	// unfilter (+ Adam7 de-interlace) into one byte-per-sample-aligned pixel array
	const size_t stride = (size_t(W) * bitsPerPixel + 7) / 8;
	std::vector<u8> pix(stride * size_t(H), 0);
	if (!interlace)
	{
		size_t used = 0;
		if (!pngUnfilter(raw.data(), raw.size(), W, H, bitsPerPixel, pix, used)) return fail("PNG: corrupt image data");
	}
	else
	{
		static const int x0[7] = {0, 4, 0, 2, 0, 1, 0}, y0[7] = {0, 0, 4, 0, 2, 0, 1}, dx[7] = {8, 8, 4, 4, 2, 2, 1}, dy[7] = {8, 8, 8, 4, 4, 2, 2};
		size_t off = 0;
		for (int pass = 0; pass < 7; ++pass)
		{
			const int pw = (W - x0[pass] + dx[pass] - 1) / dx[pass], ph = (H - y0[pass] + dy[pass] - 1) / dy[pass];
			if (pw <= 0 || ph <= 0) continue;
			std::vector<u8> sub;
			size_t used = 0;
			if (off > raw.size() || !pngUnfilter(raw.data() + off, raw.size() - off, pw, ph, bitsPerPixel, sub, used)) return fail("PNG: corrupt interlaced data");
			off += used;
			const size_t subStride = (size_t(pw) * bitsPerPixel + 7) / 8;
			for (int y = 0; y < ph; ++y)
				for (int x = 0; x < pw; ++x)
				{
					const int tx = x0[pass] + x * dx[pass], ty = y0[pass] + y * dy[pass];
					if (bitsPerPixel >= 8)
						std::memcpy(&pix[stride * size_t(ty) + size_t(tx) * (bitsPerPixel / 8)], &sub[subStride * size_t(y) + size_t(x) * (bitsPerPixel / 8)], size_t(bitsPerPixel / 8));
					else
					{
						const int sbit = x * bitsPerPixel, tbit = tx * bitsPerPixel;
						const int v = (sub[subStride * size_t(y) + size_t(sbit / 8)] >> (8 - bitsPerPixel - (sbit & 7))) & ((1 << bitsPerPixel) - 1);
						pix[stride * size_t(ty) + size_t(tbit / 8)] |= u8(v << (8 - bitsPerPixel - (tbit & 7)));
					}
				}
		}
	}
	// to RGBA8
	rgba.assign(size_t(W) * H * 4, 255);
	const int bytesPerSample = depth == 16 ? 2 : 1;
	auto sample = [&](const u8* row, int x, int c) -> int
	{
		if (depth >= 8) return row[(size_t(x) * channels + c) * bytesPerSample]; // 16 bit: the high byte
		const int bit = x * depth;
		return (row[bit / 8] >> (8 - depth - (bit & 7))) & ((1 << depth) - 1);
	};
	auto sample16 = [&](const u8* row, int x, int c) -> int
	{
		const u8* s = row + (size_t(x) * channels + c) * 2;
		return (s[0] << 8) | s[1];
	};
	static const int scaleLow[5] = {0, 0xff, 0x55, 0, 0x11};
	for (int y = 0; y < H; ++y)
	{
		const u8* row = pix.data() + stride * size_t(y);
		u8* dst = rgba.data() + size_t(y) * W * 4;
		for (int x = 0; x < W; ++x, dst += 4)
		{
			if (colorType == 3)
			{
				const size_t i = size_t(sample(row, x, 0));
				if (3 * i + 2 < palette.size()) { dst[0] = palette[3 * i]; dst[1] = palette[3 * i + 1]; dst[2] = palette[3 * i + 2]; }
				else dst[0] = dst[1] = dst[2] = 0;
				dst[3] = i < trns.size() ? trns[i] : 255;
			}
			else if (colorType == 0 || colorType == 4)
			{
				int g = sample(row, x, 0);
				if (depth < 8) g *= scaleLow[depth];
				dst[0] = dst[1] = dst[2] = u8(g);
				if (colorType == 4) dst[3] = u8(sample(row, x, 1));
				else if (trns.size() >= 2)
				{
					const int key = (trns[0] << 8) | trns[1];
					const int v = depth == 16 ? sample16(row, x, 0) : sample(row, x, 0);
					if (v == key) dst[3] = 0;
				}
			}
			else
			{
				for (int c = 0; c < 3; ++c) dst[c] = u8(sample(row, x, c));
				if (colorType == 6) dst[3] = u8(sample(row, x, 3));
				else if (trns.size() >= 6)
				{
					bool eq = true;
					for (int c = 0; c < 3; ++c)
					{
						const int key = (trns[2 * c] << 8) | trns[2 * c + 1];
						const int v = depth == 16 ? sample16(row, x, c) : sample(row, x, c);
						eq = eq && v == key;
					}
					if (eq) dst[3] = 0;
				}
			}
		}
	}
	return true;
}

// ------------------------------------------------------------------------------------------------ BMP
int shiftOf(uint32_t mask) { int s = 0; if (!mask) return 0; while (!(mask & 1u)) { mask >>= 1; ++s; } return s; }
int bitsOf(uint32_t mask) { int b = 0; while (mask) { b += int(mask & 1u); mask >>= 1; } return b; }
u8 expandBits(uint32_t v, int bits)
{
	if (bits <= 0) return 0;
	if (bits >= 8) return u8(v >> (bits - 8));
	// replicate the bit pattern to fill 8 bits
	uint32_t out = 0; int have = 0;
	while (have < 8) { out = (out << bits) | v; have += bits; }
	return u8(out >> (have - 8));
}

bool decodeBmp(Reader r, int& W, int& H, std::vector<u8>& rgba, const Fail& fail)
{
	r.pos = 10;
	const uint32_t offset = r.le32(), hsz = r.le32();
	int w, h, bpp; uint32_t compress = 0, mr = 0, mg = 0, mb = 0, ma = 0;
	if (hsz == 12) { w = int(r.le16()); h = int(r.le16()); r.le16(); bpp = int(r.le16()); }
	else if (hsz == 40 || hsz == 52 || hsz == 56 || hsz == 108 || hsz == 124)
	{
		w = int(int32_t(r.le32())); h = int(int32_t(r.le32())); r.le16(); bpp = int(r.le16()); compress = r.le32();
		r.pos += 20; // image size, resolution, palette counts
		if (compress == 1 || compress == 2) return fail("BMP: RLE compression is not supported");
		if (hsz == 40 && compress == 3) { mr = r.le32(); mg = r.le32(); mb = r.le32(); }
		else if (hsz > 40) { mr = r.le32(); mg = r.le32(); mb = r.le32(); if (hsz >= 56) ma = r.le32(); }
		if (compress == 0)
		{
			if (bpp == 16) { mr = 31u << 10; mg = 31u << 5; mb = 31u; ma = 0; }
			else if (bpp == 32) { mr = 0xffu << 16; mg = 0xffu << 8; mb = 0xffu; ma = 0xffu << 24; }
		}
		else if (compress != 3) return fail("BMP: unknown compression");
	}
	else return fail("BMP: unknown header size");
	if (bpp != 1 && bpp != 2 && bpp != 4 && bpp != 8 && bpp != 16 && bpp != 24 && bpp != 32) return fail("BMP: unsupported bit depth");
	const bool flip = h > 0;
	if (h == INT32_MIN) return fail("BMP: bad size");
	h = std::abs(h);
	if (!sizeOk(w, h)) return fail("BMP: bad size");
	W = w; H = h;
	rgba.assign(size_t(w) * h * 4, 255);
	std::vector<u8> pal;
	if (bpp <= 8)
	{
		const size_t entry = hsz == 12 ? 3 : 4;
		const size_t count = std::min<size_t>((offset - 14 - hsz) / entry, 256);
		r.pos = 14 + hsz;
		pal.resize(count * 4, 255);
		for (size_t i = 0; i < count; ++i) { pal[4 * i + 2] = u8(r.get8()); pal[4 * i + 1] = u8(r.get8()); pal[4 * i] = u8(r.get8()); if (entry == 4) r.get8(); }
		if (count == 0) return fail("BMP: palette image without palette");
	}
	const size_t rowBytes = ((size_t(w) * bpp + 31) / 32) * 4;
	if (size_t(offset) + rowBytes * h > r.n) return fail("BMP: truncated");
	uint32_t anyAlpha = 0;
	for (int y = 0; y < h; ++y)
	{
		const u8* row = r.p + offset + rowBytes * size_t(y);
		u8* dst = rgba.data() + size_t(flip ? h - 1 - y : y) * w * 4;
		for (int x = 0; x < w; ++x, dst += 4)
		{
			if (bpp <= 8)
			{
				const int bit = x * bpp;
				size_t i = (row[bit / 8] >> (8 - bpp - (bit & 7))) & ((1 << bpp) - 1);
				if (4 * i + 2 >= pal.size()) i = 0;
				dst[0] = pal[4 * i]; dst[1] = pal[4 * i + 1]; dst[2] = pal[4 * i + 2];
			}
			else if (bpp == 24) { dst[2] = row[3 * x]; dst[1] = row[3 * x + 1]; dst[0] = row[3 * x + 2]; }
			else if (bpp == 16 || bpp == 32)
			{
				const uint32_t v = bpp == 16 ? uint32_t(row[2 * x] | (row[2 * x + 1] << 8)) : uint32_t(row[4 * x] | (row[4 * x + 1] << 8) | (row[4 * x + 2] << 16) | (uint32_t(row[4 * x + 3]) << 24));
				dst[0] = expandBits((v & mr) >> shiftOf(mr), bitsOf(mr));
				dst[1] = expandBits((v & mg) >> shiftOf(mg), bitsOf(mg));
				dst[2] = expandBits((v & mb) >> shiftOf(mb), bitsOf(mb));
				if (ma) { dst[3] = expandBits((v & ma) >> shiftOf(ma), bitsOf(ma)); anyAlpha |= dst[3]; }
			}
			else return fail("BMP: unsupported bit depth");
		}
	}
	// a 32-bit file whose alpha channel is all zero carries no alpha
	if (ma && anyAlpha == 0) for (size_t i = 3; i < rgba.size(); i += 4) rgba[i] = 255;
	return true;
}

// ------------------------------------------------------------------------------------------------ TGA
bool decodeTga(Reader r, int& W, int& H, std::vector<u8>& rgba, const Fail& fail)
{
	r.pos = 0;
	const uint32_t idLen = r.get8(), mapType = r.get8(), type = r.get8();
	const uint32_t mapFirst = r.le16(), mapLen = r.le16(), mapBits = r.get8();
	r.le16(); r.le16();
	const int w = int(r.le16()), h = int(r.le16());
	const int bits = int(r.get8()), desc = int(r.get8());
	const bool rle = type >= 8;
	const int kind = int(type & 7); // 1 colour-mapped, 2 true colour, 3 grey
	if (!sizeOk(w, h) || (kind != 1 && kind != 2 && kind != 3)) return fail("TGA: unsupported image type");
	if (kind == 1 && (!mapType || (bits != 8 && bits != 16))) return fail("TGA: bad colour map image");
	if (kind == 2 && bits != 15 && bits != 16 && bits != 24 && bits != 32) return fail("TGA: unsupported pixel depth");
	if (kind == 3 && bits != 8 && bits != 16) return fail("TGA: unsupported grey depth");
	r.pos += idLen;
	auto texel = [](const u8* s, int b, u8* out)
	{
		out[3] = 255;
		if (b == 8) { out[0] = out[1] = out[2] = s[0]; }
		else if (b == 15 || b == 16)
		{
			const uint32_t v = uint32_t(s[0] | (s[1] << 8));
			out[0] = u8((((v >> 10) & 31u) * 255u) / 31u); out[1] = u8((((v >> 5) & 31u) * 255u) / 31u); out[2] = u8(((v & 31u) * 255u) / 31u);
		}
		else { out[2] = s[0]; out[1] = s[1]; out[0] = s[2]; if (b == 32) out[3] = s[3]; }
	};
	std::vector<u8> map;
	if (mapType)
	{
		const size_t eb = (mapBits + 7) / 8;
		if (!r.has(eb * mapLen)) return fail("TGA: truncated colour map");
		if (kind == 1)
		{
			map.resize(size_t(mapLen) * 4);
			for (uint32_t i = 0; i < mapLen; ++i) texel(r.p + r.pos + eb * i, int(mapBits), &map[4 * size_t(i)]);
		}
		r.pos += eb * mapLen;
	}
	W = w; H = h;
	rgba.assign(size_t(w) * h * 4, 255);
	const size_t pb = size_t(bits + 7) / 8;
	const bool topDown = (desc & 0x20) != 0, rightLeft = (desc & 0x10) != 0;
	size_t count = 0, total = size_t(w) * h;
	u8 px[4] = {0, 0, 0, 255};
	auto decodeOne = [&](const u8* s)
	{
		if (kind == 1)
		{
			size_t i = bits == 8 ? s[0] : size_t(s[0] | (s[1] << 8));
			i = i >= mapFirst ? i - mapFirst : 0;
			if (4 * i + 3 < map.size()) std::memcpy(px, &map[4 * i], 4); else { px[0] = px[1] = px[2] = 0; px[3] = 255; }
		}
		else if (kind == 3 && bits == 16) { px[0] = px[1] = px[2] = s[0]; px[3] = s[1]; }
		else texel(s, bits, px);
	};
	auto put = [&]()
	{
		const size_t x = count % size_t(w), y = count / size_t(w);
		u8* dst = rgba.data() + ((topDown ? y : size_t(h) - 1 - y) * size_t(w) + (rightLeft ? size_t(w) - 1 - x : x)) * 4;
		std::memcpy(dst, px, 4);
		++count;
	};
	while (count < total)
	{
		if (!rle)
		{
			if (!r.has(pb)) return fail("TGA: truncated");
			decodeOne(r.p + r.pos); r.pos += pb; put();
		}
		else
		{
			if (!r.has(1)) return fail("TGA: truncated");
			const uint32_t head = r.get8(), run = (head & 127u) + 1u;
			if (head & 128u)
			{
				if (!r.has(pb)) return fail("TGA: truncated");
				decodeOne(r.p + r.pos); r.pos += pb;
				for (uint32_t i = 0; i < run && count < total; ++i) put();
			}
			else
				for (uint32_t i = 0; i < run && count < total; ++i)
				{
					if (!r.has(pb)) return fail("TGA: truncated");
					decodeOne(r.p + r.pos); r.pos += pb; put();
				}
		}
	}
	return true;
}

// ------------------------------------------------------------------------------------------------ JPEG (baseline / extended sequential, Huffman, 8 bit)
struct Huff
{
	// canonical code tables (T.81 Annex C): per code length the smallest code, the largest code and the index of its first value
	int minCode[17], maxCode[18], valPtr[17];
	u8 values[256];
	bool present = false;
	void build(const u8 counts[16])
	{
		int code = 0, k = 0;
		for (int len = 1; len <= 16; ++len)
		{
			valPtr[len] = k; minCode[len] = code;
			code += counts[len - 1]; k += counts[len - 1];
			maxCode[len] = counts[len - 1] ? code - 1 : -1;
			code <<= 1;
		}
		maxCode[17] = 0x7fffffff;
		present = true;
	}
};

struct JpegDecoder
{
	Reader r;
	const Fail& fail;
	int W = 0, H = 0, ncomp = 0;
	struct Comp { int id, h, v, tq, td, ta, pred; int bw, bh; std::vector<u8> plane; int stride; } comp[3];
	uint16_t quant[4][64];
	Huff dc[4], ac[4];
	int restartInterval = 0;
	// bit reader over the entropy-coded segment (byte stuffing: FF 00 -> FF; a marker ends the data)
	uint32_t bitBuf = 0; int bitCnt = 0; bool hitMarker = false;

	JpegDecoder(Reader rr, const Fail& f) : r(rr), fail(f) { std::memset(quant, 0, sizeof(quant)); }

	void fill()
	{
		while (bitCnt <= 24)
		{
			uint32_t b = 0;
			if (!hitMarker && r.pos < r.n)
			{
				b = r.p[r.pos];
				if (b == 0xff)
				{
					const uint32_t nx = r.u8at(r.pos + 1);
					if (nx == 0) r.pos += 2;
					else { hitMarker = true; b = 0; }
				}
				else ++r.pos;
			}
			bitBuf |= b << (24 - bitCnt);
			bitCnt += 8;
		}
	}
	int getBits(int n)
	{
		if (n == 0) return 0;
		if (bitCnt < n) fill();
		const int v = int(bitBuf >> (32 - n));
		bitBuf <<= n; bitCnt -= n;
		return v;
	}
	int decodeSymbol(const Huff& h)
	{
		if (bitCnt < 16) fill();
		int code = 0;
		for (int len = 1; len <= 16; ++len)
		{
			code = (code << 1) | int(bitBuf >> 31);
			bitBuf <<= 1; --bitCnt;
			if (h.maxCode[len] >= 0 && code <= h.maxCode[len] && code >= h.minCode[len]) return h.values[h.valPtr[len] + code - h.minCode[len]];
		}
		return -1;
	}
	static int extend(int v, int bits) { return v < (1 << (bits - 1)) ? v - (1 << bits) + 1 : v; }

	bool decodeBlock(Comp& c, int block[64])
	{
		static const u8 zigzag[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
		                              35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
		std::memset(block, 0, 64 * sizeof(int));
		const int t = decodeSymbol(dc[c.td]);
		if (t < 0 || t > 15) return false;
		const int diff = t ? extend(getBits(t), t) : 0;
		c.pred += diff;
		block[0] = c.pred * quant[c.tq][0];
		for (int k = 1; k < 64;)
		{
			const int rs = decodeSymbol(ac[c.ta]);
			if (rs < 0) return false;
			const int run = rs >> 4, size = rs & 15;
			if (size == 0)
			{
				if (run != 15) break; // end of block
				k += 16;
				continue;
			}
			k += run;
			if (k > 63) return false;
			block[zigzag[k]] = extend(getBits(size), size) * quant[c.tq][k];
			++k;
		}
		return true;
	}

	// separable inverse DCT in double precision (the definition, T.81 A.3.3), level shift and clamp
	static void idct(const int in[64], u8* out, int stride)
	{
		static double cosTable[8][8];
		static bool init = false;
		if (!init)
		{
			for (int x = 0; x < 8; ++x)
				for (int u = 0; u < 8; ++u) cosTable[x][u] = (u == 0 ? std::sqrt(0.125) : 0.5) * std::cos((2 * x + 1) * u * 3.14159265358979323846 / 16.0);
			init = true;
		}
		double tmp[64];
		for (int v = 0; v < 8; ++v)
			for (int x = 0; x < 8; ++x)
			{
				double s = 0.0;
				for (int u = 0; u < 8; ++u) s += cosTable[x][u] * in[v * 8 + u];
				tmp[v * 8 + x] = s;
			}
		for (int x = 0; x < 8; ++x)
			for (int y = 0; y < 8; ++y)
			{
				double s = 0.0;
				for (int v = 0; v < 8; ++v) s += cosTable[y][v] * tmp[v * 8 + x];
				const int val = int(std::floor(s + 128.5));
				out[y * stride + x] = u8(val < 0 ? 0 : (val > 255 ? 255 : val));
			}
	}

	bool readSegments(bool& sawScan)
	{
		r.pos = 2;
		for (;;)
		{
			if (!r.has(4)) return fail("JPEG: truncated");
			uint32_t m = r.get8();
			if (m != 0xff) return fail("JPEG: marker expected");
			while ((m = r.get8()) == 0xff) {}
			if (m == 0xd8 || (m >= 0xd0 && m <= 0xd7) || m == 0x01) continue;
			if (m == 0xd9) return fail("JPEG: no image data");
			const size_t len = r.be16();
			if (len < 2 || !r.has(len - 2)) return fail("JPEG: truncated segment");
			Reader s{r.p + r.pos, len - 2, 0};
			r.pos += len - 2;
			if (m == 0xdb) // DQT
			{
				while (s.pos < s.n)
				{
					const uint32_t pq = s.get8();
					const int t = int(pq & 15u);
					if (t > 3) return fail("JPEG: bad quantisation table");
					for (int i = 0; i < 64; ++i) quant[t][i] = uint16_t((pq >> 4) ? s.be16() : s.get8());
				}
			}
			else if (m == 0xc4) // DHT
			{
				while (s.pos < s.n)
				{
					const uint32_t tc = s.get8();
					const int t = int(tc & 15u);
					if (t > 3) return fail("JPEG: bad Huffman table");
					u8 counts[16]; int total = 0;
					for (int i = 0; i < 16; ++i) { counts[i] = u8(s.get8()); total += counts[i]; }
					if (total > 256) return fail("JPEG: bad Huffman table");
					Huff& h = (tc >> 4) ? ac[t] : dc[t];
					for (int i = 0; i < total; ++i) h.values[i] = u8(s.get8());
					h.build(counts);
				}
			}
			else if (m == 0xc0 || m == 0xc1) // SOF0 / SOF1
			{
				if (s.get8() != 8) return fail("JPEG: only 8-bit samples are supported");
				H = int(s.be16()); W = int(s.be16()); ncomp = int(s.get8());
				if (!sizeOk(W, H) || (ncomp != 1 && ncomp != 3)) return fail("JPEG: unsupported frame");
				for (int i = 0; i < ncomp; ++i)
				{
					comp[i].id = int(s.get8());
					const uint32_t hv = s.get8();
					comp[i].h = int(hv >> 4); comp[i].v = int(hv & 15u); comp[i].tq = int(s.get8());
					if (comp[i].h < 1 || comp[i].h > 4 || comp[i].v < 1 || comp[i].v > 4 || comp[i].tq > 3) return fail("JPEG: bad component");
				}
			}
			else if (m == 0xc2 || m == 0xc6 || m == 0xca || m == 0xce) return fail("JPEG: progressive streams are not supported");
			else if (m == 0xc3 || (m >= 0xc5 && m <= 0xcf && m != 0xc8 && m != 0xcc)) return fail("JPEG: lossless / arithmetic / hierarchical streams are not supported");
			else if (m == 0xdd) restartInterval = int(s.be16());
			else if (m == 0xda) // SOS
			{
				if (!ncomp) return fail("JPEG: scan before frame header");
				const int ns = int(s.get8());
				if (ns != ncomp) return fail("JPEG: non-interleaved scans are not supported");
				for (int i = 0; i < ns; ++i)
				{
					const int id = int(s.get8());
					const uint32_t t = s.get8();
					int which = -1;
					for (int c = 0; c < ncomp; ++c) if (comp[c].id == id) which = c;
					if (which < 0) return fail("JPEG: scan references an unknown component");
					comp[which].td = int(t >> 4); comp[which].ta = int(t & 15u);
					if (comp[which].td > 3 || comp[which].ta > 3 || !dc[comp[which].td].present || !ac[comp[which].ta].present) return fail("JPEG: missing Huffman table");
				}
				sawScan = true;
				return true;
			}
		}
	}

	bool decode(int& outW, int& outH, std::vector<u8>& rgba)
	{
		bool sawScan = false;
		if (!readSegments(sawScan) || !sawScan) return false;
		int hmax = 1, vmax = 1;
		for (int i = 0; i < ncomp; ++i) { hmax = std::max(hmax, comp[i].h); vmax = std::max(vmax, comp[i].v); }
		const int mcuW = 8 * hmax, mcuH = 8 * vmax, mcusX = (W + mcuW - 1) / mcuW, mcusY = (H + mcuH - 1) / mcuH;
		// the upsampling below replicates / filters by whole factors (1, 2, 3, 4): a component sampled 2 : 3 against the largest one has none
		for (int i = 0; i < ncomp; ++i)
			if (hmax % comp[i].h || vmax % comp[i].v) return fail("JPEG: fractional sampling ratios are not supported");
		for (int i = 0; i < ncomp; ++i)
		{
			comp[i].bw = mcusX * comp[i].h; comp[i].bh = mcusY * comp[i].v;
			comp[i].stride = comp[i].bw * 8;
			comp[i].plane.assign(size_t(comp[i].stride) * comp[i].bh * 8, 0);
			comp[i].pred = 0;
		}
		int block[64], untilRestart = restartInterval;
		for (int my = 0; my < mcusY; ++my)
			for (int mx = 0; mx < mcusX; ++mx)
			{
				for (int i = 0; i < ncomp; ++i)
					for (int by = 0; by < comp[i].v; ++by)
						for (int bx = 0; bx < comp[i].h; ++bx)
						{
							if (!decodeBlock(comp[i], block)) return fail("JPEG: corrupt entropy-coded data");
							idct(block, comp[i].plane.data() + size_t((my * comp[i].v + by) * 8) * comp[i].stride + size_t(mx * comp[i].h + bx) * 8, comp[i].stride);
						}
				if (restartInterval && --untilRestart == 0 && !(my == mcusY - 1 && mx == mcusX - 1))
				{
					// byte-align, expect RSTn
					bitBuf = 0; bitCnt = 0;
					if (hitMarker) { hitMarker = false; }
					while (r.pos + 1 < r.n && !(r.p[r.pos] == 0xff && r.p[r.pos + 1] >= 0xd0 && r.p[r.pos + 1] <= 0xd7)) ++r.pos;
					if (r.pos + 1 < r.n) r.pos += 2;
					for (int i = 0; i < ncomp; ++i) comp[i].pred = 0;
					untilRestart = restartInterval;
				}
			}
		// upsample (triangle filter for 2:1, replication otherwise) + colour conversion
		outW = W; outH = H;
		rgba.assign(size_t(W) * H * 4, 255);
		std::vector<u8> full[3];
		for (int i = 0; i < ncomp; ++i)
		{
			const int sx = hmax / comp[i].h, sy = vmax / comp[i].v;
			const int cw = (W * comp[i].h + hmax - 1) / hmax, ch = (H * comp[i].v + vmax - 1) / vmax; // meaningful samples of the plane
			full[i].resize(size_t(W) * H);
			const u8* src = comp[i].plane.data();
			const int st = comp[i].stride;
			auto at = [&](int x, int y) -> int { x = x < 0 ? 0 : (x >= cw ? cw - 1 : x); y = y < 0 ? 0 : (y >= ch ? ch - 1 : y); return src[size_t(y) * st + x]; };
			for (int y = 0; y < H; ++y)
				for (int x = 0; x < W; ++x)
				{
					int v;
					if (sx == 1 && sy == 1) v = src[size_t(y) * st + x];
					else if (sx == 2 && sy == 1)
					{
						const int cx = x >> 1, nb = (x & 1) ? cx + 1 : cx - 1;
						v = (3 * at(cx, y) + at(nb, y) + 2) >> 2;
						// The reference's decoder (stb_image v2.15, its 2:1 horizontal resampler) weights the LAST chroma sample's left output 1 : 3
						// instead of 3 : 1; textures are what the reference sees, so the quirk is kept (one column of 4:2:2 images).
						if (!(x & 1) && cx == cw - 1 && cw >= 2) v = (3 * at(cw - 2, y) + at(cw - 1, y) + 2) >> 2;
					}
					else if (sx == 1 && sy == 2)
					{
						const int cy = y >> 1, nb = (y & 1) ? cy + 1 : cy - 1;
						v = (3 * at(x, cy) + at(x, nb) + 2) >> 2;
					}
					else if (sx == 2 && sy == 2)
					{
						const int cx = x >> 1, cy = y >> 1, nx = (x & 1) ? cx + 1 : cx - 1, ny = (y & 1) ? cy + 1 : cy - 1;
						v = (9 * at(cx, cy) + 3 * at(nx, cy) + 3 * at(cx, ny) + at(nx, ny) + 8) >> 4;
					}
					else v = at(x / sx, y / sy);
					full[i][size_t(y) * W + x] = u8(v);
				}
		}
		for (size_t p = 0; p < size_t(W) * H; ++p)
		{
			u8* dst = &rgba[4 * p];
			if (ncomp == 1) { dst[0] = dst[1] = dst[2] = full[0][p]; continue; }
			const double Y = full[0][p], cb = double(full[1][p]) - 128.0, cr = double(full[2][p]) - 128.0;
			const double rr = Y + 1.402 * cr, gg = Y - 0.344136 * cb - 0.714136 * cr, bb = Y + 1.772 * cb;
			auto clamp = [](double v) { const int i = int(std::floor(v + 0.5)); return u8(i < 0 ? 0 : (i > 255 ? 255 : i)); };
			dst[0] = clamp(rr); dst[1] = clamp(gg); dst[2] = clamp(bb);
		}
		return true;
	}
};

// ------------------------------------------------------------------------------------------------ PNM / GITX
bool decodePnm(Reader r, int& W, int& H, std::vector<u8>& rgba, const Fail& fail)
{
	const bool colour = r.p[1] == '6';
	r.pos = 2;
	auto number = [&]() -> int
	{
		for (;;)
		{
			while (r.pos < r.n && (r.p[r.pos] == ' ' || r.p[r.pos] == '\t' || r.p[r.pos] == '\n' || r.p[r.pos] == '\r')) ++r.pos;
			if (r.pos < r.n && r.p[r.pos] == '#') { while (r.pos < r.n && r.p[r.pos] != '\n') ++r.pos; continue; }
			break;
		}
		int v = -1;
		while (r.pos < r.n && r.p[r.pos] >= '0' && r.p[r.pos] <= '9') { v = (v < 0 ? 0 : v * 10) + (r.p[r.pos] - '0'); ++r.pos; }
		return v;
	};
	const int w = number(), h = number(), maxv = number();
	++r.pos; // the single whitespace after maxval
	if (!sizeOk(w, h) || maxv != 255) return fail("PNM: only 8-bit binary P5 / P6 files are supported");
	const size_t need = size_t(w) * h * (colour ? 3 : 1);
	if (!r.has(need)) return fail("PNM: truncated");
	W = w; H = h;
	rgba.assign(size_t(w) * h * 4, 255);
	const u8* s = r.p + r.pos;
	for (size_t i = 0; i < size_t(w) * h; ++i)
	{
		if (colour) { rgba[4 * i] = s[3 * i]; rgba[4 * i + 1] = s[3 * i + 1]; rgba[4 * i + 2] = s[3 * i + 2]; }
		else rgba[4 * i] = rgba[4 * i + 1] = rgba[4 * i + 2] = s[i];
	}
	return true;
}

// format dispatch by signature
bool decodeAny(const unsigned char* data, size_t size, Reader r, int& width, int& height, std::vector<unsigned char>& rgba, const Fail& fail)
{
	static const u8 pngSig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
	if (!std::memcmp(data, pngSig, 8)) return decodePng(r, width, height, rgba, fail);
	if (data[0] == 0xff && data[1] == 0xd8) { JpegDecoder d(r, fail); return d.decode(width, height, rgba); }
	if (data[0] == 'B' && data[1] == 'M') return decodeBmp(r, width, height, rgba, fail);
	if (data[0] == 'P' && (data[1] == '5' || data[1] == '6')) return decodePnm(r, width, height, rgba, fail);
	if (!std::memcmp(data, "GITX", 4))
	{
		int32_t dims[2];
		std::memcpy(dims, data + 4, 8);
		if (!sizeOk(dims[0], dims[1]) || size < 12 + size_t(dims[0]) * dims[1] * 4) return fail("GITX: truncated");
		width = dims[0]; height = dims[1];
		rgba.assign(data + 12, data + 12 + size_t(width) * height * 4);
		return true;
	}
	// TGA has no signature: accept when the header is plausible
	{
		const uint32_t mapType = data[1], type = data[2], bits = data[16];
		if (mapType <= 1 && (type == 1 || type == 2 || type == 3 || type == 9 || type == 10 || type == 11) && (bits == 8 || bits == 15 || bits == 16 || bits == 24 || bits == 32))
			return decodeTga(r, width, height, rgba, fail);
	}
	return fail("unknown image format");
}

} // namespace

bool DecodeImageRGBA8(const unsigned char* data, size_t size, int& width, int& height, std::vector<unsigned char>& rgba, std::string* error)
{
	const Fail fail{error};
	Reader r{data, size, 0};
	if (size < 18) return fail("image file too short");
	try { return decodeAny(data, size, r, width, height, rgba, fail); }
	catch (const std::bad_alloc&) { return fail("out of memory while decoding"); }
}

bool LoadImageRGBA8(const std::string& file, int& width, int& height, std::vector<unsigned char>& rgba, std::string* error)
{
	std::ifstream f(file, std::ios::binary);
	if (!f) { if (error) *error = "cannot open '" + file + "'"; return false; }
	std::vector<unsigned char> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
	std::string err;
	if (!DecodeImageRGBA8(bytes.data(), bytes.size(), width, height, rgba, &err))
	{
		if (error) *error = "'" + file + "': " + err;
		return false;
	}
	return true;
}

} // namespace gi
