// gltypes.hpp -- stand-ins for the two glhelper types that appear in the Renderer interface
// (renderer.hpp:14-18,38,40): gl::Texture2D (the RGBA32F backbuffer) and gl::TextureCubemap (the
// sky).  They carry no GL state; the pixels live in the GiContext behind the C ABI.
#pragma once
#include "../../include/gpugi_b200.h"
#include <cstdint>
#include <vector>

namespace gl
{
	class Texture2D
	{
	public:
		Texture2D(GiContext* ctx, int width, int height) : m_ctx(ctx), m_width(width), m_height(height) {}
		int GetWidth() const { return m_width; }
		int GetHeight() const { return m_height; }
		/// gl::Texture::ClearToZero
		void ClearToZero(int = 0) const { gi_clear_accumulator(m_ctx); }
		/// gl::Texture::ReadImage(RGBA, FLOAT): width*height*4 floats, the un-normalised sum
		bool ReadImage(float* rgba) const { return gi_read_accumulator(m_ctx, rgba) == GI_OK; }
		/// Progressive display read (gi_read_accumulator_async): does not stall or drain the renderer; `rgba` is complete
		/// after WaitReads() / the read after next.  iterationsInFlight: submitted iterations not yet in the image.
		bool ReadImageAsync(float* rgba, uint32_t* iterationsInFlight) const { return gi_read_accumulator_async(m_ctx, rgba, iterationsInFlight) == GI_OK; }
		bool WaitReads() const { return gi_wait_reads(m_ctx) == GI_OK; }
		/// Multi-GPU read-out (gi_comm_*): the sum of all ranks' backbuffers, this rank's rows only (gi_comm_slice) -- collective.
		bool ReadImageReduced(float* rgbaRows) const { return gi_read_accumulator_reduced(m_ctx, rgbaRows) == GI_OK; }
		bool ReadImageReducedAsync(float* rgbaRows, uint32_t* iterationsInFlight) const { return gi_read_accumulator_reduced_async(m_ctx, rgbaRows, iterationsInFlight) == GI_OK; }
		GiContext* GetDevice() const { return m_ctx; }

	private:
		GiContext* m_ctx;
		int m_width, m_height;
	};

	/// Six size*size RGBA8 (sRGB) faces in the order -X,+X,-Y,+Y,-Z,+Z (renderersystem.cpp:41-47).
	class TextureCubemap
	{
	public:
		explicit TextureCubemap(int size) : m_size(size), m_faces(6, std::vector<uint8_t>(size_t(size) * size * 4, 0)) {}
		int GetSize() const { return m_size; }
		std::vector<uint8_t>& Face(int i) { return m_faces[i]; }
		const std::vector<uint8_t>& Face(int i) const { return m_faces[i]; }

	private:
		int m_size;
		std::vector<std::vector<uint8_t>> m_faces;
	};
}
