// texturemse.hpp -- host mirror of gpugi/imageut/texturemse.hpp: MSE of the current backbuffer against a reference PFM
// (sqerror.frag:16-17 + the mip-chain mean, texturemse.cpp:39-66), evaluated on the device by gi_compute_mse.
#pragma once
#include "gltypes.hpp"
#include "hdrimage.hpp"
#include <limits>
#include <memory>
#include <string>

class TextureMSE
{
public:
	void LoadReferenceFromPFM(const std::string& _filename)
	{
		ei::IVec2 size;
		auto data = LoadPfm(_filename, size);
		if (data) { m_reference = std::move(data); m_size = size; }
	}
	bool HasValidReferenceImage() const { return m_reference != nullptr; }
	/// \param _divisor the iteration count: the backbuffer holds a sum (texturemse.cpp:39).  Infinity without a (matching) reference.
	float ComputeCurrentMSE(const gl::Texture2D& _texture, unsigned int _divisor) const
	{
		float mse = std::numeric_limits<float>::infinity();
		if (m_reference && _divisor > 0 && m_size.x == _texture.GetWidth() && m_size.y == _texture.GetHeight())
			if (gi_compute_mse(_texture.GetDevice(), &m_reference[0].x, _divisor, &mse) != GI_OK) mse = std::numeric_limits<float>::infinity();
		return mse;
	}

private:
	std::unique_ptr<ei::Vec3[]> m_reference;
	ei::IVec2 m_size;
};
