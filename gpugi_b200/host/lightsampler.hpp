// lightsampler.hpp -- host mirror of gpugi/scene/lightsampler.hpp:9-35.
#pragma once
#include "../../include/gpugi_b200.h"
#include <cstdint>
#include <memory>

class Scene;

class LightSampler
{
public:
	LightSampler();

	void SetScene(std::shared_ptr<const Scene> _scene);

	/// Same record as lightsampler.hpp:19-25 (32 bytes).
	typedef GiLightSample LightSample;

	/// \param _positionBias moves the sample along the triangle normal (lightsampler.hpp:29-31).
	void GenerateRandomSamples(LightSample* _destinationBuffer, unsigned int _numSamples, float _positionBias = 1.0e-4f);

	/// The same samples generated on the device straight into the context's light-sample ring (gi_generate_light_samples):
	/// no host loop, no upload.  The random stream advances exactly as in GenerateRandomSamples, so both can be mixed.
	bool GenerateRandomSamplesOnDevice(GiContext* _device, unsigned int _numSamples, float _positionBias = 1.0e-4f);

	/// Deviation from the reference: the stream can be re-seeded per iteration so that any
	/// rank can render any iteration.  The reference's free-running stream is the default (seed 123).
	void SetSeed(std::uint64_t _seed) { m_randomSeed = _seed; }
	std::uint64_t GetSeed() const { return m_randomSeed; }

private:
	std::shared_ptr<const Scene> m_scene;
	std::uint64_t m_randomSeed;
	GiContext* m_tablesDevice = nullptr;   // where the light tables were uploaded last, and which version
	unsigned m_tablesVersion = 0;
	const Scene* m_tablesScene = nullptr;
};
