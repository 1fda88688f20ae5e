// bidirectionalpathtracer.hpp -- host mirror of gpugi/renderer/bidirectionalpathtracer.hpp.
#pragma once
#include "renderer.hpp"
#include <cstdint>

class BidirectionalPathtracer : public Renderer
{
public:
	BidirectionalPathtracer(RendererSystem& _rendererSystem);

	std::string GetName() const override { return "BPT"; }

	void SetScene(std::shared_ptr<Scene> _scene) override;
	void SetScreenSize(const gl::Texture2D& _newBackbuffer) override;

	void Draw() override;

	int GetNumRaysPerLightSample() const { return m_numRaysPerLightSample; }
	uint32_t GetLightCacheCapacity() const { return m_lightCacheCapacity; }
	int GetAverageLightPathLength() const { return m_averageLightPathLength; }

private:
	/// Estimates light cache capacity and creates a new cache (bidirectionalpathtracer.cpp:65-132).
	void CreateLightCacheWithCapacityEstimate();
	void RecompileShaders(const std::string& _additionalDefines);

	bool m_needToDetermineNeededLightCacheCapacity;
	int m_maxPathLength;
	int m_numRaysPerLightSample;
	uint32_t m_lightCacheCapacity;
	int m_averageLightPathLength;

	static const unsigned int m_localSizeLightPathtracer;
};
