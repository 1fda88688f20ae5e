// lightpathtracer.hpp -- host mirror of gpugi/renderer/lightpathtracer.hpp.
#pragma once
#include "renderer.hpp"

class LightPathtracer : public Renderer
{
public:
	LightPathtracer(RendererSystem& _rendererSystem);

	std::string GetName() const override { return "LPT"; }

	void SetScene(std::shared_ptr<Scene> _scene) override;
	void SetScreenSize(const gl::Texture2D& _newBackbuffer) override;

	void Draw() override;

	int GetNumRaysPerLightSample() const { return m_numRaysPerLightSample; }

private:
	int m_maxPathLength;
	int m_numRaysPerLightSample;
	static const unsigned int m_localSizeLightPathtracer;

	void RecompileShaders(const std::string& _additionalDefines);
};
