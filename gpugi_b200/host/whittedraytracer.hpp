// whittedraytracer.hpp -- host mirror of gpugi/renderer/whittedraytracer.hpp.
#pragma once
#include "renderer.hpp"

class WhittedRayTracer : public Renderer
{
public:
	WhittedRayTracer(RendererSystem& _rendererSystem);

	std::string GetName() const override { return "WPT"; } // whittedraytracer.hpp:19

	void SetScene(std::shared_ptr<Scene> _scene) override;
	void SetScreenSize(const gl::Texture2D& _newBackbuffer) override;

	void Draw() override;

private:
	int m_maxPathLength; ///< the MAX_PATHLENGTH #define baked in by RecompileShaders (whittedraytracer.cpp:46)

	void RecompileShaders(const std::string& _additionalDefines);
};
