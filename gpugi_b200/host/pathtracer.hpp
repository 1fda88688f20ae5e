// pathtracer.hpp -- host mirror of gpugi/renderer/pathtracer.hpp:13-33.
#pragma once
#include "renderer.hpp"

class Pathtracer : public Renderer
{
public:
	Pathtracer(RendererSystem& _rendererSystem);

	std::string GetName() const override { return "PT"; }

	void SetScene(std::shared_ptr<Scene> _scene) override;
	void SetScreenSize(const gl::Texture2D& _newBackbuffer) override;
	void SetEnvironmentMap(std::shared_ptr<gl::TextureCubemap> _envMap) override;

	void Draw() override;

private:
	int m_maxPathLength; ///< the MAX_PATHLENGTH #define baked in by RecompileShaders (pathtracer.cpp:61)

	void RecompileShaders(const std::string& _additionalDefines);
};
