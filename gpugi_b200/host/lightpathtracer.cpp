// lightpathtracer.cpp -- host mirror of gpugi/renderer/lightpathtracer.cpp:10-70.
#include "lightpathtracer.hpp"
#include "globalconfig.hpp"
#include "gltypes.hpp"
#include "logger.hpp"
#include "renderersystem.hpp"
#include "scene.hpp"

#include <algorithm>

const unsigned int LightPathtracer::m_localSizeLightPathtracer = 8 * 8;

LightPathtracer::LightPathtracer(RendererSystem& _rendererSystem) : Renderer(_rendererSystem), m_maxPathLength(16), m_numRaysPerLightSample(0)
{
	_rendererSystem.SetNumInitialLightSamples(256);
}

void LightPathtracer::SetScene(std::shared_ptr<Scene> _scene)
{
	RecompileShaders(_scene ? _scene->GetBvhTypeDefineString() : "");
}

void LightPathtracer::SetScreenSize(const gl::Texture2D& _newBackbuffer)
{
	// Rule: every block (64 paths) works with the same initial light sample (lightpathtracer.cpp:32-34).
	int numPixels = _newBackbuffer.GetWidth() * _newBackbuffer.GetHeight();
	m_numRaysPerLightSample = std::max(m_localSizeLightPathtracer,
	                                   (numPixels / m_rendererSystem.GetNumInitialLightSamples() / m_localSizeLightPathtracer) * m_localSizeLightPathtracer);
}

void LightPathtracer::Draw()
{
	GiIterationParams p = {};
	p.renderer = GI_RENDERER_LPT;
	p.frame_seed = m_rendererSystem.GetFrameSeed();
	p.max_path_length = m_maxPathLength;
	p.num_rays_per_light_sample = m_numRaysPerLightSample;
	if (gi_render_iteration(m_rendererSystem.GetDevice(), &p) != GI_OK)
		LOG_ERROR("LightPathtracer::Draw failed: " << gi_last_error(m_rendererSystem.GetDevice()));

	m_rendererSystem.DispatchShowLightCacheShader();
}

void LightPathtracer::RecompileShaders(const std::string&)
{
	m_maxPathLength = GlobalConfig::GetParameter("pathLength")[0].As<int>();
}
