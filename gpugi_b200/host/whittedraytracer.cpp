// whittedraytracer.cpp -- host mirror of gpugi/renderer/whittedraytracer.cpp:13-50.  glDispatchCompute of whittedraytracer.comp
// (STOP_ON_DIFFUSE_BOUNCE build, :38-49) becomes gi_render_iteration(GI_RENDERER_WHITTED).
#include "whittedraytracer.hpp"
#include "globalconfig.hpp"
#include "gltypes.hpp"
#include "logger.hpp"
#include "renderersystem.hpp"
#include "scene.hpp"

WhittedRayTracer::WhittedRayTracer(RendererSystem& _rendererSystem) : Renderer(_rendererSystem), m_maxPathLength(16)
{
	m_rendererSystem.SetNumInitialLightSamples(128);
}

void WhittedRayTracer::SetScene(std::shared_ptr<Scene> _scene)
{
	RecompileShaders(_scene ? _scene->GetBvhTypeDefineString() : "");
}

void WhittedRayTracer::SetScreenSize(const gl::Texture2D&)
{
	// the backbuffer is bound inside the device context (gi_set_screen_size)
}

void WhittedRayTracer::Draw()
{
	GiIterationParams p = {};
	p.renderer = GI_RENDERER_WHITTED;
	p.frame_seed = m_rendererSystem.GetFrameSeed();
	p.max_path_length = m_maxPathLength;
	if (gi_render_iteration(m_rendererSystem.GetDevice(), &p) != GI_OK)
		LOG_ERROR("WhittedRayTracer::Draw failed: " << gi_last_error(m_rendererSystem.GetDevice()));
}

void WhittedRayTracer::RecompileShaders(const std::string&)
{
	m_maxPathLength = GlobalConfig::GetParameter("pathLength")[0].As<int>();
}
