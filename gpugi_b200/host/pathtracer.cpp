// pathtracer.cpp -- host mirror of gpugi/renderer/pathtracer.cpp:14-66.  glDispatchCompute of
// pathtracer.comp (:50) becomes gi_render_iteration(GI_RENDERER_PT).
#include "pathtracer.hpp"
#include "globalconfig.hpp"
#include "gltypes.hpp"
#include "logger.hpp"
#include "renderersystem.hpp"
#include "scene.hpp"

Pathtracer::Pathtracer(RendererSystem& _rendererSystem) : Renderer(_rendererSystem), m_maxPathLength(16)
{
	m_rendererSystem.SetNumInitialLightSamples(128);
}

void Pathtracer::SetScene(std::shared_ptr<Scene> _scene)
{
	RecompileShaders(_scene ? _scene->GetBvhTypeDefineString() : "");
}

void Pathtracer::SetScreenSize(const gl::Texture2D&)
{
	// the backbuffer is bound inside the device context (gi_set_screen_size)
}

void Pathtracer::SetEnvironmentMap(std::shared_ptr<gl::TextureCubemap> _envMap)
{
	if (!_envMap) return;
	const uint8_t* faces[6];
	for (int i = 0; i < 6; ++i) faces[i] = _envMap->Face(i).data();
	if (gi_set_environment_map(m_rendererSystem.GetDevice(), _envMap->GetSize(), faces) != GI_OK)
		LOG_ERROR("gi_set_environment_map failed: " << gi_last_error(m_rendererSystem.GetDevice()));
}

void Pathtracer::Draw()
{
	GiIterationParams p = {};
	p.renderer = GI_RENDERER_PT;
	p.frame_seed = m_rendererSystem.GetFrameSeed();
	p.max_path_length = m_maxPathLength;
	if (gi_render_iteration(m_rendererSystem.GetDevice(), &p) != GI_OK)
		LOG_ERROR("Pathtracer::Draw failed: " << gi_last_error(m_rendererSystem.GetDevice()));
}

void Pathtracer::RecompileShaders(const std::string&)
{
	// The shader #defines are kernel parameters here; only MAX_PATHLENGTH is latched (pathtracer.cpp:61).
	m_maxPathLength = GlobalConfig::GetParameter("pathLength")[0].As<int>();
}
