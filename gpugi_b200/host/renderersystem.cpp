// renderersystem.cpp -- host mirror of gpugi/renderer/renderersystem.cpp.  The control flow is the
// reference's (who resets what, when light samples are regenerated); every GL call is a gi_* call.
#include "renderersystem.hpp"
#include "logger.hpp"
#include "random.hpp"
#include "scene.hpp"
#include "imageio.hpp"

#include <chrono>
#include <fstream>
#include <stdexcept>

namespace {
// splitmix64: decorrelates the per-iteration light-sampler seeds in sharded mode.
inline uint64_t Mix64(uint64_t x)
{
	x += 0x9E3779B97F4A7C15ull;
	x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
	x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
	return x ^ (x >> 31);
}
}

RendererSystem::RendererSystem(int _device)
	: m_device(nullptr), m_iterationCount(0), m_renderTime(0), m_frameSeed(WangHash(0)), m_numInitialLightSamples(0), m_rank(0), m_worldSize(1),
	  m_synchronousDraw(true), m_activeRenderer(nullptr)
{
	int st = gi_create(_device, &m_device);
	if (st != GI_OK)
		throw std::runtime_error(std::string("gpugi_b200: cannot create the device context (no CPU fallback): ") + gi_last_error(nullptr));
	// black 1x1 fall back env-map (renderersystem.cpp:39-47)
	m_envMap = std::make_shared<gl::TextureCubemap>(1);
	PerIterationBufferUpdate();
	m_iterationCount = 0;
	m_renderTime = 0;
}

RendererSystem::~RendererSystem()
{
	delete m_activeRenderer;
	gi_destroy(m_device);
}

void RendererSystem::Check(int status, const char* what) const
{
	if (status != GI_OK) LOG_ERROR(what << " failed: " << gi_last_error(m_device));
}

void RendererSystem::SetIterationShard(unsigned int _rank, unsigned int _worldSize)
{
	m_rank = _rank;
	m_worldSize = _worldSize == 0 ? 1 : _worldSize;
	ResetIterationCount();
	PerIterationBufferUpdate(false);
}

void RendererSystem::SetStreaming(bool _streaming)
{
	Check(gi_set_option(m_device, GI_OPT_STREAMING, _streaming ? 1 : 0), "gi_set_option(GI_OPT_STREAMING)");
	if (_streaming) m_synchronousDraw = false;
}

const std::vector<LightSampler::LightSample>& RendererSystem::GetInitialLightSamples()
{
	if (m_deviceLightSampling && !m_lightSamples.empty())
		Check(gi_get_light_samples(m_device, m_lightSamples.data(), uint32_t(m_lightSamples.size())), "gi_get_light_samples");
	return m_lightSamples;
}

void RendererSystem::ResetIterationCount()
{
	m_iterationCount = 0;
	m_renderTime = 0;
	m_frameSeed = WangHash(static_cast<std::uint32_t>(GetGlobalIterationIndex()));
}

void RendererSystem::PerIterationBufferUpdate(bool _iterationIncrement)
{
	if (_iterationIncrement)
		++m_iterationCount;
	const uint64_t global = GetGlobalIterationIndex();
	m_frameSeed = WangHash(static_cast<std::uint32_t>(global));

	if (m_numInitialLightSamples > 0 && m_scene && m_scene->IsValid())
	{
		if (m_worldSize > 1)
		{
			uint64_t seed = Mix64(global) ^ 123ull;
			m_lightSampler.SetSeed(seed ? seed : 123ull);
		}
		if (m_deviceLightSampling && m_lightSampler.GenerateRandomSamplesOnDevice(m_device, m_numInitialLightSamples))
			return;
		m_lightSampler.GenerateRandomSamples(m_lightSamples.data(), m_numInitialLightSamples);
		Check(gi_set_light_samples(m_device, m_lightSamples.data(), m_numInitialLightSamples), "gi_set_light_samples");
	}
}

void RendererSystem::SetNumInitialLightSamples(unsigned int _numInitialLightSamples)
{
	m_numInitialLightSamples = _numInitialLightSamples;
	m_lightSamples.assign(_numInitialLightSamples, LightSampler::LightSample());
	PerIterationBufferUpdate(false);
}

void RendererSystem::SetScene(std::shared_ptr<Scene> _scene)
{
	m_scene = _scene;
	if (m_scene && m_scene->IsValid())
	{
		GiSceneDesc desc = m_scene->GetSceneDesc();
		Check(gi_upload_scene(m_device, &desc), "gi_upload_scene");
	}
	m_lightSampler.SetScene(m_scene);
	PerIterationBufferUpdate(false);

	m_iterationCount = 0;
	m_renderTime = 0;
	m_frameSeed = WangHash(static_cast<std::uint32_t>(GetGlobalIterationIndex()));
	if (m_backbuffer)
		m_backbuffer->ClearToZero(0);
	if (m_activeRenderer)
		m_activeRenderer->SetScene(m_scene);
}

namespace {
// one sky face: any format imageio.cpp decodes (the reference: gl::TextureCubemap faces through stb_image), of the given size
bool LoadSkyFace(const std::string& file, int size, std::vector<uint8_t>& rgba)
{
	std::vector<unsigned char> px;
	int w = 0, h = 0;
	if (!gi::LoadImageRGBA8(file, w, h, px) || w != size || h != size) return false;
	rgba.assign(px.begin(), px.end());
	return true;
}
}

void RendererSystem::SetEnvironmentMap(int _size, const std::string& _xneg, const std::string& _xpos, const std::string& _yneg, const std::string& _ypos,
                                       const std::string& _zneg, const std::string& _zpos)
{
	auto env = std::make_shared<gl::TextureCubemap>(_size);
	const std::string* files[6] = {&_xneg, &_xpos, &_yneg, &_ypos, &_zneg, &_zpos};
	for (int i = 0; i < 6; ++i)
		if (!LoadSkyFace(*files[i], _size, env->Face(i)))
			LOG_ERROR("Failed to load environment face '" + *files[i] + "' (an image of the given size expected); face stays black.");
	SetEnvironmentMap(env);
}

void RendererSystem::SetEnvironmentMap(std::shared_ptr<gl::TextureCubemap> _envMap)
{
	m_envMap = _envMap;
	if (m_activeRenderer)
		m_activeRenderer->SetEnvironmentMap(m_envMap);
}

void RendererSystem::UploadCamera()
{
	ei::Vec3 camU, camV, camW;
	m_camera.ComputeCameraParams(camU, camV, camW);
	GiCamera c;
	c.position[0] = m_camera.GetPosition().x; c.position[1] = m_camera.GetPosition().y; c.position[2] = m_camera.GetPosition().z;
	c.u[0] = camU.x; c.u[1] = camU.y; c.u[2] = camU.z;
	c.v[0] = camV.x; c.v[1] = camV.y; c.v[2] = camV.z;
	c.w[0] = camW.x; c.w[1] = camW.y; c.w[2] = camW.z;
	// renderersystem.cpp:208
	c.pixel_area = ei::len(camU) * ei::len(camV) * 4.0f / (m_backbuffer->GetWidth() * m_backbuffer->GetHeight());
	Check(gi_set_camera(m_device, &c), "gi_set_camera");
}

void RendererSystem::SetCamera(const Camera& _camera)
{
	m_iterationCount = 0;
	m_renderTime = 0;
	m_frameSeed = WangHash(static_cast<std::uint32_t>(GetGlobalIterationIndex()));
	m_camera = _camera;
	if (m_backbuffer)
	{
		UploadCamera(); // also clears the accumulator
	}
	if (m_activeRenderer)
		m_activeRenderer->SetCamera(_camera);
}

void RendererSystem::SetScreenSize(const ei::IVec2& _newSize)
{
	Check(gi_set_screen_size(m_device, _newSize.x, _newSize.y), "gi_set_screen_size");
	m_backbuffer.reset(new gl::Texture2D(m_device, _newSize.x, _newSize.y));
	m_iterationCount = 0;
	m_renderTime = 0;
	m_frameSeed = WangHash(static_cast<std::uint32_t>(GetGlobalIterationIndex()));
	if (m_activeRenderer)
		m_activeRenderer->SetScreenSize(*m_backbuffer);
}

void RendererSystem::Draw()
{
	if (!m_activeRenderer)
		return;
	if (m_synchronousDraw) gi_synchronize(m_device);
	auto begin = std::chrono::steady_clock::now();
	m_activeRenderer->Draw();
	PerIterationBufferUpdate();
	if (m_synchronousDraw) gi_synchronize(m_device);
	auto end = std::chrono::steady_clock::now();
	m_renderTime += uint64_t(std::chrono::duration_cast<std::chrono::milliseconds>(end - begin).count());
}

void RendererSystem::DispatchShowLightCacheShader()
{
	Check(gi_show_light_caches(m_device), "gi_show_light_caches");
}
