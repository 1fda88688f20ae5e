// bidirectionalpathtracer.cpp -- host mirror of gpugi/renderer/bidirectionalpathtracer.cpp:16-201.
#include "bidirectionalpathtracer.hpp"
#include "globalconfig.hpp"
#include "gltypes.hpp"
#include "logger.hpp"
#include "renderersystem.hpp"
#include "scene.hpp"

#include <algorithm>

const unsigned int BidirectionalPathtracer::m_localSizeLightPathtracer = 8 * 8;

BidirectionalPathtracer::BidirectionalPathtracer(RendererSystem& _rendererSystem)
	: Renderer(_rendererSystem), m_needToDetermineNeededLightCacheCapacity(false), m_maxPathLength(16), m_numRaysPerLightSample(0), m_lightCacheCapacity(0),
	  m_averageLightPathLength(0)
{
	m_rendererSystem.SetNumInitialLightSamples(128);
}

void BidirectionalPathtracer::SetScene(std::shared_ptr<Scene> _scene)
{
	RecompileShaders(_scene ? _scene->GetBvhTypeDefineString() : "");
	m_needToDetermineNeededLightCacheCapacity = true; // a new scene changes the light path statistics
}

void BidirectionalPathtracer::SetScreenSize(const gl::Texture2D& _newBackbuffer)
{
	int numPixels = _newBackbuffer.GetWidth() * _newBackbuffer.GetHeight();
	m_numRaysPerLightSample = std::max(m_localSizeLightPathtracer,
	                                   (numPixels / m_rendererSystem.GetNumInitialLightSamples() / m_localSizeLightPathtracer) * m_localSizeLightPathtracer);
	m_needToDetermineNeededLightCacheCapacity = true;
}

void BidirectionalPathtracer::CreateLightCacheWithCapacityEstimate()
{
	GiContext* dev = m_rendererSystem.GetDevice();
	LOG_LVL2("Starting bpt warmup to estimate needed light cache capacity...");
	static const unsigned int numWarmupRuns = 8;
	const unsigned int blockCountPerRun = m_rendererSystem.GetNumInitialLightSamples() * m_numRaysPerLightSample / m_localSizeLightPathtracer;
	std::uint64_t numCacheEntries = 0;
	std::uint64_t summedPathLength = 0;
	for (unsigned int run = 0; run < numWarmupRuns; ++run)
	{
		GiIterationParams p = {};
		p.renderer = GI_RENDERER_BPT;
		p.frame_seed = m_rendererSystem.GetFrameSeed();
		p.max_path_length = m_maxPathLength;
		p.num_rays_per_light_sample = m_numRaysPerLightSample;
		p.light_cache_capacity = 0x7fffffff; // not known yet (bidirectionalpathtracer.cpp:39)
		p.warmup = 1;
		if (gi_render_iteration(dev, &p) != GI_OK)
			LOG_ERROR("BPT warm-up failed: " << gi_last_error(dev));
		GiCounters c;
		gi_get_counters(dev, &c); // synchronises
		numCacheEntries += c.light_cache_entries;
		summedPathLength += c.light_path_length_sum;
		// Important to change random seed
		m_rendererSystem.PerIterationBufferUpdate();
	}
	m_rendererSystem.ResetIterationCount();
	if (m_rendererSystem.GetWorldSize() > 1)
		m_rendererSystem.PerIterationBufferUpdate(false); // sharded mode: light samples are a function of the global iteration

	m_averageLightPathLength = static_cast<int>(summedPathLength / numWarmupRuns / blockCountPerRun / m_localSizeLightPathtracer + 1);
	numCacheEntries /= numWarmupRuns;
	m_lightCacheCapacity = static_cast<unsigned int>(numCacheEntries + numCacheEntries / 10); // 10% safety margin
	LOG_LVL2("... bpt warmup done. Reserving " << (uint64_t(m_lightCacheCapacity) * 64 / 1024 / 1024) << "mb light sample cache. Average light path length "
	                                          << m_averageLightPathLength);
	if (gi_alloc_light_cache(dev, m_lightCacheCapacity) != GI_OK)
		LOG_ERROR("gi_alloc_light_cache failed: " << gi_last_error(dev));
}

void BidirectionalPathtracer::RecompileShaders(const std::string&)
{
	m_maxPathLength = GlobalConfig::GetParameter("pathLength")[0].As<int>();
}

void BidirectionalPathtracer::Draw()
{
	if (m_needToDetermineNeededLightCacheCapacity)
	{
		CreateLightCacheWithCapacityEstimate();
		m_needToDetermineNeededLightCacheCapacity = false;
	}
	// resetcounter -> light pass (SAVE_LIGHT_CACHE) -> eye pass: one call, ordered on the stream
	GiIterationParams p = {};
	p.renderer = GI_RENDERER_BPT;
	p.frame_seed = m_rendererSystem.GetFrameSeed();
	p.max_path_length = m_maxPathLength;
	p.num_rays_per_light_sample = m_numRaysPerLightSample;
	p.light_cache_capacity = int32_t(m_lightCacheCapacity);
	if (gi_render_iteration(m_rendererSystem.GetDevice(), &p) != GI_OK)
		LOG_ERROR("BidirectionalPathtracer::Draw failed: " << gi_last_error(m_rendererSystem.GetDevice()));
}
