// renderersystem.hpp -- host mirror of gpugi/renderer/renderersystem.hpp:27-199.  Same public
// methods; the GL resources (backbuffer texture, four UBOs, texture-buffer views, light-sample
// buffer) are replaced by one GiContext behind the C ABI of include/gpugi_b200.h.
#pragma once

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "../../include/gpugi_b200.h"
#include "camera.hpp"
#include "gltypes.hpp"
#include "lightsampler.hpp"
#include "renderer.hpp"
#include "vec.hpp"

class Scene;

class RendererSystem
{
public:
	/// \param _device CUDA device ordinal (one process per GPU).  Throws std::runtime_error if the
	///        device context cannot be created -- there is no CPU fallback.
	explicit RendererSystem(int _device = 0);
	~RendererSystem();

	/// Sets scene; active renderer is notified; resets iteration count (renderersystem.cpp:132-168).
	void SetScene(std::shared_ptr<Scene> _scene);
	const std::shared_ptr<Scene>& GetScene() const { return m_scene; }

	/// Replaces the cubemap for the sky (renderersystem.cpp:173-186).  Faces are binary PPM files.
	void SetEnvironmentMap(int _size, const std::string& _xneg, const std::string& _xpos, const std::string& _yneg, const std::string& _ypos,
	                       const std::string& _zneg, const std::string& _zpos);
	/// Same, from memory.
	void SetEnvironmentMap(std::shared_ptr<gl::TextureCubemap> _envMap);

	/// Sets camera; resets iteration count (renderersystem.cpp:188-219).
	void SetCamera(const Camera& camera);

	/// Sets back buffer size; resets iteration count (renderersystem.cpp:221-233).
	void SetScreenSize(const ei::IVec2& newSize);

	gl::Texture2D& GetBackbuffer() { return *m_backbuffer; }
	const gl::Texture2D& GetBackbuffer() const { return *m_backbuffer; }

	Renderer* GetActiveRenderer() { return m_activeRenderer; }

	/// Destroys old renderer, creates new renderer and makes it active (renderersystem.hpp:74-86).
	template <typename RendererType, typename... Args>
	void SetRenderer(Args&&... _arguments)
	{
		delete m_activeRenderer;
		m_activeRenderer = nullptr;
		ResetIterationCount();
		if (m_backbuffer) m_backbuffer->ClearToZero();

		m_activeRenderer = new RendererType(*this, _arguments...);
		if (m_backbuffer) m_activeRenderer->SetScreenSize(*m_backbuffer);
		m_activeRenderer->SetCamera(m_camera);
		if (m_scene) m_activeRenderer->SetScene(m_scene);
		m_activeRenderer->SetEnvironmentMap(m_envMap);
	}

	/// Performs a draw iteration (renderersystem.cpp:235-258).
	void Draw();

	unsigned long GetIterationCount() const { return m_iterationCount; }
	/// Duration of all passed iterations in milliseconds.
	uint64_t GetRenderTime() const { return m_renderTime; }

	void SetNumInitialLightSamples(unsigned int _numInitialLightSamples);
	unsigned int GetNumInitialLightSamples() const { return m_numInitialLightSamples; }

	void ResetIterationCount();
	void PerIterationBufferUpdate(bool _iterationIncrement = true);
	void DispatchShowLightCacheShader();

	// ---- additions over the reference (multi-GPU sample sharding) ------------------
	/// This process renders global iterations rank, rank+worldSize, ...; FrameSeed and the light
	/// sample stream are then functions of the GLOBAL iteration index (SURVEY 9 deviation).
	void SetIterationShard(unsigned int _rank, unsigned int _worldSize);
	/// Global index of the iteration the next Draw() renders.
	uint64_t GetGlobalIterationIndex() const { return uint64_t(m_iterationCount) * m_worldSize + m_rank; }
	/// FrameSeed of the next Draw().
	uint32_t GetFrameSeed() const { return m_frameSeed; }
	/// If false, Draw() does not synchronise the device (the reference's glFinish, :249,:253);
	/// GetRenderTime() then only counts host time.
	void SetSynchronousDraw(bool _sync) { m_synchronousDraw = _sync; }
	/// Streaming Draw() (GI_OPT_STREAMING): consecutive Draw() calls of the path tracer overlap wave by wave on the device
	/// instead of running each iteration to its last bounce; the accumulator is exact again after any synchronous read.
	/// Implies SetSynchronousDraw(false) while on.
	void SetStreaming(bool _streaming);
	/// Light samples of every iteration are drawn on the device (LightSampler::GenerateRandomSamplesOnDevice) instead of on the host
	/// + 4 KB upload; GetInitialLightSamples() then reads them back on demand.
	void SetDeviceLightSampling(bool _onDevice) { m_deviceLightSampling = _onDevice; }
	unsigned int GetWorldSize() const { return m_worldSize; }
	GiContext* GetDevice() const { return m_device; }
	const Camera& GetCamera() const { return m_camera; }
	const std::vector<LightSampler::LightSample>& GetInitialLightSamples();

private:
	GiContext* m_device;
	std::unique_ptr<gl::Texture2D> m_backbuffer;
	Camera m_camera;

	unsigned long m_iterationCount;
	uint64_t m_renderTime;
	uint32_t m_frameSeed;

	std::shared_ptr<Scene> m_scene;
	std::shared_ptr<gl::TextureCubemap> m_envMap;

	unsigned int m_numInitialLightSamples;
	std::vector<LightSampler::LightSample> m_lightSamples;
	LightSampler m_lightSampler;

	unsigned int m_rank, m_worldSize;
	bool m_synchronousDraw;
	bool m_deviceLightSampling = false;

	Renderer* m_activeRenderer;

	void UploadCamera();
	void Check(int status, const char* what) const;
};
