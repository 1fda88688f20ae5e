// application.hpp -- headless host mirror of gpugi/application.hpp/.cpp: owns the RendererSystem, the camera, the scene and the
// script processor, registers the reference's script commands (application.cpp:77-193) and runs the update / draw loop
// (:287-344).  No window: what the reference displays every frame is available through RendererSystem::GetBackbuffer(); the
// loop ends at "exit" or when the script has run out of commands (the reference keeps drawing until its window is closed).
#pragma once

#include <chrono>
#include <memory>
#include <string>

#include "camera.hpp"
#include "globalconfig.hpp"
#include "renderersystem.hpp"
#include "scriptprocessing.hpp"
#include "texturemse.hpp"

class Scene;

class Application
{
public:
	/// \param scriptFilename run this command script (argv[1] of the reference); may be empty
	explicit Application(const std::string& scriptFilename = "", int cudaDevice = 0);
	~Application();

	/// Update / Draw until "exit" or until no command and no wait is pending.  maxIterations > 0 bounds the loop.
	void Run(unsigned long maxIterations = 0);
	/// One console line (what the reference reads from std::cin), executed inside the next Update.
	void Command(const std::string& line) { m_scriptProcessing.PushCommand(line); }

	RendererSystem& GetRendererSystem() { return *m_rendererSystem; }
	const std::string& GetLastScreenshot() const { return m_lastScreenshot; }
	float GetLastMSE() const { return m_lastMSE; }
	bool HasShutDown() const { return m_shutdown; }

private:
	void RegisterScriptCommands();
	void SwitchRenderer(const GlobalConfig::ParameterType& p);
	void Update();
	void Draw();
	void MSEChecks();
	void SaveImage(const std::string& name = "");

	std::unique_ptr<RendererSystem> m_rendererSystem;
	std::unique_ptr<TextureMSE> m_textureMSE;
	std::unique_ptr<Camera> m_camera;
	std::shared_ptr<Scene> m_scene;
	ScriptProcessing m_scriptProcessing;
	bool m_shutdown;
	bool m_scriptWaitsForMSE;
	unsigned int m_iterationSinceLastMSECheck;
	double m_lastRenderTimeStamp;
	std::chrono::steady_clock::time_point m_stopwatch, m_lastUpdate;
	std::string m_lastScreenshot;
	float m_lastMSE;
};
