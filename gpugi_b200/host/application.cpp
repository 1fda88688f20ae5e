// application.cpp -- see application.hpp.  Command names, value counts and defaults are the reference's (application.cpp:77-193,
// camera.cpp:41-60, outputwindow.cpp for "resolution"); renderers 3-7 of its list are outside this library and answer with an error.
#include "application.hpp"
#include "bidirectionalpathtracer.hpp"
#include "hdrimage.hpp"
#include "lightpathtracer.hpp"
#include "logger.hpp"
#include "pathtracer.hpp"
#include "whittedraytracer.hpp"
#include "scene.hpp"

#include <cmath>
#include <ctime>
#include <iostream>
#include <sys/stat.h>

namespace {
const char* const kParams[] = {"pause", "help", "exit", "tic", "toc", "screenshot", "namedscreenshot", "screenshotDir", "cameraPos", "cameraLookAt", "cameraFOV",
                               "renderer", "debugrenderer", "resolution", "bvhType", "sceneFilename", "envMap", "addPointLight", "setPointLight", "referenceImage",
                               "mseCheckInterval", "waitMSE", "waitSeconds", "waitIterations", "streaming"};

std::string UIntToMinLengthString(int _number, int _minDigits)
{
	int zeros = std::max(0, _minDigits - static_cast<int>(std::ceil(std::log10(_number + 1))));
	return std::string(size_t(zeros), '0') + std::to_string(_number);
}
}

Application::Application(const std::string& scriptFilename, int cudaDevice)
	: m_shutdown(false), m_scriptWaitsForMSE(false), m_iterationSinceLastMSECheck(0), m_lastRenderTimeStamp(0.0), m_lastMSE(std::numeric_limits<float>::infinity())
{
	m_stopwatch = m_lastUpdate = std::chrono::steady_clock::now();
	// application.cpp:40-41: start camera + 1024 x 768 window
	m_camera.reset(new Camera(ei::Vec3(0.0f), ei::Vec3(0.0f, 0.0f, 1.0f), 1024.0f / 768.0f, 70.0f));
	m_rendererSystem.reset(new RendererSystem(cudaDevice));
	m_textureMSE.reset(new TextureMSE());
	RegisterScriptCommands();
	m_rendererSystem->SetScreenSize(ei::IVec2(1024, 768));
	if (!scriptFilename.empty()) m_scriptProcessing.RunScript(scriptFilename);
}

Application::~Application()
{
	if (!m_shutdown && m_rendererSystem->GetActiveRenderer() && m_rendererSystem->GetIterationCount() > 0) SaveImage(); // application.cpp:196-198
	for (const char* p : kParams) GlobalConfig::RemoveParameter(p);
}

void Application::RegisterScriptCommands()
{
	for (const char* p : kParams) GlobalConfig::RemoveParameter(p); // a second Application in one process starts from the defaults
	GlobalConfig::AddParameter("pause", {false}, "Set to true to pause drawing.");
	GlobalConfig::AddParameter("help", {}, "Dumps all available parameter/events.");
	GlobalConfig::AddListener("help", "HelpDumpFunc", [](const GlobalConfig::ParameterType&) { std::cout << GlobalConfig::GetEntryDescriptions() << std::endl; });
	GlobalConfig::AddParameter("exit", {}, "Exits the application. Does NOT take a screenshot.");
	GlobalConfig::AddListener("exit", "exit application", [this](const GlobalConfig::ParameterType&) { m_shutdown = true; });
	GlobalConfig::AddParameter("tic", {}, "Resets stopwatch.");
	GlobalConfig::AddListener("tic", "stopwatch reset", [this](const GlobalConfig::ParameterType&) { m_stopwatch = std::chrono::steady_clock::now(); });
	GlobalConfig::AddParameter("toc", {}, "Gets current stopwatch time.");
	GlobalConfig::AddListener("toc", "stopwatch reset", [this](const GlobalConfig::ParameterType&) {
		gi_synchronize(m_rendererSystem->GetDevice());
		std::cout << std::chrono::duration<double>(std::chrono::steady_clock::now() - m_stopwatch).count() << "s" << std::endl;
	});

	// screenshots (PFM, divided by the iteration count)
	GlobalConfig::AddParameter("screenshot", {}, "Saves a screenshot.");
	GlobalConfig::AddListener("screenshot", "SaveScreenshot", [this](const GlobalConfig::ParameterType&) { SaveImage(); });
	GlobalConfig::AddParameter("namedscreenshot", {std::string("")}, "Saves a screenshot with the given name.");
	GlobalConfig::AddListener("namedscreenshot", "SaveScreenshot", [this](const GlobalConfig::ParameterType& p) { SaveImage(p[0].As<std::string>()); });
	GlobalConfig::AddParameter("screenshotDir", {std::string("../screenshots/")}, "Directory screenshots are written to (the reference's fixed ../screenshots/).");

	// camera (camera.cpp:41-60)
	GlobalConfig::AddParameter("cameraPos", {m_camera->GetPosition().x, m_camera->GetPosition().y, m_camera->GetPosition().z}, "Global camera's position");
	GlobalConfig::AddListener("cameraPos", "global camera", [this](const GlobalConfig::ParameterType& p) { m_camera->SetPosition(ei::Vec3(p[0].As<float>(), p[1].As<float>(), p[2].As<float>())); });
	GlobalConfig::AddParameter("cameraLookAt", {m_camera->GetLookAt().x, m_camera->GetLookAt().y, m_camera->GetLookAt().z}, "Global camera's look at");
	GlobalConfig::AddListener("cameraLookAt", "global camera", [this](const GlobalConfig::ParameterType& p) { m_camera->SetLookAt(ei::Vec3(p[0].As<float>(), p[1].As<float>(), p[2].As<float>())); });
	GlobalConfig::AddParameter("cameraFOV", {m_camera->GetVFov()}, "Global camera's vertical FOV");
	GlobalConfig::AddListener("cameraFOV", "global camera", [this](const GlobalConfig::ParameterType& p) { m_camera->SetVFov(p[0].As<float>()); });
	auto updateCamera = [this](const GlobalConfig::ParameterType&) { m_rendererSystem->SetCamera(*m_camera); };
	GlobalConfig::AddListener("cameraPos", "update renderer cam", updateCamera);
	GlobalConfig::AddListener("cameraLookAt", "update renderer cam", updateCamera);
	GlobalConfig::AddListener("cameraFOV", "update renderer cam", updateCamera);

	GlobalConfig::AddParameter("renderer", {0}, "Change this value to change the active renderer (resets previous results!).\n0: Pathtracer\n1: LightPathtracer\n2: BidirectionalPathtracer\n4: Whitted Pathtracer");
	GlobalConfig::AddListener("renderer", "ChangeRenderer", [this](const GlobalConfig::ParameterType& p) { SwitchRenderer(p); });
	GlobalConfig::AddParameter("debugrenderer", {-1}, "Debug renderers are not part of this library; only -1 (off) is accepted.");
	GlobalConfig::AddListener("debugrenderer", "Renderer", [](const GlobalConfig::ParameterType& p) { if (p[0].As<int>() != -1) LOG_ERROR("debug renderers are not part of gpugi_b200"); });
	if (!GlobalConfig::HasParameter("pathLength")) GlobalConfig::AddParameter("pathLength", {16}, "Change the maximum path length of almost any renderer.");

	GlobalConfig::AddParameter("resolution", {1024, 768}, "The window's width and height.");
	GlobalConfig::AddListener("resolution", "global camera aspect", [this](const GlobalConfig::ParameterType& p) {
		m_camera->SetAspectRatio(p[0].As<float>() / p[1].As<float>());
		m_rendererSystem->SetCamera(*m_camera);
		m_rendererSystem->SetScreenSize(ei::IVec2(p[0].As<int>(), p[1].As<int>()));
		m_rendererSystem->SetCamera(*m_camera); // (pixel area depends on the new size)
	});

	GlobalConfig::AddParameter("bvhType", {0}, "Use AABoxes (0) or OBoxes (1) for the BVH. This parameter must be set before the scene is loaded!");
	GlobalConfig::AddParameter("sceneFilename", {std::string("")}, "Change this value to load a new scene.");
	GlobalConfig::AddListener("sceneFilename", "LoadScene", [this](const GlobalConfig::ParameterType& p) {
		std::string sceneFilename = p[0].As<std::string>();
		std::cout << "Loading scene " << sceneFilename << std::endl;
		ei::Types3D bvhType = GlobalConfig::GetParameter("bvhType")[0].As<int>() == 0 ? ei::Types3D::BOX : ei::Types3D::OBOX;
		m_scene = std::make_shared<Scene>(sceneFilename, bvhType);
		m_rendererSystem->SetScene(m_scene);
	});

	GlobalConfig::AddParameter("envMap", {512, std::string(""), std::string(""), std::string(""), std::string(""), std::string(""), std::string("")},
	                           "Replace the environment map: size, then 6 faces in the order x-, x+, y-, y+, z-, z+ (binary PPM).");
	GlobalConfig::AddListener("envMap", "LoadEnvmap", [this](const GlobalConfig::ParameterType& p) {
		m_rendererSystem->SetEnvironmentMap(p[0].As<int>(), p[1].As<std::string>(), p[2].As<std::string>(), p[3].As<std::string>(), p[4].As<std::string>(),
		                                    p[5].As<std::string>(), p[6].As<std::string>());
	});

	GlobalConfig::AddParameter("addPointLight", {0.0f, 0.0f, 0.0f, 1.0f, 1.0f, 1.0f}, "Create a new point light source.");
	GlobalConfig::AddListener("addPointLight", "addPointLight", [this](const GlobalConfig::ParameterType& p) {
		if (!m_scene) { LOG_ERROR("addPointLight: no scene loaded"); return; }
		Scene::PointLight light;
		light.position = ei::Vec3(p[0].As<float>(), p[1].As<float>(), p[2].As<float>());
		light.intensity = ei::Vec3(p[3].As<float>(), p[4].As<float>(), p[5].As<float>());
		m_scene->AddPointLight(light);
	});
	GlobalConfig::AddParameter("setPointLight", {0, 0.0f, 0.0f, 0.0f, 1.0f, 1.0f, 1.0f}, "Change setting of an existing point light. The index (param 1) depends on creation order.");
	GlobalConfig::AddListener("setPointLight", "setPointLight", [this](const GlobalConfig::ParameterType& p) {
		if (!m_scene) { LOG_ERROR("setPointLight: no scene loaded"); return; }
		Scene::PointLight light;
		light.position = ei::Vec3(p[1].As<float>(), p[2].As<float>(), p[3].As<float>());
		light.intensity = ei::Vec3(p[4].As<float>(), p[5].As<float>(), p[6].As<float>());
		m_scene->SetPointLight(p[0].As<unsigned int>(), light);
	});

	GlobalConfig::AddParameter("referenceImage", {std::string("")}, "Reference for all mse checks.");
	GlobalConfig::AddListener("referenceImage", "LoadReference", [this](const GlobalConfig::ParameterType& p) { m_textureMSE->LoadReferenceFromPFM(p[0].As<std::string>()); });
	GlobalConfig::AddParameter("mseCheckInterval", {0}, "MSE will be computed every xth iteration. 0 for no MSE computation");
	GlobalConfig::AddParameter("waitMSE", {0}, "Script execution halts until the MSE is lower than the given value. Only used if mseCheckInterval is >0.");
	GlobalConfig::AddListener("waitMSE", "ScriptWait", [this](const GlobalConfig::ParameterType&) {
		if (GlobalConfig::GetParameter("mseCheckInterval")[0].As<int>() <= 0) LOG_ERROR("Can't activate waitMSE as long the mseCheckInterval is 0!");
		else if (!m_textureMSE->HasValidReferenceImage()) LOG_ERROR("Can't activate waitMSE as long as there is no reference image loaded!");
		else m_scriptWaitsForMSE = true;
	});
	m_scriptProcessing.AddProcessingPauseCommand("waitMSE");
	GlobalConfig::AddParameter("waitSeconds", {std::string("")}, "Halt script execution for next x seconds.");
	GlobalConfig::AddParameter("waitIterations", {std::string("")}, "Halt script execution for next x iterations.");

	// addition: streaming Draw (GI_OPT_STREAMING) -- consecutive path-tracer iterations overlap on the device; exact again at every read
	GlobalConfig::AddParameter("streaming", {true}, "Overlap consecutive Draw() calls of the path tracer on the device (results are unchanged).");
	GlobalConfig::AddListener("streaming", "SetStreaming", [this](const GlobalConfig::ParameterType& p) { m_rendererSystem->SetStreaming(p[0].As<bool>()); });
	m_rendererSystem->SetStreaming(true);
}

void Application::SwitchRenderer(const GlobalConfig::ParameterType& p)
{
	switch (p[0].As<int>())
	{
	case 0: m_rendererSystem->SetRenderer<Pathtracer>(); break;
	case 1: m_rendererSystem->SetRenderer<LightPathtracer>(); break;
	case 2: m_rendererSystem->SetRenderer<BidirectionalPathtracer>(); break;
	case 4: m_rendererSystem->SetRenderer<WhittedRayTracer>(); break;
	default: LOG_ERROR("renderer " << p[0].As<int>() << " is not part of gpugi_b200 (0: Pathtracer, 1: LightPathtracer, 2: BidirectionalPathtracer, 4: Whitted)"); return;
	}
}

void Application::Run(unsigned long maxIterations)
{
	unsigned long loops = 0;
	while (!m_shutdown)
	{
		Update();
		if (m_shutdown) break;
		if (m_scriptProcessing.IsIdle() && !m_scriptWaitsForMSE) break;
		if (!GlobalConfig::GetParameter("pause")[0].As<bool>()) Draw();
		if (maxIterations && ++loops >= maxIterations) break;
	}
}

void Application::Update()
{
	if (!m_scriptWaitsForMSE)
	{
		// application.cpp:307-313 feeds the script clock with the renderer's accumulated time; Draw() does not block here, so the
		// wall clock between updates takes its place
		auto now = std::chrono::steady_clock::now();
		double delta = std::chrono::duration<double>(now - m_lastUpdate).count();
		m_lastUpdate = now;
		m_scriptProcessing.ProcessCommandQueue(delta);
	}
}

void Application::Draw()
{
	if (!m_rendererSystem->GetActiveRenderer() || !m_scene) return;
	m_rendererSystem->Draw();
	MSEChecks();
}

void Application::MSEChecks()
{
	++m_iterationSinceLastMSECheck;
	int mseCheckInterval = GlobalConfig::GetParameter("mseCheckInterval")[0].As<int>();
	if (mseCheckInterval > 0 && m_iterationSinceLastMSECheck >= static_cast<unsigned int>(mseCheckInterval))
	{
		m_iterationSinceLastMSECheck = 0;
		float mse = m_textureMSE->ComputeCurrentMSE(m_rendererSystem->GetBackbuffer(), static_cast<unsigned int>(m_rendererSystem->GetIterationCount()));
		m_lastMSE = mse;
		LOG_LVL2("Iteration " << m_rendererSystem->GetIterationCount() << " MSE: " << mse);
		if (m_scriptWaitsForMSE)
		{
			float goalMSE = GlobalConfig::GetParameter("waitMSE")[0].As<float>();
			if (goalMSE > 0)
			{
				if (mse != mse) { LOG_ERROR("MSE is NaN. Resuming script."); m_scriptWaitsForMSE = false; }
				else if (goalMSE > mse) m_scriptWaitsForMSE = false;
			}
			else
				m_scriptWaitsForMSE = false;
		}
	}
}

void Application::SaveImage(const std::string& name)
{
	if (!m_rendererSystem->GetActiveRenderer()) return;
	time_t t = time(0);
	struct tm* now = localtime(&t);
	std::string date = UIntToMinLengthString(now->tm_mon + 1, 2) + "." + UIntToMinLengthString(now->tm_mday, 2) + " " + UIntToMinLengthString(now->tm_hour, 2) + ";" +
	                   UIntToMinLengthString(now->tm_min, 2) + ";" + std::to_string(now->tm_sec) + " ";
	std::string dir = GlobalConfig::GetParameter("screenshotDir")[0].As<std::string>();
	if (!dir.empty() && dir.back() != '/') dir += '/';
	if (!dir.empty()) mkdir(dir.c_str(), 0777);
	gl::Texture2D& backbuffer = m_rendererSystem->GetBackbuffer();
	std::unique_ptr<ei::Vec4[]> imageData(new ei::Vec4[size_t(backbuffer.GetWidth()) * backbuffer.GetHeight()]);
	if (!backbuffer.ReadImage(&imageData[0].x)) { LOG_ERROR("screenshot: reading the backbuffer failed"); return; }
	std::string filename = dir + date + m_rendererSystem->GetActiveRenderer()->GetName() + " " + std::to_string(m_rendererSystem->GetIterationCount()) + "it " +
	                       std::to_string(m_rendererSystem->GetRenderTime()) + "ms.pfm";
	if (!name.empty()) filename = dir + name + " " + filename.substr(dir.size());
	if (WritePfm(imageData.get(), ei::IVec2(backbuffer.GetWidth(), backbuffer.GetHeight()), filename, static_cast<std::uint32_t>(m_rendererSystem->GetIterationCount())))
	{
		m_lastScreenshot = filename;
		std::cout << "Wrote screenshot \"" << filename << "\"" << std::endl;
	}
}
