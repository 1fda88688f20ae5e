// scriptprocessing.hpp -- host mirror of gpugi/control/scriptprocessing.hpp:13-58: "name = v0, v1" command lines from a script file
// (scenes/*.txt) or from the caller, executed through GlobalConfig.  Script files have two extra commands, waitSeconds and
// waitIterations, that delay the following commands.  (The reference's console thread reads std::cin; here PushCommand is the
// thread-safe entry for a front end to feed lines.)
#pragma once

#include <mutex>
#include <queue>
#include <string>
#include <unordered_set>

class ScriptProcessing
{
public:
	ScriptProcessing();

	/// Loads a sequence of commands from the given script file (scriptprocessing.cpp:15-27).
	bool RunScript(const std::string& _scriptFilename);
	/// Queues one console command (the role of CommandWindowThread, :29-42).
	void PushCommand(const std::string& _line);
	/// Parses and processes all commands in the queues (:160-183).
	void ProcessCommandQueue(double timeDelta, unsigned int iterationDelta = 1);
	/// Adds a command that causes ProcessCommandQueue to stop processing and return (:38).
	void AddProcessingPauseCommand(const std::string& _processPauseCommand) { m_processPauseCommand.insert(_processPauseCommand); }
	/// No command and no wait pending (headless runs end here; the reference keeps drawing until the window closes).
	bool IsIdle();
	/// One line, executed at once (:66-158).
	void ParseCommand(std::string command, bool fromScriptFile, bool& _stopProcessing);

private:
	std::queue<std::string> m_scriptCommandQueue;
	double m_scriptWaitSeconds;
	unsigned int m_scriptWaitIterations;
	std::mutex m_consoleCommandQueueMutex;
	std::queue<std::string> m_consoleCommandQueue;
	std::unordered_set<std::string> m_processPauseCommand;
};
