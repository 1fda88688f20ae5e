// scriptprocessing.cpp -- see scriptprocessing.hpp.  Line syntax (scriptprocessing.cpp:66-110): everything after '#' or "//" is a
// comment, blanks and tabs are dropped, "name=v0,v1,..." sets a parameter / fires a command, a bare name with a non-empty
// current value prints it.
#include "scriptprocessing.hpp"
#include "globalconfig.hpp"
#include "logger.hpp"

#include <algorithm>
#include <fstream>
#include <iostream>

ScriptProcessing::ScriptProcessing() : m_scriptWaitSeconds(0.0), m_scriptWaitIterations(0) {}

bool ScriptProcessing::RunScript(const std::string& _scriptFilename)
{
	std::ifstream file(_scriptFilename);
	if (file.bad() || file.fail())
	{
		LOG_ERROR("Failed to load script \"" + _scriptFilename + "\"");
		return false;
	}
	std::string line;
	while (std::getline(file, line))
	{
		if (!line.empty() && line.back() == '\r') line.pop_back(); // the reference's scripts have DOS line ends
		m_scriptCommandQueue.push(line);
	}
	return true;
}

void ScriptProcessing::PushCommand(const std::string& _line)
{
	if (_line.empty()) return;
	std::lock_guard<std::mutex> lock(m_consoleCommandQueueMutex);
	m_consoleCommandQueue.push(_line);
}

bool ScriptProcessing::IsIdle()
{
	std::lock_guard<std::mutex> lock(m_consoleCommandQueueMutex);
	return m_scriptCommandQueue.empty() && m_consoleCommandQueue.empty() && m_scriptWaitSeconds <= 0.0 && m_scriptWaitIterations == 0;
}

void ScriptProcessing::ParseCommand(std::string _commandLine, bool _fromScriptFile, bool& _stopProcessing)
{
	auto commentStart = _commandLine.find('#');
	if (commentStart == std::string::npos) commentStart = _commandLine.find("//");
	if (commentStart != std::string::npos) _commandLine.erase(commentStart);
	_commandLine.erase(std::remove(_commandLine.begin(), _commandLine.end(), ' '), _commandLine.end());
	_commandLine.erase(std::remove(_commandLine.begin(), _commandLine.end(), '\t'), _commandLine.end());
	if (_commandLine.empty()) return;

	const std::string::size_type equalityIdx = _commandLine.find('=');
	const std::string name = _commandLine.substr(0, equalityIdx);
	GlobalConfig::ParameterType argumentList;
	if (equalityIdx != std::string::npos)
	{
		std::string::size_type lastComma = equalityIdx, nextComma;
		do
		{
			nextComma = _commandLine.find(',', lastComma + 1);
			argumentList.push_back(GlobalConfig::ParameterElementType(nextComma == std::string::npos ? _commandLine.substr(lastComma + 1)
			                                                                                          : _commandLine.substr(lastComma + 1, nextComma - lastComma - 1)));
			lastComma = nextComma;
		} while (nextComma != std::string::npos);
	}

	if (_fromScriptFile)
	{
		if (name == "waitSeconds")
		{
			if (argumentList.size() != 1) LOG_ERROR("The waitSeconds command expects a single number as argument.");
			else m_scriptWaitSeconds = argumentList[0].As<float>();
			_stopProcessing = true;
			return;
		}
		if (name == "waitIterations")
		{
			if (argumentList.size() != 1) LOG_ERROR("The waitIterations command expects a single number as argument.");
			else m_scriptWaitIterations = argumentList[0].As<unsigned int>();
			_stopProcessing = true;
			return;
		}
	}

	// an empty argument list on a parameter that has a value is a getter
	try
	{
		GlobalConfig::ParameterType currentValue = GlobalConfig::GetParameter(name);
		if (argumentList.empty() && !currentValue.empty())
		{
			std::cout << "{";
			for (size_t i = 0; i + 1 < currentValue.size(); ++i) std::cout << currentValue[i].As<std::string>() << ", ";
			std::cout << currentValue.back().As<std::string>() << " }\n";
		}
		else
			GlobalConfig::SetParameter(name, argumentList);
	}
	catch (const std::invalid_argument& e)
	{
		LOG_ERROR(e.what());
	}
	_stopProcessing = m_processPauseCommand.find(name) != m_processPauseCommand.end();
}

void ScriptProcessing::ProcessCommandQueue(double timeDelta, unsigned int iterationDelta)
{
	m_scriptWaitSeconds -= timeDelta;
	if (m_scriptWaitIterations > 0) m_scriptWaitIterations -= std::min(iterationDelta, m_scriptWaitIterations);
	bool processingStop = false;
	while (m_scriptWaitSeconds <= 0.0 && m_scriptWaitIterations == 0 && !m_scriptCommandQueue.empty() && !processingStop)
	{
		ParseCommand(m_scriptCommandQueue.front(), true, processingStop);
		m_scriptCommandQueue.pop();
	}
	std::lock_guard<std::mutex> lock(m_consoleCommandQueueMutex);
	while (!m_consoleCommandQueue.empty() && !processingStop)
	{
		ParseCommand(m_consoleCommandQueue.front(), false, processingStop);
		m_consoleCommandQueue.pop();
	}
}
