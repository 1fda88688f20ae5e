// scriptprocessing.cpp -- command-line front end of the host layer (interface: scriptprocessing.hpp).
//
// Written against the BEHAVIOUR of gpugi/control/scriptprocessing.cpp:66-183 (the scenes/*.txt syntax), not its text:
//   * a line is cut at the first '#', else at the first "//"; blanks and tabs carry no meaning anywhere in a line;
//   * "name=a,b,c" fires GlobalConfig::SetParameter(name, {a,b,c}); an empty field between two commas is an empty string value;
//   * a bare "name" whose parameter currently holds values prints them, otherwise it fires the parameter with no values (a command);
//   * script files (not the console) own two delay commands, waitSeconds=<s> and waitIterations=<n>, which suspend the script queue;
//   * names registered with AddProcessingPauseCommand end the current ProcessCommandQueue round after they ran.
// There is no std::cin thread here: PushCommand is the console's entry point.
#include "scriptprocessing.hpp"
#include "globalconfig.hpp"
#include "logger.hpp"

#include <fstream>
#include <iostream>

namespace {

// comment removal + whitespace removal in one pass over the line
std::string significantCharacters(const std::string& line)
{
	size_t cut = line.find('#');
	if (cut == std::string::npos) cut = line.find("//");
	std::string out;
	out.reserve(line.size());
	for (size_t i = 0; i < line.size() && i < cut; ++i)
		if (line[i] != ' ' && line[i] != '\t') out.push_back(line[i]);
	return out;
}

struct Statement
{
	std::string name;
	GlobalConfig::ParameterType values;
	bool assigns = false; // an '=' was present
};

Statement splitStatement(const std::string& text)
{
	Statement s;
	const size_t eq = text.find('=');
	s.name = text.substr(0, eq);
	if (eq == std::string::npos) return s;
	s.assigns = true;
	std::string field;
	for (size_t i = eq + 1; i <= text.size(); ++i)
	{
		if (i == text.size() || text[i] == ',')
		{
			s.values.emplace_back(field);
			field.clear();
		}
		else field.push_back(text[i]);
	}
	return s;
}

void printValues(const GlobalConfig::ParameterType& v)
{
	std::string text = "{";
	for (size_t i = 0; i < v.size(); ++i) text += v[i].As<std::string>() + (i + 1 < v.size() ? ", " : " }");
	std::cout << text << "\n";
}

} // namespace

ScriptProcessing::ScriptProcessing() : m_scriptWaitSeconds(0.0), m_scriptWaitIterations(0) {}

bool ScriptProcessing::RunScript(const std::string& _scriptFilename)
{
	std::ifstream in(_scriptFilename);
	if (!in.good())
	{
		LOG_ERROR("Script file \"" + _scriptFilename + "\" could not be opened");
		return false;
	}
	for (std::string line; std::getline(in, line);)
	{
		while (!line.empty() && (line.back() == '\r' || line.back() == '\n')) line.pop_back(); // scenes/*.txt use DOS line ends
		m_scriptCommandQueue.push(std::move(line));
	}
	return true;
}

void ScriptProcessing::PushCommand(const std::string& _line)
{
	if (_line.empty()) return;
	std::lock_guard<std::mutex> guard(m_consoleCommandQueueMutex);
	m_consoleCommandQueue.push(_line);
}

bool ScriptProcessing::IsIdle()
{
	std::lock_guard<std::mutex> guard(m_consoleCommandQueueMutex);
	const bool waiting = m_scriptWaitSeconds > 0.0 || m_scriptWaitIterations != 0;
	return !waiting && m_scriptCommandQueue.empty() && m_consoleCommandQueue.empty();
}

void ScriptProcessing::ParseCommand(std::string _commandLine, bool _fromScriptFile, bool& _stopProcessing)
{
	const std::string text = significantCharacters(_commandLine);
	if (text.empty()) return;
	const Statement st = splitStatement(text);

	// the two delay commands exist for script files only; either one suspends the queue even when its argument is malformed
	if (_fromScriptFile && (st.name == "waitSeconds" || st.name == "waitIterations"))
	{
		if (st.values.size() != 1) LOG_ERROR(st.name + " takes exactly one number");
		else if (st.name == "waitSeconds") m_scriptWaitSeconds = st.values[0].As<float>();
		else m_scriptWaitIterations = st.values[0].As<unsigned int>();
		_stopProcessing = true;
		return;
	}

	try
	{
		const GlobalConfig::ParameterType held = GlobalConfig::GetParameter(st.name); // throws for unknown names
		if (st.values.empty() && !held.empty()) printValues(held);
		else GlobalConfig::SetParameter(st.name, st.values);
	}
	catch (const std::invalid_argument& problem)
	{
		LOG_ERROR(problem.what());
	}
	_stopProcessing = m_processPauseCommand.count(st.name) != 0;
}

void ScriptProcessing::ProcessCommandQueue(double timeDelta, unsigned int iterationDelta)
{
	// pending delays run down first
	m_scriptWaitSeconds -= timeDelta;
	m_scriptWaitIterations = iterationDelta >= m_scriptWaitIterations ? 0u : m_scriptWaitIterations - iterationDelta;

	bool stop = false;
	auto scriptMayRun = [&] { return !stop && m_scriptWaitSeconds <= 0.0 && m_scriptWaitIterations == 0 && !m_scriptCommandQueue.empty(); };
	while (scriptMayRun())
	{
		const std::string line = std::move(m_scriptCommandQueue.front());
		m_scriptCommandQueue.pop();
		ParseCommand(line, true, stop);
	}
	std::lock_guard<std::mutex> guard(m_consoleCommandQueueMutex);
	while (!stop && !m_consoleCommandQueue.empty())
	{
		const std::string line = std::move(m_consoleCommandQueue.front());
		m_consoleCommandQueue.pop();
		ParseCommand(line, false, stop);
	}
}
