// globalconfig.hpp -- host mirror of gpugi/control/globalconfig.hpp:8-75 / globalconfig.cpp:12-87: named parameters holding a short
// list of float/int/bool/string variants, with listeners that fire on every SetParameter ("misuse as event system is
// appreciated": an empty value list is a command).  The script / console front end (scriptprocessing.hpp) drives it.
#pragma once
#include <cstdlib>
#include <functional>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

/// utilities/variant.hpp: one of float / int / bool / string, convertible on read.  Script arguments arrive as strings.
class Variant
{
public:
	Variant() : m_kind(INT), m_i(0), m_f(0.0f) {}
	Variant(int v) : m_kind(INT), m_i(v), m_f(float(v)) {}
	Variant(unsigned int v) : m_kind(INT), m_i(int(v)), m_f(float(v)) {}
	Variant(float v) : m_kind(FLOAT), m_i(int(v)), m_f(v) {}
	Variant(double v) : m_kind(FLOAT), m_i(int(v)), m_f(float(v)) {}
	Variant(bool v) : m_kind(BOOL), m_i(v ? 1 : 0), m_f(v ? 1.0f : 0.0f) {}
	Variant(const std::string& v) : m_kind(STRING), m_i(0), m_f(0.0f), m_s(v) {}
	Variant(const char* v) : m_kind(STRING), m_i(0), m_f(0.0f), m_s(v) {}
	template <typename T> T As() const;

private:
	enum Kind { INT, FLOAT, BOOL, STRING } m_kind;
	int m_i;
	float m_f;
	std::string m_s;
};
template <> inline float Variant::As<float>() const { return m_kind == STRING ? std::strtof(m_s.c_str(), nullptr) : (m_kind == FLOAT ? m_f : float(m_i)); }
template <> inline int Variant::As<int>() const
{
	if (m_kind == STRING) { char* end = nullptr; long v = std::strtol(m_s.c_str(), &end, 10); return (end && *end == '\0') ? int(v) : int(std::strtof(m_s.c_str(), nullptr)); }
	return m_kind == FLOAT ? int(m_f) : m_i;
}
template <> inline unsigned int Variant::As<unsigned int>() const { return static_cast<unsigned int>(As<int>()); }
template <> inline bool Variant::As<bool>() const { return m_kind == STRING ? (m_s == "true" || m_s == "1" || m_s == "True") : m_i != 0; }
template <> inline std::string Variant::As<std::string>() const
{
	if (m_kind == STRING) return m_s;
	std::ostringstream os;
	if (m_kind == FLOAT) os << m_f; else if (m_kind == BOOL) os << (m_i ? "true" : "false"); else os << m_i;
	return os.str();
}

namespace GlobalConfig
{
	typedef Variant ParameterElementType;
	typedef std::vector<Variant> ParameterType;
	typedef std::function<void(const ParameterType&)> ListenerFunc;

	struct Entry
	{
		ParameterType value;
		std::string description;
		std::map<std::string, ListenerFunc> listeners;
	};

	inline std::map<std::string, Entry>& Registry()
	{
		static std::map<std::string, Entry> params = {{"pathLength", Entry{{Variant(16)}, "Change the maximum path length of almost any renderer.", {}}}}; // application.cpp:123
		return params;
	}
	/// \throws std::invalid_argument if the parameter already exists (globalconfig.cpp:20-27)
	inline void AddParameter(const std::string& name, const ParameterType& value, const std::string& description = "<no desc>")
	{
		if (Registry().count(name)) throw std::invalid_argument("Config parameter with name \"" + name + "\" already exists!");
		Registry()[name] = Entry{value, description, {}};
	}
	inline bool HasParameter(const std::string& name) { return Registry().count(name) != 0; }
	inline void RemoveParameter(const std::string& name) { Registry().erase(name); }
	inline void AddListener(const std::string& name, const std::string& listenerName, const ListenerFunc& f)
	{
		auto it = Registry().find(name);
		if (it == Registry().end()) throw std::invalid_argument("Config parameter with name \"" + name + "\" does not exist!");
		if (it->second.listeners.count(listenerName)) throw std::invalid_argument("Listener \"" + listenerName + "\" already exists for \"" + name + "\"!");
		it->second.listeners[listenerName] = f;
	}
	inline void RemovesListener(const std::string& name, const std::string& listenerName)
	{
		auto it = Registry().find(name);
		if (it == Registry().end() || !it->second.listeners.erase(listenerName)) throw std::invalid_argument("Config parameter or listener does not exist!");
	}
	inline ParameterType GetParameter(const std::string& name)
	{
		auto it = Registry().find(name);
		if (it == Registry().end()) throw std::invalid_argument("Config parameter with name \"" + name + "\" does not exist!");
		return it->second.value;
	}
	/// Sets the value and calls all listeners; the number of values must match the parameter's (globalconfig.cpp:66-78).
	inline void SetParameter(const std::string& name, const ParameterType& value)
	{
		auto it = Registry().find(name);
		if (it == Registry().end()) throw std::invalid_argument("Config parameter with name \"" + name + "\" does not exist!");
		if (it->second.value.size() != value.size())
			throw std::invalid_argument("Config parameter \"" + name + "\" expects " + std::to_string(it->second.value.size()) + " value(s)!");
		it->second.value = value;
		auto listeners = it->second.listeners; // a listener may add / remove parameters
		for (auto& l : listeners) l.second(value);
	}
	inline std::string GetEntryDescriptions()
	{
		std::string out;
		for (auto& e : Registry())
		{
			out += e.first + " (" + std::to_string(e.second.value.size()) + " value(s)): " + e.second.description + "\n";
		}
		return out;
	}
}
