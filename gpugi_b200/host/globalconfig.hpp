// globalconfig.hpp -- the sliver of gpugi/control/globalconfig.hpp the in-scope renderers use:
// named parameters holding a short list of int/float/bool/string variants (globalconfig.cpp:12-87).
// The script/console front end and the listeners are out of scope.
#pragma once
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

class Variant
{
public:
	Variant() : m_kind(INT), m_i(0), m_f(0.0f) {}
	Variant(int v) : m_kind(INT), m_i(v), m_f(float(v)) {}
	Variant(float v) : m_kind(FLOAT), m_i(int(v)), m_f(v) {}
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
template <> inline int Variant::As<int>() const { return m_kind == FLOAT ? int(m_f) : m_i; }
template <> inline float Variant::As<float>() const { return m_kind == FLOAT ? m_f : float(m_i); }
template <> inline bool Variant::As<bool>() const { return m_i != 0; }
template <> inline std::string Variant::As<std::string>() const { return m_s; }

namespace GlobalConfig
{
	typedef std::vector<Variant> ParameterType;

	inline std::map<std::string, ParameterType>& Registry()
	{
		static std::map<std::string, ParameterType> params = {{"pathLength", {Variant(16)}}}; // application.cpp:123
		return params;
	}
	inline void AddParameter(const std::string& name, const ParameterType& value, const std::string& = "") { Registry()[name] = value; }
	inline void SetParameter(const std::string& name, const ParameterType& value) { Registry()[name] = value; }
	inline ParameterType GetParameter(const std::string& name)
	{
		auto it = Registry().find(name);
		if (it == Registry().end()) throw std::invalid_argument("Config parameter with name \"" + name + "\" does not exist!");
		return it->second;
	}
	inline void RemoveParameter(const std::string& name) { Registry().erase(name); }
}
