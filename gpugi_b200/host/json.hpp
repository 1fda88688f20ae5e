// json.hpp -- minimal JSON reader for the scene environment files (scenes/*/*.json of the reference:
// objects, arrays, strings, numbers, true/false/null).  Objects keep insertion order.
#pragma once

#include <cstdlib>
#include <string>
#include <utility>
#include <vector>

namespace gi {
namespace json {

struct Value
{
	enum Type { NUL, BOOL, NUMBER, STRING, ARRAY, OBJECT } type = NUL;
	bool boolean = false;
	double number = 0.0;
	std::string str;
	std::vector<Value> array;
	std::vector<std::pair<std::string, Value>> object;

	const Value* find(const std::string& key) const
	{
		for (const auto& kv : object)
			if (kv.first == key) return &kv.second;
		return nullptr;
	}
};

namespace detail {
struct Parser
{
	const std::string& s;
	size_t p = 0;
	std::string err;

	explicit Parser(const std::string& text) : s(text) {}

	void skip()
	{
		for (;;)
		{
			while (p < s.size() && (s[p] == ' ' || s[p] == '\t' || s[p] == '\n' || s[p] == '\r')) ++p;
			// tolerate a UTF-8 BOM and // comments (some scene files carry them)
			if (p + 2 < s.size() && (unsigned char)s[p] == 0xEF && (unsigned char)s[p + 1] == 0xBB && (unsigned char)s[p + 2] == 0xBF) { p += 3; continue; }
			if (p + 1 < s.size() && s[p] == '/' && s[p + 1] == '/') { while (p < s.size() && s[p] != '\n') ++p; continue; }
			break;
		}
	}
	bool fail(const char* msg) { if (err.empty()) err = std::string(msg) + " at offset " + std::to_string(p); return false; }

	bool parseString(std::string& out)
	{
		if (s[p] != '"') return fail("expected string");
		++p;
		out.clear();
		while (p < s.size() && s[p] != '"')
		{
			if (s[p] == '\\' && p + 1 < s.size())
			{
				++p;
				switch (s[p])
				{
				case 'n': out += '\n'; break;
				case 't': out += '\t'; break;
				case 'r': out += '\r'; break;
				case 'b': out += '\b'; break;
				case 'f': out += '\f'; break;
				case 'u': p += 4; out += '?'; break;
				default: out += s[p];
				}
				++p;
			}
			else out += s[p++];
		}
		if (p >= s.size()) return fail("unterminated string");
		++p;
		return true;
	}

	// nesting guard: the parser recurses per level, scene files nest 3-4 deep
	int depth = 0;
	struct Level { int& d; explicit Level(int& depth) : d(depth) { ++d; } ~Level() { --d; } };

	bool parseValue(Value& v)
	{
		Level level(depth);
		if (depth > 200) return fail("nested deeper than 200 levels");
		skip();
		if (p >= s.size()) return fail("unexpected end");
		char c = s[p];
		if (c == '{')
		{
			v.type = Value::OBJECT;
			++p; skip();
			if (p < s.size() && s[p] == '}') { ++p; return true; }
			for (;;)
			{
				skip();
				std::string key;
				if (p >= s.size() || !parseString(key)) return fail("expected key");
				skip();
				if (p >= s.size() || s[p] != ':') return fail("expected ':'");
				++p;
				Value child;
				if (!parseValue(child)) return false;
				v.object.emplace_back(std::move(key), std::move(child));
				skip();
				if (p < s.size() && s[p] == ',') { ++p; continue; }
				if (p < s.size() && s[p] == '}') { ++p; return true; }
				return fail("expected ',' or '}'");
			}
		}
		if (c == '[')
		{
			v.type = Value::ARRAY;
			++p; skip();
			if (p < s.size() && s[p] == ']') { ++p; return true; }
			for (;;)
			{
				Value child;
				if (!parseValue(child)) return false;
				v.array.push_back(std::move(child));
				skip();
				if (p < s.size() && s[p] == ',') { ++p; continue; }
				if (p < s.size() && s[p] == ']') { ++p; return true; }
				return fail("expected ',' or ']'");
			}
		}
		if (c == '"') { v.type = Value::STRING; return parseString(v.str); }
		if (s.compare(p, 4, "true") == 0) { v.type = Value::BOOL; v.boolean = true; p += 4; return true; }
		if (s.compare(p, 5, "false") == 0) { v.type = Value::BOOL; v.boolean = false; p += 5; return true; }
		if (s.compare(p, 4, "null") == 0) { v.type = Value::NUL; p += 4; return true; }
		const char* start = s.c_str() + p;
		char* end = nullptr;
		double d = std::strtod(start, &end);
		if (end == start) return fail("unexpected character");
		v.type = Value::NUMBER;
		v.number = d;
		p += size_t(end - start);
		return true;
	}
};
} // namespace detail

inline bool parse(const std::string& text, Value& out, std::string* error = nullptr)
{
	detail::Parser parser(text);
	bool ok = parser.parseValue(out);
	if (!ok && error) *error = parser.err;
	return ok;
}

} // namespace json
} // namespace gi
