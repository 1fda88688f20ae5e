// logger.hpp -- stand-in for gpugi/utilities/logger.hpp: the reference reports errors with LOG_ERROR
// and carries on (scene.cpp:31,42); messages go to stderr and are kept for the C API.
#pragma once
#include <iostream>
#include <sstream>
#include <string>

namespace gi { std::string& LastHostError(); }

#define LOG_ERROR(msg) do { std::ostringstream gi_os_; gi_os_ << msg; gi::LastHostError() = gi_os_.str(); std::cerr << "[gpugi_b200 error] " << gi_os_.str() << std::endl; } while (0)
#define LOG_LVL2(msg) do { if (std::getenv("GPUGI_VERBOSE")) std::cerr << "[gpugi_b200] " << msg << std::endl; } while (0)
