// Stand-in for <jofilelib.hpp> (binary-only dependency; bvhmake.hpp holds one Jo::Files::MetaFileWrapper member for the material file).  TEST INFRASTRUCTURE ONLY.
#pragma once
namespace Jo { namespace Files { class MetaFileWrapper {}; } }
