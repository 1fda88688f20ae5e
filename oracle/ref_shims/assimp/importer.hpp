// Stand-in for <assimp/importer.hpp> (a binary-only Windows dependency of the reference's bvhmake tool).  TEST INFRASTRUCTURE ONLY.
// bvhmake/bvhmake.hpp includes it for the importer that feeds BVHBuilder; the build methods compiled into oracle/_ref/libref_bvhmake.so
// (buildmethods/kdtree.cpp, sweep.cpp, fitmethods/aaboxfit.cpp) never touch it.
#pragma once
struct aiScene;
struct aiNode;
