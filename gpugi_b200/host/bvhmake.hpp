// bvhmake.hpp -- the reference's legacy scene pipeline: the `.rawscene` container of bvhmake (bvhmake/filedef.hpp:11-89, written by
// bvhmake/bvhbuilder.cpp:174-211,569-640 in the order materialref, vertices, bounding_aabox, hierarchy, triangles[, approx_sggx]) and its three
// hierarchy builders over 8-triangle leaves (bvhmake/buildmethods/{kdtree,sweep,lds}.cpp with the axis-aligned box fit of fitmethods/aaboxfit.cpp).
//
// Written from the format and the algorithms, not from the reference's text.  The current gpugi reads `.bim` only (scene.cpp:15-59; its rawscene
// loaders are commented out, scene.hpp:182-184) and its shaders hard-code two triangles per leaf (scenedata.glsl:5), so a raw scene enters the
// renderer through ToModel(): same vertices, same triangles, the file's hierarchy with every 8-triangle leaf expanded into a small subtree of
// 2-triangle leaves.
#pragma once

#include "bim.hpp"

#include <cstdint>
#include <string>
#include <vector>

namespace gi {
namespace bvhmake {

// filedef.hpp:20-25 / :28-40 / :53-58 / :66-71 / :74-79
struct Vertex { bim::Vec3 position; bim::Vec3 normal; bim::Vec2 texcoord; };  // 32 bytes
struct Triangle { uint32_t vertices[3]; uint32_t material; };                   // 16 bytes; all-zero = unused slot
struct MaterialRef { char material[32]; };
struct Node { uint32_t parent, firstChild, escape; };                           // firstChild: MSB = leaf, low bits = leaf index
enum : uint32_t { NUM_PRIMITIVES = 8 };
struct Leaf { Triangle triangles[NUM_PRIMITIVES]; };

struct RawScene
{
	std::vector<MaterialRef> materials;   // "materialref"
	std::vector<Vertex> vertices;         // "vertices"
	std::vector<bim::Box> boxes;          // "bounding_aabox", one per node
	std::vector<Node> nodes;              // "hierarchy", preorder (RecursiveWriteHierarchy, bvhbuilder.cpp:706-718)
	std::vector<Leaf> leaves;             // "triangles": the leaf array
	std::vector<std::string> skipped;     // names of arrays the reader does not interpret (approx_sggx, bounding_sphere, ...)
};

enum class BuildMethod { KDTREE, SWEEP, LDS };

/// Sequence of {char name[32]; u32 numElements; u32 elementSize} headers, each followed by its data (filedef.hpp:11-16).
bool LoadRawScene(const std::string& file, RawScene& out, std::string* error = nullptr);
bool StoreRawScene(const std::string& file, const RawScene& scene, std::string* error = nullptr);

/// BVHBuilder::BuildBVH (bvhbuilder.cpp:526-555) with the chosen build method and the aabox fit: fills nodes / boxes / leaves of `scene` from its
/// vertices and `triangles` (the importer's triangle list, in import order).
void BuildHierarchy(RawScene& scene, const std::vector<Triangle>& triangles, BuildMethod method);

/// The raw scene as a bim model the Scene class takes: vertex arrays, the triangle list in leaf order, and the file's hierarchy with every leaf
/// of up to 8 triangles replaced by a balanced subtree of 2-triangle leaves (boxes of the new nodes fitted bottom-up; inner boxes as stored).
/// Materials: TRIANGLE_MAT index i = materialref entry i; `envJson` supplies their definitions (the .json bvhmake writes next to the scene).
bool ToModel(const RawScene& scene, const std::string& envJson, bim::BinaryModel& model, std::string* error = nullptr);

} // namespace bvhmake
} // namespace gi
