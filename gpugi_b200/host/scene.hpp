// scene.hpp -- host mirror of gpugi/scene/scene.hpp:17-204.  Same public surface; the GL buffers and
// bindless texture handles become host arrays in the layout of include/gpugi_b200.h (GiSceneDesc),
// which gi_upload_scene() copies to the device.
#pragma once

#include "../../include/gpugi_b200.h"
#include "bim.hpp"
#include "vec.hpp"

#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

class Scene
{
public:
	/// Load a scene: `_file` is the JSON environment file that names the .bim (scene.cpp:15-59).
	/// Errors are logged and leave an empty scene, like the reference (scene.cpp:31,42).
	Scene(const std::string& _file, ei::Types3D _bvhType);
	/// Adopt an in-memory model (synthetic scenes); textures are resolved relative to _sourceDirectory.
	Scene(std::unique_ptr<gi::bim::BinaryModel> _model, const std::string& _sourceDirectory, ei::Types3D _bvhType);
	~Scene();

	/// scene.hpp:91-105
	struct LightTriangle
	{
		ei::Triangle triangle;
		ei::Vec3 luminance;
	};
	struct PointLight
	{
		ei::Vec3 position;
		ei::Vec3 intensity;
	};

	bool IsValid() const { return m_valid; }
	const gi::bim::BinaryModel& GetModel() const { return *m_model; }
	/// Constant emissivity per material (zero for textured emissivity), scene.cpp:173-177,253-255.
	const std::vector<ei::Vec3>& GetConstantEmissivity() const { return m_emissivity; }

	// Buffers (replace GetVertexPositionBuffer/GetVertexInfoBuffer/GetTriangleBuffer/GetHierarchyBuffer, scene.hpp:107-110)
	const std::vector<float>& GetVertexPositions() const { return m_vertexPositions; }
	const std::vector<GiVertexInfo>& GetVertexInfo() const { return m_vertexInfo; }
	const std::vector<GiTriangle>& GetTriangles() const { return m_leafTriangles; }
	const std::vector<GiNode>& GetHierarchy() const { return m_hierarchy; }
	const std::vector<GiOBoxNode>& GetOBoxHierarchy() const { return m_oboxHierarchy; }
	const std::vector<GiMaterial>& GetMaterials() const { return m_materials; }
	/// Everything gi_upload_scene needs; pointers stay valid while the Scene lives.
	GiSceneDesc GetSceneDesc() const;

	const gi::bim::Node* GetHierarchyRAM() const { return m_model->chunk().getHierarchy(); }
	const uint32_t* GetHierarchyParentsRAM() const { return m_model->chunk().getHierarchyParents(); }
	const std::vector<PointLight>& GetPointLights() const { return m_pointLights; }

	ei::Types3D GetBoundingVolumeType() const { return ei::Types3D::BOX; }
	uint32_t GetNumTrianglesPerLeaf() const { return m_model->getNumTrianglesPerLeaf(); }
	uint32_t GetNumLeafTriangles() const { return m_model->chunk().getNumLeafNodes() * m_model->getNumTrianglesPerLeaf(); }
	uint32_t GetNumInnerNodes() const { return m_model->chunk().getNumNodes(); }
	uint32_t GetNumMaterials() const { return m_model->getNumUsedMaterials(); }
	uint32_t GetNumVertices() const { return m_model->chunk().getNumVertices(); }
	uint32_t GetNumLightTriangles() const { return static_cast<uint32_t>(m_lightTriangles.size()); }
	uint32_t GetNumTreeLevels() const { return m_model->chunk().getNumTreeLevels(); }
	const LightTriangle* GetLightTriangles() const { return m_lightTriangles.data(); }
	const float* GetLightSummedAreaTable() const { return m_lightSummedArea.data(); }
	const float* GetPointLightSummedFluxTable() const { return m_pointLightSummedFlux.data(); }
	float GetLightAreaSum() const { return m_lightAreaSum; }
	float GetTotalAreaLightFlux() const { return m_totalAreaLightFlux; }
	float GetTotalPointLightFlux() const { return m_totalPointLightFlux; }
	/// Changes whenever the light tables do (LoadLightSources, ComputePointLightTable): device copies are refreshed lazily.
	unsigned GetLightTablesVersion() const { return m_lightTablesVersion; }
	const std::vector<PointLight>& GetPointLightsRef() const { return m_pointLights; }
	ei::Box GetBoundingBox() const;

	void AddPointLight(const PointLight& _light) { m_pointLights.push_back(_light); ComputePointLightTable(); }
	/// Returns false if no light with the index exists.  (The reference returns before refreshing the
	/// table, scene.hpp:146; here the table is refreshed.)
	bool SetPointLight(size_t _index, const PointLight& _light);

	ei::Types3D GetBvhType() const { return m_bvhType; }
	const char* GetBvhTypeDefineString() const;

private:
	std::unique_ptr<gi::bim::BinaryModel> m_model;
	bool m_valid = false;

	std::vector<float> m_vertexPositions;
	std::vector<GiVertexInfo> m_vertexInfo;
	std::vector<GiTriangle> m_leafTriangles;
	std::vector<GiNode> m_hierarchy;
	std::vector<GiOBoxNode> m_oboxHierarchy; // bvhType = 1: TreeNode<ei::OBox> (scene.hpp:40-49), else empty

	unsigned m_lightTablesVersion = 0;
	std::vector<LightTriangle> m_lightTriangles;
	std::vector<PointLight> m_pointLights;
	std::vector<float> m_lightSummedArea;
	std::vector<float> m_pointLightSummedFlux;
	std::vector<GiMaterial> m_materials;
	std::vector<ei::Vec3> m_emissivity; ///< constant emissivity per material, used to find light triangles
	float m_lightAreaSum = 0.0f;
	float m_totalAreaLightFlux = 0.0f, m_totalPointLightFlux = 0.0f;

	// "Bindless handles" become indices into m_textures; payloads are owned here.
	struct TextureStorage { std::vector<unsigned char> bytes; };
	std::vector<GiTexture> m_textures;
	std::vector<std::unique_ptr<TextureStorage>> m_textureStorage;
	std::unordered_map<std::string, int32_t> m_textureIndex;

	std::string m_sourceDirectory;
	ei::Types3D m_bvhType;

	void Init();
	void UploadGeometry();
	void UploadHierarchy(ei::Types3D _bvhType);
	void LoadLightSources();
	void ComputePointLightTable();
	void LoadMaterial(const gi::bim::Material& _material);
	int32_t GetBindlessHandle(const std::string& _name);
	int32_t GetBindlessHandle(const ei::Vec3& _data, bool _hdr);
	int32_t GetBindlessHandle(const ei::Vec4& _data, bool _hdr);
	int32_t AddTexture(const std::string& key, int w, int h, int format, std::vector<unsigned char>&& bytes);
};
