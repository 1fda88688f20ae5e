// lightsampler.cpp -- host mirror of gpugi/scene/lightsampler.cpp:8-95: N "initial light samples"
// per iteration, split between area and point lights by flux, drawn with the host xorshift64*.
#include "lightsampler.hpp"
#include "random.hpp"
#include "scene.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

LightSampler::LightSampler() : m_randomSeed(123) {}

void LightSampler::SetScene(std::shared_ptr<const Scene> _scene) { m_scene = _scene; }

void LightSampler::GenerateRandomSamples(LightSample* dst, unsigned int _numSamples, float _positionBias)
{
	const float areaFlux = m_scene->GetTotalAreaLightFlux();
	const float pointFlux = m_scene->GetTotalPointLightFlux();
	// no uniform light source at all: black dummy lights (lightsampler.cpp:21-31)
	if (areaFlux + pointFlux == 0.0f)
	{
		std::memset(dst, 0, sizeof(LightSample) * _numSamples);
		return;
	}
	float pointLightPercentage = pointFlux / (areaFlux + pointFlux);
	unsigned numPointSamples = unsigned(pointLightPercentage * _numSamples);
	unsigned numAreaSamples = _numSamples - numPointSamples;
	float sampleWeight = m_scene->GetLightAreaSum() / numAreaSamples;

	const float* areaTable = m_scene->GetLightSummedAreaTable();
	for (unsigned s = 0; s < numAreaSamples; ++s, ++dst)
	{
		float pick = XorshiftF(m_randomSeed);
		unsigned triangleIdx = unsigned(std::lower_bound(areaTable, areaTable + m_scene->GetNumLightTriangles(), pick) - areaTable);
		// uniform point on the triangle (sqrt warp)
		float xi0 = XorshiftF(m_randomSeed);
		float xi1 = XorshiftF(m_randomSeed);
		xi0 = sqrt(xi0);
		float alpha = (1.0f - xi0);
		float beta = xi0 * (1.0f - xi1);
		float gamma = xi0 * xi1;
		const Scene::LightTriangle& lt = m_scene->GetLightTriangles()[triangleIdx];
		ei::Vec3 position = lt.triangle.v0 * alpha + lt.triangle.v1 * beta + lt.triangle.v2 * gamma;
		ei::Vec3 normal = ei::normalize(ei::cross(lt.triangle.v0 - lt.triangle.v1, lt.triangle.v0 - lt.triangle.v2));
		position += normal * (_positionBias);
		dst->position[0] = position.x; dst->position[1] = position.y; dst->position[2] = position.z;
		dst->normal_phi = atan2(normal.y, normal.x);
		ei::Vec3 intensity = lt.luminance * sampleWeight;
		dst->intensity[0] = intensity.x; dst->intensity[1] = intensity.y; dst->intensity[2] = intensity.z;
		dst->normal_theta_cos = normal.z;
	}

	const float* fluxTable = m_scene->GetPointLightSummedFluxTable();
	const auto& lights = m_scene->GetPointLights();
	for (unsigned s = 0; s < numPointSamples; ++s, ++dst)
	{
		float pick = XorshiftF(m_randomSeed);
		const float* entry = std::lower_bound(fluxTable, fluxTable + lights.size(), pick);
		unsigned lightIdx = unsigned(entry - fluxTable);
		sampleWeight = *entry;
		if (lightIdx > 0) sampleWeight -= fluxTable[lightIdx - 1];
		dst->position[0] = lights[lightIdx].position.x; dst->position[1] = lights[lightIdx].position.y; dst->position[2] = lights[lightIdx].position.z;
		dst->normal_phi = 0.0;
		dst->normal_theta_cos = 2.0f; // marks an omnidirectional light for the shaders
		ei::Vec3 intensity = lights[lightIdx].intensity / (numPointSamples * sampleWeight);
		dst->intensity[0] = intensity.x; dst->intensity[1] = intensity.y; dst->intensity[2] = intensity.z;
	}
}

bool LightSampler::GenerateRandomSamplesOnDevice(GiContext* _device, unsigned int _numSamples, float _positionBias)
{
	if (!m_scene || !_device) return false;
	if (m_tablesDevice != _device || m_tablesScene != m_scene.get() || m_tablesVersion != m_scene->GetLightTablesVersion())
	{
		static_assert(sizeof(Scene::LightTriangle) == sizeof(GiLightTriangle), "LightTriangle layout");
		std::vector<float> pos, inten;
		for (const Scene::PointLight& l : m_scene->GetPointLightsRef())
		{
			pos.push_back(l.position.x); pos.push_back(l.position.y); pos.push_back(l.position.z);
			inten.push_back(l.intensity.x); inten.push_back(l.intensity.y); inten.push_back(l.intensity.z);
		}
		GiLightTables t;
		std::memset(&t, 0, sizeof(t));
		t.light_triangles = reinterpret_cast<const GiLightTriangle*>(m_scene->GetLightTriangles());
		t.summed_area = m_scene->GetLightSummedAreaTable();
		t.num_light_triangles = m_scene->GetNumLightTriangles();
		t.light_area_sum = m_scene->GetLightAreaSum();
		t.point_positions = pos.data(); t.point_intensities = inten.data();
		t.point_summed_flux = m_scene->GetPointLightSummedFluxTable();
		t.num_point_lights = uint32_t(m_scene->GetPointLightsRef().size());
		t.total_area_flux = m_scene->GetTotalAreaLightFlux();
		t.total_point_flux = m_scene->GetTotalPointLightFlux();
		if (gi_set_light_tables(_device, &t) != GI_OK) return false;
		m_tablesDevice = _device; m_tablesScene = m_scene.get(); m_tablesVersion = m_scene->GetLightTablesVersion();
	}
	uint64_t state = m_randomSeed;
	if (gi_generate_light_samples(_device, &state, _numSamples, _positionBias) != GI_OK) return false;
	m_randomSeed = state;
	return true;
}
