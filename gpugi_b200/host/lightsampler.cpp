// lightsampler.cpp -- host mirror of gpugi/scene/lightsampler.cpp:8-95: N "initial light samples"
// per iteration, split between area and point lights by flux, drawn with the host xorshift64*.
#include "lightsampler.hpp"
#include "random.hpp"
#include "scene.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>

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
