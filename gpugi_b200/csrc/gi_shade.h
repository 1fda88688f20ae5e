// gi_shade.h -- per-lane shading steps of the wavefront light tracer and bidirectional path tracer (gi_wavefront.cuh).
// One call = one path vertex: what the body of the bounce loop of lighttracer.comp:93-192 / pathtracer_bidir.comp:42-244
// does between two TraceRay calls, with the any-hit queries (camera connection, light-cache connection) returned as
// shadow-ray records instead of being traced in place.  Plain per-thread code (GI_HD) so that tests/emul can run exactly
// this logic on the CPU against the oracle (test infrastructure only -- the product has no CPU path).
#pragma once

#include "gi_core.h"

namespace gik {

// light pass variants (lighttracer.comp #defines): SAVE_LIGHT_CACHE, SAVE_LIGHT_CACHE_WARMUP
enum { WF_LIGHT_SAVE = 1u, WF_LIGHT_WARMUP = 2u };

// what a shaded vertex hands to the scheduler
struct WfPush
{
	bool ray;            // the path continues: (origin, dir) is the next extension ray
	bool shadow;         // an any-hit query is pending: (sOrigin, sDir, sLen) -> adds sValue to its target when the ray is free
	F3 origin, dir;
	F3 sOrigin, sDir;
	float sLen;
	F3 sValue;
	int px, py;          // light tracer: target pixel of the splat (already bounds-checked; with a reconstruction filter: the centre pixel, which may lie up to 2 pixels outside)
	float posX, posY;    // ... and the unrounded position the filter weights refer to
};

// Path state records (HBM): T = (pathThroughput.xyz, RNG state bits), C = (radiance sum.xyz, alpha increment),
// M = lastConnectionProbability of the BPT eye path.

// lighttracer.comp:49-91: start of light path `gid`.  Writes the light-origin cache entry in SAVE mode (:56-61).
GI_HD void wfGenerateLightLane(const DFrame& f, uint32_t lightFlags, uint32_t gid, F3& origin, F3& dir, F4& stateT, PathCounters& pc)
{
	const bool save = (lightFlags & WF_LIGHT_SAVE) != 0u, warmup = (lightFlags & WF_LIGHT_WARMUP) != 0u;
	uint32_t seed = initRandomSeed(f.frameSeed, gid);
	int lightSampleIndex = int(gid / uint32_t(f.numRaysPerLightSample));
	F4 l0 = ld4(f.lightSamples + 2 * lightSampleIndex), l1 = ld4(f.lightSamples + 2 * lightSampleIndex + 1);
	F3 intensity = f3(l1.x, l1.y, l1.z);
	++pc.lightPaths;
	if (save && !warmup)
	{
		F2 pn; pn.x = l0.w; pn.y = l1.w;
		F2 z; z.x = 0.0f; z.y = 0.0f;
		saveLightCache(f, gid, intensity / GI_PI_2 / float(f.numRaysPerLightSample), -1, f3(l0.x, l0.y, l0.z), pn, z, f3(0.0f), 0);
	}
	F3 pathThroughput = intensity * (GI_PI / float(f.numRaysPerLightSample));
	origin = f3(l0.x, l0.y, l0.z);
	if (l1.w > 1.5f)
	{
		dir = sampleDirection(random2(seed));
		pathThroughput = pathThroughput * 4.0f;
	}
	else
	{
		F3 normal = unpackNormal(l0.w, l1.w);
		F3 U, V;
		createONB(normal, U, V);
		dir = sampleHemisphereCosine(random2(seed), U, V, normal);
	}
	stateT.x = pathThroughput.x; stateT.y = pathThroughput.y; stateT.z = pathThroughput.z; stateT.w = i2f(int(seed));
}

// winning triangle of a closest-hit record: barycentrics + geometric normal as TraceRay returns them (traceClosest)
GI_HD void wfHitGeometry(const DScene& s, F3 o, F3 d, int leafTri, F3& bary, F3& gn, I4& tri)
{
	F3 p0, p1, p2; int tv;
	loadLeafTri(s, leafTri, p0, p1, p2, tv);
	float tt;
	intersectTriangle(o, d, p0, p1, p2, tt, bary, gn);
	tri = ld4(s.leafTris + leafTri);
}

// lighttracer.comp:93-192, loop body `bounce` of one light path after its TraceRay returned (tHit, leafTri).
GI_HD void wfShadeLightLane(const DFrame& f, uint32_t lightFlags, F3 origin, F3 dir, int bounce, float tHit, int leafTri, F4& stateT, WfPush& out, PathCounters& pc)
{
	out.ray = false; out.shadow = false;
	if (leafTri < 0) return; // the path left the scene (:97-98)
	const bool save = (lightFlags & WF_LIGHT_SAVE) != 0u, warmup = (lightFlags & WF_LIGHT_WARMUP) != 0u;
	F3 pathThroughput = f3(stateT.x, stateT.y, stateT.z);
	uint32_t seed = uint32_t(f2i(stateT.w));
	F3 bary, gn; I4 tri;
	wfHitGeometry(f.scene, origin, dir, leafTri, bary, gn, tri);
	origin = origin + tHit * dir;
	F3 shadingNormal; F2 tc;
	getTriangleHitInfo(f.scene, tri, bary, shadingNormal, tc);
	MaterialData md = sampleMaterialData(f.scene, tri.w, tc);
	++pc.shadedVertices;
	if (save && md.Reflectiveness.w < 100.0f)
	{
		uint32_t slot = allocLightCacheSlot(f, warmup);
		if (!warmup) saveLightCache(f, slot, pathThroughput, tri.w, origin, packNormal(shadingNormal), tc, dir, bounce + 1);
	}
	if (!warmup)
	{
		ScreenSplatQuery q;
		// (a filtered splat whose centre lies just outside the frame still reaches pixels inside)
		const int border = f.reconFilter ? 2 : 0;
		if (prepareScreenSplat(f, origin, dir, gn, shadingNormal, md, pathThroughput, save, q) && q.px >= -border && q.py >= -border && q.px < f.width + border &&
		    q.py < f.height + border)
		{
			++pc.anyRays;
			out.shadow = true;
			out.sOrigin = q.origin; out.sDir = q.dir; out.sLen = q.len; out.sValue = q.value;
			out.px = q.px; out.py = q.py; out.posX = q.posX; out.posY = q.posY;
		}
	}
	F3 throughput = f3(1.0f);
	float samplePDF;
	F3 newDir = sampleBSDF(dir, md, seed, shadingNormal, throughput, true, samplePDF);
	throughput = throughput * adjointShadingNormalCorrection(dir, newDir, gn, shadingNormal);
	pathThroughput = pathThroughput * throughput;
	float cont = saturate(luminance(throughput));
	if (randomFloat(seed) >= cont) return;
	if (bounce + 1 >= f.maxPathLength) return;
	pathThroughput = pathThroughput / cont;
	out.ray = true;
	++pc.closestRays;
	out.dir = newDir;
	out.origin = origin + newDir * GI_RAY_HIT_EPSILON;
	stateT.x = pathThroughput.x; stateT.y = pathThroughput.y; stateT.z = pathThroughput.z; stateT.w = i2f(int(seed));
}

// pathtracer_bidir.comp:42-244, loop body `bounce` of the eye path of pixel `pixel` after its TraceRay returned.
// numEntries = NumLightCacheEntries after the light pass.  stateC is always written (radiance sum so far, alpha += 1).
GI_HD void wfShadeEyeLane(const DFrame& f, uint32_t numEntries, F3 origin, F3 dir, int bounce, uint32_t pixel, float tHit, int leafTri, F4& stateT, F4& stateC,
                          float& stateM, WfPush& out, PathCounters& pc)
{
	out.ray = false; out.shadow = false;
	F3 color = f3(0.0f), pathThroughput = f3(1.0f);
	float lastConnectionProbability = 1.0f; // InitialMIS()
	uint32_t seed;
	if (bounce == 0)
	{
		seed = initRandomSeed(f.frameSeed, pixel);
		random2(seed); // the two draws of the sub-pixel jitter (pathtracer_bidir.comp:23 via cameraRayDirection)
	}
	else
	{
		pathThroughput = f3(stateT.x, stateT.y, stateT.z); seed = uint32_t(f2i(stateT.w));
		color = f3(stateC.x, stateC.y, stateC.z);
		lastConnectionProbability = stateM;
	}
	if (leafTri >= 0)
	{
		const uint32_t actualNumEntries = numEntries < f.lightCacheCapacity ? numEntries : f.lightCacheCapacity;
		const float lightSampleWeight = float(actualNumEntries);
		const int i = bounce;
		const float rayHit = tHit;
		F3 bary, gn; I4 tri;
		wfHitGeometry(f.scene, origin, dir, leafTri, bary, gn, tri);
		origin = origin + rayHit * dir;
		F3 hitNormal; F2 tc;
		getTriangleHitInfo(f.scene, tri, bary, hitNormal, tc);
		MaterialData md = sampleMaterialData(f.scene, tri.w, tc);
		++pc.shadedVertices;
		float mis;
		{
			float incidentCosSat = saturate(-dot(hitNormal, dir));
			float lightToEye = distanceMISHeuristic(rayHit) * misHeuristic(incidentCosSat / GI_PI);
			mis = (i == 0) ? 1.0f : (lastConnectionProbability / fmaxf(lastConnectionProbability + lightToEye, GI_DIVISOR_EPSILON));
			color = color + mis * pathThroughput * md.Emissivity;
		}
		{
			ConnectionQuery cq;
			if (prepareConnection(f, actualNumEntries, lightSampleWeight, origin, dir, hitNormal, md, i, rayHit, lastConnectionProbability, pathThroughput, seed, cq))
			{
				++pc.anyRays;
				out.shadow = true;
				out.sOrigin = cq.origin; out.sDir = cq.dir; out.sLen = cq.len; out.sValue = cq.value;
			}
		}
		float samplePDF;
		F3 throughput = f3(1.0f);
		dir = sampleBSDF(dir, md, seed, hitNormal, throughput, false, samplePDF);
		origin = origin + dir * GI_RAY_HIT_EPSILON;
		pathThroughput = pathThroughput * throughput;
		float cont = saturate(luminance(throughput));
		if (!(randomFloat(seed) >= cont) && bounce + 1 < f.maxPathLength)
		{
			pathThroughput = pathThroughput / cont;
			float reverseProb = misHeuristic(samplePDF);
			if (i == 0)
				mis = reverseProb / fmaxf(reverseProb + lastConnectionProbability, GI_DIVISOR_EPSILON);
			else
			{
				float distMIS = distanceMISHeuristic(rayHit);
				reverseProb *= distMIS;
				mis = lastConnectionProbability / fmaxf(lastConnectionProbability + distMIS * reverseProb, GI_DIVISOR_EPSILON);
				if (md.Reflectiveness.w >= 100.0f) mis = 1.0f;
			}
			pathThroughput = pathThroughput * mis;
			out.ray = true;
			++pc.closestRays;
			out.origin = origin; out.dir = dir;
			stateT.x = pathThroughput.x; stateT.y = pathThroughput.y; stateT.z = pathThroughput.z; stateT.w = i2f(int(seed));
			stateM = reverseProb;
		}
	}
	stateC.x = color.x; stateC.y = color.y; stateC.z = color.z; stateC.w = 1.0f; // Out[px] += vec4(color, 1) (pathtracer_bidir.comp:248-250)
}

} // namespace gik
