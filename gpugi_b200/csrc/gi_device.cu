// gi_device.cu -- CUDA kernels (sm_100a) and the C ABI of include/gpugi_b200.h.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
// There is no CPU path in this file: every compute entry point needs a live sm_100 device.
#include "gi_core.h"
#include "gi_layout.h"
#include "gi_wavefront.cuh"
#include "gi_lbvh.cuh"
#include "gi_comm.cuh"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <chrono>
#include <vector>

using namespace gik;

// =====================================================================================================
// kernels
// =====================================================================================================

// 64-bit device counters (GiCounters order)
enum { CNT_CLOSEST = 0, CNT_ANY, CNT_CAMERA, CNT_LIGHT, CNT_SHADED, CNT_SPLATS, CNT_FALLBACK, CNT_NODES, CNT_TRIS, CNT_INNER_STEPS, CNT_LEAF_STEPS, CNT_COUNT };
static_assert(int(CNT_COUNT) == int(WFC_COUNT), "device counter slots = per-warp counter slots");

__device__ __forceinline__ void flushCounters(unsigned long long* counters, const PathCounters& pc, const TraceStats& st)
{
	// warp-aggregate, one atomic per warp and counter
	const unsigned mask = __activemask();
	unsigned v[CNT_COUNT] = {pc.closestRays, pc.anyRays, pc.cameraPaths, pc.lightPaths, pc.shadedVertices, pc.splats, pc.fallbackRays, st.nodes, st.tris, 0u, 0u};
	const int lane = threadIdx.x & 31;
	const int leader = __ffs(mask) - 1;
#pragma unroll
	for (int i = 0; i < CNT_COUNT; ++i)
	{
		unsigned s = __reduce_add_sync(mask, v[i]);
		if (lane == leader && s) atomicAdd(counters + i, (unsigned long long)s);
	}
}

// TraceRay / TraceRayAnyHit over a ray batch (traceray.glsl:19-180): one thread per ray.
__global__ void __launch_bounds__(128) k_trace(DScene s, const GiRay* __restrict__ rays, uint32_t n, GiHit* __restrict__ hits, int mode,
                                               unsigned long long* counters)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	PathCounters pc = {};
	TraceStats st = {0, 0};
	if (i < n)
	{
		const F4 r0 = ld4(reinterpret_cast<const F4*>(rays) + 2 * size_t(i));
		const F4 r1 = ld4(reinterpret_cast<const F4*>(rays) + 2 * size_t(i) + 1);
		const F3 o = f3(r0.x, r0.y, r0.z), d = f3(r1.x, r1.y, r1.z);
		const float tMax = r0.w;
		const uint32_t flags = uint32_t(f2i(r1.w));
		F4 out;
		if (flags & 1u)
		{
			++pc.anyRays;
			bool occluded = traceAny(s, mode, o, d, tMax, &st);
			out.x = tMax; out.y = i2f(occluded ? 1 : 0); out.z = 0.0f; out.w = 0.0f;
		}
		else
		{
			++pc.closestRays;
			Hit h; bool fb = false;
			traceClosest(s, mode, o, d, tMax, h, &st, &fb);
			if (fb) ++pc.fallbackRays;
			out.x = h.t; out.y = i2f(h.leafTri); out.z = h.leafTri >= 0 ? h.bary.y : 0.0f; out.w = h.leafTri >= 0 ? h.bary.z : 0.0f;
		}
		*reinterpret_cast<float4*>(hits + i) = make_float4(out.x, out.y, out.z, out.w);
	}
	flushCounters(counters, pc, st);
}

// pathtracer.comp: 8x8 work groups, grid (W/8, H/8) (pathtracer.cpp:50).
__global__ void __launch_bounds__(64) k_pathtrace(DFrame f, unsigned long long* counters)
{
	const int gx = blockIdx.x * 8 + (threadIdx.x & 7);
	const int gy = blockIdx.y * 8 + (threadIdx.x >> 3);
	PathCounters pc = {};
	TraceStats st = {0, 0};
	pathtracePixel(f, gx, gy, pc, &st);
	flushCounters(counters, pc, st);
}

// whittedraytracer.comp: same launch shape (whittedraytracer.cpp:31-36).
__global__ void __launch_bounds__(64) k_whitted(DFrame f, unsigned long long* counters)
{
	const int gx = blockIdx.x * 8 + (threadIdx.x & 7);
	const int gy = blockIdx.y * 8 + (threadIdx.x >> 3);
	PathCounters pc = {};
	TraceStats st = {0, 0};
	whittedPixel(f, gx, gy, pc, &st);
	flushCounters(counters, pc, st);
}

// pathtracer_bidir.comp: same launch shape (bidirectionalpathtracer.cpp:200).
__global__ void __launch_bounds__(64) k_bidir_eye(DFrame f, unsigned long long* counters)
{
	const int gx = blockIdx.x * 8 + (threadIdx.x & 7);
	const int gy = blockIdx.y * 8 + (threadIdx.x >> 3);
	PathCounters pc = {};
	TraceStats st = {0, 0};
	bidirPathtracePixel(f, gx, gy, pc, &st);
	flushCounters(counters, pc, st);
}

// lighttracer.comp: 64 paths per group, NumInitialLightSamples*NumRaysPerLightSample/64 groups (lightpathtracer.cpp:48).
template <bool SAVE, bool WARMUP>
__global__ void __launch_bounds__(64) k_lighttrace(DFrame f, unsigned long long* counters)
{
	const uint32_t gid = blockIdx.x * 64u + threadIdx.x;
	PathCounters pc = {};
	TraceStats st = {0, 0};
	lighttracePath<SAVE, WARMUP>(f, gid, pc, &st);
	flushCounters(counters, pc, st);
}

// showlightcaches.comp restated with NumInitialLightSamples threads.
__global__ void __launch_bounds__(64) k_show_light_caches(DFrame f, unsigned long long* counters)
{
	const int i = blockIdx.x * 64 + threadIdx.x;
	PathCounters pc = {};
	TraceStats st = {0, 0};
	if (i < f.numLightSamples) showLightCache(f, i, pc, &st);
	flushCounters(counters, pc, st);
}

// resetcounter.comp:10-11
__global__ void k_reset_counter(uint32_t* lightCacheCount, uint32_t* lightPathLengthSum, uint32_t initial)
{
	*lightCacheCount = initial;
	*lightPathLengthSum = 0u;
}

// LightSampler::GenerateRandomSamples (lightsampler.cpp:18-95) for one iteration.  The xorshift64* recurrence of utilities/random.cpp:25-43
// stores the multiplied value back into its state, so the stream has no jump-ahead: thread 0 draws the 3 * numArea + numPoint numbers
// of the iteration into shared memory, then one thread per sample does the rest.  Every +, -, *, /, sqrt is the host's (no contraction);
// two things are not: atan2f is the device's (normal_phi may differ from the host's libm in the last bit), and the table index is clamped
// to the last entry where the host asserts (a pick of exactly 1.0 on a table whose last entry rounds below 1).
struct DLightTables
{
	const float* tri;        // 12 floats per light triangle: v0, v1, v2, luminance
	const float* summedArea;
	uint32_t numTri;
	float areaSum;
	const float* pointPos; const float* pointInt; const float* pointSummedFlux;
	uint32_t numPoint;
	float areaFlux, pointFlux;  // totals
};

__device__ __forceinline__ uint32_t lowerBound(const float* table, uint32_t n, float key)
{
	uint32_t lo = 0, hi = n;
	while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (table[mid] < key) lo = mid + 1; else hi = mid; }
	return lo;
}

__global__ void __launch_bounds__(256) k_light_samples(DLightTables t, unsigned long long state, uint32_t n, float positionBias, F4* __restrict__ out)
{
	extern __shared__ float draws[];
	if (t.areaFlux + t.pointFlux == 0.0f)
	{
		for (uint32_t s = threadIdx.x; s < 2u * n; s += blockDim.x) *reinterpret_cast<float4*>(out + s) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
		return;
	}
	const float pointLightPercentage = __fdiv_rn(t.pointFlux, __fadd_rn(t.areaFlux, t.pointFlux));
	const uint32_t numPoint = uint32_t(__fmul_rn(pointLightPercentage, float(n)));
	const uint32_t numArea = n - numPoint;
	if (threadIdx.x == 0)
	{
		const uint32_t numDraws = 3u * numArea + numPoint;
		for (uint32_t i = 0; i < numDraws; ++i)
		{
			state ^= state >> 12; state ^= state << 25; state ^= state >> 27;
			state *= 2685821657736338717ull;
			draws[i] = __fdiv_rn(float(state % 8388593ull), 8388593.0f);
		}
	}
	__syncthreads();
	for (uint32_t s = threadIdx.x; s < n; s += blockDim.x)
	{
		float4 a, b;
		if (s < numArea)
		{
			const float sampleWeight = __fdiv_rn(t.areaSum, float(numArea));
			const float pick = draws[3u * s];
			float xi0 = draws[3u * s + 1u];
			const float xi1 = draws[3u * s + 2u];
			const uint32_t ti = min(lowerBound(t.summedArea, t.numTri, pick), t.numTri - 1u);
			xi0 = __fsqrt_rn(xi0);
			const float alpha = __fsub_rn(1.0f, xi0), beta = __fmul_rn(xi0, __fsub_rn(1.0f, xi1)), gamma = __fmul_rn(xi0, xi1);
			const float* v = t.tri + 12 * size_t(ti);
			float pos[3], e1[3], e2[3];
			for (int k = 0; k < 3; ++k)
			{
				pos[k] = __fadd_rn(__fadd_rn(__fmul_rn(v[k], alpha), __fmul_rn(v[3 + k], beta)), __fmul_rn(v[6 + k], gamma));
				e1[k] = __fsub_rn(v[k], v[3 + k]); e2[k] = __fsub_rn(v[k], v[6 + k]);
			}
			float nrm[3] = {__fsub_rn(__fmul_rn(e1[1], e2[2]), __fmul_rn(e1[2], e2[1])), __fsub_rn(__fmul_rn(e1[2], e2[0]), __fmul_rn(e1[0], e2[2])),
			                __fsub_rn(__fmul_rn(e1[0], e2[1]), __fmul_rn(e1[1], e2[0]))};
			const float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(nrm[0], nrm[0]), __fmul_rn(nrm[1], nrm[1])), __fmul_rn(nrm[2], nrm[2])));
			for (int k = 0; k < 3; ++k) { nrm[k] = __fdiv_rn(nrm[k], len); pos[k] = __fadd_rn(pos[k], __fmul_rn(nrm[k], positionBias)); }
			a = make_float4(pos[0], pos[1], pos[2], atan2f(nrm[1], nrm[0]));
			b = make_float4(__fmul_rn(v[9], sampleWeight), __fmul_rn(v[10], sampleWeight), __fmul_rn(v[11], sampleWeight), nrm[2]);
		}
		else
		{
			const float pick = draws[3u * numArea + (s - numArea)];
			const uint32_t li = min(lowerBound(t.pointSummedFlux, t.numPoint, pick), t.numPoint - 1u);
			float w = t.pointSummedFlux[li];
			if (li > 0) w = __fsub_rn(w, t.pointSummedFlux[li - 1]);
			const float div = __fmul_rn(float(numPoint), w);
			a = make_float4(t.pointPos[3 * li], t.pointPos[3 * li + 1], t.pointPos[3 * li + 2], 0.0f);
			b = make_float4(__fdiv_rn(t.pointInt[3 * li], div), __fdiv_rn(t.pointInt[3 * li + 1], div), __fdiv_rn(t.pointInt[3 * li + 2], div), 2.0f);
		}
		*reinterpret_cast<float4*>(out + 2 * size_t(s)) = a;
		*reinterpret_cast<float4*>(out + 2 * size_t(s) + 1) = b;
	}
}

// ---- wavefront scheduler (gi_wavefront.cuh) -------------------------------------------------------------
#ifndef WF_TRACE_THREADS
#define WF_TRACE_THREADS 128
#endif
#ifndef WF_TRACE_MIN_BLOCKS
#define WF_TRACE_MIN_BLOCKS (WF_CENTRE_EXTENT ? 9 : 7)   // centre / half-extent records: 56 registers, 9 x 4 warps per SM (B200 A/B, ms per PT iteration C2 / C5: 7 -> 1.669 / 15.04,
#endif                                                  // 8 -> 1.649 / 14.74, 9 -> 1.645 / 14.66, 10 -> 1.705 / -); 16-bit grid records: 72 registers, 7 blocks (round 1: 6 -> 2.378, 7 -> 2.137, 8 -> 2.168, 9 -> 2.255)
#define WF_SHADE_THREADS 128
#ifndef WF_SHADE_MIN_BLOCKS
#define WF_SHADE_MIN_BLOCKS 8   // k_wf_shade_pt: 64 registers, 8 blocks per SM (B200 A/B, PT batch: 1 -> 2.246, 8 -> 2.195, 10 -> 2.208, 12 -> 2.241 ms/iteration)
#endif

#ifndef WF_SHADE_LIGHT_MIN_BLOCKS
#define WF_SHADE_LIGHT_MIN_BLOCKS 8   // B200 A/B (ms per iteration, 4 / 5 / 6 / 8 / 10 blocks per SM = 118 / 96 / 80 / 64 / 48 registers): C4 LPT 4K 6.92 / 6.78 / 6.62 / 6.49 / 6.62
#endif
#ifndef WF_SHADE_LIGHT_COMPACT
#define WF_SHADE_LIGHT_COMPACT 1
#endif
#ifndef WF_SHADE_EYE_MIN_BLOCKS
#define WF_SHADE_EYE_MIN_BLOCKS 8     // C3 BPT 1080p 7.64 / 7.55 / 7.46 / 7.43 / 7.64 (both kernels bounded together)
#endif

// Camera rays of all W/8 x H/8 work groups (pathtracer.comp:36-46) at the head of this wave's extension queue.
__global__ void __launch_bounds__(256) k_wf_generate(WfFrame w, unsigned long long* counters)
{
	const uint32_t stride = gridDim.x * blockDim.x;
	for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < w.numPaths; slot += stride)
	{
		const uint32_t pixel = wfSlotToPixel(w, slot);
		const int gx = int(pixel % uint32_t(w.f.width)), gy = int(pixel / uint32_t(w.f.width));
		uint32_t seed = initRandomSeed(w.f.frameSeed, pixel);
		const F3 dir = cameraRayDirection(w.f, seed, gx, gy);
		F4* rp = w.q.rays[w.wave & 1] + 2 * size_t(slot);
		qst4(rp, make_float4(w.f.cam.position[0], w.f.cam.position[1], w.f.cam.position[2], i2f(0)));
		qst4(rp + 1, make_float4(dir.x, dir.y, dir.z, i2f(int(pixel))));
	}
	if (blockIdx.x == 0 && threadIdx.x == 0)
	{
		atomicAdd(counters + CNT_CLOSEST, (unsigned long long)w.numPaths);
		atomicAdd(counters + CNT_CAMERA, (unsigned long long)w.numPaths);
	}
}

// trace(w): extension rays of wave w + shadow rays pushed by shade(w-1); persistent warps, dynamic refill.
template <bool REFERENCE, bool STEPS>
__global__ void __launch_bounds__(WF_TRACE_THREADS, WF_TRACE_MIN_BLOCKS) k_wf_trace(WfFrame w, unsigned long long* counters)
{
	uint32_t* counts = wfCounts(w, w.wave);
	WfQueueIO io;
	io.q.rays = w.q.rays[w.wave & 1];
	io.q.hits = w.q.hits;
	io.q.shadow = w.q.shadow;
	io.q.stateC = w.q.stateC;
	io.q.accum = w.f.accum;
	io.q.width = w.f.width; io.q.height = w.f.height;
	io.q.numClosest = w.numCam + counts[0];
	io.q.numShadow = counts[1];
	io.q.rayBase = w.rayBase;
	io.q.splatBase = w.splatBase;
	io.q.splatStride = size_t(w.f.width) * size_t(w.f.height) * 4u;
	io.q.shadowPos = w.q.shadowPos;
	io.q.reconFilter = w.q.shadowPos ? w.f.reconFilter : 0;
	WfCounters<STEPS> cnt;
	if (REFERENCE) wfTraceLoopReference(w.f.scene, io, counts + 2, cnt);
	else wfTraceLoop(w.f.scene, io, counts + 2, w.tune, cnt);
	flushWarpCounters(counters, cnt.v, WFC_COUNT);
}

// gi_trace_device on the same persistent loop.
template <bool STEPS>
__global__ void __launch_bounds__(WF_TRACE_THREADS, WF_TRACE_MIN_BLOCKS) k_wf_trace_batch(DScene s, const GiRay* __restrict__ rays, uint32_t n, GiHit* __restrict__ hits, WfTune tune,
                                                                      uint32_t* workCounter, unsigned long long* counters)
{
	WfBatchIO io;
	io.rays = reinterpret_cast<const F4*>(rays);
	io.hits = reinterpret_cast<float4*>(hits);
	io.n = n;
	WfCounters<STEPS> cnt;
	wfTraceLoop(s, io, workCounter, tune, cnt);
	flushWarpCounters(counters, cnt.v, WFC_COUNT);
}

// Work distribution of the shade kernels: 32 consecutive queue slots per warp and fetch.  (A software-pipelined variant that prefetched the
// ray records, leaf record and leaf triangle of the next fetch measured slower on B200 -- C2 k_wf_shade_pt 0.326 -> 0.344 / 0.358 ms, C3 1.740 -> 1.833 /
// 1.932, C4 1.505 -> 1.552 / 1.658 -- and was removed)
GI_D uint32_t wfGrab32(uint32_t* counter)
{
	uint32_t base = 0;
	if (laneId() == 0) base = atomicAdd(counter, 32u);
	return __shfl_sync(0xffffffffu, base, 0);
}

template <class BODY>
GI_D void wfShadeLoop(const WfFrame& w, uint32_t* counter, uint32_t n, BODY body)
{
	for (;;)
	{
		const uint32_t base = wfGrab32(counter);
		if (base >= n) break;
		const uint32_t slot = base + laneId();
		body(w.rayBase + slot, slot < n);
	}
}

// shade(w) of the path tracer: one lane per hit record, 32 consecutive queue slots per warp and fetch.
__global__ void __launch_bounds__(WF_SHADE_THREADS, WF_SHADE_MIN_BLOCKS) k_wf_shade_pt(WfFrame w, unsigned long long* counters)
{
	uint32_t* counts = wfCounts(w, w.wave);
	const uint32_t n = w.numCam + counts[0];
	unsigned cnt[WFC_COUNT] = {};
	wfShadeLoop(w, counts + 3, n, [&](uint32_t slot, bool valid) { wfShadePT(w, slot, valid, cnt); });
	flushWarpCounters(counters, cnt, WFC_COUNT);
}

// generate(w) / shade(w) of the light tracer and shade(w) of the BPT eye pass (gi_shade.h); same work distribution as k_wf_shade_pt.
__global__ void __launch_bounds__(256) k_wf_generate_light(WfFrame w, unsigned long long* counters)
{
	PathCounters pc = {};
	const uint32_t stride = gridDim.x * blockDim.x;
	for (uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x; gid < w.numPaths; gid += stride) wfGenerateLight(w, gid, pc);
	unsigned cnt[WFC_COUNT] = {};
	wfAddCounters(cnt, pc);
	flushWarpCounters(counters, cnt, WFC_COUNT);
}

__global__ void __launch_bounds__(WF_SHADE_THREADS, WF_SHADE_LIGHT_MIN_BLOCKS) k_wf_shade_light(WfFrame w, unsigned long long* counters)
{
	uint32_t* counts = wfCounts(w, w.wave);
	const uint32_t n = w.numCam + counts[0];
	PathCounters pc = {};
#if WF_SHADE_LIGHT_COMPACT
	// A light path that left the scene has nothing to shade (lighttracer.comp:97-98) -- in an open scene that is a third of the hit records (C4: 36 %),
	// and their lanes would idle through the whole vertex.  Each warp therefore collects the slots of real hits in a small pool (shared memory)
	// and shades 32 of them at a time.  B200 A/B (k_wf_shade_light per iteration / whole iteration, ms): C4 1.488 -> 1.185 / 6.495 -> 6.196.
	__shared__ uint32_t pool[WF_SHADE_THREADS / 32][64];
	uint32_t* mine = pool[threadIdx.x >> 5];
	const unsigned lane = laneId();
	uint32_t pooled = 0;
	bool more = true;
	for (;;)
	{
		while (more && pooled < 32u)
		{
			const uint32_t base = wfGrab32(counts + 3);
			if (base >= n) { more = false; break; }
			const uint32_t slot = base + lane;
			const bool hit = slot < n && f2i(qld2(w.q.hits + w.rayBase + slot).y) >= 0;
			const unsigned m = __ballot_sync(0xffffffffu, hit);
			if (m == 0xffffffffu && pooled == 0u) { wfShadeLight(w, w.rayBase + slot, true, pc); continue; } // a full warp of hits (closed scenes): no detour through the pool
			if (hit) mine[pooled + __popc(m & lanemaskLt())] = w.rayBase + slot;
			pooled += uint32_t(__popc(m));
			__syncwarp();
		}
		if (pooled == 0u) break;
		const uint32_t take = min(32u, pooled);
		const bool valid = lane < take;
		const uint32_t slot = valid ? mine[pooled - take + lane] : 0u;
		__syncwarp();
		pooled -= take;
		wfShadeLight(w, slot, valid, pc);
	}
#else
	wfShadeLoop(w, counts + 3, n, [&](uint32_t slot, bool valid) { wfShadeLight(w, slot, valid, pc); });
#endif
	unsigned cnt[WFC_COUNT] = {};
	wfAddCounters(cnt, pc);
	flushWarpCounters(counters, cnt, WFC_COUNT);
}

__global__ void __launch_bounds__(WF_SHADE_THREADS, WF_SHADE_EYE_MIN_BLOCKS) k_wf_shade_eye(WfFrame w, unsigned long long* counters)
{
	uint32_t* counts = wfCounts(w, w.wave);
	const uint32_t n = w.numCam + counts[0];
	PathCounters pc = {};
	wfShadeLoop(w, counts + 3, n, [&](uint32_t slot, bool valid) { wfShadeEye(w, slot, valid, pc); });
	unsigned cnt[WFC_COUNT] = {};
	wfAddCounters(cnt, pc);
	flushWarpCounters(counters, cnt, WFC_COUNT);
}

// Final fold: the radiance sums of the iterations still held in state slots join the accumulator, per pixel in
// iteration order (pathtracer.comp:107-109: Out[px] += vec4(color, 0)).
__global__ void __launch_bounds__(256) k_wf_fold(float4* __restrict__ accum, const F4* __restrict__ stateC, uint32_t numIds, uint32_t numSlots,
                                                  uint32_t firstIteration, uint32_t endIteration)
{
	for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < numIds; p += gridDim.x * blockDim.x)
	{
		float4 a = accum[p];
		for (uint32_t k = firstIteration; k < endIteration; ++k)
		{
			const F4 c = ld4(stateC + size_t(k % numSlots) * numIds + p);
			a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w;
		}
		accum[p] = a;
	}
}

// A stream of light paths keeps the splats of every iteration in flight in that iteration's own image; when the iteration is complete the
// image joins the accumulator and is cleared for the iteration that takes the ring slot next.
__global__ void __launch_bounds__(256) k_wf_fold_splats(float4* __restrict__ accum, float4* __restrict__ image, uint32_t numPixels)
{
	for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < numPixels; p += gridDim.x * blockDim.x)
	{
		const float4 s = image[p];
		if (s.x != 0.0f || s.y != 0.0f || s.z != 0.0f || s.w != 0.0f)
		{
			float4 a = accum[p];
			a.x += s.x; a.y += s.y; a.z += s.z; a.w += s.w;
			accum[p] = a;
			image[p] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
		}
	}
}

// sqerror.frag:16-17 + mean (texturemse.cpp:58-62): per-block partial sums in double.
__global__ void __launch_bounds__(256) k_mse(const float4* __restrict__ accum, const float* __restrict__ ref, uint32_t numPixels, float invDivisor,
                                              double* result)
{
	double sum = 0.0;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < numPixels; i += gridDim.x * blockDim.x)
	{
		float4 a = accum[i];
		float d0 = a.x * invDivisor - ref[3 * size_t(i)], d1 = a.y * invDivisor - ref[3 * size_t(i) + 1], d2 = a.z * invDivisor - ref[3 * size_t(i) + 2];
		sum += double(d0 * d0 + d1 * d1 + d2 * d2);
	}
	for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
	__shared__ double warpSums[8];
	if ((threadIdx.x & 31) == 0) warpSums[threadIdx.x >> 5] = sum;
	__syncthreads();
	if (threadIdx.x == 0)
	{
		double s = 0.0;
		for (int w = 0; w < 8; ++w) s += warpSums[w];
		atomicAdd(result, s);
	}
}

// =====================================================================================================
// context
// =====================================================================================================
enum { WF_PASS_PT = 0, WF_PASS_LIGHT, WF_PASS_EYE };
#define WF_STREAM_COUNTS_RING 8u
#define GI_LIGHT_RING 64u            // light-sample sets kept on the device: a stream has up to maxPathLength + 1 iterations in flight

struct WfRun
{
	bool open = false, streaming = false;
	int pass = 0;
	uint32_t count = 0;      // iterations started
	uint32_t wave = 0;       // next wave
	uint32_t maxLen = 0, numSlots = 0;
	int renderer = 0, numRaysPerLightSample = 0;
	// splat ring of a light-path stream: iteration k splats into image k % splatSlots and is folded into the accumulator after the trace of
	// wave k + foldDelay (LPT: its last camera connection is traced in wave k + maxLen; BPT: together with the eye radiance of iteration k)
	uint32_t splatSlots = 0, foldDelay = 0, splatFolded = 0;
	WfFrame w;
};

// pipelined BPT (renderPipelinedBPT / streaming Draw of the bidirectional path tracer)
struct BptPipe
{
	bool open = false, streaming = false;
	uint32_t count = 0;      // iterations submitted
	uint32_t wave = 0;       // next global wave
	uint32_t expected = 0;   // batch: total iterations
	uint32_t ring = 0, D = 0, maxLen = 0, capacity = 0, numLightPaths = 0;
	int numRaysPerLightSample = 0;
	std::vector<uint32_t> seeds; // FrameSeed of iteration k at [k % size]: its eye pass starts D waves later
};

struct GiContext
{
	int device = 0;
	cudaStream_t ownStream = nullptr;
	cudaStream_t stream = nullptr;
	std::string error;
	uint64_t launches = 0;

	// scene
	bool haveScene = false;
	DScene scene = {};
	std::vector<void*> sceneAllocs;
	std::vector<cudaArray_t> arrays;
	std::vector<cudaTextureObject_t> texObjects;
	bool fastValid = false;

	// environment
	cudaArray_t envArray = nullptr;
	cudaTextureObject_t envTex = 0;
	F3 envColor = {0.0f, 0.0f, 0.0f};

	// frame
	GiCamera cam = {};
	bool haveCamera = false;
	int width = 0, height = 0;
	float* accum = nullptr;
	F4* lightSamples = nullptr;
	const F4* externalLightSamples = nullptr; // gi_set_light_samples_device
	uint32_t lightSampleCapacity = 0;
	int numLightSamples = 0;

	// device-side light sampling (gi_set_light_tables)
	DLightTables lightTables = {};
	std::vector<void*> lightTableAllocs;

	// light cache
	F4* lightCache = nullptr;
	uint32_t lightCacheCapacity = 0;
	uint32_t* bptCounters = nullptr; // [0] NumLightCacheEntries, [1] LightPathLengthSum
	// pipelined BPT batches: a ring of caches + counter pairs, one per iteration in flight (up to 2 x maxPathLength)
	F4* cacheRing = nullptr;
	uint64_t cacheRingEntries = 0;   // allocated entries in total
	uint32_t* ringCounters = nullptr;
	uint32_t ringCountersLen = 0;

	// counters
	unsigned long long* counters = nullptr;
	double* mseResult = nullptr;

	// options
	int traversal = GI_TRAVERSAL_FAST;
	int textureFilterExact = 0;
	int scheduler = GI_SCHEDULER_WAVEFRONT;
	int countSteps = 0;              // GI_OPT_COUNT_STEPS
	int reconFilter = 0;             // GI_OPT_RECONSTRUCTION_FILTER
	int chainedMIS = 0;              // GI_OPT_CHAINED_MIS
	WfTune tune = {24, 16, 0};     // B200 sweep (tools/gpu/run_variants.sh TUNES=...): refill:minInner 16:20 2.191, 24:20 2.153, 24:16 2.126, 28:16 2.143, 32:16 2.222 ms/iteration

	// wavefront queues
	BptPipe bpt;
	WfRun run[2];                    // [0]: PT / LPT / BPT light pass / stream, [1]: BPT eye pass of a pipelined batch
	int streamingOption = 0;         // GI_OPT_STREAMING
	int deviceBuilder = GI_BUILDER_PLOC; // GI_OPT_DEVICE_BUILDER
	int uploadTopology = GI_TOPOLOGY_PLOC; // GI_OPT_UPLOAD_TOPOLOGY (what the next gi_upload_scene builds)
	int sceneTopology = GI_TOPOLOGY_REFERENCE; // what the uploaded scene's ordered layout is built on
	double topologyBuildMs = 0.0;
	int plocRadius = 4; // GI_OPT_PLOC_RADIUS (clustering passes of the upload topology; B200 sweep: 4 beats 1..3 and 6..64 on C2, 2..4 on C5)
	uint32_t lightSlot = 0;          // ring slot of the light-sample set the next iteration uses
	uint32_t lightSetIteration = 0;  // stream iteration (run[0].count at the time) whose ring slot holds a set installed by the caller
	float* splatRing = nullptr;      // per-iteration splat images of an open LPT / BPT stream
	uint32_t splatRingSlots = 0;
	size_t splatRingStride = 0;      // floats per image
	// asynchronous read-out (gi_read_accumulator_async)
	cudaStream_t copyStream = nullptr;
	float* snapshot[2] = {nullptr, nullptr};
	size_t snapshotBytes = 0;
	cudaEvent_t snapTaken[2] = {nullptr, nullptr}, snapCopied[2] = {nullptr, nullptr};
	bool snapUsed[2] = {false, false};
	int snapNext = 0;
	WfQueues wf[2] = {};
	uint32_t wfCountsLen[2] = {0, 0};
	uint64_t wfStates[2] = {0, 0};   // allocated path state records (numSlots x numIds)
	uint32_t* traceWork = nullptr;   // work counter of k_wf_trace_batch
	int numSMs = 148;
	uint64_t residentBytes[4] = {0, 0, 0, 0}; // gi_scene_resident_bytes
	int traceBlocksPerSM = 1, shadeBlocksPerSM = 1, shadeLightBlocksPerSM = 1, shadeEyeBlocksPerSM = 1;

	// multi-GPU read-out (gi_comm.cuh): exchange block shared with the peer ranks over CUDA IPC
	int commRank = 0, commWorld = 0;          // world 0: no communicator
	bool commConnected = false;
	unsigned char* commBlock = nullptr;       // own exchange block
	unsigned char* commPeer[GI_COMM_MAX_WORLD] = {};
	int commWidth = 0, commHeight = 0;
	uint32_t commEpoch = 0;
	float* commOut[2] = {nullptr, nullptr};   // this rank's rows of the summed image
	uint32_t* commWords = nullptr;            // [0] arrival counter of k_comm_snapshot, [1] of k_comm_reduce, [2] status (0 = ok)
	cudaEvent_t commSnapped[2] = {nullptr, nullptr}, commReduced[2] = {nullptr, nullptr};
	bool commUsed[2] = {false, false};

	// GI_OPT_KERNEL_TIMING: CUDA-event duration of every launch of an iteration, summed per kernel class
	int kernelTiming = 0;
	double kernelMs[GI_KERNEL_CLASS_COUNT] = {};
	uint64_t kernelLaunches[GI_KERNEL_CLASS_COUNT] = {};
	std::vector<cudaEvent_t> timingEvents;   // pool, reused across iterations
	std::vector<int> timingClass;            // class of the launch that ends at event i+1
	size_t timingUsed = 0;
};

// Kernel timing: timingMark(c, cls) after a launch closes the interval since the previous mark and books it to `cls`
// (cls < 0 opens the first interval); timingCollect() synchronises and sums.  Events sit on the launching stream.
static void timingMark(GiContext* c, int cls)
{
	if (!c->kernelTiming) return;
	if (c->timingUsed == c->timingEvents.size()) { cudaEvent_t e; cudaEventCreate(&e); c->timingEvents.push_back(e); c->timingClass.push_back(-1); }
	cudaEventRecord(c->timingEvents[c->timingUsed], c->stream);
	c->timingClass[c->timingUsed] = cls;
	++c->timingUsed;
}

static void timingCollect(GiContext* c)
{
	if (!c->kernelTiming || c->timingUsed == 0) return;
	cudaEventSynchronize(c->timingEvents[c->timingUsed - 1]);
	for (size_t i = 1; i < c->timingUsed; ++i)
	{
		const int cls = c->timingClass[i];
		if (cls < 0) continue;
		float ms = 0.0f;
		cudaEventElapsedTime(&ms, c->timingEvents[i - 1], c->timingEvents[i]);
		c->kernelMs[cls] += double(ms);
		++c->kernelLaunches[cls];
	}
	c->timingUsed = 0;
}

static std::string g_createError;

static int fail(GiContext* ctx, int code, const std::string& msg)
{
	if (ctx) ctx->error = msg;
	else g_createError = msg;
	return code;
}

#define GI_CUDA(ctx, call)                                                                                         \
	do                                                                                                             \
	{                                                                                                              \
		cudaError_t e_ = (call);                                                                                   \
		if (e_ != cudaSuccess)                                                                                     \
			return fail(ctx, GI_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                    \
	} while (0)

static void freeScene(GiContext* c)
{
	for (cudaTextureObject_t t : c->texObjects) cudaDestroyTextureObject(t);
	for (cudaArray_t a : c->arrays) cudaFreeArray(a);
	for (void* p : c->sceneAllocs) cudaFree(p);
	c->texObjects.clear(); c->arrays.clear(); c->sceneAllocs.clear();
	c->haveScene = false;
	std::memset(&c->scene, 0, sizeof(c->scene));
}

template <typename T>
static cudaError_t uploadVector(GiContext* c, const std::vector<T>& v, const T** out)
{
	void* d = nullptr;
	size_t bytes = std::max<size_t>(v.size() * sizeof(T), 16);
	cudaError_t e = cudaMalloc(&d, bytes);
	if (e != cudaSuccess) return e;
	c->sceneAllocs.push_back(d);
	if (!v.empty()) e = cudaMemcpyAsync(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, c->stream);
	*out = static_cast<const T*>(d);
	return e;
}

static DFrame makeFrame(GiContext* c, const GiIterationParams* p)
{
	DFrame f;
	std::memset(&f, 0, sizeof(f));
	f.scene = c->scene;
	f.scene.textureFilterExact = c->textureFilterExact;
	// bvhType = 1: GI_TRAVERSAL_FAST selects the ordered walk on the oriented boxes (the reference-order kernels host it)
	f.scene.oboxOrdered = (c->scene.oboxNodes && c->scene.refTreeBinary && c->traversal == GI_TRAVERSAL_FAST) ? 1 : 0;
	f.cam = c->cam;
	f.width = c->width; f.height = c->height;
	f.lightSamples = c->externalLightSamples ? c->externalLightSamples : c->lightSamples + size_t(c->lightSlot) * 2u * size_t(c->numLightSamples);
	f.numLightSamples = c->numLightSamples;
	f.frameSeed = p ? p->frame_seed : 0u;
	f.maxPathLength = p ? p->max_path_length : 16;
	f.traversalMode = c->fastValid ? c->traversal : GI_TRAVERSAL_REFERENCE;
	f.accum = c->accum;
	f.envColor = c->envColor;
	f.envCube = (unsigned long long)c->envTex;
	f.numRaysPerLightSample = p ? p->num_rays_per_light_sample : 0;
	f.lightCache = c->lightCache;
	f.lightCacheCapacity = 0;
	f.lightCacheCount = c->bptCounters;
	f.lightPathLengthSum = c->bptCounters + 1;
	f.reconFilter = c->reconFilter;
	f.chainedMIS = c->chainedMIS;
	return f;
}


// ---- wavefront queue storage: 2 ray queues, hit records, shadow queue, path state, per-bounce counters
static void freeWavefront(GiContext* c)
{
	for (int set = 0; set < 2; ++set)
	{
		WfQueues& q = c->wf[set];
		cudaFree(q.rays[0]); cudaFree(q.rays[1]); cudaFree(q.hits); cudaFree(q.shadow); cudaFree(q.shadowPos); cudaFree(q.stateT); cudaFree(q.stateC); cudaFree(q.stateM);
		cudaFree(q.counts);
		q = WfQueues{};
		c->wfCountsLen[set] = 0;
		c->wfStates[set] = 0;
	}
}

// capacity: queue slots; numIds: path ids per iteration (PT/BPT: the pixel index, which is not dense when the frame size is
// not a multiple of 8); numSlots: iterations in flight; countsLen: wave counters
static int ensureWavefront(GiContext* c, int set, uint64_t capacity, uint32_t numIds, uint32_t numSlots, uint32_t countsLen)
{
	if (capacity > WF_ID_MASK || uint64_t(numIds) * numSlots > WF_ID_MASK)
		return fail(c, GI_ERR_CAPACITY, "wavefront scheduler: more than 2^30 queue slots / path states");
	const uint64_t states = uint64_t(numIds) * numSlots;
	const bool needPos = c->reconFilter != 0;
	if (c->wf[set].capacity < capacity || c->wfStates[set] < states || (needPos && !c->wf[set].shadowPos))
	{
		GI_CUDA(c, cudaStreamSynchronize(c->stream));
		const uint64_t cap = std::max<uint64_t>(capacity, c->wf[set].capacity);
		const uint64_t st = std::max<uint64_t>(states, c->wfStates[set]);
		cudaFree(c->wf[set].rays[0]); cudaFree(c->wf[set].rays[1]); cudaFree(c->wf[set].hits); cudaFree(c->wf[set].shadow); cudaFree(c->wf[set].shadowPos); cudaFree(c->wf[set].stateT); cudaFree(c->wf[set].stateC); cudaFree(c->wf[set].stateM);
		c->wf[set].rays[0] = c->wf[set].rays[1] = nullptr; c->wf[set].hits = nullptr; c->wf[set].shadow = nullptr; c->wf[set].shadowPos = nullptr; c->wf[set].stateT = c->wf[set].stateC = nullptr; c->wf[set].stateM = nullptr;
		c->wf[set].capacity = 0; c->wfStates[set] = 0;
		GI_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->wf[set].rays[0]), size_t(cap) * 32));
		GI_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->wf[set].rays[1]), size_t(cap) * 32));
		GI_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->wf[set].hits), size_t(cap) * 8));
		GI_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->wf[set].shadow), size_t(cap) * 48));
		if (needPos) GI_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->wf[set].shadowPos), size_t(cap) * 8));
		GI_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->wf[set].stateT), size_t(st) * 16));
		GI_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->wf[set].stateC), size_t(st) * 16));
		GI_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->wf[set].stateM), size_t(st) * 4));
		c->wf[set].capacity = uint32_t(cap);
		c->wfStates[set] = st;
	}
	if (c->wfCountsLen[set] < countsLen)
	{
		GI_CUDA(c, cudaStreamSynchronize(c->stream));
		cudaFree(c->wf[set].counts); c->wf[set].counts = nullptr; c->wfCountsLen[set] = 0;
		GI_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->wf[set].counts), size_t(countsLen) * sizeof(uint32_t)));
		c->wfCountsLen[set] = countsLen;
	}
	c->wf[set].numIds = numIds;
	c->wf[set].numSlots = numSlots;
	return GI_OK;
}

// Iterations of one pass as waves (gi_wavefront.cuh): wave w = generate(iteration w), trace, shade; maxPathLength drain
// waves; final fold of the per-path radiance sums (PT, BPT eye pass).
//   WF_PASS_PT     pathtracer.comp          ids = pixels, W/8 x H/8 tiles
//   WF_PASS_LIGHT  lighttracer.comp         ids = light paths; lightFlags = WF_LIGHT_SAVE / WF_LIGHT_WARMUP (gi_shade.h)
//   WF_PASS_EYE    pathtracer_bidir.comp    ids = pixels; one iteration (the pass reads the light cache its own light pass filled)
// A run is opened (wfOpen), fed one iteration per wave (wfStep) and closed (wfClose: drain + fold).  A batch
// (gi_render_iterations) does all three in one call; a stream (GI_OPT_STREAMING) keeps the run open across
// gi_render_iteration calls, so that consecutive Draw() calls overlap exactly like the iterations of a batch.
static int ensureSplatRing(GiContext* c, uint32_t slots)
{
	const size_t stride = size_t(c->width) * size_t(c->height) * 4u;
	if (c->splatRingSlots < slots || c->splatRingStride != stride)
	{
		GI_CUDA(c, cudaStreamSynchronize(c->stream));
		cudaFree(c->splatRing); c->splatRing = nullptr; c->splatRingSlots = 0; c->splatRingStride = 0;
		GI_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->splatRing), size_t(slots) * stride * sizeof(float)));
		c->splatRingSlots = slots; c->splatRingStride = stride;
	}
	GI_CUDA(c, cudaMemsetAsync(c->splatRing, 0, size_t(slots) * stride * sizeof(float), c->stream));
	return GI_OK;
}

// the splat image of stream iteration k joins the accumulator
static int foldSplatImage(GiContext* c, WfRun& r, uint32_t iteration)
{
	float* image = c->splatRing + size_t(iteration % r.splatSlots) * c->splatRingStride;
	k_wf_fold_splats<<<unsigned(c->numSMs) * 8u, 256, 0, c->stream>>>(reinterpret_cast<float4*>(c->accum), reinterpret_cast<float4*>(image), uint32_t(c->width) * uint32_t(c->height));
	++c->launches;
	timingMark(c, GI_KERNEL_OTHER);
	return GI_OK;
}

static int wfOpen(GiContext* c, int set, const DFrame& f, int pass, uint32_t lightFlags, uint32_t expectedCount, bool streaming, uint32_t lightStride, uint32_t lightBase,
                  uint32_t lightRing, uint32_t splatSlots = 0u, uint32_t foldDelay = 0u)
{
	WfRun& r = c->run[set];
	WfFrame& w = r.w;
	std::memset(&w, 0, sizeof(w));
	w.f = f;
	w.lightFlags = lightFlags;
	w.tilesX = c->width / 8; w.tilesY = c->height / 8; // integer division, pathtracer.cpp:50
	uint32_t numIds;
	if (pass == WF_PASS_LIGHT)
	{
		w.numPaths = (uint32_t(f.numLightSamples) * uint32_t(f.numRaysPerLightSample) / 64u) * 64u; // whole groups of 64, lightpathtracer.cpp:48
		numIds = w.numPaths;
	}
	else
	{
		w.numPaths = uint32_t(w.tilesX) * uint32_t(w.tilesY) * 64u;
		numIds = uint32_t(c->width) * uint32_t(c->height);
	}
	w.lightStride = lightStride; w.lightBase = lightBase; w.lightRing = lightRing ? lightRing : 0x7fffffffu;
	r.open = false;
	r.pass = pass; r.streaming = streaming; r.count = 0; r.wave = 0;
	r.maxLen = uint32_t(f.maxPathLength);
	r.splatSlots = 0u; r.foldDelay = foldDelay; r.splatFolded = 0u;
	if (w.numPaths == 0) return GI_OK;
	if (splatSlots)
	{
		int sst = ensureSplatRing(c, splatSlots);
		if (sst != GI_OK) return sst;
		r.splatSlots = splatSlots;
		w.splatBase = c->splatRing; w.splatSlots = splatSlots;
	}
	const bool many = streaming || expectedCount > 1u;
	r.numSlots = streaming ? r.maxLen + 1u : std::min(expectedCount, r.maxLen + 1u);
	// a batch has one counter entry per wave (zeroed here); a stream cycles through a ring
	const uint32_t countsRing = streaming ? WF_STREAM_COUNTS_RING : expectedCount + r.maxLen + 1u;
	// worst case (no path ever ends before maxPathLength): maxLen generations of rays share a wave.  Any batch / stream
	// reserves that much, so that runs of different length never reallocate.
	int st = ensureWavefront(c, set, uint64_t(w.numPaths) * (many ? r.maxLen : 1u), numIds, many ? r.maxLen + 1u : 1u, WF_COUNTS_PER_WAVE * countsRing);
	if (st != GI_OK) return st;
	w.q = c->wf[set];
	w.q.numSlots = r.numSlots;
	w.q.countsRing = countsRing;
	w.tune = c->tune;
	GI_CUDA(c, cudaMemsetAsync(c->wf[set].counts, 0, size_t(WF_COUNTS_PER_WAVE) * countsRing * sizeof(uint32_t), c->stream));
	// ids outside the W/8 x H/8 tiles are never written by shade(); the fold reads them
	if (pass != WF_PASS_LIGHT && w.numPaths != w.q.numIds) GI_CUDA(c, cudaMemsetAsync(c->wf[set].stateC, 0, size_t(r.numSlots) * w.q.numIds * 16, c->stream));
	r.open = true;
	timingMark(c, -1);
	return GI_OK;
}

// One wave.  generate: iteration r.count starts in this wave (FrameSeed `seed`); nextGenerates: the next wave will start an
// iteration too (a stream cannot know and always reserves the head of the next queue); shade: false for the very last wave.
static int wfStep(GiContext* c, int set, bool generate, uint32_t seed, bool nextGenerates, bool shade)
{
	WfRun& r = c->run[set];
	WfFrame& w = r.w;
	if (!r.open) return GI_OK;
	const int pass = r.pass;
	w.wave = int(r.wave);
	w.numCam = generate ? w.numPaths : 0u;
	w.pushBase = (r.streaming || nextGenerates) ? w.numPaths : 0u;
	w.rayBase = (r.streaming && !generate) ? w.numPaths : 0u;
	const unsigned traceGrid = unsigned(c->numSMs * c->traceBlocksPerSM);
	const unsigned shadeGrid = unsigned(c->numSMs * (pass == WF_PASS_PT ? c->shadeBlocksPerSM : pass == WF_PASS_LIGHT ? c->shadeLightBlocksPerSM : c->shadeEyeBlocksPerSM));
	const bool ref = w.f.traversalMode == GI_TRAVERSAL_REFERENCE;
	if (r.streaming)
		GI_CUDA(c, cudaMemsetAsync(c->wf[set].counts + WF_COUNTS_PER_WAVE * size_t((r.wave + 1u) % w.q.countsRing), 0, WF_COUNTS_PER_WAVE * sizeof(uint32_t), c->stream));
	if (generate)
	{
		w.f.frameSeed = seed;
		if (pass == WF_PASS_LIGHT) k_wf_generate_light<<<unsigned(c->numSMs) * 8u, 256, 0, c->stream>>>(w, c->counters);
		else k_wf_generate<<<unsigned(c->numSMs) * 8u, 256, 0, c->stream>>>(w, c->counters);
		++c->launches;
		++r.count;
		timingMark(c, GI_KERNEL_GENERATE);
	}
	if (ref) k_wf_trace<true, true><<<traceGrid, WF_TRACE_THREADS, 0, c->stream>>>(w, c->counters);
	else if (c->countSteps) k_wf_trace<false, true><<<traceGrid, WF_TRACE_THREADS, 0, c->stream>>>(w, c->counters);
	else k_wf_trace<false, false><<<traceGrid, WF_TRACE_THREADS, 0, c->stream>>>(w, c->counters);
	++c->launches;
	timingMark(c, GI_KERNEL_TRACE);
	if (r.splatSlots && r.wave >= r.foldDelay && r.wave - r.foldDelay < r.count && r.splatFolded == r.wave - r.foldDelay)
	{
		// every splat of iteration wave - foldDelay has landed in its image
		int fst = foldSplatImage(c, r, r.splatFolded);
		if (fst != GI_OK) return fst;
		++r.splatFolded;
	}
	if (shade)
	{
		if (pass == WF_PASS_PT) k_wf_shade_pt<<<shadeGrid, WF_SHADE_THREADS, 0, c->stream>>>(w, c->counters);
		else if (pass == WF_PASS_LIGHT) k_wf_shade_light<<<shadeGrid, WF_SHADE_THREADS, 0, c->stream>>>(w, c->counters);
		else k_wf_shade_eye<<<shadeGrid, WF_SHADE_THREADS, 0, c->stream>>>(w, c->counters);
		++c->launches;
		timingMark(c, GI_KERNEL_SHADE);
	}
	++r.wave;
	return GI_OK;
}

// Drain: maxPathLength waves without generate (the last one carries shadow rays only), then the fold.
static int wfClose(GiContext* c, int set, bool drain = true)
{
	WfRun& r = c->run[set];
	if (!r.open) return GI_OK;
	if (r.count > 0)
	{
		for (uint32_t k = 0; drain && k < r.maxLen; ++k)
		{
			int st = wfStep(c, set, false, 0u, false, k + 1u < r.maxLen);
			if (st != GI_OK) return st;
		}
		while (r.splatSlots && r.splatFolded < r.count)
		{
			int fst = foldSplatImage(c, r, r.splatFolded);
			if (fst != GI_OK) return fst;
			++r.splatFolded;
		}
		if (r.pass != WF_PASS_LIGHT)
		{
			const uint32_t first = r.count > r.numSlots ? r.count - r.numSlots : 0u;
			k_wf_fold<<<unsigned(c->numSMs) * 8u, 256, 0, c->stream>>>(reinterpret_cast<float4*>(c->accum), c->wf[set].stateC, r.w.q.numIds, r.numSlots, first, r.count);
			++c->launches;
			timingMark(c, GI_KERNEL_OTHER);
		}
	}
	r.open = false;
	// a stream leaves the set of the NEXT iteration at ring slot lightBase + count
	if (r.streaming) c->lightSlot = (r.w.lightBase + r.count) % GI_LIGHT_RING;
	timingCollect(c);
	GI_CUDA(c, cudaGetLastError());
	return GI_OK;
}

static int bptClose(GiContext* c);

// an open stream must be drained before anything reads or redefines what its iterations depend on
static int flushStream(GiContext* c)
{
	if (c && c->bpt.open && c->bpt.streaming) return bptClose(c);
	if (c && c->run[0].open && c->run[0].streaming) return wfClose(c, 0);
	return GI_OK;
}

static int renderWavefront(GiContext* c, const DFrame& f, int pass, uint32_t lightFlags, const uint32_t* seeds, uint32_t count, uint32_t lightStride)
{
	if (count == 0) return GI_OK;
	if (pass == WF_PASS_EYE && count != 1u) return fail(c, GI_ERR_INVALID, "wavefront BPT eye pass: one iteration at a time");
	int st = wfOpen(c, 0, f, pass, lightFlags, count, false, lightStride, 0u, 0u);
	for (uint32_t k = 0; st == GI_OK && k < count; ++k) st = wfStep(c, 0, true, seeds[k], k + 1u < count, true);
	if (st == GI_OK) st = wfClose(c, 0);
	c->run[0].open = false;
	return st;
}


// A batch of BPT iterations as ONE pipeline of waves.  Global wave w carries
//   light run (queue set 0):  generate(light paths of iteration w), light bounce b of iteration w - b
//   eye run   (queue set 1):  generate(camera rays of iteration w - D), eye bounce b of iteration w - D - b,   D = maxPathLength
// The eye pass of iteration k may only start when its light cache is complete (last light shade: wave k + maxPathLength - 1),
// hence the offset D; cache k is written in waves k .. k+D-1 and read in waves k+D .. k+2D-1, so 2 D caches (each with its own
// NumLightCacheEntries counter) are alive at a time: a ring in HBM (C3 at 1080p: 32 x 0.4 GB).  Every wave then has the
// same amount of work and the per-iteration tails of both passes are paid once per batch.  Returns GI_ERR_CAPACITY without
// having rendered anything when the ring does not fit the free device memory (the caller then renders one by one).
// bptOpen: allocate the ring (expectedCount = 0: a stream, which always needs the full 2 D caches) and open both runs.
static int bptOpen(GiContext* c, const GiIterationParams* p, uint32_t expectedCount, uint32_t lightStride, uint32_t lightBase, uint32_t lightRing)
{
	BptPipe& b = c->bpt;
	const bool streaming = expectedCount == 0u;
	const uint32_t maxLen = uint32_t(p->max_path_length);
	const uint32_t ring = streaming ? 2u * maxLen : std::min(expectedCount, 2u * maxLen);
	const uint32_t capacity = uint32_t(p->light_cache_capacity);
	const uint32_t numLightPaths = uint32_t(c->numLightSamples) * uint32_t(p->num_rays_per_light_sample);
	const uint64_t entries = uint64_t(ring) * capacity;
	if (c->cacheRingEntries < entries)
	{
		GI_CUDA(c, cudaStreamSynchronize(c->stream));
		cudaFree(c->cacheRing); c->cacheRing = nullptr; c->cacheRingEntries = 0;
		size_t freeB = 0, totalB = 0;
		GI_CUDA(c, cudaMemGetInfo(&freeB, &totalB));
		// the two queue sets of the runs are allocated after this: leave room for them (and 2 GB of slack)
		const uint64_t numPix = uint64_t(c->width) * c->height;
		const uint64_t queues = (uint64_t(numLightPaths) + numPix) * (uint64_t(maxLen) * 120u + uint64_t(maxLen + 1u) * 36u);
		if (entries * 64u + queues + (2ull << 30) > freeB) return fail(c, GI_ERR_CAPACITY, "pipelined BPT: the light-cache ring does not fit the device memory");
		GI_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->cacheRing), size_t(entries) * 64u));
		c->cacheRingEntries = entries;
	}
	if (c->ringCountersLen < 2u * ring)
	{
		GI_CUDA(c, cudaStreamSynchronize(c->stream));
		cudaFree(c->ringCounters); c->ringCounters = nullptr; c->ringCountersLen = 0;
		GI_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->ringCounters), size_t(2u * ring) * sizeof(uint32_t)));
		c->ringCountersLen = 2u * ring;
	}
	DFrame f = makeFrame(c, p);
	if (streaming) f.lightSamples = c->lightSamples; // ring base (see gi_render_iteration)
	f.lightCache = c->cacheRing;
	f.lightCacheCapacity = capacity;
	f.lightCacheCount = c->ringCounters;
	f.lightPathLengthSum = c->ringCounters + 1;
	f.cacheRing = ring;
	// a stream can be read while it runs (gi_read_accumulator_async): the camera-connection splats of iteration k wait in their own image
	// until the eye radiance of iteration k joins the accumulator (global wave k + D + maxLen + 1), so a snapshot never holds half an iteration
	int st = wfOpen(c, 0, f, WF_PASS_LIGHT, WF_LIGHT_SAVE, expectedCount, streaming, lightStride, lightBase, lightRing, streaming ? 2u * maxLen + 2u : 0u, 2u * maxLen + 1u);
	if (st == GI_OK) st = wfOpen(c, 1, f, WF_PASS_EYE, 0u, expectedCount, streaming, 0u, 0u, 0u);
	if (st != GI_OK) { c->run[0].open = c->run[1].open = false; return st; }
	b.open = true; b.streaming = streaming;
	b.count = 0; b.wave = 0; b.expected = expectedCount;
	b.ring = ring; b.D = maxLen; b.maxLen = maxLen; b.capacity = capacity; b.numLightPaths = numLightPaths;
	b.numRaysPerLightSample = p->num_rays_per_light_sample;
	b.seeds.assign(size_t(maxLen) + 2u, 0u);
	return GI_OK;
}

// One global wave.  submit: a new iteration (FrameSeed `seed`) starts its light pass in this wave; otherwise the wave belongs to the
// drain of a pipeline that got `finalCount` iterations.
static int bptWave(GiContext* c, bool submit, uint32_t seed)
{
	BptPipe& b = c->bpt;
	const uint32_t w = b.wave, D = b.D, maxLen = b.maxLen;
	if (submit) { b.seeds[b.count % b.seeds.size()] = seed; ++b.count; }
	const uint32_t C = submit ? (b.streaming ? 0xffffffffu : b.expected) : b.count; // total iterations, if known
	const uint32_t runWaves = C == 0xffffffffu ? 0xffffffffu : C + maxLen;           // waves of one run; its last wave carries shadow rays only
	int st = GI_OK;
	if (w < runWaves)
	{
		if (submit)
		{
			// resetcounter.comp for the cache of this iteration (bidirectionalpathtracer.cpp:190)
			const uint32_t slot = w % b.ring;
			k_reset_counter<<<1, 1, 0, c->stream>>>(c->ringCounters + 2u * slot, c->ringCounters + 2u * slot + 1u, b.numLightPaths);
			++c->launches;
		}
		st = wfStep(c, 0, submit, seed, submit && w + 1u < C, runWaves == 0xffffffffu || w + 1u < runWaves);
	}
	if (st == GI_OK && w >= D)
	{
		const uint32_t e = w - D; // the eye pass of iteration e starts D waves after its light pass
		const bool gen = e < b.count;
		if (gen || e < runWaves)
			st = wfStep(c, 1, gen, gen ? b.seeds[e % b.seeds.size()] : 0u, e + 1u < (C == 0xffffffffu ? b.count + 1u : C), runWaves == 0xffffffffu || e + 1u < runWaves);
	}
	++b.wave;
	return st;
}

// Drain (maxPathLength + D more waves), fold, publish the last iteration's counters.
static int bptClose(GiContext* c)
{
	BptPipe& b = c->bpt;
	if (!b.open) return GI_OK;
	int st = GI_OK;
	if (b.count > 0)
	{
		const uint32_t end = b.count + b.maxLen + b.D;
		while (st == GI_OK && b.wave < end) st = bptWave(c, false, 0u);
	}
	if (st == GI_OK) st = wfClose(c, 0, false);
	if (st == GI_OK) st = wfClose(c, 1, false);
	c->run[0].open = c->run[1].open = false;
	b.open = false;
	if (st != GI_OK) return st;
	if (b.count > 0)
		GI_CUDA(c, cudaMemcpyAsync(c->bptCounters, c->ringCounters + 2u * ((b.count - 1u) % b.ring), 2 * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
	return GI_OK;
}

static int renderPipelinedBPT(GiContext* c, const GiIterationParams* p, const uint32_t* seeds, uint32_t count, uint32_t lightStride)
{
	int st = bptOpen(c, p, count, lightStride, 0u, 0u);
	if (st != GI_OK) return st;
	for (uint32_t k = 0; st == GI_OK && k < count; ++k) st = bptWave(c, true, seeds[k]);
	if (st == GI_OK) st = bptClose(c);
	c->bpt.open = false; c->run[0].open = c->run[1].open = false;
	return st;
}


// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {

int gi_version(void) { return 100; }

const char* gi_last_error(const GiContext* ctx) { return ctx ? ctx->error.c_str() : g_createError.c_str(); }

int gi_create(int device, GiContext** out_ctx)
{
	if (!out_ctx) return fail(nullptr, GI_ERR_INVALID, "gi_create: out_ctx is NULL");
	*out_ctx = nullptr;
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0)
		return fail(nullptr, GI_ERR_CUDA, std::string("no CUDA device available (") + cudaGetErrorString(e) + "); gpugi_b200 has no CPU fallback");
	if (device < 0 || device >= count) return fail(nullptr, GI_ERR_INVALID, "gi_create: device ordinal out of range");
	cudaDeviceProp prop;
	GI_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
	if (prop.major != 10)
		return fail(nullptr, GI_ERR_CUDA, std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
		                                      "; this library carries sm_100a code only");
	GI_CUDA(nullptr, cudaSetDevice(device));
	GiContext* c = new GiContext();
	c->device = device;
	e = cudaStreamCreateWithFlags(&c->ownStream, cudaStreamNonBlocking);
	if (e != cudaSuccess) { delete c; return fail(nullptr, GI_ERR_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e)); }
	c->stream = c->ownStream;
	bool ok = cudaMalloc(reinterpret_cast<void**>(&c->counters), CNT_COUNT * sizeof(unsigned long long)) == cudaSuccess &&
	          cudaMalloc(reinterpret_cast<void**>(&c->bptCounters), 2 * sizeof(uint32_t)) == cudaSuccess &&
	          cudaMalloc(reinterpret_cast<void**>(&c->mseResult), sizeof(double)) == cudaSuccess &&
	          cudaMemset(c->counters, 0, CNT_COUNT * sizeof(unsigned long long)) == cudaSuccess && cudaMemset(c->bptCounters, 0, 2 * sizeof(uint32_t)) == cudaSuccess;
	ok = ok && cudaMalloc(reinterpret_cast<void**>(&c->traceWork), sizeof(uint32_t)) == cudaSuccess;
	if (!ok) { gi_destroy(c); return fail(nullptr, GI_ERR_CUDA, "gi_create: device allocation failed"); }
	// persistent kernels: one resident wave, grid = SMs x resident blocks per SM
	c->numSMs = prop.multiProcessorCount;
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->traceBlocksPerSM, k_wf_trace<false, false>, WF_TRACE_THREADS, 0);
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->shadeBlocksPerSM, k_wf_shade_pt, WF_SHADE_THREADS, 0);
	c->traceBlocksPerSM = std::max(c->traceBlocksPerSM, 1);
	c->shadeBlocksPerSM = std::max(c->shadeBlocksPerSM, 1);
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->shadeLightBlocksPerSM, k_wf_shade_light, WF_SHADE_THREADS, 0);
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->shadeEyeBlocksPerSM, k_wf_shade_eye, WF_SHADE_THREADS, 0);
	c->shadeLightBlocksPerSM = std::max(c->shadeLightBlocksPerSM, 1);
	c->shadeEyeBlocksPerSM = std::max(c->shadeEyeBlocksPerSM, 1);
	// k_light_samples keeps up to 3 x 8192 draws in dynamic shared memory: above 48 KB a kernel has to opt in
	cudaFuncSetAttribute(k_light_samples, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 8192 * int(sizeof(float)));
	*out_ctx = c;
	return GI_OK;
}

void gi_destroy(GiContext* c)
{
	if (!c) return;
	cudaSetDevice(c->device);
	cudaDeviceSynchronize();
	gi_comm_destroy(c); // (uses the copy stream: before it is destroyed)
	freeScene(c);
	freeWavefront(c);
	cudaFree(c->traceWork);
	cudaFree(c->snapshot[0]); cudaFree(c->snapshot[1]);
	for (int i = 0; i < 2; ++i) { if (c->snapTaken[i]) cudaEventDestroy(c->snapTaken[i]); if (c->snapCopied[i]) cudaEventDestroy(c->snapCopied[i]); }
	if (c->copyStream) cudaStreamDestroy(c->copyStream);
	if (c->envTex) cudaDestroyTextureObject(c->envTex);
	if (c->envArray) cudaFreeArray(c->envArray);
	cudaFree(c->cacheRing); cudaFree(c->ringCounters); cudaFree(c->splatRing);
	for (void* p : c->lightTableAllocs) cudaFree(p);
	cudaFree(c->accum); cudaFree(c->lightSamples); cudaFree(c->lightCache); cudaFree(c->bptCounters); cudaFree(c->counters); cudaFree(c->mseResult);
	for (cudaEvent_t e : c->timingEvents) cudaEventDestroy(e);
	if (c->ownStream) cudaStreamDestroy(c->ownStream);
	delete c;
}

int gi_set_stream(GiContext* c, void* cuda_stream)
{
	if (!c) return GI_ERR_INVALID;
	int fst = flushStream(c);
	if (fst != GI_OK) return fst;
	GI_CUDA(c, cudaStreamSynchronize(c->stream));
	c->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c->ownStream;
	return GI_OK;
}

int gi_set_option(GiContext* c, int option, int value)
{
	if (!c) return GI_ERR_INVALID;
	int fst = flushStream(c);
	if (fst != GI_OK) return fst;
	switch (option)
	{
	case GI_OPT_STREAMING: c->streamingOption = value ? 1 : 0; return GI_OK;
	case GI_OPT_DEVICE_BUILDER:
		if (value != GI_BUILDER_LBVH && value != GI_BUILDER_PLOC) return fail(c, GI_ERR_INVALID, "GI_OPT_DEVICE_BUILDER: unknown builder");
		c->deviceBuilder = value; return GI_OK;
	case GI_OPT_PLOC_RADIUS: c->plocRadius = std::min(std::max(value, 1), PLOC_MAX_RADIUS); return GI_OK;
	case GI_OPT_UPLOAD_TOPOLOGY:
		if (value != GI_TOPOLOGY_REFERENCE && value != GI_TOPOLOGY_PLOC) return fail(c, GI_ERR_INVALID, "GI_OPT_UPLOAD_TOPOLOGY: unknown topology");
		c->uploadTopology = value; return GI_OK;
	case GI_OPT_TRAVERSAL:
		if (value != GI_TRAVERSAL_FAST && value != GI_TRAVERSAL_REFERENCE) return fail(c, GI_ERR_INVALID, "GI_OPT_TRAVERSAL: unknown mode");
		c->traversal = value; return GI_OK;
	case GI_OPT_TEXTURE_FILTER: c->textureFilterExact = value ? 1 : 0; return GI_OK;
	case GI_OPT_SCHEDULER:
		if (value != GI_SCHEDULER_WAVEFRONT && value != GI_SCHEDULER_MEGAKERNEL) return fail(c, GI_ERR_INVALID, "GI_OPT_SCHEDULER: unknown scheduler");
		c->scheduler = value; return GI_OK;
	case GI_OPT_COUNT_STEPS: c->countSteps = value ? 1 : 0; return GI_OK;
	case GI_OPT_RECONSTRUCTION_FILTER:
		if (value < 0 || value > 3) return fail(c, GI_ERR_INVALID, "GI_OPT_RECONSTRUCTION_FILTER: 0 (box) .. 3");
		c->reconFilter = value; return GI_OK;
	case GI_OPT_CHAINED_MIS: c->chainedMIS = value ? 1 : 0; return GI_OK;
	case GI_OPT_KERNEL_TIMING:
		c->kernelTiming = value ? 1 : 0;
		for (int i = 0; i < GI_KERNEL_CLASS_COUNT; ++i) { c->kernelMs[i] = 0.0; c->kernelLaunches[i] = 0; }
		return GI_OK;
	case GI_OPT_WF_REFILL_THRESHOLD: c->tune.refillThreshold = std::min(std::max(value, 1), 32); return GI_OK;
	case GI_OPT_WF_MIN_INNER: c->tune.minInner = std::min(std::max(value, 1), 32); return GI_OK;
	case GI_OPT_WF_SHADE_BLOCKS_PER_SM:
		if (value > 0) c->shadeBlocksPerSM = c->shadeLightBlocksPerSM = c->shadeEyeBlocksPerSM = value;
		else
		{
			cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->shadeBlocksPerSM, k_wf_shade_pt, WF_SHADE_THREADS, 0);
			cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->shadeLightBlocksPerSM, k_wf_shade_light, WF_SHADE_THREADS, 0);
			cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->shadeEyeBlocksPerSM, k_wf_shade_eye, WF_SHADE_THREADS, 0);
		}
		return GI_OK;
	case GI_OPT_WF_BLOCKS_PER_SM:
		if (value > 0) c->traceBlocksPerSM = value;
		else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->traceBlocksPerSM, k_wf_trace<false, false>, WF_TRACE_THREADS, 0);
		return GI_OK;
	}
	return fail(c, GI_ERR_INVALID, "gi_set_option: unknown option");
}

int gi_clear_accumulator(GiContext* c)
{
	if (!c) return GI_ERR_INVALID;
	int fst = flushStream(c);
	if (fst != GI_OK) return fst;
	if (c->accum) GI_CUDA(c, cudaMemsetAsync(c->accum, 0, size_t(c->width) * c->height * 4 * sizeof(float), c->stream));
	return GI_OK;
}

// GI_OPT_UPLOAD_TOPOLOGY = GI_TOPOLOGY_PLOC: re-cluster the reference's leaves on the device (gi_lbvh.cuh: Morton order of the leaf
// boxes + PLOC passes) and hand the tree back to the host encoder (gi_layout.h: applyLeafTree).  GI_OK with t.numLeaves == 0 when the
// scene has fewer than two leaves.
static int buildLeafTreeOnDevice(GiContext* c, const GiSceneDesc* d, const HostLayout& L, LeafTree& t)
{
	const uint32_t nl = d->num_leaf_triangles / 2u;
	t = LeafTree();
	if (nl < 2u) return GI_OK;
	const uint32_t N = 2u * nl - 1u;
	std::vector<float> hMin(size_t(nl) * 3), hMax(size_t(nl) * 3);
	for (uint32_t l = 0; l < nl; ++l)
	{
		const GiNode& ln = d->nodes[L.leafNode[l]];
		std::memcpy(&hMin[size_t(l) * 3], ln.min, 12); std::memcpy(&hMax[size_t(l) * 3], ln.max, 12);
	}
	cudaStream_t st = c->stream;
	std::vector<void*> allocs;
	auto dalloc = [&](size_t bytes) -> void* { void* p = nullptr; if (cudaMalloc(&p, std::max<size_t>(bytes, 16)) != cudaSuccess) return nullptr; allocs.push_back(p); return p; };
	auto freeAll = [&]() { for (void* p : allocs) cudaFree(p); };
	float* dLeafMin = static_cast<float*>(dalloc(size_t(nl) * 12)); float* dLeafMax = static_cast<float*>(dalloc(size_t(nl) * 12));
	unsigned long long* dCodes = static_cast<unsigned long long*>(dalloc(size_t(nl) * 8));
	unsigned long long* dCodesSorted = static_cast<unsigned long long*>(dalloc(size_t(nl) * 8));
	uint32_t* dIds = static_cast<uint32_t*>(dalloc(size_t(nl) * 4)); uint32_t* dIdsSorted = static_cast<uint32_t*>(dalloc(size_t(nl) * 4));
	PlocTree pt;
	pt.boxMin = static_cast<float*>(dalloc(size_t(N) * 12)); pt.boxMax = static_cast<float*>(dalloc(size_t(N) * 12));
	pt.left = static_cast<uint32_t*>(dalloc(size_t(N) * 4)); pt.right = static_cast<uint32_t*>(dalloc(size_t(N) * 4));
	pt.parent = static_cast<uint32_t*>(dalloc(size_t(N) * 4)); pt.size = static_cast<uint32_t*>(dalloc(size_t(N) * 4)); pt.leaves = static_cast<uint32_t*>(dalloc(size_t(N) * 4));
	uint32_t* clA = static_cast<uint32_t*>(dalloc(size_t(nl) * 4)); uint32_t* clB = static_cast<uint32_t*>(dalloc(size_t(nl) * 4));
	uint32_t* dNearest = static_cast<uint32_t*>(dalloc(size_t(nl) * 4)); uint32_t* dMerged = static_cast<uint32_t*>(dalloc(size_t(nl) * 4));
	uint32_t* dValid = static_cast<uint32_t*>(dalloc(size_t(nl) * 4)); uint32_t* dOffset = static_cast<uint32_t*>(dalloc(size_t(nl) * 4));
	uint32_t* dCounters = static_cast<uint32_t*>(dalloc(8)); // [0] next inner node id, [1] cluster count after the pass
	size_t sortBytes = 0, scanBytes = 0;
	cub::DeviceRadixSort::SortPairs(nullptr, sortBytes, dCodes, dCodesSorted, dIds, dIdsSorted, int(nl), 0, 63, st);
	cub::DeviceScan::ExclusiveSum(nullptr, scanBytes, dValid, dOffset, int(nl), st);
	void* dSortTmp = dalloc(sortBytes); void* dScanTmp = dalloc(scanBytes);
	if (!(dLeafMin && dLeafMax && dCodes && dCodesSorted && dIds && dIdsSorted && pt.boxMin && pt.boxMax && pt.left && pt.right && pt.parent && pt.size && pt.leaves && clA && clB &&
	      dNearest && dMerged && dValid && dOffset && dCounters && dSortTmp && dScanTmp))
	{
		freeAll();
		return fail(c, GI_ERR_CUDA, "gi_upload_scene: device allocation failed (topology build)");
	}
	cudaError_t e = cudaSuccess;
	auto step = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
	step(cudaMemcpyAsync(dLeafMin, hMin.data(), size_t(nl) * 12, cudaMemcpyHostToDevice, st));
	step(cudaMemcpyAsync(dLeafMax, hMax.data(), size_t(nl) * 12, cudaMemcpyHostToDevice, st));
	const uint32_t initCounters[2] = {nl, nl};
	step(cudaMemcpyAsync(dCounters, initCounters, 8, cudaMemcpyHostToDevice, st));
	const unsigned B = 256;
	if (e == cudaSuccess)
	{
		k_leafbox_morton<<<(nl + B - 1) / B, B, 0, st>>>(dLeafMin, dLeafMax, nl, f3(d->nodes[0].min), f3(d->nodes[0].max), dCodes, dIds);
		step(cub::DeviceRadixSort::SortPairs(dSortTmp, sortBytes, dCodes, dCodesSorted, dIds, dIdsSorted, int(nl), 0, 63, st));
		k_ploc_init_gather<<<(nl + B - 1) / B, B, 0, st>>>(nl, dIdsSorted, dLeafMin, dLeafMax, pt, clA);
		c->launches += 2;
		uint32_t n = nl;
		auto pass = [&](bool forced)
		{
			const unsigned g = (n + B - 1) / B;
			if (forced) k_ploc_force_pairs<<<g, B, 0, st>>>(n, dNearest);
			else k_ploc_nearest<<<g, B, 0, st>>>(n, clA, pt, dNearest, c->plocRadius);
			k_ploc_merge<<<g, B, 0, st>>>(n, clA, dNearest, pt, dCounters, dMerged, dValid);
			step(cub::DeviceScan::ExclusiveSum(dScanTmp, scanBytes, dValid, dOffset, int(n), st));
			k_ploc_compact<<<g, B, 0, st>>>(n, dMerged, dValid, dOffset, clB, dCounters + 1);
			c->launches += 3;
			uint32_t next = n;
			step(cudaMemcpyAsync(&next, dCounters + 1, 4, cudaMemcpyDeviceToHost, st));
			step(cudaStreamSynchronize(st));
			return next;
		};
		// A pass normally merges 35-45 % of the clusters.  Many coincident boxes (duplicated geometry) can leave a single mutual pair per pass;
		// a pass that removed less than 1/16 of a list of more than 1024 clusters is therefore followed by a forced pairing of neighbours, which halves the list.
		bool forceNext = false;
		while (n > 1u && e == cudaSuccess)
		{
			uint32_t next = pass(forceNext);
			forceNext = false;
			if (next >= n && e == cudaSuccess) next = pass(true);
			if (next >= n) { if (e == cudaSuccess) e = cudaErrorUnknown; break; }
			if (n > 1024u && uint64_t(n - next) * 16u < uint64_t(n)) forceNext = true; // (short lists near the root always merge slowly: leave them to the area rule)
			std::swap(clA, clB);
			n = next;
		}
		step(cudaGetLastError());
	}
	uint32_t counters[2] = {0, 0};
	if (e == cudaSuccess)
	{
		t.left.resize(N); t.right.resize(N); t.boxMin.resize(size_t(N) * 3); t.boxMax.resize(size_t(N) * 3); t.leafOf.resize(nl);
		step(cudaMemcpyAsync(t.left.data(), pt.left, size_t(N) * 4, cudaMemcpyDeviceToHost, st));
		step(cudaMemcpyAsync(t.right.data(), pt.right, size_t(N) * 4, cudaMemcpyDeviceToHost, st));
		step(cudaMemcpyAsync(t.boxMin.data(), pt.boxMin, size_t(N) * 12, cudaMemcpyDeviceToHost, st));
		step(cudaMemcpyAsync(t.boxMax.data(), pt.boxMax, size_t(N) * 12, cudaMemcpyDeviceToHost, st));
		step(cudaMemcpyAsync(t.leafOf.data(), dIdsSorted, size_t(nl) * 4, cudaMemcpyDeviceToHost, st));
		step(cudaMemcpyAsync(counters, dCounters, 8, cudaMemcpyDeviceToHost, st));
		step(cudaStreamSynchronize(st));
	}
	freeAll();
	if (e != cudaSuccess) return fail(c, GI_ERR_CUDA, std::string("gi_upload_scene (topology build): ") + cudaGetErrorString(e));
	if (counters[0] != N) return fail(c, GI_ERR_CUDA, "gi_upload_scene (topology build): the clustering did not end in one root");
	t.numLeaves = nl;
	t.root = N - 1u; // the last merge
	return GI_OK;
}

int gi_upload_scene(GiContext* c, const GiSceneDesc* d)
{
	if (!c || !d) return GI_ERR_INVALID;
	if (!d->nodes || d->num_nodes == 0 || !d->leaf_triangles || d->num_leaf_triangles == 0 || (d->num_leaf_triangles & 1u) || !d->positions ||
	    !d->vertex_info || d->num_vertices == 0 || !d->materials || d->num_materials == 0)
		return fail(c, GI_ERR_INVALID, "gi_upload_scene: incomplete scene description");
	GI_CUDA(c, cudaSetDevice(c->device));
	int fst = flushStream(c);
	if (fst != GI_OK) return fst;
	GI_CUDA(c, cudaStreamSynchronize(c->stream));
	freeScene(c);

	// ---- validate + re-lay out on the host (gi_layout.h), then upload
	HostLayout L;
	std::string lerr;
	if (!buildHostLayout(d, L, lerr)) return fail(c, GI_ERR_INVALID, "gi_upload_scene: " + lerr);
	c->fastValid = L.binary;
	c->sceneTopology = GI_TOPOLOGY_REFERENCE;
	c->topologyBuildMs = 0.0;
	if (L.binary && c->uploadTopology == GI_TOPOLOGY_PLOC && d->num_leaf_triangles >= 4u)
	{
		const auto t0 = std::chrono::steady_clock::now();
		LeafTree tree;
		const int tst = buildLeafTreeOnDevice(c, d, L, tree);
		if (tst != GI_OK) return tst;
		std::string why;
		if (applyLeafTree(d, L, tree, why)) c->sceneTopology = GI_TOPOLOGY_PLOC; // else: the reference's own topology stays (always valid)
		c->topologyBuildMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
	}

	// ---- textures: CUDA arrays + texture objects (hardware filter) and a linear copy (fp32 filter)
	std::vector<cudaTextureObject_t> objOfTexture(d->num_textures, 0);
	std::vector<const void*> rawOfTexture(d->num_textures, nullptr);
	for (uint32_t ti = 0; ti < d->num_textures; ++ti)
	{
		const GiTexture& t = d->textures[ti];
		if (t.width == 1 && t.height == 1) continue; // folded into the material record
		const bool isFloat = t.format == GI_TEX_RGBA32F;
		const size_t texelBytes = isFloat ? 16 : 4;
		cudaChannelFormatDesc cd = isFloat ? cudaCreateChannelDesc<float4>() : cudaCreateChannelDesc<uchar4>();
		cudaArray_t arr = nullptr;
		GI_CUDA(c, cudaMallocArray(&arr, &cd, size_t(t.width), size_t(t.height)));
		c->arrays.push_back(arr);
		GI_CUDA(c, cudaMemcpy2DToArrayAsync(arr, 0, 0, t.pixels, t.width * texelBytes, t.width * texelBytes, size_t(t.height), cudaMemcpyHostToDevice, c->stream));
		cudaResourceDesc rd;
		std::memset(&rd, 0, sizeof(rd));
		rd.resType = cudaResourceTypeArray;
		rd.res.array.array = arr;
		cudaTextureDesc td;
		std::memset(&td, 0, sizeof(td));
		td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeWrap; // REPEAT (scene.cpp:20)
		td.filterMode = cudaFilterModeLinear;                                             // mag LINEAR, LOD 0
		td.readMode = isFloat ? cudaReadModeElementType : cudaReadModeNormalizedFloat;
		td.sRGB = t.format == GI_TEX_SRGB8_ALPHA8 ? 1 : 0;
		td.normalizedCoords = 1;
		cudaTextureObject_t obj = 0;
		GI_CUDA(c, cudaCreateTextureObject(&obj, &rd, &td, nullptr));
		c->texObjects.push_back(obj);
		objOfTexture[ti] = obj;
		void* raw = nullptr;
		GI_CUDA(c, cudaMalloc(&raw, size_t(t.width) * t.height * texelBytes));
		c->sceneAllocs.push_back(raw);
		GI_CUDA(c, cudaMemcpyAsync(raw, t.pixels, size_t(t.width) * t.height * texelBytes, cudaMemcpyHostToDevice, c->stream));
		rawOfTexture[ti] = raw;
	}
	for (uint32_t mi = 0; mi < d->num_materials; ++mi)
		for (int sl = 0; sl < 4; ++sl)
		{
			const int32_t ti = L.materialTexture[size_t(mi) * 4 + sl];
			L.texels[size_t(mi) * 4 + sl].pixels = rawOfTexture[ti];
			L.materials[mi].tex[sl] = (unsigned long long)objOfTexture[ti]; // 0 for folded constants
		}

	// ---- upload
	DScene s = L.scene;
	GI_CUDA(c, uploadVector(c, L.refNodes, &s.refNodes));
	if (!L.oboxNodes.empty()) GI_CUDA(c, uploadVector(c, L.oboxNodes, &s.oboxNodes));
	GI_CUDA(c, uploadVector(c, L.fastNodes, &s.fastNodes));
	GI_CUDA(c, uploadVector(c, L.qNodes, &s.qNodes));
	c->residentBytes[0] = L.qNodes.size() * sizeof(uint32_t);
	c->residentBytes[1] = L.leafPos.size() * sizeof(F4);
	c->residentBytes[2] = L.refNodes.size() * sizeof(F4);
	c->residentBytes[3] = L.leafTris.size() * sizeof(I4) + L.vertexInfo.size() * sizeof(F4) + L.materials.size() * sizeof(DMaterial);
	GI_CUDA(c, uploadVector(c, L.leafPos, &s.leafPos));
	GI_CUDA(c, uploadVector(c, L.leafTris, &s.leafTris));
	GI_CUDA(c, uploadVector(c, L.vertexInfo, &s.vertexInfo));
	GI_CUDA(c, uploadVector(c, L.materials, &s.materials));
	GI_CUDA(c, uploadVector(c, L.texels, &s.texels));
	GI_CUDA(c, cudaStreamSynchronize(c->stream)); // host staging vectors die here
	c->scene = s;
	c->haveScene = true;
	return gi_clear_accumulator(c);
}

int gi_set_environment_map(GiContext* c, int size, const uint8_t* const* faces)
{
	if (!c) return GI_ERR_INVALID;
	GI_CUDA(c, cudaSetDevice(c->device));
	int fst = flushStream(c);
	if (fst != GI_OK) return fst;
	GI_CUDA(c, cudaStreamSynchronize(c->stream));
	if (c->envTex) { cudaDestroyTextureObject(c->envTex); c->envTex = 0; }
	if (c->envArray) { cudaFreeArray(c->envArray); c->envArray = nullptr; }
	c->envColor = f3(0.0f);
	if (!faces || size <= 0) return GI_OK;
	if (size == 1)
	{
		// a 1x1 cube is a constant per face; the reference's default is all-black (renderersystem.cpp:39-47).
		bool same = true;
		for (int f = 1; f < 6; ++f) same = same && std::memcmp(faces[0], faces[f], 3) == 0;
		if (same) { c->envColor = f3(srgbDecodeHost(faces[0][0]), srgbDecodeHost(faces[0][1]), srgbDecodeHost(faces[0][2])); return GI_OK; }
	}
	cudaChannelFormatDesc cd = cudaCreateChannelDesc<uchar4>();
	GI_CUDA(c, cudaMalloc3DArray(&c->envArray, &cd, make_cudaExtent(size_t(size), size_t(size), 6), cudaArrayCubemap));
	// CUDA cubemap layer order: +X,-X,+Y,-Y,+Z,-Z; ours: -X,+X,-Y,+Y,-Z,+Z
	const int order[6] = {1, 0, 3, 2, 5, 4};
	std::vector<uint8_t> staging(size_t(size) * size * 4 * 6);
	for (int l = 0; l < 6; ++l) std::memcpy(staging.data() + size_t(l) * size * size * 4, faces[order[l]], size_t(size) * size * 4);
	cudaMemcpy3DParms cp;
	std::memset(&cp, 0, sizeof(cp));
	cp.srcPtr = make_cudaPitchedPtr(staging.data(), size_t(size) * 4, size_t(size), size_t(size));
	cp.dstArray = c->envArray;
	cp.extent = make_cudaExtent(size_t(size), size_t(size), 6);
	cp.kind = cudaMemcpyHostToDevice;
	GI_CUDA(c, cudaMemcpy3D(&cp));
	cudaResourceDesc rd;
	std::memset(&rd, 0, sizeof(rd));
	rd.resType = cudaResourceTypeArray;
	rd.res.array.array = c->envArray;
	cudaTextureDesc td;
	std::memset(&td, 0, sizeof(td));
	td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
	td.filterMode = cudaFilterModeLinear;
	td.readMode = cudaReadModeNormalizedFloat;
	td.sRGB = 1;
	td.normalizedCoords = 1;
	GI_CUDA(c, cudaCreateTextureObject(&c->envTex, &rd, &td, nullptr));
	return GI_OK;
}

int gi_set_camera(GiContext* c, const GiCamera* cam)
{
	if (!c || !cam) return GI_ERR_INVALID;
	int fst = flushStream(c);
	if (fst != GI_OK) return fst;
	c->cam = *cam;
	c->haveCamera = true;
	return gi_clear_accumulator(c);
}

int gi_set_screen_size(GiContext* c, int width, int height)
{
	if (!c || width <= 0 || height <= 0) return fail(c, GI_ERR_INVALID, "gi_set_screen_size: bad size");
	GI_CUDA(c, cudaSetDevice(c->device));
	int fst = flushStream(c);
	if (fst != GI_OK) return fst;
	if (c->copyStream) GI_CUDA(c, cudaStreamSynchronize(c->copyStream));
	GI_CUDA(c, cudaStreamSynchronize(c->stream));
	if (c->accum) { cudaFree(c->accum); c->accum = nullptr; }
	c->width = width; c->height = height;
	GI_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->accum), size_t(width) * height * 4 * sizeof(float)));
	return gi_clear_accumulator(c);
}

// Light-sample sets live in a ring of GI_LIGHT_RING sets on the device: a stream (GI_OPT_STREAMING) has up to maxPathLength + 1
// iterations in flight, each shading with its own set.  Outside a stream the ring position does not move.
static int lightSlotForNextIteration(GiContext* c, uint32_t n, uint32_t* slot)
{
	if (c->run[0].open && c->run[0].streaming && int(n) != c->numLightSamples)
	{
		int st = flushStream(c);
		if (st != GI_OK) return st;
	}
	if (n > c->lightSampleCapacity)
	{
		int st = flushStream(c);
		if (st != GI_OK) return st;
		GI_CUDA(c, cudaStreamSynchronize(c->stream));
		if (c->lightSamples) cudaFree(c->lightSamples);
		c->lightSamples = nullptr;
		GI_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->lightSamples), size_t(n) * GI_LIGHT_RING * sizeof(GiLightSample)));
		c->lightSampleCapacity = n;
	}
	*slot = (c->run[0].open && c->run[0].streaming) ? (c->run[0].w.lightBase + c->run[0].count) % GI_LIGHT_RING : c->lightSlot;
	return GI_OK;
}

// Streams: the set just installed belongs to the iteration the next gi_render_iteration call submits.
static void markLightSet(GiContext* c) { c->lightSetIteration = (c->run[0].open && c->run[0].streaming) ? c->run[0].count : 0u; }

// A caller who installs ONE set and then renders many iterations gets that set for all of them outside a stream (the ring position does not
// move).  Inside a stream every iteration reads its own ring slot: when no set was installed for the iteration about to be submitted, the
// previous iteration's set is copied forward (device to device, stream ordered), so both modes render the same thing.
static int carryLightSetForward(GiContext* c)
{
	const WfRun& r = c->run[0];
	if (!(r.open && r.streaming) || r.count == 0u || c->lightSetIteration == r.count) return GI_OK;
	const size_t setElems = 2u * size_t(c->numLightSamples);
	const uint32_t src = (r.w.lightBase + r.count - 1u) % GI_LIGHT_RING, dst = (r.w.lightBase + r.count) % GI_LIGHT_RING;
	GI_CUDA(c, cudaMemcpyAsync(c->lightSamples + size_t(dst) * setElems, c->lightSamples + size_t(src) * setElems, setElems * sizeof(F4), cudaMemcpyDeviceToDevice, c->stream));
	c->lightSetIteration = r.count;
	return GI_OK;
}

int gi_set_light_samples(GiContext* c, const GiLightSample* samples, uint32_t n)
{
	if (!c || !samples || n == 0) return fail(c, GI_ERR_INVALID, "gi_set_light_samples: no samples");
	GI_CUDA(c, cudaSetDevice(c->device));
	uint32_t slot = 0;
	int st = lightSlotForNextIteration(c, n, &slot);
	if (st != GI_OK) return st;
	// pageable source: the runtime stages the bytes before returning, so the caller may reuse `samples`;
	// stream order keeps the previous iteration's reads ahead of this write.
	GI_CUDA(c, cudaMemcpyAsync(c->lightSamples + size_t(slot) * 2u * n, samples, size_t(n) * sizeof(GiLightSample), cudaMemcpyHostToDevice, c->stream));
	c->lightSlot = slot;
	c->externalLightSamples = nullptr;
	c->numLightSamples = int(n);
	markLightSet(c);
	return GI_OK;
}

int gi_set_light_samples_device(GiContext* c, const GiLightSample* d_samples, uint32_t n)
{
	if (!c || !d_samples || n == 0) return fail(c, GI_ERR_INVALID, "gi_set_light_samples_device: no samples");
	if (c->streamingOption)
	{
		// a stream keeps using the set after this call returns: copy it into the ring
		GI_CUDA(c, cudaSetDevice(c->device));
		uint32_t slot = 0;
		int st = lightSlotForNextIteration(c, n, &slot);
		if (st != GI_OK) return st;
		GI_CUDA(c, cudaMemcpyAsync(c->lightSamples + size_t(slot) * 2u * n, d_samples, size_t(n) * sizeof(GiLightSample), cudaMemcpyDeviceToDevice, c->stream));
		c->lightSlot = slot;
		c->externalLightSamples = nullptr;
		c->numLightSamples = int(n);
		markLightSet(c);
		return GI_OK;
	}
	c->externalLightSamples = reinterpret_cast<const F4*>(d_samples);
	c->numLightSamples = int(n);
	return GI_OK;
}

int gi_set_light_tables(GiContext* c, const GiLightTables* t)
{
	if (!c || !t) return GI_ERR_INVALID;
	if ((t->num_light_triangles && (!t->light_triangles || !t->summed_area)) || (t->num_point_lights && (!t->point_positions || !t->point_intensities || !t->point_summed_flux)))
		return fail(c, GI_ERR_INVALID, "gi_set_light_tables: incomplete tables");
	GI_CUDA(c, cudaSetDevice(c->device));
	int fst = flushStream(c);
	if (fst != GI_OK) return fst;
	GI_CUDA(c, cudaStreamSynchronize(c->stream));
	for (void* p : c->lightTableAllocs) cudaFree(p);
	c->lightTableAllocs.clear();
	c->lightTables = DLightTables{};
	auto up = [&](const void* src, size_t bytes, const float** dst) -> cudaError_t
	{
		void* d = nullptr;
		cudaError_t e = cudaMalloc(&d, std::max<size_t>(bytes, 16));
		if (e != cudaSuccess) return e;
		c->lightTableAllocs.push_back(d);
		if (bytes) e = cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, c->stream);
		*dst = static_cast<const float*>(d);
		return e;
	};
	DLightTables d = {};
	GI_CUDA(c, up(t->light_triangles, size_t(t->num_light_triangles) * sizeof(GiLightTriangle), &d.tri));
	GI_CUDA(c, up(t->summed_area, size_t(t->num_light_triangles) * 4, &d.summedArea));
	GI_CUDA(c, up(t->point_positions, size_t(t->num_point_lights) * 12, &d.pointPos));
	GI_CUDA(c, up(t->point_intensities, size_t(t->num_point_lights) * 12, &d.pointInt));
	GI_CUDA(c, up(t->point_summed_flux, size_t(t->num_point_lights) * 4, &d.pointSummedFlux));
	d.numTri = t->num_light_triangles; d.areaSum = t->light_area_sum; d.numPoint = t->num_point_lights;
	d.areaFlux = t->num_light_triangles ? t->total_area_flux : 0.0f;
	d.pointFlux = t->num_point_lights ? t->total_point_flux : 0.0f;
	GI_CUDA(c, cudaStreamSynchronize(c->stream));
	c->lightTables = d;
	return GI_OK;
}

int gi_generate_light_samples(GiContext* c, uint64_t* rng_state, uint32_t n, float position_bias)
{
	if (!c || !rng_state || n == 0) return fail(c, GI_ERR_INVALID, "gi_generate_light_samples: no samples");
	if (n > 8192u) return fail(c, GI_ERR_CAPACITY, "gi_generate_light_samples: more than 8192 samples per iteration"); // 3 draws x 4 B x 8192 = 96 KB of shared memory (opted in at gi_create)
	GI_CUDA(c, cudaSetDevice(c->device));
	uint32_t slot = 0;
	int st = lightSlotForNextIteration(c, n, &slot);
	if (st != GI_OK) return st;
	const DLightTables& t = c->lightTables;
	k_light_samples<<<1, 256, size_t(3u * n) * sizeof(float), c->stream>>>(t, (unsigned long long)*rng_state, n, position_bias, c->lightSamples + size_t(slot) * 2u * n);
	++c->launches;
	GI_CUDA(c, cudaGetLastError());
	// the host copy of the stream advances exactly as the kernel's (lightsampler.cpp:33-35: 3 draws per area sample, 1 per point sample)
	if (t.areaFlux + t.pointFlux != 0.0f)
	{
		const float pointLightPercentage = t.pointFlux / (t.areaFlux + t.pointFlux);
		const uint32_t numPoint = uint32_t(pointLightPercentage * float(n));
		const uint32_t numDraws = 3u * (n - numPoint) + numPoint;
		uint64_t s = *rng_state;
		for (uint32_t i = 0; i < numDraws; ++i) { s ^= s >> 12; s ^= s << 25; s ^= s >> 27; s *= 2685821657736338717ull; }
		*rng_state = s;
	}
	c->lightSlot = slot;
	c->externalLightSamples = nullptr;
	c->numLightSamples = int(n);
	markLightSet(c);
	return GI_OK;
}

int gi_get_light_samples(GiContext* c, GiLightSample* out, uint32_t n)
{
	if (!c || !out) return GI_ERR_INVALID;
	if (int(n) != c->numLightSamples || n == 0) return fail(c, GI_ERR_INVALID, "gi_get_light_samples: n differs from the current set");
	const F4* src = c->externalLightSamples ? c->externalLightSamples : c->lightSamples + size_t(c->lightSlot) * 2u * n;
	GI_CUDA(c, cudaMemcpyAsync(out, src, size_t(n) * sizeof(GiLightSample), cudaMemcpyDeviceToHost, c->stream));
	GI_CUDA(c, cudaStreamSynchronize(c->stream));
	return GI_OK;
}

int gi_get_light_cache(GiContext* c, void* entries, uint32_t max_entries, uint32_t* num_entries)
{
	if (!c || !num_entries) return GI_ERR_INVALID;
	int fst = flushStream(c);
	if (fst != GI_OK) return fst;
	if (!c->lightCache) return fail(c, GI_ERR_INVALID, "gi_get_light_cache: no light cache allocated");
	GI_CUDA(c, cudaSetDevice(c->device));
	uint32_t counters[2] = {0u, 0u};
	GI_CUDA(c, cudaMemcpyAsync(counters, c->bptCounters, sizeof(counters), cudaMemcpyDeviceToHost, c->stream));
	GI_CUDA(c, cudaStreamSynchronize(c->stream));
	*num_entries = counters[0];
	const uint32_t n = std::min(std::min(counters[0], c->lightCacheCapacity), max_entries);
	if (entries && n)
	{
		GI_CUDA(c, cudaMemcpyAsync(entries, c->lightCache, size_t(n) * 64u, cudaMemcpyDeviceToHost, c->stream));
		GI_CUDA(c, cudaStreamSynchronize(c->stream));
	}
	return GI_OK;
}

int gi_alloc_light_cache(GiContext* c, uint32_t capacity)
{
	if (!c || capacity == 0) return fail(c, GI_ERR_INVALID, "gi_alloc_light_cache: zero capacity");
	GI_CUDA(c, cudaSetDevice(c->device));
	int fst = flushStream(c);
	if (fst != GI_OK) return fst;
	GI_CUDA(c, cudaStreamSynchronize(c->stream));
	if (c->lightCache) { cudaFree(c->lightCache); c->lightCache = nullptr; c->lightCacheCapacity = 0; }
	GI_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->lightCache), size_t(capacity) * 64));
	c->lightCacheCapacity = capacity;
	return GI_OK;
}

static int checkRenderState(GiContext* c)
{
	if (!c->haveScene) return fail(c, GI_ERR_NOSCENE, "no scene uploaded");
	if (!c->accum) return fail(c, GI_ERR_INVALID, "gi_set_screen_size was not called");
	if (!c->haveCamera) return fail(c, GI_ERR_INVALID, "gi_set_camera was not called");
	if (c->numLightSamples <= 0) return fail(c, GI_ERR_INVALID, "gi_set_light_samples was not called");
	return GI_OK;
}

int gi_render_iteration(GiContext* c, const GiIterationParams* p)
{
	if (!c || !p) return GI_ERR_INVALID;
	int st = checkRenderState(c);
	if (st != GI_OK) return st;
	if (p->max_path_length <= 0) return fail(c, GI_ERR_INVALID, "max_path_length must be positive");
	if (c->chainedMIS && p->renderer == GI_RENDERER_BPT && p->max_path_length > GI_MAX_PATHLENGTH)
		return fail(c, GI_ERR_INVALID, "GI_OPT_CHAINED_MIS: max_path_length must not exceed 64");
	GI_CUDA(c, cudaSetDevice(c->device));
	// Streams and the light-sample ring (GI_LIGHT_RING sets): iteration k reads its set in the waves that shade it with next-event estimation (PT:
	// waves k .. k + maxPathLength) or that start its light paths (LPT / BPT: generate wave k only -- light vertices carry their flux, and the BPT
	// eye pass takes its light vertices from the cache, never from InitialLightSamples).  maxPathLength + 1 <= GI_LIGHT_RING therefore covers all
	// three, although a BPT stream keeps 2 x maxPathLength + 1 iterations in flight.
	const bool canStream = c->streamingOption && c->scheduler == GI_SCHEDULER_WAVEFRONT && !c->externalLightSamples && uint32_t(p->max_path_length) + 1u <= GI_LIGHT_RING;
	// the wave index is an int inside the kernels: a stream that has been open for 2^30 waves is drained and reopened
	if (canStream && ((c->run[0].open && c->run[0].streaming && c->run[0].wave >= 0x40000000u) || (c->bpt.open && c->bpt.wave >= 0x40000000u)))
	{
		st = flushStream(c);
		if (st != GI_OK) return st;
	}
	if (canStream && (p->renderer == GI_RENDERER_PT || (p->renderer == GI_RENDERER_LPT && p->num_rays_per_light_sample > 0)))
	{
		// GI_OPT_STREAMING: this iteration is one more wave of the open run (opened here if there is none)
		WfRun& r = c->run[0];
		if (c->bpt.open || (r.open && r.streaming && (r.renderer != p->renderer || r.maxLen != uint32_t(p->max_path_length) ||
		                                              (p->renderer == GI_RENDERER_LPT && r.numRaysPerLightSample != p->num_rays_per_light_sample))))
		{
			st = flushStream(c);
			if (st != GI_OK) return st;
		}
		if (!(r.open && r.streaming))
		{
			DFrame sf = makeFrame(c, p);
			sf.lightSamples = c->lightSamples; // ring base; iteration i of the stream reads set (lightSlot + i) % GI_LIGHT_RING
			// LPT: the splats of every iteration in flight go to that iteration's own image and join the accumulator when its last camera
			// connection has been traced (wave k + maxPathLength), so that a progressive read-out only ever sees complete iterations
			if (p->renderer == GI_RENDERER_PT) st = wfOpen(c, 0, sf, WF_PASS_PT, 0u, 0u, true, 2u * uint32_t(c->numLightSamples), c->lightSlot, GI_LIGHT_RING);
			else st = wfOpen(c, 0, sf, WF_PASS_LIGHT, 0u, 0u, true, 2u * uint32_t(c->numLightSamples), c->lightSlot, GI_LIGHT_RING, uint32_t(p->max_path_length) + 1u, uint32_t(p->max_path_length));
			if (st != GI_OK) return st;
			r.renderer = p->renderer;
			r.numRaysPerLightSample = p->num_rays_per_light_sample;
			c->lightSetIteration = 0u;
		}
		st = carryLightSetForward(c);
		if (st != GI_OK) return st;
		st = wfStep(c, 0, true, p->frame_seed, true, true);
		if (st != GI_OK) return st;
		GI_CUDA(c, cudaGetLastError());
		return GI_OK;
	}
	if (canStream && !c->chainedMIS && p->renderer == GI_RENDERER_BPT && !p->warmup && p->num_rays_per_light_sample > 0 && p->light_cache_capacity > 0 &&
	    uint64_t(p->light_cache_capacity) >= uint64_t(c->numLightSamples) * uint64_t(p->num_rays_per_light_sample))
	{
		// streaming BPT: one global wave of the light / eye pipeline per call (renderPipelinedBPT's schedule, kept open)
		BptPipe& b = c->bpt;
		if ((c->run[0].open && c->run[0].streaming && !b.open) ||
		    (b.open && (b.maxLen != uint32_t(p->max_path_length) || b.capacity != uint32_t(p->light_cache_capacity) || b.numRaysPerLightSample != p->num_rays_per_light_sample)))
		{
			st = flushStream(c);
			if (st != GI_OK) return st;
		}
		if (!b.open)
		{
			st = bptOpen(c, p, 0u, 2u * uint32_t(c->numLightSamples), c->lightSlot, GI_LIGHT_RING);
			if (st != GI_OK && st != GI_ERR_CAPACITY) return st;
			if (st == GI_OK) c->lightSetIteration = 0u;
		}
		if (b.open)
		{
			st = carryLightSetForward(c);
			if (st != GI_OK) return st;
			st = bptWave(c, true, p->frame_seed);
			if (st != GI_OK) return st;
			GI_CUDA(c, cudaGetLastError());
			return GI_OK;
		}
		// the ring does not fit: one iteration at a time, below
	}
	st = flushStream(c);
	if (st != GI_OK) return st;
	DFrame f = makeFrame(c, p);
	const dim3 pixelGrid(unsigned(c->width / 8), unsigned(c->height / 8), 1); // integer division, pathtracer.cpp:50
	switch (p->renderer)
	{
	case GI_RENDERER_PT:
		if (c->scheduler == GI_SCHEDULER_WAVEFRONT)
		{
			st = renderWavefront(c, f, WF_PASS_PT, 0u, &p->frame_seed, 1u, 0u);
			if (st != GI_OK) return st;
		}
		else if (pixelGrid.x && pixelGrid.y)
		{
			k_pathtrace<<<pixelGrid, 64, 0, c->stream>>>(f, c->counters);
			++c->launches;
		}
		break;
	case GI_RENDERER_LPT:
	{
		if (p->num_rays_per_light_sample <= 0) return fail(c, GI_ERR_INVALID, "num_rays_per_light_sample must be positive");
		const unsigned groups = unsigned(c->numLightSamples) * unsigned(p->num_rays_per_light_sample) / 64u; // lightpathtracer.cpp:48
		if (c->scheduler == GI_SCHEDULER_WAVEFRONT)
		{
			st = renderWavefront(c, f, WF_PASS_LIGHT, 0u, &p->frame_seed, 1u, 0u);
			if (st != GI_OK) return st;
		}
		else if (groups)
		{
			k_lighttrace<false, false><<<groups, 64, 0, c->stream>>>(f, c->counters);
			++c->launches;
		}
		break;
	}
	case GI_RENDERER_BPT:
	{
		if (p->num_rays_per_light_sample <= 0) return fail(c, GI_ERR_INVALID, "num_rays_per_light_sample must be positive");
		const unsigned numLightPaths = unsigned(c->numLightSamples) * unsigned(p->num_rays_per_light_sample);
		const unsigned groups = numLightPaths / 64u;
		// CHAINED_MIS keeps g_mis[MAX_PATHLENGTH] per eye path and two more weights per light path: the one-thread-per-path kernels carry them
		const bool wavefront = c->scheduler == GI_SCHEDULER_WAVEFRONT && !(c->chainedMIS && !p->warmup);
		// resetcounter.comp (bidirectionalpathtracer.cpp:190)
		k_reset_counter<<<1, 1, 0, c->stream>>>(c->bptCounters, c->bptCounters + 1, numLightPaths);
		++c->launches;
		if (p->warmup)
		{
			f.lightCacheCapacity = 0x7fffffffu;
			if (wavefront)
			{
				st = renderWavefront(c, f, WF_PASS_LIGHT, WF_LIGHT_SAVE | WF_LIGHT_WARMUP, &p->frame_seed, 1u, 0u);
				if (st != GI_OK) return st;
			}
			else if (groups) { k_lighttrace<true, true><<<groups, 64, 0, c->stream>>>(f, c->counters); ++c->launches; }
			break;
		}
		if (!c->lightCache) return fail(c, GI_ERR_CAPACITY, "BPT needs gi_alloc_light_cache first");
		if (p->light_cache_capacity <= 0 || uint32_t(p->light_cache_capacity) > c->lightCacheCapacity)
			return fail(c, GI_ERR_CAPACITY, "light_cache_capacity exceeds the allocated light cache");
		if (uint32_t(p->light_cache_capacity) < numLightPaths)
			return fail(c, GI_ERR_CAPACITY, "light_cache_capacity is smaller than the number of light paths (their origins occupy the first slots)");
		f.lightCacheCapacity = uint32_t(p->light_cache_capacity);
		if (wavefront)
		{
			// light pass to completion (fills the cache), then the eye pass (bidirectionalpathtracer.cpp:190-200)
			st = renderWavefront(c, f, WF_PASS_LIGHT, WF_LIGHT_SAVE, &p->frame_seed, 1u, 0u);
			if (st == GI_OK) st = renderWavefront(c, f, WF_PASS_EYE, 0u, &p->frame_seed, 1u, 0u);
			if (st != GI_OK) return st;
			break;
		}
		if (groups) { k_lighttrace<true, false><<<groups, 64, 0, c->stream>>>(f, c->counters); ++c->launches; }
		if (pixelGrid.x && pixelGrid.y) { k_bidir_eye<<<pixelGrid, 64, 0, c->stream>>>(f, c->counters); ++c->launches; }
		break;
	}
	case GI_RENDERER_WHITTED:
		if (pixelGrid.x && pixelGrid.y)
		{
			k_whitted<<<pixelGrid, 64, 0, c->stream>>>(f, c->counters);
			++c->launches;
		}
		break;
	default: return fail(c, GI_ERR_INVALID, "unknown renderer");
	}
	GI_CUDA(c, cudaGetLastError());
	return GI_OK;
}

int gi_render_iterations(GiContext* c, const GiIterationParams* p, const uint32_t* frame_seeds, uint32_t count, const GiLightSample* d_light_samples,
                         uint32_t samples_per_iteration)
{
	if (!c || !p || (count && !frame_seeds)) return GI_ERR_INVALID;
	if (count == 0) return GI_OK;
	{
		int fst = flushStream(c);
		if (fst != GI_OK) return fst;
	}
	if (d_light_samples)
	{
		if (samples_per_iteration == 0) return fail(c, GI_ERR_INVALID, "gi_render_iterations: samples_per_iteration must be positive");
		c->externalLightSamples = reinterpret_cast<const F4*>(d_light_samples);
		c->numLightSamples = int(samples_per_iteration);
	}
	int st = checkRenderState(c);
	if (st != GI_OK) return st;
	if (p->max_path_length <= 0) return fail(c, GI_ERR_INVALID, "max_path_length must be positive");
	GI_CUDA(c, cudaSetDevice(c->device));
	if ((p->renderer == GI_RENDERER_PT || p->renderer == GI_RENDERER_LPT) && c->scheduler == GI_SCHEDULER_WAVEFRONT)
	{
		// PT / LPT iterations are independent: their waves overlap.  (BPT: the eye pass of an iteration needs its own complete
		// light cache -> one iteration after the other, below.)
		if (p->renderer == GI_RENDERER_LPT && p->num_rays_per_light_sample <= 0) return fail(c, GI_ERR_INVALID, "num_rays_per_light_sample must be positive");
		DFrame f = makeFrame(c, p);
		st = renderWavefront(c, f, p->renderer == GI_RENDERER_PT ? WF_PASS_PT : WF_PASS_LIGHT, 0u, frame_seeds, count, d_light_samples ? 2u * samples_per_iteration : 0u);
		if (st != GI_OK) return st;
		GI_CUDA(c, cudaGetLastError());
		return GI_OK;
	}
	if (p->renderer == GI_RENDERER_BPT && c->scheduler == GI_SCHEDULER_WAVEFRONT && !c->chainedMIS && count > 1u && !p->warmup && p->num_rays_per_light_sample > 0 &&
	    p->light_cache_capacity > 0 && uint64_t(p->light_cache_capacity) >= uint64_t(c->numLightSamples) * uint64_t(p->num_rays_per_light_sample))
	{
		// BPT: light and eye passes of consecutive iterations pipelined over a ring of light caches
		st = renderPipelinedBPT(c, p, frame_seeds, count, d_light_samples ? 2u * samples_per_iteration : 0u);
		if (st == GI_OK) { GI_CUDA(c, cudaGetLastError()); return GI_OK; }
		if (st != GI_ERR_CAPACITY) return st;
	}
	// every other renderer / scheduler: one iteration after the other
	for (uint32_t k = 0; k < count; ++k)
	{
		GiIterationParams q = *p;
		q.frame_seed = frame_seeds[k];
		if (d_light_samples) c->externalLightSamples = reinterpret_cast<const F4*>(d_light_samples + size_t(k) * samples_per_iteration);
		st = gi_render_iteration(c, &q);
		if (st != GI_OK) return st;
	}
	return GI_OK;
}

int gi_show_light_caches(GiContext* c)
{
	if (!c) return GI_ERR_INVALID;
	const WfRun& r0 = c->run[0];
	if (r0.open && r0.streaming && r0.pass == WF_PASS_LIGHT && r0.splatSlots && !c->bpt.open && r0.count > 0u && c->haveScene)
	{
		// LightPathtracer::Draw (lightpathtracer.cpp:45-50) shows the light samples of the iteration it has just started: inside a stream that is
		// iteration count - 1, whose splats are collected in its own image and whose samples sit in its own ring slot -- no drain
		GI_CUDA(c, cudaSetDevice(c->device));
		DFrame f = makeFrame(c, nullptr);
		const uint32_t k = r0.count - 1u;
		f.accum = c->splatRing + size_t(k % r0.splatSlots) * c->splatRingStride;
		f.lightSamples = c->lightSamples + size_t((r0.w.lightBase + k) % GI_LIGHT_RING) * 2u * size_t(c->numLightSamples);
		k_show_light_caches<<<(c->numLightSamples + 63) / 64, 64, 0, c->stream>>>(f, c->counters);
		++c->launches;
		GI_CUDA(c, cudaGetLastError());
		return GI_OK;
	}
	{
		int fst = flushStream(c);
		if (fst != GI_OK) return fst;
	}
	int st = checkRenderState(c);
	if (st != GI_OK) return st;
	GI_CUDA(c, cudaSetDevice(c->device));
	DFrame f = makeFrame(c, nullptr);
	k_show_light_caches<<<(c->numLightSamples + 63) / 64, 64, 0, c->stream>>>(f, c->counters);
	++c->launches;
	GI_CUDA(c, cudaGetLastError());
	return GI_OK;
}

int gi_trace_device(GiContext* c, const GiRay* d_rays, uint32_t n, GiHit* d_hits)
{
	if (!c || (n && (!d_rays || !d_hits))) return GI_ERR_INVALID;
	if (!c->haveScene) return fail(c, GI_ERR_NOSCENE, "no scene uploaded");
	if (n == 0) return GI_OK;
	GI_CUDA(c, cudaSetDevice(c->device));
	DScene s = c->scene;
	s.oboxOrdered = (s.oboxNodes && s.refTreeBinary && c->traversal == GI_TRAVERSAL_FAST) ? 1 : 0;
	const int mode = c->fastValid ? c->traversal : GI_TRAVERSAL_REFERENCE;
	if (mode == GI_TRAVERSAL_FAST && c->scheduler == GI_SCHEDULER_WAVEFRONT)
	{
		GI_CUDA(c, cudaMemsetAsync(c->traceWork, 0, sizeof(uint32_t), c->stream));
		const unsigned grid = std::min(unsigned(c->numSMs * c->traceBlocksPerSM), (n + WF_TRACE_THREADS - 1) / WF_TRACE_THREADS);
		if (c->countSteps) k_wf_trace_batch<true><<<grid, WF_TRACE_THREADS, 0, c->stream>>>(s, d_rays, n, d_hits, c->tune, c->traceWork, c->counters);
		else k_wf_trace_batch<false><<<grid, WF_TRACE_THREADS, 0, c->stream>>>(s, d_rays, n, d_hits, c->tune, c->traceWork, c->counters);
		++c->launches;
	}
	else
	{
		k_trace<<<(n + 127) / 128, 128, 0, c->stream>>>(s, d_rays, n, d_hits, mode, c->counters);
		++c->launches;
	}
	GI_CUDA(c, cudaGetLastError());
	return GI_OK;
}

int gi_trace(GiContext* c, const GiRay* rays, uint32_t n, GiHit* hits)
{
	if (!c || (n && (!rays || !hits))) return GI_ERR_INVALID;
	if (!c->haveScene) return fail(c, GI_ERR_NOSCENE, "no scene uploaded");
	if (n == 0) return GI_OK;
	GI_CUDA(c, cudaSetDevice(c->device));
	GiRay* dr = nullptr; GiHit* dh = nullptr;
	GI_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&dr), size_t(n) * sizeof(GiRay)));
	cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&dh), size_t(n) * sizeof(GiHit));
	if (e != cudaSuccess) { cudaFree(dr); return fail(c, GI_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e)); }
	int st = GI_OK;
	e = cudaMemcpyAsync(dr, rays, size_t(n) * sizeof(GiRay), cudaMemcpyHostToDevice, c->stream);
	if (e == cudaSuccess) st = gi_trace_device(c, dr, n, dh);
	if (e == cudaSuccess && st == GI_OK) e = cudaMemcpyAsync(hits, dh, size_t(n) * sizeof(GiHit), cudaMemcpyDeviceToHost, c->stream);
	if (e == cudaSuccess && st == GI_OK) e = cudaStreamSynchronize(c->stream);
	cudaFree(dr); cudaFree(dh);
	if (e != cudaSuccess) return fail(c, GI_ERR_CUDA, std::string("gi_trace: ") + cudaGetErrorString(e));
	return st;
}

// Device-side hierarchy build (gi_lbvh.cuh).  Host buffers in and out; the result is in the reference's layout.
int gi_build_hierarchy(GiContext* c, const float* positions, uint32_t num_vertices, const uint32_t* triangles, const uint32_t* triangle_materials,
                       uint32_t num_triangles, GiNode* out_nodes, GiTriangle* out_leaf_triangles, uint32_t* out_parents, uint32_t* out_num_nodes,
                       uint32_t* out_num_tree_levels)
{
	if (!c || !positions || !triangles || !out_nodes || !out_leaf_triangles || !out_parents || !out_num_nodes || !out_num_tree_levels) return GI_ERR_INVALID;
	if (num_vertices == 0 || num_triangles == 0) return fail(c, GI_ERR_INVALID, "gi_build_hierarchy: empty mesh");
	if (num_triangles > 0x7ffffff0u) return fail(c, GI_ERR_CAPACITY, "gi_build_hierarchy: too many triangles");
	for (size_t i = 0; i < size_t(num_triangles) * 3; ++i)
		if (triangles[i] >= num_vertices) return fail(c, GI_ERR_INVALID, "gi_build_hierarchy: vertex index out of range");
	GI_CUDA(c, cudaSetDevice(c->device));
	const uint32_t T = num_triangles;
	const bool ploc = c->deviceBuilder == GI_BUILDER_PLOC;
	const uint32_t Lcap = ploc ? T : (T + 1u) / 2u;   // PLOC: leaves by mutual nearest neighbours, anything between T/2 and T of them
	uint32_t L = (T + 1u) / 2u, N = 2u * L - 1u, I = L - 1u;
	cudaStream_t st = c->stream;
	std::vector<void*> allocs;
	auto dalloc = [&](size_t bytes) -> void* { void* p = nullptr; if (cudaMalloc(&p, std::max<size_t>(bytes, 16)) != cudaSuccess) return nullptr; allocs.push_back(p); return p; };
	auto freeAll = [&]() { for (void* p : allocs) cudaFree(p); };
	float* dPos = static_cast<float*>(dalloc(size_t(num_vertices) * 12));
	uint32_t* dTri = static_cast<uint32_t*>(dalloc(size_t(T) * 12));
	uint32_t* dMat = triangle_materials ? static_cast<uint32_t*>(dalloc(size_t(T) * 4)) : nullptr;
	unsigned* dBounds = static_cast<unsigned*>(dalloc(6 * 4));
	unsigned long long* dCodes = static_cast<unsigned long long*>(dalloc(size_t(T) * 8));
	unsigned long long* dCodesSorted = static_cast<unsigned long long*>(dalloc(size_t(T) * 8));
	uint32_t* dIds = static_cast<uint32_t*>(dalloc(size_t(T) * 4));
	uint32_t* dIdsSorted = static_cast<uint32_t*>(dalloc(size_t(T) * 4));
	unsigned long long* dLeafCodes = static_cast<unsigned long long*>(dalloc(size_t(Lcap) * 8));
	I4* dLeafTris = static_cast<I4*>(dalloc(size_t(Lcap) * 2 * 16));
	LbvhTree tr;
	tr.left = static_cast<uint32_t*>(dalloc(size_t(Lcap) * 4)); tr.right = static_cast<uint32_t*>(dalloc(size_t(Lcap) * 4));
	tr.parent = static_cast<uint32_t*>(dalloc(size_t(Lcap) * 4)); tr.first = static_cast<uint32_t*>(dalloc(size_t(Lcap) * 4)); tr.last = static_cast<uint32_t*>(dalloc(size_t(Lcap) * 4));
	tr.leafParent = static_cast<uint32_t*>(dalloc(size_t(Lcap) * 4));
	tr.boxMin = static_cast<float*>(dalloc(size_t(Lcap) * 12)); tr.boxMax = static_cast<float*>(dalloc(size_t(Lcap) * 12));
	tr.leafMin = static_cast<float*>(dalloc(size_t(Lcap) * 12)); tr.leafMax = static_cast<float*>(dalloc(size_t(Lcap) * 12));
	tr.visits = static_cast<uint32_t*>(dalloc(size_t(Lcap) * 4));
	uint32_t* dPreInner = static_cast<uint32_t*>(dalloc(size_t(Lcap) * 4));
	uint32_t* dPreLeaf = static_cast<uint32_t*>(dalloc(size_t(Lcap) * 4));
	uint32_t* dDepth = static_cast<uint32_t*>(dalloc(4));
	F4* dNodes = static_cast<F4*>(dalloc(size_t(2u * Lcap) * 32));
	uint32_t* dParents = static_cast<uint32_t*>(dalloc(size_t(2u * Lcap) * 4));
	size_t sortBytes = 0;
	cub::DeviceRadixSort::SortPairs(nullptr, sortBytes, dCodes, dCodesSorted, dIds, dIdsSorted, int(T), 0, 63, st);
	void* dSortTmp = dalloc(sortBytes);
	bool ok = dPos && dTri && (!triangle_materials || dMat) && dBounds && dCodes && dCodesSorted && dIds && dIdsSorted && dLeafCodes && dLeafTris && tr.left && tr.right &&
	          tr.parent && tr.first && tr.last && tr.leafParent && tr.boxMin && tr.boxMax && tr.leafMin && tr.leafMax && tr.visits && dPreInner && dPreLeaf && dDepth && dNodes &&
	          dParents && dSortTmp;
	if (!ok) { freeAll(); return fail(c, GI_ERR_CUDA, "gi_build_hierarchy: device allocation failed"); }
	cudaError_t e = cudaSuccess;
	auto step = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
	step(cudaMemcpyAsync(dPos, positions, size_t(num_vertices) * 12, cudaMemcpyHostToDevice, st));
	step(cudaMemcpyAsync(dTri, triangles, size_t(T) * 12, cudaMemcpyHostToDevice, st));
	if (dMat) step(cudaMemcpyAsync(dMat, triangle_materials, size_t(T) * 4, cudaMemcpyHostToDevice, st));
	const unsigned initBounds[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
	step(cudaMemcpyAsync(dBounds, initBounds, sizeof(initBounds), cudaMemcpyHostToDevice, st));
	step(cudaMemsetAsync(tr.visits, 0, std::max<size_t>(size_t(Lcap) * 4, 4), st));
	step(cudaMemsetAsync(dDepth, 0, 4, st));
	if (e == cudaSuccess)
	{
		const unsigned B = 256;
		k_lbvh_bounds<<<unsigned(c->numSMs) * 4u, B, 0, st>>>(dPos, num_vertices, dBounds);
		k_lbvh_morton<<<(T + B - 1) / B, B, 0, st>>>(dPos, dTri, T, dBounds, dCodes, dIds);
		step(cub::DeviceRadixSort::SortPairs(dSortTmp, sortBytes, dCodes, dCodesSorted, dIds, dIdsSorted, int(T), 0, 63, st));
		if (!ploc)
			k_lbvh_leaves<<<(L + B - 1) / B, B, 0, st>>>(dPos, dTri, dMat, dIdsSorted, dCodesSorted, T, L, dLeafTris, dLeafCodes, tr);
		else
		{
			// leaves by mutual nearest neighbours (gi_lbvh.cuh); scratch: the not yet used tree arrays
			float* tmn = static_cast<float*>(dalloc(size_t(T) * 12)); float* tmx = static_cast<float*>(dalloc(size_t(T) * 12));
			uint32_t* pNearest = static_cast<uint32_t*>(dalloc(size_t(T) * 4)); uint32_t* pValid = static_cast<uint32_t*>(dalloc(size_t(T) * 4));
			uint32_t* pOffset = static_cast<uint32_t*>(dalloc(size_t(T) * 4)); uint32_t* pCount = static_cast<uint32_t*>(dalloc(4));
			size_t pScanBytes = 0;
			cub::DeviceScan::ExclusiveSum(nullptr, pScanBytes, pValid, pOffset, int(T), st);
			void* pScanTmp = dalloc(pScanBytes);
			if (!(tmn && tmx && pNearest && pValid && pOffset && pCount && pScanTmp)) { freeAll(); return fail(c, GI_ERR_CUDA, "gi_build_hierarchy: device allocation failed"); }
			PlocTree view = {};
			view.boxMin = tmn; view.boxMax = tmx;
			k_pair_boxes<<<(T + B - 1) / B, B, 0, st>>>(dPos, dTri, dIdsSorted, T, tmn, tmx);
			k_ploc_nearest<<<(T + B - 1) / B, B, 0, st>>>(T, nullptr, view, pNearest);
			k_pair_flags<<<(T + B - 1) / B, B, 0, st>>>(T, pNearest, pValid);
			step(cub::DeviceScan::ExclusiveSum(pScanTmp, pScanBytes, pValid, pOffset, int(T), st));
			k_pair_emit<<<(T + B - 1) / B, B, 0, st>>>(T, dTri, dMat, dIdsSorted, dCodesSorted, pNearest, pValid, pOffset, tmn, tmx, dLeafTris, dLeafCodes, tr, pCount);
			c->launches += 4;
			step(cudaMemcpyAsync(&L, pCount, 4, cudaMemcpyDeviceToHost, st));
			step(cudaStreamSynchronize(st));
			if (e == cudaSuccess) { N = 2u * L - 1u; I = L - 1u; }
		}
		c->launches += 4;
		if (I > 0 && c->deviceBuilder == GI_BUILDER_LBVH)
		{
			k_lbvh_tree<<<(I + B - 1) / B, B, 0, st>>>(dLeafCodes, int(L), tr);
			k_lbvh_boxes<<<(L + B - 1) / B, B, 0, st>>>(int(L), tr);
			k_lbvh_preorder<<<(N + B - 1) / B, B, 0, st>>>(int(L), tr, dPreInner, dPreLeaf, dDepth);
			k_lbvh_emit<<<(N + B - 1) / B, B, 0, st>>>(int(L), tr, dPreInner, dPreLeaf, dNodes, dParents);
			c->launches += 4;
		}
		else if (I > 0)
		{
			// PLOC over the Morton-ordered leaves (gi_lbvh.cuh)
			PlocTree pt;
			pt.boxMin = static_cast<float*>(dalloc(size_t(N) * 12)); pt.boxMax = static_cast<float*>(dalloc(size_t(N) * 12));
			pt.left = static_cast<uint32_t*>(dalloc(size_t(N) * 4)); pt.right = static_cast<uint32_t*>(dalloc(size_t(N) * 4));
			pt.parent = static_cast<uint32_t*>(dalloc(size_t(N) * 4)); pt.size = static_cast<uint32_t*>(dalloc(size_t(N) * 4)); pt.leaves = static_cast<uint32_t*>(dalloc(size_t(N) * 4));
			uint32_t* clA = static_cast<uint32_t*>(dalloc(size_t(L) * 4)); uint32_t* clB = static_cast<uint32_t*>(dalloc(size_t(L) * 4));
			uint32_t* dNearest = static_cast<uint32_t*>(dalloc(size_t(L) * 4)); uint32_t* dMerged = static_cast<uint32_t*>(dalloc(size_t(L) * 4));
			uint32_t* dValid = static_cast<uint32_t*>(dalloc(size_t(L) * 4)); uint32_t* dOffset = static_cast<uint32_t*>(dalloc(size_t(L) * 4));
			uint32_t* dCounters = static_cast<uint32_t*>(dalloc(8)); // [0] next inner node id, [1] cluster count after the pass
			uint32_t* dPre = static_cast<uint32_t*>(dalloc(size_t(N) * 4)); uint32_t* dLeafIndex = static_cast<uint32_t*>(dalloc(size_t(L) * 4));
			I4* dLeafTrisOut = static_cast<I4*>(dalloc(size_t(L) * 2 * 16));
			size_t scanBytes = 0;
			cub::DeviceScan::ExclusiveSum(nullptr, scanBytes, dValid, dOffset, int(L), st);
			void* dScanTmp = dalloc(scanBytes);
			if (!(pt.boxMin && pt.boxMax && pt.left && pt.right && pt.parent && pt.size && pt.leaves && clA && clB && dNearest && dMerged && dValid && dOffset && dCounters &&
			      dPre && dLeafIndex && dLeafTrisOut && dScanTmp))
			{
				freeAll();
				return fail(c, GI_ERR_CUDA, "gi_build_hierarchy: device allocation failed");
			}
			const uint32_t initCounters[2] = {L, L};
			step(cudaMemcpyAsync(dCounters, initCounters, 8, cudaMemcpyHostToDevice, st));
			k_ploc_init<<<(L + B - 1) / B, B, 0, st>>>(L, tr, pt, clA);
			++c->launches;
			uint32_t n = L;
			auto pass = [&](bool forced)
			{
				const unsigned g = (n + B - 1) / B;
				if (forced) k_ploc_force_pairs<<<g, B, 0, st>>>(n, dNearest);
				else k_ploc_nearest<<<g, B, 0, st>>>(n, clA, pt, dNearest);
				k_ploc_merge<<<g, B, 0, st>>>(n, clA, dNearest, pt, dCounters, dMerged, dValid);
				step(cub::DeviceScan::ExclusiveSum(dScanTmp, scanBytes, dValid, dOffset, int(n), st));
				k_ploc_compact<<<g, B, 0, st>>>(n, dMerged, dValid, dOffset, clB, dCounters + 1);
				c->launches += 3;
				uint32_t next = n;
				step(cudaMemcpyAsync(&next, dCounters + 1, 4, cudaMemcpyDeviceToHost, st));
				step(cudaStreamSynchronize(st));
				return next;
			};
			bool forceNext = false; // (see buildLeafTreeOnDevice: coincident boxes must not cost one pass per cluster)
			while (n > 1u && e == cudaSuccess)
			{
				uint32_t next = pass(forceNext);
				forceNext = false;
				if (next >= n && e == cudaSuccess) next = pass(true); // (never taken with the tie rule of k_ploc_nearest)
				if (next >= n) { if (e == cudaSuccess) e = cudaErrorUnknown; break; }
				if (n > 1024u && uint64_t(n - next) * 16u < uint64_t(n)) forceNext = true; // (short lists near the root always merge slowly: leave them to the area rule)
				std::swap(clA, clB);
				n = next;
			}
			if (e == cudaSuccess)
			{
				k_ploc_preorder<<<(N + B - 1) / B, B, 0, st>>>(N, L, pt, dPre, dLeafIndex, dDepth);
				k_ploc_emit<<<(N + B - 1) / B, B, 0, st>>>(N, L, pt, dPre, dLeafIndex, dLeafTris, dNodes, dParents, dLeafTrisOut);
				c->launches += 2;
				dLeafTris = dLeafTrisOut; // leaves renumbered in preorder
			}
		}
		step(cudaGetLastError());
	}
	uint32_t depth = 1;
	if (I > 0)
	{
		step(cudaMemcpyAsync(out_nodes, dNodes, size_t(N) * 32, cudaMemcpyDeviceToHost, st));
		step(cudaMemcpyAsync(out_parents, dParents, size_t(N) * 4, cudaMemcpyDeviceToHost, st));
		step(cudaMemcpyAsync(&depth, dDepth, 4, cudaMemcpyDeviceToHost, st));
	}
	step(cudaMemcpyAsync(out_leaf_triangles, dLeafTris, size_t(L) * 2 * 16, cudaMemcpyDeviceToHost, st));
	float leafBox[6];
	if (I == 0)
	{
		step(cudaMemcpyAsync(leafBox, tr.leafMin, 12, cudaMemcpyDeviceToHost, st));
		step(cudaMemcpyAsync(leafBox + 3, tr.leafMax, 12, cudaMemcpyDeviceToHost, st));
	}
	step(cudaStreamSynchronize(st));
	freeAll();
	if (e != cudaSuccess) return fail(c, GI_ERR_CUDA, std::string("gi_build_hierarchy: ") + cudaGetErrorString(e));
	if (I == 0)
	{
		// one leaf: the root is a leaf node (bim_sahspilt.cpp:36)
		std::memcpy(out_nodes[0].min, leafBox, 12); std::memcpy(out_nodes[0].max, leafBox + 3, 12);
		out_nodes[0].first_child = 0x80000000u; out_nodes[0].escape = 0u;
		out_parents[0] = 0u;
	}
	*out_num_nodes = N;
	*out_num_tree_levels = depth;
	return GI_OK;
}

int gi_scene_topology(const GiContext* c, double* build_ms)
{
	if (!c || !c->haveScene) return -1;
	if (build_ms) *build_ms = c->topologyBuildMs;
	return c->sceneTopology;
}

int gi_scene_resident_bytes(const GiContext* c, uint64_t out[4])
{
	if (!c || !c->haveScene || !out) return GI_ERR_INVALID;
	for (int i = 0; i < 4; ++i) out[i] = c->residentBytes[i];
	return GI_OK;
}

int gi_synchronize(GiContext* c)
{
	if (!c) return GI_ERR_INVALID;
	{
		int fst = flushStream(c);
		if (fst != GI_OK) return fst;
	}
	GI_CUDA(c, cudaStreamSynchronize(c->stream));
	return GI_OK;
}

int gi_read_accumulator(GiContext* c, float* rgba)
{
	if (!c || !rgba) return GI_ERR_INVALID;
	{
		int fst = flushStream(c);
		if (fst != GI_OK) return fst;
	}
	if (!c->accum) return fail(c, GI_ERR_INVALID, "gi_set_screen_size was not called");
	GI_CUDA(c, cudaMemcpyAsync(rgba, c->accum, size_t(c->width) * c->height * 4 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
	GI_CUDA(c, cudaStreamSynchronize(c->stream));
	return GI_OK;
}

// Progressive read-out that does not drain an open stream: a device-to-device snapshot of the accumulator in stream order
// (after the waves submitted so far), copied to the host on a second stream while the next waves run.
int gi_read_accumulator_async(GiContext* c, float* rgba, uint32_t* iterations_in_flight)
{
	if (!c || !rgba) return GI_ERR_INVALID;
	if (!c->accum) return fail(c, GI_ERR_INVALID, "gi_set_screen_size was not called");
	GI_CUDA(c, cudaSetDevice(c->device));
	const size_t bytes = size_t(c->width) * c->height * 4 * sizeof(float);
	if (!c->copyStream)
	{
		GI_CUDA(c, cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking));
		for (int i = 0; i < 2; ++i)
		{
			GI_CUDA(c, cudaEventCreateWithFlags(&c->snapTaken[i], cudaEventDisableTiming));
			GI_CUDA(c, cudaEventCreateWithFlags(&c->snapCopied[i], cudaEventDisableTiming));
		}
	}
	if (c->snapshotBytes != bytes)
	{
		GI_CUDA(c, cudaStreamSynchronize(c->copyStream));
		GI_CUDA(c, cudaStreamSynchronize(c->stream));
		for (int i = 0; i < 2; ++i) { cudaFree(c->snapshot[i]); c->snapshot[i] = nullptr; c->snapUsed[i] = false; }
		c->snapshotBytes = 0;
		for (int i = 0; i < 2; ++i) GI_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->snapshot[i]), bytes));
		c->snapshotBytes = bytes;
	}
	const int i = c->snapNext;
	c->snapNext ^= 1;
	if (c->snapUsed[i]) GI_CUDA(c, cudaStreamWaitEvent(c->stream, c->snapCopied[i], 0)); // the read before last has left this buffer
	GI_CUDA(c, cudaMemcpyAsync(c->snapshot[i], c->accum, bytes, cudaMemcpyDeviceToDevice, c->stream));
	GI_CUDA(c, cudaEventRecord(c->snapTaken[i], c->stream));
	GI_CUDA(c, cudaStreamWaitEvent(c->copyStream, c->snapTaken[i], 0));
	GI_CUDA(c, cudaMemcpyAsync(rgba, c->snapshot[i], bytes, cudaMemcpyDeviceToHost, c->copyStream));
	GI_CUDA(c, cudaEventRecord(c->snapCopied[i], c->copyStream));
	c->snapUsed[i] = true;
	// PT stream: iteration k joins the accumulator when its state slot is recycled, maxPathLength + 1 waves after it started
	if (iterations_in_flight)
	{
		// BPT: the eye pass starts D waves late; the light-pass splats of an iteration wait in its splat image until its eye radiance is folded
		if (c->bpt.open && c->bpt.streaming) *iterations_in_flight = std::min(c->bpt.count, c->bpt.D + c->bpt.maxLen + 1u);
		else if (c->run[0].open && c->run[0].streaming && c->run[0].pass == WF_PASS_LIGHT) *iterations_in_flight = c->run[0].count - c->run[0].splatFolded;
		else *iterations_in_flight = (c->run[0].open && c->run[0].streaming) ? std::min(c->run[0].count, c->run[0].numSlots) : 0u;
	}
	return GI_OK;
}

static int commCheckStatus(GiContext* c);

int gi_wait_reads(GiContext* c)
{
	if (!c) return GI_ERR_INVALID;
	if (c->copyStream) GI_CUDA(c, cudaStreamSynchronize(c->copyStream));
	return commCheckStatus(c);
}

// ---- multi-GPU read-out (gi_comm.cuh)
int gi_comm_slice(int height, int rank, int world, uint32_t* first_row, uint32_t* num_rows)
{
	if (height <= 0 || world <= 0 || rank < 0 || rank >= world || !first_row || !num_rows) return GI_ERR_INVALID;
	const uint64_t a = uint64_t(height) * uint64_t(rank) / uint64_t(world), b = uint64_t(height) * uint64_t(rank + 1) / uint64_t(world);
	*first_row = uint32_t(a); *num_rows = uint32_t(b - a);
	return GI_OK;
}

int gi_comm_destroy(GiContext* c)
{
	if (!c) return GI_ERR_INVALID;
	if (c->commWorld == 0) return GI_OK;
	cudaSetDevice(c->device);
	if (c->copyStream) cudaStreamSynchronize(c->copyStream);
	cudaStreamSynchronize(c->stream);
	for (int r = 0; r < c->commWorld; ++r)
		if (r != c->commRank && c->commPeer[r]) cudaIpcCloseMemHandle(c->commPeer[r]);
	for (int r = 0; r < GI_COMM_MAX_WORLD; ++r) c->commPeer[r] = nullptr;
	cudaFree(c->commBlock); c->commBlock = nullptr;
	cudaFree(c->commOut[0]); cudaFree(c->commOut[1]); c->commOut[0] = c->commOut[1] = nullptr;
	cudaFree(c->commWords); c->commWords = nullptr;
	for (int i = 0; i < 2; ++i)
	{
		if (c->commSnapped[i]) cudaEventDestroy(c->commSnapped[i]);
		if (c->commReduced[i]) cudaEventDestroy(c->commReduced[i]);
		c->commSnapped[i] = c->commReduced[i] = nullptr; c->commUsed[i] = false;
	}
	c->commWorld = 0; c->commConnected = false; c->commEpoch = 0;
	return GI_OK;
}

int gi_comm_create(GiContext* c, int rank, int world, GiCommHandle* out_handle)
{
	if (!c || !out_handle) return GI_ERR_INVALID;
	if (world < 1 || world > GI_COMM_MAX_WORLD || rank < 0 || rank >= world) return fail(c, GI_ERR_INVALID, "gi_comm_create: rank / world out of range (1..16 ranks)");
	if (!c->accum) return fail(c, GI_ERR_INVALID, "gi_comm_create: gi_set_screen_size was not called");
	static_assert(sizeof(GiCommHandle) >= sizeof(cudaIpcMemHandle_t), "GiCommHandle holds a cudaIpcMemHandle_t");
	GI_CUDA(c, cudaSetDevice(c->device));
	gi_comm_destroy(c);
	const size_t imageBytes = size_t(c->width) * size_t(c->height) * 16u;
	uint32_t firstRow = 0, rows = 0;
	gi_comm_slice(c->height, rank, world, &firstRow, &rows);
	GI_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->commBlock), GI_COMM_FLAG_BYTES + 2u * imageBytes));
	GI_CUDA(c, cudaMemset(c->commBlock, 0, GI_COMM_FLAG_BYTES));
	for (int i = 0; i < 2; ++i) GI_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->commOut[i]), std::max<size_t>(size_t(rows) * size_t(c->width) * 16u, 16u)));
	GI_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->commWords), 4 * sizeof(uint32_t)));
	GI_CUDA(c, cudaMemset(c->commWords, 0, 4 * sizeof(uint32_t)));
	for (int i = 0; i < 2; ++i)
	{
		GI_CUDA(c, cudaEventCreateWithFlags(&c->commSnapped[i], cudaEventDisableTiming));
		GI_CUDA(c, cudaEventCreateWithFlags(&c->commReduced[i], cudaEventDisableTiming));
	}
	std::memset(out_handle, 0, sizeof(*out_handle));
	cudaIpcMemHandle_t h;
	if (world > 1)
	{
		GI_CUDA(c, cudaIpcGetMemHandle(&h, c->commBlock));
		std::memcpy(out_handle, &h, sizeof(h));
	}
	c->commRank = rank; c->commWorld = world; c->commWidth = c->width; c->commHeight = c->height;
	c->commPeer[rank] = c->commBlock;
	c->commConnected = world == 1;
	c->commEpoch = 0;
	return GI_OK;
}

int gi_comm_connect(GiContext* c, const GiCommHandle* handles)
{
	if (!c || !handles) return GI_ERR_INVALID;
	if (c->commWorld == 0) return fail(c, GI_ERR_INVALID, "gi_comm_connect: gi_comm_create was not called");
	GI_CUDA(c, cudaSetDevice(c->device));
	for (int r = 0; r < c->commWorld; ++r)
	{
		if (r == c->commRank || c->commPeer[r]) continue;
		cudaIpcMemHandle_t h;
		std::memcpy(&h, handles + r, sizeof(h));
		void* p = nullptr;
		cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
		if (e != cudaSuccess)
			return fail(c, GI_ERR_CUDA, std::string("gi_comm_connect: cannot map the exchange block of rank ") + std::to_string(r) + " (" + cudaGetErrorString(e) +
			                                "); the ranks must be processes on GPUs of one NVLink / PCIe peer group");
		c->commPeer[r] = static_cast<unsigned char*>(p);
	}
	c->commConnected = true;
	return GI_OK;
}

static int commRead(GiContext* c, float* rgba_rows, uint32_t* iterations_in_flight, bool drain)
{
	if (!c || !rgba_rows) return GI_ERR_INVALID;
	if (c->commWorld == 0 || !c->commConnected) return fail(c, GI_ERR_INVALID, "reduced read-out: gi_comm_create / gi_comm_connect first");
	if (!c->accum || c->commWidth != c->width || c->commHeight != c->height) return fail(c, GI_ERR_INVALID, "reduced read-out: the screen size changed since gi_comm_create");
	GI_CUDA(c, cudaSetDevice(c->device));
	if (drain)
	{
		int fst = flushStream(c);
		if (fst != GI_OK) return fst;
	}
	if (!c->copyStream) GI_CUDA(c, cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking));
	const uint32_t epoch = c->commEpoch++;
	const int i = int(epoch & 1u);
	const uint32_t numPixels = uint32_t(c->width) * uint32_t(c->height);
	const size_t imageBytes = size_t(numPixels) * 16u;
	uint32_t firstRow = 0, rows = 0;
	gi_comm_slice(c->height, c->commRank, c->commWorld, &firstRow, &rows);
	CommPeers peers;
	for (int r = 0; r < GI_COMM_MAX_WORLD; ++r) peers.block[r] = c->commPeer[r];
	peers.world = c->commWorld; peers.rank = c->commRank;
	// snapshot e & 1 is still being read by this rank's own reduce of read-out e - 2 until commReduced[i]
	if (c->commUsed[i]) GI_CUDA(c, cudaStreamWaitEvent(c->stream, c->commReduced[i], 0));
	k_comm_snapshot<<<unsigned(c->numSMs) * 2u, 256, 0, c->stream>>>(reinterpret_cast<const float4*>(c->accum), peers, epoch, numPixels, imageBytes, c->commWords, c->commWords + 2);
	++c->launches;
	GI_CUDA(c, cudaEventRecord(c->commSnapped[i], c->stream));
	GI_CUDA(c, cudaStreamWaitEvent(c->copyStream, c->commSnapped[i], 0));
	// a small grid: the kernel may sit next to the persistent trace kernel while it waits for the slowest peer
	k_comm_reduce<<<unsigned(c->numSMs) / 2u, 256, 0, c->copyStream>>>(peers, epoch, firstRow * uint32_t(c->width), rows * uint32_t(c->width), imageBytes,
	                                                                   reinterpret_cast<float4*>(c->commOut[i]), c->commWords + 1, c->commWords + 2);
	++c->launches;
	GI_CUDA(c, cudaEventRecord(c->commReduced[i], c->copyStream));
	c->commUsed[i] = true;
	if (rows) GI_CUDA(c, cudaMemcpyAsync(rgba_rows, c->commOut[i], size_t(rows) * size_t(c->width) * 16u, cudaMemcpyDeviceToHost, c->copyStream));
	GI_CUDA(c, cudaGetLastError());
	if (iterations_in_flight)
	{
		if (c->bpt.open && c->bpt.streaming) *iterations_in_flight = std::min(c->bpt.count, c->bpt.D + c->bpt.maxLen + 1u);
		else if (c->run[0].open && c->run[0].streaming && c->run[0].pass == WF_PASS_LIGHT) *iterations_in_flight = c->run[0].count - c->run[0].splatFolded;
		else *iterations_in_flight = (c->run[0].open && c->run[0].streaming) ? std::min(c->run[0].count, c->run[0].numSlots) : 0u;
	}
	return GI_OK;
}

static int commCheckStatus(GiContext* c)
{
	if (c->commWorld == 0 || !c->commWords) return GI_OK;
	uint32_t status = 0;
	GI_CUDA(c, cudaMemcpy(&status, c->commWords + 2, sizeof(status), cudaMemcpyDeviceToHost));
	if (status != 0u)
	{
		cudaMemset(c->commWords + 2, 0, sizeof(uint32_t));
		return fail(c, GI_ERR_CUDA, status == 1u ? "reduced read-out: a peer rank did not publish its snapshot within 4 s (ranks must issue the same sequence of reduced reads)"
		                                         : "reduced read-out: a peer rank did not finish reading within 4 s");
	}
	return GI_OK;
}

int gi_read_accumulator_reduced_async(GiContext* c, float* rgba_rows, uint32_t* iterations_in_flight) { return commRead(c, rgba_rows, iterations_in_flight, false); }

int gi_read_accumulator_reduced(GiContext* c, float* rgba_rows)
{
	int st = commRead(c, rgba_rows, nullptr, true);
	if (st != GI_OK) return st;
	GI_CUDA(c, cudaStreamSynchronize(c->copyStream));
	return commCheckStatus(c);
}

int gi_accumulator_device(GiContext* c, void** device_ptr, size_t* num_bytes)
{
	if (!c || !device_ptr) return GI_ERR_INVALID;
	{
		int fst = flushStream(c);
		if (fst != GI_OK) return fst;
	}
	if (!c->accum) return fail(c, GI_ERR_INVALID, "gi_set_screen_size was not called");
	*device_ptr = c->accum;
	if (num_bytes) *num_bytes = size_t(c->width) * c->height * 4 * sizeof(float);
	return GI_OK;
}

int gi_compute_mse(GiContext* c, const float* reference_rgb, uint32_t divisor, float* mse)
{
	if (!c || !reference_rgb || !mse || divisor == 0) return GI_ERR_INVALID;
	{
		int fst = flushStream(c);
		if (fst != GI_OK) return fst;
	}
	if (!c->accum) return fail(c, GI_ERR_INVALID, "gi_set_screen_size was not called");
	GI_CUDA(c, cudaSetDevice(c->device));
	const uint32_t numPixels = uint32_t(c->width) * uint32_t(c->height);
	float* dref = nullptr;
	GI_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&dref), size_t(numPixels) * 3 * sizeof(float)));
	cudaError_t e = cudaMemcpyAsync(dref, reference_rgb, size_t(numPixels) * 3 * sizeof(float), cudaMemcpyHostToDevice, c->stream);
	if (e == cudaSuccess) e = cudaMemsetAsync(c->mseResult, 0, sizeof(double), c->stream);
	if (e == cudaSuccess)
	{
		k_mse<<<296, 256, 0, c->stream>>>(reinterpret_cast<const float4*>(c->accum), dref, numPixels, 1.0f / float(divisor), c->mseResult);
		++c->launches;
		e = cudaGetLastError();
	}
	double sum = 0.0;
	if (e == cudaSuccess) e = cudaMemcpyAsync(&sum, c->mseResult, sizeof(double), cudaMemcpyDeviceToHost, c->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
	cudaFree(dref);
	if (e != cudaSuccess) return fail(c, GI_ERR_CUDA, std::string("gi_compute_mse: ") + cudaGetErrorString(e));
	*mse = float(sum / double(numPixels));
	return GI_OK;
}

int gi_get_counters(GiContext* c, GiCounters* out)
{
	if (!c || !out) return GI_ERR_INVALID;
	{
		int fst = flushStream(c);
		if (fst != GI_OK) return fst;
	}
	unsigned long long h[CNT_COUNT];
	uint32_t b[2];
	GI_CUDA(c, cudaMemcpyAsync(h, c->counters, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
	GI_CUDA(c, cudaMemcpyAsync(b, c->bptCounters, sizeof(b), cudaMemcpyDeviceToHost, c->stream));
	GI_CUDA(c, cudaStreamSynchronize(c->stream));
	out->closest_rays = h[CNT_CLOSEST]; out->any_rays = h[CNT_ANY]; out->camera_paths = h[CNT_CAMERA]; out->light_paths = h[CNT_LIGHT];
	out->shaded_vertices = h[CNT_SHADED]; out->splats = h[CNT_SPLATS]; out->exact_fallback_rays = h[CNT_FALLBACK];
	out->node_tests = h[CNT_NODES]; out->triangle_tests = h[CNT_TRIS];
	out->inner_node_steps = h[CNT_INNER_STEPS]; out->leaf_steps = h[CNT_LEAF_STEPS];
	out->light_cache_entries = b[0]; out->light_path_length_sum = b[1];
	return GI_OK;
}

int gi_reset_counters(GiContext* c)
{
	if (!c) return GI_ERR_INVALID;
	GI_CUDA(c, cudaMemsetAsync(c->counters, 0, CNT_COUNT * sizeof(unsigned long long), c->stream));
	return GI_OK;
}

uint64_t gi_kernel_launches(const GiContext* c) { return c ? c->launches : 0; }

int gi_get_kernel_times(GiContext* c, GiKernelTimes* out)
{
	if (!c || !out) return GI_ERR_INVALID;
	for (int i = 0; i < GI_KERNEL_CLASS_COUNT; ++i) { out->ms[i] = c->kernelMs[i]; out->launches[i] = c->kernelLaunches[i]; }
	return GI_OK;
}

} // extern "C"
