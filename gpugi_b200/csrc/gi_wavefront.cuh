// gi_wavefront.cuh -- wavefront scheduling of the three integrators (device only, sm_100a).
//
// The reference runs one GL compute invocation per path from the first ray to the last bounce
// (pathtracer.comp:16-110, lighttracer.comp:49-193, pathtracer_bidir.comp:16-251).  On a GPU without RT cores
// that shape is divergence-bound: Russian roulette ends most paths after 1-2 bounces while a few run to 16, and
// inner-node / leaf / shading code serialise inside a warp (profiles/r01_ncu_pathtrace.md: 6.95 of 32 lanes active).
// Here the work is a sequence of WAVES over HBM-resident queues, and a wave mixes all iterations (sample passes) that are
// in flight:
//
//   wave w:  generate(w)   camera rays of iteration w (while w < count)            -> extension queue of wave w
//            trace(w)      extension rays of wave w + shadow rays pushed by shade(w-1)
//            shade(w)      one lane per hit: bounce b of iteration w-b             -> extension / shadow queue of wave w+1
//
//   trace     persistent warps; every lane owns one ray and refills itself from the queue as soon as enough lanes of its
//             warp ran dry (ballot + popc, one atomicAdd per warp).  Inner-node steps and leaf steps are scheduled per
//             warp by vote so that the two code paths never serialise against each other.  Same tree, same box/triangle
//             arithmetic, same ambiguity rule as traceFast() (gi_core.h) -- the result per ray is bit-identical.
//   shade     GetTriangleHitInfo, SampleMaterialData, emission, next-event estimation (pushes a shadow ray carrying its
//             contribution), SampleBSDF, Russian roulette, and a ballot-compacted push of the surviving ray.
//
// Russian roulette ends most paths after 1-2 bounces while a few run to 16: rendered one iteration at a time, bounces
// 4..16 are ~30 launches that keep a handful of warps busy (1 ms of a 3.6 ms iteration for 1.5 % of the rays).  With the
// iterations of a batch overlapped (gi_render_iterations), every wave carries the bounce-0 rays of iteration w, the
// bounce-1 rays of iteration w-1, ... : constant work per wave, three launches per iteration, and the tail is paid once
// per batch.  gi_render_iteration is the same code with count = 1.
//
// Determinism.  Every path keeps the reference's private RNG stream and its floating-point operation order; the radiance
// sum of a path lives in its state record (slot = iteration mod S, S = maxPathLength + 1 slots) and is added to the
// accumulator when the slot is recycled S waves later (or by the final fold), i.e. per pixel in iteration order.  A batch is
// therefore bit-identical to the same iterations rendered one by one.
#pragma once

#include "gi_core.h"
#include "gi_lane.h"
#include "gi_shade.h"

namespace gik {

// ------------------------------------------------------------------ queues (HBM)
// ray record      2 x F4: (origin.xyz, bounce bits) (direction.xyz, path word); extension rays always start with RAY_MAX
// hit record      float2: (t, leafTri bits), leafTri = -1: miss; always the reference's exact answer
// shadow record   3 x F4: (origin.xyz, rayLength) (direction.xyz, target word) (value.rgb, -)
// path word       bits 0..29 path id (pixel index / light path index), bit 30 didNEE
// target word     bits 0..29 state index (slot * numIds + path id) or pixel index, bit 30 splat (atomic add into the
//                 accumulator instead of the path's radiance sum)
#define WF_ID_MASK 0x3fffffffu
#define WF_FLAG_NEE 0x40000000u
#define WF_FLAG_SPLAT 0x40000000u
#define WF_COUNTS_PER_WAVE 4 // [0] extension rays pushed by shade(w-1), [1] shadow rays pushed by shade(w-1), [2] trace work counter, [3] shade work counter

struct WfQueues
{
	F4* rays[2];
	float2* hits;
	F4* shadow;
	float2* shadowPos;   // reconstruction filters only (else nullptr): unrounded pixel position of a splat, one per shadow-queue slot
	F4* stateT;          // numSlots x numIds: (throughput.xyz, seed bits)
	F4* stateC;          // numSlots x numIds: (radiance sum.xyz, alpha increment)
	float* stateM;       // numSlots x numIds: lastConnectionProbability of a BPT eye path
	uint32_t* counts;    // WF_COUNTS_PER_WAVE x countsRing; wave w uses entry w % countsRing (a batch: one entry per wave, zeroed up front;
	                     // a stream: a ring, entry w + 1 is zeroed before wave w starts)
	uint32_t countsRing;
	uint32_t capacity;   // slots per queue
	uint32_t numIds;     // path ids per iteration (PT/BPT: width x height)
	uint32_t numSlots;   // S iterations in flight
};

// tuning knobs (uniform kernel arguments so that one build can be swept)
struct WfTune
{
	int refillThreshold; // refill a warp's idle lanes once this many are idle
	int minInner;        // keep taking inner-node steps while this many (of 32 active) lanes are at inner nodes
	int reserved;
};

// ---- cache policy of the queue traffic (WF_STREAM_HINTS)
// Ray / hit / shadow records and path states stream through the GPU once per wave (several hundred MB, more than the 126 MB L2), while
// the hierarchy, the leaf block, vertices, materials and textures are gathered again and again.  Queue loads and stores therefore carry
// the streaming hint (ld.global.cs / st.global.cs: evict-first in L1 and L2) so that they do not push the scene out of the L2.
#ifndef WF_STREAM_HINTS
#define WF_STREAM_HINTS 1
#endif
GI_D float4 qld4(const void* p)
{
#if WF_STREAM_HINTS
	return __ldcs(reinterpret_cast<const float4*>(p));
#else
	return *reinterpret_cast<const float4*>(p);
#endif
}
GI_D F4 qldF4(const F4* p)
{
#if WF_STREAM_HINTS
	const float4 v = __ldcs(reinterpret_cast<const float4*>(p)); F4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r;
#else
	return ld4(p);
#endif
}
GI_D void qst4(void* p, float4 v)
{
#if WF_STREAM_HINTS
	__stcs(reinterpret_cast<float4*>(p), v);
#else
	*reinterpret_cast<float4*>(p) = v;
#endif
}
GI_D float2 qld2(const float2* p)
{
#if WF_STREAM_HINTS
	return __ldcs(p);
#else
	return *p;
#endif
}
GI_D void qst2(float2* p, float2 v)
{
#if WF_STREAM_HINTS
	__stcs(p, v);
#else
	*p = v;
#endif
}
GI_D float qld1(const float* p)
{
#if WF_STREAM_HINTS
	return __ldcs(p);
#else
	return *p;
#endif
}
GI_D void qst1(float* p, float v)
{
#if WF_STREAM_HINTS
	__stcs(p, v);
#else
	*p = v;
#endif
}

GI_D F4 f4zero() { F4 r; r.x = r.y = r.z = r.w = 0.0f; return r; }
GI_D unsigned laneId() { return threadIdx.x & 31u; }
GI_D unsigned lanemaskLt() { unsigned m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

// ballot-compacted slot allocation: one atomicAdd per warp (all 32 lanes must call)
GI_D uint32_t warpPush(uint32_t* counter, bool pred)
{
	const unsigned m = __ballot_sync(0xffffffffu, pred);
	if (m == 0u) return 0u;
	uint32_t base = 0;
	const int leader = __ffs(m) - 1;
	if (int(laneId()) == leader) base = atomicAdd(counter, uint32_t(__popc(m)));
	base = __shfl_sync(0xffffffffu, base, leader);
	return base + __popc(m & lanemaskLt());
}

GI_D void flushWarpCounters(unsigned long long* counters, const unsigned* v, int n)
{
	for (int i = 0; i < n; ++i)
	{
		unsigned s = __reduce_add_sync(0xffffffffu, v[i]);
		if (laneId() == 0 && s) atomicAdd(counters + i, (unsigned long long)s);
	}
}

// Work source / sink of the wavefront trace kernel: extension rays of bounce b, then shadow rays of bounce b-1.
struct WfTraceIO
{
	const F4* rays;
	float2* hits;
	const F4* shadow;
	F4* stateC;
	float* accum;
	int width, height;
	uint32_t numClosest, numShadow;
	uint32_t rayBase;   // extension ray i of this wave sits in queue slot rayBase + i
	float* splatBase;   // streams of light paths: ring of per-iteration splat images (nullptr: splats go straight into accum)
	size_t splatStride; // floats per image
	const float2* shadowPos; // RECONSTRUCTION_FILTER != 0: where each splat falls before rounding
	int reconFilter;
};

GI_D void wfSplat(float* accum, int width, int height, uint32_t pixel, F3 v)
{
	splatReduce(accum + 4 * size_t(pixel), v);
}

// An unoccluded shadow ray delivers what it carries: s1 = (direction, target word), v = (value, splat ring slot) of shadow record `slot`.
template <class CNT>
GI_D void wfDeliverShadow(const WfTraceIO& io, uint32_t slot, F4 s1, F4 v, CNT& cnt)
{
	const uint32_t word = uint32_t(f2i(s1.w));
	const uint32_t id = word & WF_ID_MASK;
	if (word & WF_FLAG_SPLAT)
	{
		cnt.template add<WFC_SPLATS>(1u);
		// a stream keeps the splats of an iteration in that iteration's own image (ring slot in v.w) until the iteration is complete
		float* image = io.splatBase ? io.splatBase + size_t(uint32_t(f2i(v.w))) * io.splatStride : io.accum;
		if (io.reconFilter)
		{
			// screenreproject.glsl:186-223: 5 x 5 weighted splats around the rounded position
			const float2 pos = io.shadowPos[slot];
			PathCounters dummy = {};
			splatFiltered(image, io.width, io.height, io.reconFilter, pos.x, pos.y, f3(v.x, v.y, v.z), dummy);
			return;
		}
		wfSplat(image, io.width, io.height, id, f3(v.x, v.y, v.z));
		return;
	}
	// the path's radiance sum: one shadow ray per path and wave, and shade() is another kernel -> exclusive
	float4 c = qld4(io.stateC + id);
	c.x += v.x; c.y += v.y; c.z += v.z;
	qst4(io.stateC + id, c);
}

template <class CNT>
GI_D void wfFinishShadow(const WfTraceIO& io, uint32_t slot, bool occluded, CNT& cnt)
{
	if (occluded) return;
	const F4 s1 = qldF4(io.shadow + 3 * size_t(slot) + 1);
	const F4 v = qldF4(io.shadow + 3 * size_t(slot) + 2);
	wfDeliverShadow(io, slot, s1, v, cnt);
}

// The persistent trace loop.  IO provides: total(), load(index, lane state), finish(lane state).
// Lane states are encoded in L.node: >= 0 at an inner node, < 0 at a leaf, WF_NODE_DONE finished (active: waiting to be
// retired) or idle.  One outer iteration retires finished rays, refills idle lanes, runs ONE leaf step for the lanes
// waiting at a leaf and then inner-node steps for as long as enough lanes of the warp are at inner nodes (one vote per
// step), so leaf code, retire/refill code and inner-node code never serialise against each other inside the hot loop.
// Traversal stack of a trace lane: entries 0 .. WF_SHARED_STACK-1 in shared memory (entry-major, one 8-byte word per thread and
// entry: conflict-free, and the hot entries no longer compete with BVH nodes for L1 lines or cost write-through traffic to
// L2), deeper entries in local memory.
#ifndef WF_SHARED_STACK
#define WF_SHARED_STACK 0
#endif
#ifndef WF_TRACE_THREADS
#define WF_TRACE_THREADS 128
#endif
struct WfStackShared
{
	uint2* sh;          // &smem[threadIdx.x]
	StackEntry* loc;
	GI_D StackEntry load(int i) const
	{
		StackEntry r;
		if (i < WF_SHARED_STACK) { const uint2 v = sh[i * WF_TRACE_THREADS]; r.node = int(v.x); r.first = __uint_as_float(v.y); }
		else r = loc[i - WF_SHARED_STACK];
		return r;
	}
	GI_D void store(int i, int node, float first) const
	{
		if (i < WF_SHARED_STACK) sh[i * WF_TRACE_THREADS] = make_uint2(uint32_t(node), __float_as_uint(first));
		else { loc[i - WF_SHARED_STACK].node = node; loc[i - WF_SHARED_STACK].first = first; }
	}
};

#ifndef WF_SHADE_HOIST
#define WF_SHADE_HOIST 0  // k_wf_shade_pt: leaf record + triangle requested before the path state
#endif
#ifndef WF_BATCH_RETIRE
#define WF_BATCH_RETIRE 1 // finished lanes are retired together when the refill is due (B200: 1.699 -> 1.680 ms per PT iteration)
#endif
template <class IO, class CNT>
GI_D void wfTraceLoop(const DScene& s, const IO& io, uint32_t* workCounter, WfTune tune, CNT& cnt)
{
#if WF_SHARED_STACK > 0
	__shared__ uint2 wfSharedStack[WF_SHARED_STACK * WF_TRACE_THREADS];
	StackEntry entries[GI_STACK_SIZE - WF_SHARED_STACK];
	WfStackShared stack; stack.sh = wfSharedStack + threadIdx.x; stack.loc = entries;
#else
	StackEntry entries[GI_STACK_SIZE];
	WfStackArray stack; stack.e = entries;
#endif
	WfCand cand[WF_MAX_CAND];
	WfLane L;
	L.node = WF_NODE_DONE;
	bool active = false;
	bool poolEmpty = false;
	const uint32_t total = io.total();
	const unsigned lane = laneId();
	for (;;)
	{
		// ---- retire finished rays
#if WF_BATCH_RETIRE
		// A finished lane has nothing to do until the next refill, so its result is written when the refill is due (or nothing else is
		// left to do): the retire code -- end-of-walk checks, hit record / shadow-ray finish -- then runs for many lanes at once
		// instead of for the one or two lanes that happen to finish in an outer iteration.
		{
			const bool finished = active && wfFinished(L);
			const unsigned waiting = __ballot_sync(0xffffffffu, finished || !active);
			if ((poolEmpty || __popc(waiting) >= tune.refillThreshold || waiting == 0xffffffffu) && finished && wfEndOfWalk(L, s, cand, cnt))
			{
				io.finish(s, L, cnt);
				active = false;
			}
		}
#else
		if (active && wfFinished(L) && wfEndOfWalk(L, s, cand, cnt))
		{
			io.finish(s, L, cnt);
			active = false;
		}
#endif
		// ---- refill idle lanes from the queue
		const unsigned idle = __ballot_sync(0xffffffffu, !active);
		if (!poolEmpty && (__popc(idle) >= tune.refillThreshold || idle == 0xffffffffu))
		{
			const int n = __popc(idle);
			const int leader = __ffs(idle) - 1;
			uint32_t base = 0;
			if (int(lane) == leader) base = atomicAdd(workCounter, uint32_t(n));
			base = __shfl_sync(0xffffffffu, base, leader);
			const uint32_t mine = base + __popc(idle & lanemaskLt());
			if (!active && mine < total)
			{
				io.load(s, mine, L, cnt);
				active = true;
			}
			if (base + uint32_t(n) >= total) poolEmpty = true;
		}
		const unsigned activeMask = __ballot_sync(0xffffffffu, active);
		if (activeMask == 0u)
		{
			if (poolEmpty) break;
			continue;
		}
		// ---- one leaf step for the lanes waiting at a leaf
		if (active) wfStash(L, stack); // (a freshly loaded / restarted ray may stand at a leaf: single-leaf scenes)
		if (active && wfAtLeaf(L)) wfLeafStep(L, s, stack, cand, cnt);
		// ---- inner-node steps while enough lanes are at inner nodes
		unsigned innerMask = __ballot_sync(0xffffffffu, L.node >= 0);
		if (innerMask == 0u) continue;
#if WF_BATCH_RETIRE
		const int keepGoing = max(1, (__popc(__ballot_sync(0xffffffffu, active && !wfFinished(L))) * tune.minInner) >> 5); // lanes still walking
#else
		const int keepGoing = max(1, (__popc(activeMask) * tune.minInner) >> 5);
#endif
#if WF_ANY_UNSORTED && WF_CENTRE_EXTENT
		// every walking lane of the warp carries an any-hit ray (the queue hands out closest-hit rays first, then the shadow rays): unsorted steps
		const bool walkingLane = active && !wfFinished(L);
		if (__all_sync(0xffffffffu, !walkingLane || (L.flags & WFL_ANY) != 0u))
		{
			do
			{
				if (L.node >= 0) wfInnerStepAny(L, s, stack, cnt);
				innerMask = __ballot_sync(0xffffffffu, L.node >= 0);
			} while (__popc(innerMask) >= keepGoing);
			continue;
		}
#endif
		do
		{
			if (L.node >= 0) wfInnerStep(L, s, stack, cnt);
			innerMask = __ballot_sync(0xffffffffu, L.node >= 0);
		} while (__popc(innerMask) >= keepGoing);
	}
}

// Reference-order variant (GI_TRAVERSAL_REFERENCE): same work distribution, every lane walks its ray with the
// stackless escape-pointer loop of traceray.glsl from start to end.
template <class IO, class CNT>
GI_D void wfTraceLoopReference(const DScene& s, const IO& io, uint32_t* workCounter, CNT& cnt)
{
	const uint32_t total = io.total();
	for (;;)
	{
		uint32_t base = 0;
		if (laneId() == 0) base = atomicAdd(workCounter, 32u);
		base = __shfl_sync(0xffffffffu, base, 0);
		if (base >= total) break;
		const uint32_t mine = base + laneId();
		if (mine < total)
		{
			WfLane L;
			WfCounters<false> dummy;
			io.load(s, mine, L, dummy); // ray + root test (the walk below repeats the root test itself)
			TraceStats st = {0, 0};
			float len = L.tMaxInit; int leafTri = -1;
			bool fb = false;
			if (s.oboxOrdered)
			{
				// bvhType = 1 with the ordered walk (traceOrderedOBox): exact, nearest child first
				if (L.flags & WFL_ANY) { if (traceOrderedOBox<true>(s, L.o, L.d, len, leafTri, &st, nullptr)) L.flags |= WFL_OCCLUDED; }
				else traceOrderedOBox<false>(s, L.o, L.d, len, leafTri, &st, &fb);
				if (fb) cnt.template add<WFC_FALLBACK>(1u);
			}
			else if (L.flags & WFL_ANY) { if (traceReference<true>(s, L.o, L.d, len, leafTri, &st)) L.flags |= WFL_OCCLUDED; }
			else traceReference<false>(s, L.o, L.d, len, leafTri, &st);
			cnt.template add<WFC_NODES>(st.nodes); cnt.template add<WFC_TRIS>(st.tris);
			L.best = len; L.hitLeaf = leafTri; L.flags |= WFL_EXACT;
			io.finish(s, L, cnt);
		}
	}
}

struct WfQueueIO
{
	WfTraceIO q;
	GI_D uint32_t total() const { return q.numClosest + q.numShadow; }
	// the first two F4 of work item `index`: (origin.xyz, x) (direction.xyz, y)
	GI_D const F4* address(uint32_t index) const
	{
		return index >= q.numClosest ? q.shadow + 3 * size_t(index - q.numClosest) : q.rays + 2 * size_t(q.rayBase + index);
	}
	template <class CNT>
	GI_D void begin(const DScene& s, uint32_t index, F4 r0, F4 r1, WfLane& L, CNT& cnt) const
	{
		const bool isAny = index >= q.numClosest;
		wfBegin(L, s, f3(r0.x, r0.y, r0.z), f3(r1.x, r1.y, r1.z), isAny ? r0.w : GI_RAY_MAX, isAny, index, cnt);
	}
	template <class CNT>
	GI_D void load(const DScene& s, uint32_t index, WfLane& L, CNT& cnt) const
	{
		const F4* p = address(index);
		begin(s, index, qldF4(p), qldF4(p + 1), L, cnt);
	}
	template <class CNT>
	GI_D void finish(const DScene& s, WfLane& L, CNT& cnt) const
	{
		if (L.flags & WFL_ANY) wfFinishShadow(q, L.index - q.numClosest, (L.flags & WFL_OCCLUDED) != 0u, cnt);
		else qst2(q.hits + q.rayBase + L.index, make_float2(L.best, i2f(L.hitLeaf)));
	}
};

// gi_trace_device: GiRay batch in, GiHit out (the parity / Mrays/s entry point).
struct WfBatchIO
{
	const F4* rays;
	float4* hits;
	uint32_t n;
	GI_D uint32_t total() const { return n; }
	GI_D const F4* address(uint32_t index) const { return rays + 2 * size_t(index); }
	template <class CNT>
	GI_D void begin(const DScene& s, uint32_t index, F4 r0, F4 r1, WfLane& L, CNT& cnt) const
	{
		wfBegin(L, s, f3(r0.x, r0.y, r0.z), f3(r1.x, r1.y, r1.z), r0.w, (uint32_t(f2i(r1.w)) & 1u) != 0u, index, cnt);
	}
	template <class CNT>
	GI_D void load(const DScene& s, uint32_t index, WfLane& L, CNT& cnt) const
	{
		begin(s, index, ld4(address(index)), ld4(address(index) + 1), L, cnt);
	}
	template <class CNT>
	GI_D void finish(const DScene& s, WfLane& L, CNT& cnt) const
	{
		if (L.flags & WFL_ANY) cnt.template add<WFC_ANY>(1u); else cnt.template add<WFC_CLOSEST>(1u);
		if (L.flags & WFL_ANY)
		{
			hits[L.index] = make_float4(L.tMaxInit, i2f((L.flags & WFL_OCCLUDED) ? 1 : 0), 0.0f, 0.0f);
			return;
		}
		float4 out = make_float4(L.best, i2f(L.hitLeaf), 0.0f, 0.0f);
		if (L.hitLeaf >= 0)
		{
			F3 p0, p1, p2; int valid;
			loadLeafTri(s, L.hitLeaf, p0, p1, p2, valid);
			float tt; F3 bary, nn;
			intersectTriangle(L.o, L.d, p0, p1, p2, tt, bary, nn);
			out.z = bary.y; out.w = bary.z;
		}
		hits[L.index] = out;
	}
};

// ------------------------------------------------------------------ path tracer waves (pathtracer.comp:16-110)
struct WfFrame
{
	DFrame f;                // f.frameSeed: seed of the iteration generated in this wave; f.lightSamples: samples of batch iteration 0
	WfQueues q;
	WfTune tune;
	int tilesX, tilesY;      // W/8, H/8 work groups (pathtracer.cpp:50)
	uint32_t numPaths;       // paths per iteration
	int wave;                // wave index inside the batch; bounce b of this wave belongs to batch iteration wave - b
	uint32_t numCam;         // freshly generated rays at the head of this wave's extension queue (numPaths or 0)
	uint32_t pushBase;       // ... of the next wave's extension queue: pushed rays are placed behind them
	uint32_t rayBase;        // first slot of this wave's extension queue that holds a ray (0, or numPaths in a stream wave without generate)
	uint32_t lightStride;    // F4 elements between the light samples of consecutive iterations (0: the same set for all)
	uint32_t lightBase;      // iteration i reads light-sample set (lightBase + i) % lightRing
	uint32_t lightRing;
	uint32_t lightFlags;     // light tracer waves: WF_LIGHT_SAVE / WF_LIGHT_WARMUP (gi_shade.h)
	float* splatBase;        // streams of light paths: ring of splatSlots images, iteration i splats into image i % splatSlots (nullptr: into f.accum)
	uint32_t splatSlots;
};

// queue slot -> pixel: 8x8 tiles in row-major tile order, row-major inside the tile (the reference's work-group shape)
GI_D uint32_t wfSlotToPixel(const WfFrame& w, uint32_t slot)
{
	const uint32_t tile = slot >> 6, l = slot & 63u;
	const uint32_t tx = tile % uint32_t(w.tilesX), ty = tile / uint32_t(w.tilesX);
	return (tx * 8u + (l & 7u)) + (ty * 8u + (l >> 3)) * uint32_t(w.f.width);
}

GI_D uint32_t* wfCounts(const WfFrame& w, int wave) { return w.q.counts + WF_COUNTS_PER_WAVE * size_t(uint32_t(wave) % w.q.countsRing); }

// light samples of iteration `iteration` (relative to the start of the batch / stream)
GI_D const F4* wfLightSamples(const WfFrame& w, uint32_t iteration) { return w.f.lightSamples + size_t((w.lightBase + iteration) % w.lightRing) * w.lightStride; }

// Shades the hit of one extension ray of this wave.  Pushes at most one shadow ray and one extension ray
// (warp-compacted, so all 32 lanes call this; `valid` masks the tail).
GI_D void wfShadePT(const WfFrame& w, uint32_t slot, bool valid, unsigned* cnt)
{
	const DFrame& f = w.f;
	const WfQueues& q = w.q;
	bool pushShadow = false, pushRay = false;
	DirectLightQuery nee;
	F3 neeValue = f3(0.0f);
	F3 origin = f3(0.0f), dir = f3(0.0f, 0.0f, 1.0f);
	uint32_t id = 0, sidx = 0;
	int bounce = 0;
	bool didNEE = false;
	if (valid)
	{
		const F4* rp = q.rays[w.wave & 1] + 2 * size_t(slot);
		const F4 r0 = qldF4(rp), r1 = qldF4(rp + 1);
		const float2 hrec = qld2(q.hits + slot);
		const uint32_t word = uint32_t(f2i(r1.w));
		id = word & WF_ID_MASK;
		didNEE = (word & WF_FLAG_NEE) != 0u;
		bounce = f2i(r0.w);
		const uint32_t iteration = uint32_t(w.wave - bounce);         // batch-relative
		sidx = (iteration % q.numSlots) * q.numIds + id;
		origin = f3(r0.x, r0.y, r0.z); dir = f3(r1.x, r1.y, r1.z);
		F3 color = f3(0.0f), pathThroughput = f3(1.0f);
		uint32_t seed;
#if WF_SHADE_HOIST
		// the hit's leaf record and triangle are requested BEFORE the path state is read: both gathers depend on the queue entry only,
		// so their latencies overlap instead of adding up (the compiler does not hoist loads across the branches below)
		const int leafTriEarly = f2i(hrec.y);
		F3 hp0 = f3(0.0f), hp1 = f3(0.0f), hp2 = f3(0.0f); int hvalid = 0;
		I4 htri; htri.x = htri.y = htri.z = htri.w = 0;
		if (leafTriEarly >= 0)
		{
			loadLeafTri(f.scene, leafTriEarly, hp0, hp1, hp2, hvalid);
			htri = ld4(f.scene.leafTris + leafTriEarly);
		}
#endif
		if (bounce == 0)
		{
			seed = initRandomSeed(f.frameSeed, id);
			random2(seed); // the two draws of the sub-pixel jitter (pathtracer.comp:42)
			if (iteration >= q.numSlots)
			{
				// the slot is recycled: iteration - S is complete, its radiance joins the accumulator (pathtracer.comp:107-109;
				// per pixel in iteration order, alpha += 0)
				const float4 old = qld4(q.stateC + sidx);
				float4* px = reinterpret_cast<float4*>(f.accum) + id;
				float4 a = qld4(px);
				a.x += old.x; a.y += old.y; a.z += old.z; a.w += old.w;
				qst4(px, a);
			}
		}
		else
		{
			const float4 t = qld4(q.stateT + sidx);
			const float4 c = qld4(q.stateC + sidx);
			pathThroughput = f3(t.x, t.y, t.z); seed = uint32_t(f2i(t.w));
			color = f3(c.x, c.y, c.z);
		}
		const float tHit = hrec.x;
		const int leafTri = f2i(hrec.y);
		if (leafTri < 0)
			color = color + pathThroughput * environmentRadiance(f, dir);
		else
		{
#if WF_SHADE_HOIST
			const F3 p0 = hp0, p1 = hp1, p2 = hp2;
			const I4 tri = htri;
#else
			F3 p0, p1, p2; int tv;
			loadLeafTri(f.scene, leafTri, p0, p1, p2, tv);
			const I4 tri = ld4(f.scene.leafTris + leafTri);
#endif
			float tt; F3 bary, gn;
			intersectTriangle(origin, dir, p0, p1, p2, tt, bary, gn);
			origin = origin + tHit * dir;
			F3 hitNormal; F2 tc;
			getTriangleHitInfo(f.scene, tri, bary, hitNormal, tc);
			MaterialData md = sampleMaterialData(f.scene, tri.w, tc);
			++cnt[WFC_SHADED];
			if (!didNEE) color = color + pathThroughput * md.Emissivity;
			didNEE = md.Reflectiveness.w < 500.0f;
			{
				const int sampleIndex = int(randomUInt(seed) % uint32_t(f.numLightSamples));
				if (prepareDirectLight(wfLightSamples(w, iteration), origin, hitNormal, sampleIndex, dir, md, didNEE, nee))
				{
					pushShadow = true;
					neeValue = nee.radiance * pathThroughput * float(f.numLightSamples);
				}
			}
			F3 throughput = f3(1.0f);
			float pdf;
			dir = sampleBSDF(dir, md, seed, hitNormal, throughput, false, pdf);
			origin = origin + dir * GI_RAY_HIT_EPSILON;
			pathThroughput = pathThroughput * throughput;
			const float cont = saturate(luminance(throughput));
			if (!(randomFloat(seed) >= cont) && bounce + 1 < f.maxPathLength)
			{
				pathThroughput = pathThroughput / cont;
				pushRay = true;
			}
		}
		qst4(q.stateC + sidx, make_float4(color.x, color.y, color.z, 0.0f));
		if (pushRay) qst4(q.stateT + sidx, make_float4(pathThroughput.x, pathThroughput.y, pathThroughput.z, i2f(int(seed))));
	}
	uint32_t* next = wfCounts(w, w.wave + 1);
	const uint32_t sslot = warpPush(next + 1, pushShadow);
	if (pushShadow)
	{
		++cnt[WFC_ANY];
		F4* sp = q.shadow + 3 * size_t(sslot);
		qst4(sp, make_float4(nee.origin.x, nee.origin.y, nee.origin.z, nee.len));
		qst4(sp + 1, make_float4(nee.dir.x, nee.dir.y, nee.dir.z, i2f(int(sidx))));
		qst4(sp + 2, make_float4(neeValue.x, neeValue.y, neeValue.z, 0.0f));
	}
	const uint32_t rslot = w.pushBase + warpPush(next, pushRay);
	if (pushRay)
	{
		++cnt[WFC_CLOSEST];
		F4* rp = q.rays[(w.wave + 1) & 1] + 2 * size_t(rslot);
		const uint32_t word = id | (didNEE ? WF_FLAG_NEE : 0u);
		qst4(rp, make_float4(origin.x, origin.y, origin.z, i2f(bounce + 1)));
		qst4(rp + 1, make_float4(dir.x, dir.y, dir.z, i2f(int(word))));
	}
}

// ------------------------------------------------------------------ light tracer / bidirectional waves (gi_shade.h)
GI_D void wfAddCounters(unsigned* cnt, const PathCounters& pc)
{
	cnt[WFC_CLOSEST] += pc.closestRays; cnt[WFC_ANY] += pc.anyRays; cnt[WFC_CAMERA] += pc.cameraPaths; cnt[WFC_LIGHT] += pc.lightPaths;
	cnt[WFC_SHADED] += pc.shadedVertices; cnt[WFC_SPLATS] += pc.splats; cnt[WFC_FALLBACK] += pc.fallbackRays;
}

GI_D F4 ldState(const F4* p) { const float4 v = qld4(p); F4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r; }
GI_D void stState(F4* p, F4 v) { qst4(p, make_float4(v.x, v.y, v.z, v.w)); }

// warp-compacted push of what a shaded vertex produced (all 32 lanes call)
GI_D void wfPushOut(const WfFrame& w, const WfPush& out, uint32_t shadowTarget, int bounce, uint32_t id, uint32_t splatSlot = 0u)
{
	const WfQueues& q = w.q;
	uint32_t* next = wfCounts(w, w.wave + 1);
	const uint32_t sslot = warpPush(next + 1, out.shadow);
	if (out.shadow)
	{
		F4* sp = q.shadow + 3 * size_t(sslot);
		qst4(sp, make_float4(out.sOrigin.x, out.sOrigin.y, out.sOrigin.z, out.sLen));
		qst4(sp + 1, make_float4(out.sDir.x, out.sDir.y, out.sDir.z, i2f(int(shadowTarget))));
		qst4(sp + 2, make_float4(out.sValue.x, out.sValue.y, out.sValue.z, i2f(int(splatSlot))));
		if (q.shadowPos && (shadowTarget & WF_FLAG_SPLAT)) q.shadowPos[sslot] = make_float2(out.posX, out.posY);
	}
	const uint32_t rslot = w.pushBase + warpPush(next, out.ray);
	if (out.ray)
	{
		F4* rp = q.rays[(w.wave + 1) & 1] + 2 * size_t(rslot);
		qst4(rp, make_float4(out.origin.x, out.origin.y, out.origin.z, i2f(bounce + 1)));
		qst4(rp + 1, make_float4(out.dir.x, out.dir.y, out.dir.z, i2f(int(id))));
	}
}

// the light cache + counters of iteration `iteration` in a pipelined BPT batch (DFrame::cacheRing)
GI_D void wfSelectCache(DFrame& f, uint32_t iteration)
{
	if (f.cacheRing > 1u)
	{
		const uint32_t slot = iteration % f.cacheRing;
		f.lightCache += size_t(slot) * f.lightCacheCapacity * 4u;
		f.lightCacheCount += 2u * slot;
		f.lightPathLengthSum = f.lightCacheCount + 1;
	}
}

// generate(w) of the light tracer: light path `slot` of iteration w (lighttracer.comp:49-91)
GI_D void wfGenerateLight(const WfFrame& w, uint32_t gid, PathCounters& pc)
{
	DFrame f = w.f;
	wfSelectCache(f, uint32_t(w.wave));
	f.lightSamples = wfLightSamples(w, uint32_t(w.wave));
	F3 origin, dir; F4 T;
	wfGenerateLightLane(f, w.lightFlags, gid, origin, dir, T, pc);
	++pc.closestRays;
	const uint32_t sidx = (uint32_t(w.wave) % w.q.numSlots) * w.q.numIds + gid;
	stState(w.q.stateT + sidx, T);
	F4* rp = w.q.rays[w.wave & 1] + 2 * size_t(gid);
	qst4(rp, make_float4(origin.x, origin.y, origin.z, i2f(0)));
	qst4(rp + 1, make_float4(dir.x, dir.y, dir.z, i2f(int(gid))));
}

// shade(w) of the light tracer: one lane per hit record of a light path
GI_D void wfShadeLight(const WfFrame& w, uint32_t slot, bool valid, PathCounters& pc)
{
	WfPush out;
	out.ray = false; out.shadow = false; out.px = 0; out.py = 0;
	uint32_t id = 0, splatSlot = 0; int bounce = 0;
	if (valid)
	{
		const F4* rp = w.q.rays[w.wave & 1] + 2 * size_t(slot);
		const F4 r0 = qldF4(rp), r1 = qldF4(rp + 1);
		const float2 hrec = qld2(w.q.hits + slot);
		id = uint32_t(f2i(r1.w)) & WF_ID_MASK;
		bounce = f2i(r0.w);
		const uint32_t iteration = uint32_t(w.wave - bounce);
		if (w.splatSlots) splatSlot = iteration % w.splatSlots;
		const uint32_t sidx = (iteration % w.q.numSlots) * w.q.numIds + id;
		F4 T = ldState(w.q.stateT + sidx);
		DFrame f = w.f;
		wfSelectCache(f, iteration);
		wfShadeLightLane(f, w.lightFlags, f3(r0.x, r0.y, r0.z), f3(r1.x, r1.y, r1.z), bounce, hrec.x, f2i(hrec.y), T, out, pc);
		if (out.ray) stState(w.q.stateT + sidx, T);
	}
	// (with a reconstruction filter the centre pixel may lie outside the frame: the target word then only carries the flag, the position is in shadowPos)
	const bool inside = out.px >= 0 && out.py >= 0 && out.px < w.f.width && out.py < w.f.height;
	wfPushOut(w, out, WF_FLAG_SPLAT | (inside ? uint32_t(out.py) * uint32_t(w.f.width) + uint32_t(out.px) : 0u), bounce, id, splatSlot);
}

// shade(w) of the BPT eye pass: one lane per hit record of an eye path
GI_D void wfShadeEye(const WfFrame& w, uint32_t slot, bool valid, PathCounters& pc)
{
	WfPush out;
	out.ray = false; out.shadow = false;
	uint32_t id = 0, sidx = 0; int bounce = 0;
	if (valid)
	{
		const F4* rp = w.q.rays[w.wave & 1] + 2 * size_t(slot);
		const F4 r0 = qldF4(rp), r1 = qldF4(rp + 1);
		const float2 hrec = qld2(w.q.hits + slot);
		id = uint32_t(f2i(r1.w)) & WF_ID_MASK;
		bounce = f2i(r0.w);
		const uint32_t iteration = uint32_t(w.wave - bounce);
		sidx = (iteration % w.q.numSlots) * w.q.numIds + id;
		if (bounce == 0 && iteration >= w.q.numSlots)
		{
			// the slot is recycled: iteration - S is complete, its radiance joins the accumulator (pathtracer_bidir.comp:248-250)
			const float4 old = qld4(w.q.stateC + sidx);
			float4* px = reinterpret_cast<float4*>(w.f.accum) + id;
			float4 a = qld4(px);
			a.x += old.x; a.y += old.y; a.z += old.z; a.w += old.w;
			qst4(px, a);
		}
		F4 T = f4zero(), C = f4zero(); float M = 1.0f;
		if (bounce > 0) { T = ldState(w.q.stateT + sidx); C = ldState(w.q.stateC + sidx); M = qld1(w.q.stateM + sidx); }
		DFrame f = w.f;
		wfSelectCache(f, iteration);
		const uint32_t numEntries = *f.lightCacheCount; // complete: the light pass of this iteration ended before its eye pass began
		wfShadeEyeLane(f, numEntries, f3(r0.x, r0.y, r0.z), f3(r1.x, r1.y, r1.z), bounce, id, hrec.x, f2i(hrec.y), T, C, M, out, pc);
		stState(w.q.stateC + sidx, C);
		if (out.ray) { stState(w.q.stateT + sidx, T); qst1(w.q.stateM + sidx, M); }
	}
	wfPushOut(w, out, sidx, bounce, id);
}

} // namespace gik
