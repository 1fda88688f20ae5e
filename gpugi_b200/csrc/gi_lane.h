// gi_lane.h -- per-lane state machine of the persistent trace kernels (gi_wavefront.cuh): one ray, one step at a time.
// Plain per-thread code (no warp intrinsics), GI_HD so that tests/emul can run exactly this logic on the CPU against
// the oracle (test infrastructure only -- the product has no CPU path).
#pragma once

#include "gi_core.h"

namespace gik {

// per-thread work counters, flushed per warp at kernel end (order of GiCounters / CNT_* in gi_device.cu)
enum { WFC_CLOSEST = 0, WFC_ANY, WFC_CAMERA, WFC_LIGHT, WFC_SHADED, WFC_SPLATS, WFC_FALLBACK, WFC_NODES, WFC_TRIS, WFC_COUNT };

// ------------------------------------------------------------------ trace core
// Per-lane traversal state of the persistent trace kernels.  The arithmetic and the culling rules are those of
// traceFast() in gi_core.h, cut into single steps so that a warp can interleave rays.
//
// Exactness.  The reference's answer is a fold over the leaf triangles in preorder: a leaf is visited iff its own node
// box passes the slab test with first <= rayLength (its ancestors' boxes contain it, and the slab test is monotone in
// the box, so their tests are implied), and a triangle replaces the hit iff it is hit with t < rayLength
// (traceray.glsl:64-69,145).  The ordered walk below (pass 1) finds the closest candidate `best` and notices when another
// candidate lies within the window best*1e-4 + 1e-6*extent.  Only then the order of visits can matter, and only among
// candidates inside that window.  Such a ray is walked a second time (pass 2, same lane, culling limit fixed at the
// window) collecting every candidate inside the window with (t, first of its leaf box, leaf slot); the reference's fold is
// then replayed over the collected candidates in preorder (= leaf slot order, checked at upload).  Candidates in the
// upper half of the window (a guard band that keeps the collected set clearly separated from everything outside), more
// than WF_MAX_CAND candidates, a stack overflow or a scene whose leaves are not numbered in preorder fall back to the
// stackless reference walk itself.
#define WF_NODE_DONE int(0x80000000)
#define WF_MAX_CAND 8

enum { WFL_ANY = 1u, WFL_OVERFLOW = 2u, WFL_OCCLUDED = 4u, WFL_EXACT = 8u, WFL_COLLECT = 16u, WFL_GUARD = 32u };

struct WfCand { float t, first; int leaf; };

struct WfLane
{
	F3 o, d, inv;
	float tMaxInit, best, second;
	float limit;     // culling limit for box entry distances
	int hitLeaf;
	int node;        // >= 0 inner (fast index), < 0 leaf (~leaf), WF_NODE_DONE = finished
	int sp;
	uint32_t index;  // work item index
	uint32_t flags;  // WFL_*
	int numCand;
};

GI_HD float wfClosestLimit(float best, const DScene& s)
{
	const float margin = (1.0f + GI_AMBIGUITY_REL) * (1.0f + GI_AMBIGUITY_REL);
	return best * margin + s.ambiguityAbs;
}

GI_HD void wfPop(WfLane& L, const StackEntry* stack)
{
	for (;;)
	{
		if (L.sp == 0) { L.node = WF_NODE_DONE; return; }
		--L.sp;
		const StackEntry e = stack[L.sp];
		if (e.first <= L.limit) { L.node = e.node; return; }
	}
}

GI_HD void wfBegin(WfLane& L, const DScene& s, F3 o, F3 d, float tMax, bool isAny, uint32_t index, unsigned* cnt)
{
	L.o = o; L.d = d; L.inv = invDir(d);
	L.tMaxInit = tMax; L.best = tMax; L.second = GI_INF;
	L.limit = isAny ? tMax : wfClosestLimit(tMax, s);
	L.hitLeaf = -1; L.sp = 0; L.index = index;
	L.flags = isAny ? WFL_ANY : 0u; L.numCand = 0;
	// root box first, as the reference does (traceray.glsl:56-69)
	F4 a = ld4(s.refNodes), b = ld4(s.refNodes + 1);
	float f, l;
	++cnt[WFC_NODES];
	if (intersectBox(o, L.inv, a.x, a.y, a.z, b.x, b.y, b.z, f, l) && f <= tMax) L.node = s.rootIsLeaf ? ~0 : 0;
	else L.node = WF_NODE_DONE;
}

GI_HD void wfInnerStep(WfLane& L, const DScene& s, StackEntry* stack, unsigned* cnt)
{
	const F4* p = s.fastNodes + 4 * size_t(L.node);
	const F4 n0 = ld4(p), n1 = ld4(p + 1), n2 = ld4(p + 2), n3 = ld4(p + 3);
	cnt[WFC_NODES] += 2;
	float f0, l0, f1, l1;
	bool h0 = intersectBox(L.o, L.inv, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, f0, l0);
	bool h1 = intersectBox(L.o, L.inv, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, f1, l1);
	h0 = h0 && f0 <= L.limit;
	h1 = h1 && f1 <= L.limit;
	int c0 = f2i(n3.x), c1 = f2i(n3.y);
	if (h0 && h1)
	{
		if (f1 < f0) { int t = c0; c0 = c1; c1 = t; float tf = f0; f0 = f1; f1 = tf; }
		if (L.sp < GI_STACK_SIZE) { stack[L.sp].node = c1; stack[L.sp].first = f1; ++L.sp; }
		else L.flags |= WFL_OVERFLOW;
		L.node = c0;
	}
	else if (h0) L.node = c0;
	else if (h1) L.node = c1;
	else wfPop(L, stack);
}

GI_HD void wfLeafStep(WfLane& L, const DScene& s, const StackEntry* stack, WfCand* cand, unsigned* cnt)
{
	const int leaf = ~L.node;
#pragma unroll
	for (int slot = 0; slot < 2; ++slot)
	{
		const int lt = 2 * leaf + slot;
		F3 p0, p1, p2; int valid;
		loadLeafTri(s, lt, p0, p1, p2, valid);
		if (!valid) break;
		++cnt[WFC_TRIS];
		float t; F3 nb, nn;
		if (intersectTriangle(L.o, L.d, p0, p1, p2, t, nb, nn) && t < L.tMaxInit)
		{
			if (L.flags & WFL_ANY) { L.flags |= WFL_OCCLUDED; L.node = WF_NODE_DONE; return; }
			if (L.flags & WFL_COLLECT)
			{
				// pass 2: L.best is the result of pass 1 and stays fixed
				const float window = L.best * GI_AMBIGUITY_REL + s.ambiguityAbs;
				if (t <= L.best + window)
				{
					if (t > L.best + 0.5f * window) L.flags |= WFL_GUARD;
					else if (L.numCand < WF_MAX_CAND)
					{
						// first of the leaf's own node box (the reference's cull test for this leaf)
						const uint32_t nodeIdx = s.leafNode[leaf];
						const F4 a = ld4(s.refNodes + 2 * size_t(nodeIdx)), b = ld4(s.refNodes + 2 * size_t(nodeIdx) + 1);
						float f, l;
						intersectBox(L.o, L.inv, a.x, a.y, a.z, b.x, b.y, b.z, f, l);
						cand[L.numCand].t = t; cand[L.numCand].first = f; cand[L.numCand].leaf = lt;
						++L.numCand;
					}
					else L.flags |= WFL_OVERFLOW;
				}
			}
			else if (t < L.best)
			{
				if (L.hitLeaf >= 0) L.second = L.best;
				L.best = t;
				L.hitLeaf = lt;
				L.limit = wfClosestLimit(t, s);
			}
			else
				L.second = fminf(L.second, t);
		}
	}
	wfPop(L, stack);
}

// Called when a lane's walk ran out of nodes.  Returns true when the ray's result (L.best, L.hitLeaf / WFL_OCCLUDED) is
// final and equal to the reference's; false when the lane was restarted for the collecting second pass.
GI_HD bool wfEndOfWalk(WfLane& L, const DScene& s, WfCand* cand, unsigned* cnt)
{
	if (L.flags & WFL_EXACT) return true;
	TraceStats st = {0, 0};
	bool done = true;
	if (L.flags & WFL_ANY)
	{
		// any-hit is order independent; only a stack overflow (skipped subtree) needs the stackless walk
		if (!(L.flags & WFL_OCCLUDED) && (L.flags & WFL_OVERFLOW))
		{
			float len = L.tMaxInit; int dummy;
			if (traceReference<true>(s, L.o, L.d, len, dummy, &st)) L.flags |= WFL_OCCLUDED;
		}
	}
	else
	{
		bool walk = (L.flags & WFL_OVERFLOW) != 0u;
		if (!(L.flags & WFL_COLLECT))
		{
			bool ambiguous = walk;
			if (L.hitLeaf >= 0)
			{
				const float window = L.best * GI_AMBIGUITY_REL + s.ambiguityAbs;
				if (L.second <= L.best + window) ambiguous = true;   // another triangle (almost) as close: order matters
				if (L.best + window >= L.tMaxInit) ambiguous = true; // hit (almost) at the caller's length limit
			}
			if (ambiguous)
			{
				++cnt[WFC_FALLBACK];
				if (!walk && s.leafPreorder)
				{
					// pass 2: same lane, culling limit fixed at the window of pass 1's result
					L.flags |= WFL_COLLECT;
					L.numCand = 0; L.sp = 0;
					L.limit = wfClosestLimit(L.best, s);
					L.node = s.rootIsLeaf ? ~0 : 0;
					done = false;
				}
				else
					walk = true;
			}
		}
		else if (!walk && !(L.flags & WFL_GUARD))
		{
			// replay the reference's fold over the collected candidates in preorder
			for (int i = 1; i < L.numCand; ++i)
			{
				const WfCand c = cand[i];
				int j = i - 1;
				while (j >= 0 && cand[j].leaf > c.leaf) { cand[j + 1] = cand[j]; --j; }
				cand[j + 1] = c;
			}
			// (the box test gates a whole leaf: it is made once, with the rayLength at the time the leaf is reached)
#if !defined(__CUDA_ARCH__) && defined(GI_LANE_DEBUG)
			for (int i = 0; i < L.numCand; ++i) printf("cand %d: t=%.9g first=%.9g leaf=%d\n", i, cand[i].t, cand[i].first, cand[i].leaf);
			printf("pass1 best=%.9g hitLeaf=%d second=%.9g\n", L.best, L.hitLeaf, L.second);
#endif
			float r = L.tMaxInit; int hit = -1;
			int curLeaf = -1; bool entered = false;
			for (int i = 0; i < L.numCand; ++i)
			{
				if ((cand[i].leaf >> 1) != curLeaf) { curLeaf = cand[i].leaf >> 1; entered = cand[i].first <= r; }
				if (entered && cand[i].t < r) { r = cand[i].t; hit = cand[i].leaf; }
			}
			L.best = r; L.hitLeaf = hit;
		}
		else
			walk = true;
		if (walk)
		{
			L.best = L.tMaxInit; L.hitLeaf = -1;
			traceReference<false>(s, L.o, L.d, L.best, L.hitLeaf, &st);
		}
	}
	cnt[WFC_NODES] += st.nodes; cnt[WFC_TRIS] += st.tris;
	return done;
}

// One ray from start to end on one lane (what a warp does for 32 rays at once, interleaved with other rays).
GI_HD void wfWalkSingle(WfLane& L, const DScene& s, F3 o, F3 d, float tMax, bool isAny, unsigned* cnt)
{
	StackEntry stack[GI_STACK_SIZE];
	WfCand cand[WF_MAX_CAND];
	wfBegin(L, s, o, d, tMax, isAny, 0u, cnt);
	for (;;)
	{
		if (L.node == WF_NODE_DONE) { if (wfEndOfWalk(L, s, cand, cnt)) return; continue; }
		if (L.node < 0) wfLeafStep(L, s, stack, cand, cnt);
		else wfInnerStep(L, s, stack, cnt);
	}
}

} // namespace gik
