// gi_lane.h -- per-lane state machine of the persistent trace kernels (gi_wavefront.cuh): one ray, one step at a time.
// Plain per-thread code (no warp intrinsics), GI_HD so that tests/emul can run exactly this logic on the CPU against
// the oracle (test infrastructure only -- the product has no CPU path).
#pragma once

#include "gi_core.h"

namespace gik {

// per-thread work counters, flushed per warp at kernel end (order of GiCounters / CNT_* in gi_device.cu)
enum { WFC_CLOSEST = 0, WFC_ANY, WFC_CAMERA, WFC_LIGHT, WFC_SHADED, WFC_SPLATS, WFC_FALLBACK, WFC_NODES, WFC_TRIS, WFC_INNER_STEPS, WFC_LEAF_STEPS, WFC_COUNT };

// STEPS = false drops the box / triangle test counters (GiCounters::node_tests, triangle_tests) from the hot loop.
template <bool STEPS>
struct WfCounters
{
	unsigned v[WFC_COUNT];
	GI_HD WfCounters() { for (int i = 0; i < WFC_COUNT; ++i) v[i] = 0u; }
	template <int I> GI_HD void add(unsigned n) { if (STEPS || (I != WFC_NODES && I != WFC_TRIS && I != WFC_INNER_STEPS && I != WFC_LEAF_STEPS)) v[I] += n; }
};

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
// build-time switches of the trace loop, A/B measured on B200 (tools/gpu/variants.sh + run_variants.sh, PT batch, ms/iteration):
// all off 2.221 | raw prmt 2.206 | lazy leaf box 2.332 | top of stack in registers 2.265 | postponed leaf 2.267 vs 2.190 -> only raw prmt is on
#ifndef WF_RAW_PRMT
#define WF_RAW_PRMT 1      // prmt.b32 without __byte_perm's selector mask
#endif
#ifndef WF_LAZY_LEAF_BOX
#define WF_LAZY_LEAF_BOX 0 // leaf: triangle tests first, the leaf's own exact box test only when a triangle is hit
#endif
#ifndef WF_POSTPONE_LEAF
#define WF_POSTPONE_LEAF 0 // a lane that reaches a leaf stashes it and keeps stepping inner nodes until the warp's next leaf step
#endif
#ifndef WF_TOP_IN_REGS
#define WF_TOP_IN_REGS 0   // most recent stack entry in registers
#endif
#ifndef WF_WIDE4
#define WF_WIDE4 1         // 4-wide inner nodes (gi_layout.h: emitWide4): 64-byte records, two 256-bit loads, up to three pushes per step.
#endif                     // B200 A/B against the binary nodes (0), ms per iteration: C2 PT 1.810 -> 1.765, C3 BPT 8.82 -> 8.38, C4 LPT 4K 8.06 -> 7.32, C5 PT 4K 17.77 -> 15.42
#ifndef WF_CENTRE_EXTENT
#define WF_CENTRE_EXTENT 1 // 4-wide nodes with the child boxes as (centre: float, half extent: bfloat16): 96-byte records, three 256-bit loads, NO decode
#endif                     // and NO near / far selection in the slab test (needs WF_WIDE4, WF_CACHE_INV): per step 24 PRMT + 24 FFMA become 36 FFMA + 6 shifts, the
                           // lane state loses 15 registers (71 registers without spills instead of 72 with; 9 blocks per SM at 56).  B200 A/B against the 16-bit grid
                           // records (0) at 7 blocks per SM, ms per iteration: C2 1.680 -> 1.663, C3 7.462 -> 7.281, C4 6.580 -> 6.590, C5 14.82 -> 15.02 (50 % more node
                           // bytes); at 9 blocks per SM: C2 1.645, C3 7.175, C4 6.509, C5 14.66 (8: 1.649 / - / - / 14.74; 10: 1.705 / 7.468 / 6.841 / -).  +0.8 % box tests.
// (History: WF_FLOAT_PLANES -- 128-byte records with the 24 planes as floats grouped per axis so that the load address picks the near / far side,
//  seven 128-bit loads per step -- measured 1.786 vs 1.722 ms on C2 and 15.96 vs 14.95 ms on C5 in round 1: the L1 traffic cost more than the
//  saved decode.  Removed in favour of the centre / half-extent records, which save the same decode with three loads.)
#define WF_NODE_WORDS (WF_CENTRE_EXTENT ? 24 : (WF_WIDE4 ? 16 : 8)) // words per record of DScene::qNodes
#ifndef WF_STEP_OVERFLOW_CHECK
#define WF_STEP_OVERFLOW_CHECK 0 // one stack-capacity check per 4-wide step instead of one per push (B200: 1.660 vs 1.653 ms, off)
#endif
#ifndef WF_CSWAP_MINMAX
#define WF_CSWAP_MINMAX 1 // B200: 1.673 -> 1.653 ms per PT iteration
#endif
#ifndef WF_SORT_EXCHANGES
#define WF_SORT_EXCHANGES 5 // 4-wide step: 5 = children fully sorted by entry distance, 3 = nearest exact, the pushed ones roughly ordered
#endif
#ifndef WF_CACHE_INV
#define WF_CACHE_INV 1     // 1/d kept in the lane state for the leaf gate instead of three IEEE divisions per leaf step (B200: 1.712 -> 1.692 ms)
#endif
#if WF_CENTRE_EXTENT && !(WF_WIDE4 && WF_CACHE_INV)
#error "WF_CENTRE_EXTENT needs WF_WIDE4 and WF_CACHE_INV"
#endif
#ifndef WF_ANY_UNSORTED
#define WF_ANY_UNSORTED 1  // inner-node steps of a warp whose walking lanes are all any-hit rays (the queue hands out the closest-hit rays first, then the shadow
#endif                     // rays: 40-45 % of all rays) skip the sorting network -- the answer of an any-hit ray does not depend on the visit order -- and push
                           // without entry distances.  B200 A/B (ms per iteration, sorted -> unsorted): C2 1.659 -> 1.642, C3 7.423 -> 7.243, C4 6.335 -> 6.290, C5 14.53 -> 14.28
#define WF_MAX_CAND 8

enum { WFL_ANY = 1u, WFL_OVERFLOW = 2u, WFL_OCCLUDED = 4u, WFL_EXACT = 8u, WFL_COLLECT = 16u, WFL_GUARD = 32u };

struct WfCand { float t, first; int leaf; };

struct WfLane
{
	F3 o, d;
#if WF_CACHE_INV
	F3 inv;          // invDir(d), the same three IEEE quotients the leaf gate would recompute
#endif
#if WF_CENTRE_EXTENT
	// conservative inner-node test on (centre, half extent): distance of the box centre along axis k = fma(c, inv.k, negOInv.k), near / far = -/+ h * |inv.k|
	F3 negOInv;      // -(o * (1/d))
#else
	// conservative inner-node test on the 16-bit grid (wfSlabTest): distance of grid plane q along axis k
	//   near planes  fma(2^23 + q, cellInv.k, biasNear.k)      far planes  fma(2^23 + q, cellInv.k, biasFar.k)
	F3 cellInv;      // cell size * (1/d)
	F3 biasNear;     // (gridLo - o) * (1/d) - 2^23 * cellInv -/+ error bound
	F3 biasFar;
	uint32_t selNear[3], selFar[3]; // byte-permute selectors: which 16-bit half of a node word is the near / far plane (sign of d)
#endif
	float tMaxInit, best, second;
	float limit;     // culling limit for box entry distances
	int hitLeaf;
	int node;        // >= 0 inner (fast index), < 0 leaf (~leaf), WF_NODE_DONE = finished
	int pend;        // WF_POSTPONE_LEAF: a reached but not yet tested leaf (~leaf < 0), 0 = none.  A ray is finished when node is
	                 // WF_NODE_DONE and pend is 0.  Testing a leaf later only delays the shrinking of the culling limit.
	// traversal stack: the most recent entry lives in registers (topNode, topFirst), the older ones in stack[0 .. sp).  A pop
	// hands out the register entry at once and refills it with a load whose result is only needed at the NEXT pop or push,
	// so the walk does not wait on local memory unless the popped entry is culled.  Bottom = sentinel (WF_NODE_DONE, 0).
	int topNode;
	float topFirst;
	int sp;
	uint32_t index;  // work item index
	uint32_t flags;  // WFL_*
	int numCand;
};

GI_HD F3 wfInvDir(const WfLane& L)
{
#if WF_CACHE_INV
	return L.inv;
#else
	return invDir(L.d);
#endif
}

GI_HD float wfClosestLimit(float best, const DScene& s)
{
	const float margin = (1.0f + GI_AMBIGUITY_REL) * (1.0f + GI_AMBIGUITY_REL);
	return best * margin + s.ambiguityAbs;
}

// Storage of the per-lane traversal stack behind a two-function interface, so that the kernel can keep the first entries in
// shared memory (WfStackShared, gi_wavefront.cuh) while the emulation uses a plain array.
struct WfStackArray
{
	StackEntry* e;
	GI_HD StackEntry load(int i) const { return e[i]; }
	GI_HD void store(int i, int node, float first) const { e[i].node = node; e[i].first = first; }
};

GI_HD void wfStackReset(WfLane& L) { L.sp = 0; L.topNode = WF_NODE_DONE; L.topFirst = 0.0f; L.pend = 0; }

template <class STK>
GI_HD void wfPush(WfLane& L, const STK& stack, int node, float first)
{
	if (L.sp < GI_STACK_SIZE)
	{
#if WF_TOP_IN_REGS
		stack.store(L.sp, L.topNode, L.topFirst); ++L.sp;
		L.topNode = node; L.topFirst = first;
#else
		stack.store(L.sp, node, first); ++L.sp;
#endif
	}
	else L.flags |= WFL_OVERFLOW; // the subtree is skipped; the ray is answered by the stackless reference walk
}

// push without the capacity check: the caller has made sure that the entries of this step fit (WF_STEP_OVERFLOW_CHECK)
template <class STK>
GI_HD void wfPushUnchecked(WfLane& L, const STK& stack, int node, float first)
{
#if WF_TOP_IN_REGS
	wfPush(L, stack, node, first);
#else
	stack.store(L.sp, node, first); ++L.sp;
#endif
}

template <class STK>
GI_HD void wfPop(WfLane& L, const STK& stack)
{
#if WF_TOP_IN_REGS
	for (;;)
	{
		const int n = L.topNode;
		const float f = L.topFirst;
		if (n == WF_NODE_DONE) { L.node = WF_NODE_DONE; return; } // the sentinel stays
		--L.sp; // >= 0: the sentinel was stored below the first pushed entry
		const StackEntry e = stack.load(L.sp);
		L.topNode = e.node; L.topFirst = e.first;
		if (f <= L.limit) { L.node = n; return; }
	}
#else
	for (;;)
	{
		if (L.sp == 0) { L.node = WF_NODE_DONE; return; }
		--L.sp;
		const StackEntry e = stack.load(L.sp);
		if (e.first <= L.limit) { L.node = e.node; return; }
	}
#endif
}

// WF_POSTPONE_LEAF: a lane standing at a leaf with an empty stash moves the leaf into the stash and goes on with the next node
template <class STK>
GI_HD void wfStash(WfLane& L, const STK& stack)
{
#if WF_POSTPONE_LEAF
	if (L.node < 0 && L.node != WF_NODE_DONE && L.pend == 0) { L.pend = L.node; wfPop(L, stack); }
#endif
}

GI_HD bool wfFinished(const WfLane& L)
{
#if WF_POSTPONE_LEAF
	return L.node == WF_NODE_DONE && L.pend == 0;
#else
	return L.node == WF_NODE_DONE;
#endif
}

GI_HD bool wfAtLeaf(const WfLane& L)
{
#if WF_POSTPONE_LEAF
	return L.pend != 0;
#else
	return L.node < 0 && L.node != WF_NODE_DONE;
#endif
}

// ---- conservative slab test on the quantised child boxes
// A child box is stored on a 16-bit grid over the root box, rounded outwards: plane q of axis k lies at gridLo.k + q * cell.k
// (real numbers; checked in double at upload).  Per axis one 32-bit word: min plane in the low half, max plane in the high half.
// A coordinate becomes a float with ONE byte-permute, no conversion and no mask: bytes (q.lo, q.hi, 0x00, 0x4B) are the float
// 2^23 + q.  Which half is the ray's near plane depends only on the sign of the direction, so the selector is per-ray state.
//
// Error bound (u = 2^-24, all symbols per axis).  True distance of plane q: T = (gridLo + q*cell - o) * inv.  Computed:
//   S = fl(cell * inv)   B0 = fl(fl(gridLo - o) * inv)   b = fl(B0 - 2^23*S)   b' = fl(b -/+ E)   t = fl((2^23 + q) * S + b')   (one FMA)
//   |(2^23+q)*S + b - T| <= u*2^16*|S| + 3.02u*|B| + u*(|B0| + 2^23*|S|)      (B = (gridLo - o) * inv)
//   rounding of b'        <= u*(|B0| + 2^23*|S| + E)
//   rounding of the FMA   <= u*(|B| + 2^16*|S|) (1 + ...)
//   the exact-zone value the reference computes for a plane p of the real box, fl(inv * fl(p - o)), differs from its
//   real-number value by <= 2.01u*(|B| + 2^16*|S|)
// in total <= 1.016*|S| + 7.04u*|B| + u*E  <  E := 1.02*|S| + 7.5u*|B0|: about one grid cell.  Hence
//   near.k <= tmin.k and far.k >= tmax.k of IntersectBox (intersectiontests.glsl:12-15) for every box inside its grid box, and
//   (first <= last && first <= limit)  ==>  wfSlabTest accepts.
// Rays with a zero / denormal direction component or products near overflow are outside the bound's domain (inf - inf,
// 0 * inf): they are answered by the stackless reference walk (WFL_OVERFLOW).
#define WF_SEL_LO 0x7610u
#define WF_SEL_HI 0x7632u

GI_HD float wfFma(float a, float b, float c)
{
#if defined(__CUDA_ARCH__)
	return __fmaf_rn(a, b, c);
#else
	return fmaf(a, b, c);
#endif
}

GI_HD float wfPlane(uint32_t word, uint32_t sel)
{
#if defined(__CUDA_ARCH__)
	// prmt.b32 directly: __byte_perm() masks the selector with 0x7777 first (one LOP3 per plane in the hot loop); the two
	// selectors used here (WF_SEL_LO / WF_SEL_HI) have no mode bits set, so the raw instruction is the same function
#if WF_RAW_PRMT
	uint32_t r;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(word), "r"(0x4B000000u), "r"(sel));
	return __uint_as_float(r);
#else
	return __uint_as_float(__byte_perm(word, 0x4B000000u, sel));
#endif
#else
	const uint32_t q = (sel == WF_SEL_LO) ? (word & 0xffffu) : (word >> 16);
	return i2f(int(0x4B000000u | q));
#endif
}

GI_HD float wfMax3(float a, float b, float c)
{
#if defined(__CUDA_ARCH__)
	float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;
#else
	return fmaxf(fmaxf(a, b), c);
#endif
}
GI_HD float wfMin3(float a, float b, float c)
{
#if defined(__CUDA_ARCH__)
	float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;
#else
	return fminf(fminf(a, b), c);
#endif
}

#if WF_CENTRE_EXTENT
// Centre / half-extent records.  Coordinates are taken relative to the centre C0 of the root box (a float per axis, DScene::ceCentre), so that the
// bound does not depend on where the scene lies: p' = p - C0, o' = o - C0 (real numbers), P = the largest half extent of the root box.
// Per child and axis the record holds a float c and a bfloat16 h with [c - h0, c + h0] containing the child's box (in primed coordinates, real
// arithmetic) and h = h0 + e, e >= 330 u P (u = 2^-24; gi_layout.h: emitWide4 checks both in double).  With inv = fl(1/d) and
// T(p) = (p - o) inv = (p' - o') inv the real-number distance of plane p, the reference (exact zone) computes
//   t_ref(p) = fl(inv * fl(p - o)),   |t_ref(p) - T(p)| <= 2.01 u |T(p)| <= 2.01 u (P + |o'|) |inv|.
// Here:   q = fl(o - C0)   m = fl(-(q inv))   tc = fl(c inv + m)   near = fl(tc - h |inv|)   far = fl(tc + h |inv|)     (three FMAs per plane pair, no decode, no selection)
//   |q - o'| <= u |o'|,   |m + o' inv| <= 2.01 u |o' inv|,   |tc - (c - o') inv| <= 2.02 u |o' inv| + u |c - o'| |inv|
//   |near - ((c - o') inv - h |inv|)| <= 1.01 u |inv| (4 |o'| + 2 |c| + h) <= 1.01 u |inv| (4 |o'| + 3 P''),   P'' = 1.02 P >= |c| + h (checked at upload)
// A true plane p' of the child lies at least e inside [c - h, c + h], so T(p) is at least e |inv| inside [(c - o') inv -/+ h |inv|], and
//   near <= tmin.k,  far >= tmax.k of IntersectBox (intersectiontests.glsl:12-15)   as soon as   e >= u (2.01 (P + |o'|) + 4.04 |o'| + 3.1 P) ,
// which e = 330 u P satisfies for every ray with |o'.k| <= 50 P.  Hence (first <= last && first <= limit) ==> the conservative test accepts.
// Rays outside that domain (origin farther than 50 P from the centre of the scene, zero / denormal direction components, products near
// overflow or underflow) are answered by the stackless reference walk (WFL_OVERFLOW).
GI_HD bool wfSlabSetup(WfLane& L, const DScene& s, F3 o, F3 inv)
{
	const float P = s.cePmax;
	const float iv[3] = {inv.x, inv.y, inv.z};
	const float ov[3] = {GI_SUB(o.x, s.ceCentre[0]), GI_SUB(o.y, s.ceCentre[1]), GI_SUB(o.z, s.ceCentre[2])};
	float m[3];
	bool ok = true;
	for (int k = 0; k < 3; ++k)
	{
		m[k] = -GI_MUL(ov[k], iv[k]);
		const float reach = fmaxf(fabsf(ov[k]), P) * fabsf(iv[k]);
		ok = ok && (fabsf(ov[k]) <= 50.0f * P) && (reach < 1.0e30f) && (P * fabsf(iv[k]) >= 1.0e-30f); // all false for NaN
	}
	L.negOInv = f3(m[0], m[1], m[2]);
	return ok;
}
#else
// Per-ray constants of the conservative test.  Returns false when the ray is outside the domain of the error bound.
GI_HD bool wfSlabSetup(WfLane& L, const DScene& s, F3 o, F3 inv)
{
	const float u = 5.9604644775390625e-08f; // 2^-24
	const float S[3] = {s.gridScale[0] * inv.x, s.gridScale[1] * inv.y, s.gridScale[2] * inv.z};
	const float B0[3] = {(s.gridLo[0] - o.x) * inv.x, (s.gridLo[1] - o.y) * inv.y, (s.gridLo[2] - o.z) * inv.z};
	float bn[3], bf[3], m = 0.0f;
	for (int k = 0; k < 3; ++k)
	{
		const float b = B0[k] - 8388608.0f * S[k];
		const float E = 1.02f * fabsf(S[k]) + (7.5f * u) * fabsf(B0[k]) + 1.0e-37f;
		bn[k] = b - E; bf[k] = b + E;
		m = fmaxf(m, 16777216.0f * fabsf(S[k]) + fabsf(B0[k]));
	}
	L.cellInv = f3(S[0], S[1], S[2]);
	L.biasNear = f3(bn[0], bn[1], bn[2]);
	L.biasFar = f3(bf[0], bf[1], bf[2]);
	L.selNear[0] = inv.x < 0.0f ? WF_SEL_HI : WF_SEL_LO; L.selFar[0] = inv.x < 0.0f ? WF_SEL_LO : WF_SEL_HI;
	L.selNear[1] = inv.y < 0.0f ? WF_SEL_HI : WF_SEL_LO; L.selFar[1] = inv.y < 0.0f ? WF_SEL_LO : WF_SEL_HI;
	L.selNear[2] = inv.z < 0.0f ? WF_SEL_HI : WF_SEL_LO; L.selFar[2] = inv.z < 0.0f ? WF_SEL_LO : WF_SEL_HI;
	return m < 1.0e30f; // false for NaN / inf as well
}

#endif

template <class CNT>
GI_HD void wfBegin(WfLane& L, const DScene& s, F3 o, F3 d, float tMax, bool isAny, uint32_t index, CNT& cnt)
{
	const F3 inv = invDir(d);
	L.o = o; L.d = d;
#if WF_CACHE_INV
	L.inv = inv;
#endif
	L.tMaxInit = tMax; L.best = tMax; L.second = GI_INF;
	L.limit = isAny ? tMax : wfClosestLimit(tMax, s);
	L.hitLeaf = -1; wfStackReset(L); L.index = index;
	L.flags = isAny ? uint32_t(WFL_ANY) : 0u; L.numCand = 0;
	const bool inDomain = wfSlabSetup(L, s, o, inv);
	// root box first, as the reference does (traceray.glsl:56-69)
	F4 a = ld4(s.refNodes), b = ld4(s.refNodes + 1);
	float f, l;
	cnt.template add<WFC_NODES>(1u);
	if (intersectBox(o, inv, a.x, a.y, a.z, b.x, b.y, b.z, f, l) && f <= tMax) L.node = s.rootIsLeaf ? ~0 : 0;
	else L.node = WF_NODE_DONE;
	if (!inDomain && L.node != WF_NODE_DONE) { L.flags |= WFL_OVERFLOW; L.node = WF_NODE_DONE; }
}

#if !WF_CENTRE_EXTENT
// Conservative accept test of one child box (three words: x, y, z; min plane | max plane << 16): never rejects a box the
// exact-zone test (first <= last && first <= limit) accepts.  nearLow is a lower bound of the exact-zone entry distance.
GI_HD bool wfSlabTest(const WfLane& L, uint32_t wx, uint32_t wy, uint32_t wz, float& nearLow)
{
	const float nx = wfFma(wfPlane(wx, L.selNear[0]), L.cellInv.x, L.biasNear.x);
	const float ny = wfFma(wfPlane(wy, L.selNear[1]), L.cellInv.y, L.biasNear.y);
	const float nz = wfFma(wfPlane(wz, L.selNear[2]), L.cellInv.z, L.biasNear.z);
	const float fx = wfFma(wfPlane(wx, L.selFar[0]), L.cellInv.x, L.biasFar.x);
	const float fy = wfFma(wfPlane(wy, L.selFar[1]), L.cellInv.y, L.biasFar.y);
	const float fz = wfFma(wfPlane(wz, L.selFar[2]), L.cellInv.z, L.biasFar.z);
	nearLow = fmaxf(wfMax3(nx, ny, nz), 0.0f);
	return nearLow <= fminf(wfMin3(fx, fy, fz), L.limit);
}
#endif

#if WF_CENTRE_EXTENT
// 96-byte record (24 words): 0..11 centre of child k relative to the root box's centre, axis a at 3k + a (float); 12..17 half extents as bfloat16 pairs, word 12 + 3 (k >> 1) + a =
// (h of the even child | h of the odd child << 16); 18..19 unused; 20..23 child links (wide index >= 0, leaf = ~leafIndex).  The odd child's half
// extent is the word itself read as a float (the even child's bits only enlarge it, by < 2^-7), the even child's is the word shifted left by 16 --
// a multiplication by 65536 on the FMA pipe.  An unused slot has c = 0 and h = -65536 P: near > far for every ray inside the domain.
GI_HD float wfHalfExtent(uint32_t word, int child)
{
	return i2f(int((child & 1) ? word : word * 65536u));
}

GI_HD void wfCentreTest(const WfLane& L, float cx, float cy, float cz, float hx, float hy, float hz, float& nearLow, float& farHigh)
{
	const float tx = wfFma(cx, L.inv.x, L.negOInv.x), ty = wfFma(cy, L.inv.y, L.negOInv.y), tz = wfFma(cz, L.inv.z, L.negOInv.z);
	const float ax = fabsf(L.inv.x), ay = fabsf(L.inv.y), az = fabsf(L.inv.z);
	nearLow = fmaxf(wfMax3(wfFma(hx, -ax, tx), wfFma(hy, -ay, ty), wfFma(hz, -az, tz)), 0.0f);
	farHigh = fminf(wfMin3(wfFma(hx, ax, tx), wfFma(hy, ay, ty), wfFma(hz, az, tz)), L.limit);
}

// the conservative test of ONE child of the record at `node` (test hook; wfInnerStep does the same for all four)
GI_HD bool wfSlabTestChild(const WfLane& L, const uint32_t* node, int child, float& nearLow)
{
	const int hb = 12 + 3 * (child >> 1);
	float farHigh;
	wfCentreTest(L, i2f(int(node[3 * child])), i2f(int(node[3 * child + 1])), i2f(int(node[3 * child + 2])),
	             wfHalfExtent(node[hb], child), wfHalfExtent(node[hb + 1], child), wfHalfExtent(node[hb + 2], child), nearLow, farHigh);
	return nearLow <= farHigh;
}

template <class STK, class CNT>
GI_HD void wfInnerStep(WfLane& L, const DScene& s, const STK& stack, CNT& cnt)
{
	const uint32_t* nb = s.qNodes + 24 * size_t(L.node);
	const U8 a = ld8(nb), b = ld8(nb + 8), c = ld8(nb + 16);
	cnt.template add<WFC_NODES>(4u); cnt.template add<WFC_INNER_STEPS>(1u);
	float f0, f1, f2, f3v, far;
	wfCentreTest(L, i2f(int(a.v[0])), i2f(int(a.v[1])), i2f(int(a.v[2])), wfHalfExtent(b.v[4], 0), wfHalfExtent(b.v[5], 0), wfHalfExtent(b.v[6], 0), f0, far);
	if (!(f0 <= far)) f0 = GI_INF;
	wfCentreTest(L, i2f(int(a.v[3])), i2f(int(a.v[4])), i2f(int(a.v[5])), wfHalfExtent(b.v[4], 1), wfHalfExtent(b.v[5], 1), wfHalfExtent(b.v[6], 1), f1, far);
	if (!(f1 <= far)) f1 = GI_INF;
	wfCentreTest(L, i2f(int(a.v[6])), i2f(int(a.v[7])), i2f(int(b.v[0])), wfHalfExtent(b.v[7], 2), wfHalfExtent(c.v[0], 2), wfHalfExtent(c.v[1], 2), f2, far);
	if (!(f2 <= far)) f2 = GI_INF;
	wfCentreTest(L, i2f(int(b.v[1])), i2f(int(b.v[2])), i2f(int(b.v[3])), wfHalfExtent(b.v[7], 3), wfHalfExtent(c.v[0], 3), wfHalfExtent(c.v[1], 3), f3v, far);
	if (!(f3v <= far)) f3v = GI_INF;
	int c0 = int(c.v[4]), c1 = int(c.v[5]), c2 = int(c.v[6]), c3 = int(c.v[7]);
#define WF_CSWAP(fa, ca, fb, cb) { const bool sw = fb < fa; const float lo = fminf(fa, fb), hi = fmaxf(fa, fb); const int cl = sw ? cb : ca, ch = sw ? ca : cb; fa = lo; fb = hi; ca = cl; cb = ch; }
	WF_CSWAP(f0, c0, f1, c1) WF_CSWAP(f2, c2, f3v, c3) WF_CSWAP(f0, c0, f2, c2) WF_CSWAP(f1, c1, f3v, c3) WF_CSWAP(f1, c1, f2, c2)
#undef WF_CSWAP
	if (f0 < GI_INF)
	{
		if (f3v < GI_INF) wfPush(L, stack, c3, f3v);
		if (f2 < GI_INF) wfPush(L, stack, c2, f2);
		if (f1 < GI_INF) wfPush(L, stack, c1, f1);
		L.node = c0;
	}
	else wfPop(L, stack);
	wfStash(L, stack);
}


// WF_ANY_UNSORTED: the same step for an any-hit ray when the order of the visits is nobody's business: no sorting network, the first accepted
// child (in slot order) is next, the others are pushed without an entry distance (0 passes every cull; an any-hit ray's limit never shrinks)
template <class STK, class CNT>
GI_HD void wfInnerStepAny(WfLane& L, const DScene& s, const STK& stack, CNT& cnt)
{
	const uint32_t* nb = s.qNodes + 24 * size_t(L.node);
	const U8 a = ld8(nb), b = ld8(nb + 8), c = ld8(nb + 16);
	cnt.template add<WFC_NODES>(4u); cnt.template add<WFC_INNER_STEPS>(1u);
	float n0, n1, n2, n3, r0, r1, r2, r3;
	wfCentreTest(L, i2f(int(a.v[0])), i2f(int(a.v[1])), i2f(int(a.v[2])), wfHalfExtent(b.v[4], 0), wfHalfExtent(b.v[5], 0), wfHalfExtent(b.v[6], 0), n0, r0);
	wfCentreTest(L, i2f(int(a.v[3])), i2f(int(a.v[4])), i2f(int(a.v[5])), wfHalfExtent(b.v[4], 1), wfHalfExtent(b.v[5], 1), wfHalfExtent(b.v[6], 1), n1, r1);
	wfCentreTest(L, i2f(int(a.v[6])), i2f(int(a.v[7])), i2f(int(b.v[0])), wfHalfExtent(b.v[7], 2), wfHalfExtent(c.v[0], 2), wfHalfExtent(c.v[1], 2), n2, r2);
	wfCentreTest(L, i2f(int(b.v[1])), i2f(int(b.v[2])), i2f(int(b.v[3])), wfHalfExtent(b.v[7], 3), wfHalfExtent(c.v[0], 3), wfHalfExtent(c.v[1], 3), n3, r3);
	const bool p0 = n0 <= r0, p1 = n1 <= r1, p2 = n2 <= r2, p3 = n3 <= r3;
	const int c0 = int(c.v[4]), c1 = int(c.v[5]), c2 = int(c.v[6]), c3 = int(c.v[7]);
	if (p0 || p1 || p2 || p3)
	{
		if (p3 && (p0 || p1 || p2)) wfPush(L, stack, c3, 0.0f);
		if (p2 && (p0 || p1)) wfPush(L, stack, c2, 0.0f);
		if (p1 && p0) wfPush(L, stack, c1, 0.0f);
		L.node = p0 ? c0 : (p1 ? c1 : (p2 ? c2 : c3));
	}
	else wfPop(L, stack);
}
#elif WF_WIDE4
GI_HD bool wfSlabTestChild(const WfLane& L, const uint32_t* node, int child, float& nearLow) { return wfSlabTest(L, node[3 * child], node[3 * child + 1], node[3 * child + 2], nearLow); }

// 4-wide inner node: words 0..11 = child k, axis a at 3k + a (min plane | max plane << 16), words 12..15 = child links (wide index >= 0,
// leaf = ~leafIndex).  An unused slot carries the empty box (min 65535, max 0 on every axis), which the conservative test rejects for
// every ray inside its domain.  The accepted children are visited nearest first; the others go to the stack, farthest first.
template <class STK, class CNT>
GI_HD void wfInnerStep(WfLane& L, const DScene& s, const STK& stack, CNT& cnt)
{
	const U8 a = ld8(s.qNodes + 16 * size_t(L.node));
	const U8 b = ld8(s.qNodes + 16 * size_t(L.node) + 8);
	cnt.template add<WFC_NODES>(4u); cnt.template add<WFC_INNER_STEPS>(1u);
	float f0, f1, f2, f3v;
	if (!wfSlabTest(L, a.v[0], a.v[1], a.v[2], f0)) f0 = GI_INF;
	if (!wfSlabTest(L, a.v[3], a.v[4], a.v[5], f1)) f1 = GI_INF;
	if (!wfSlabTest(L, a.v[6], a.v[7], b.v[0], f2)) f2 = GI_INF;
	if (!wfSlabTest(L, b.v[1], b.v[2], b.v[3], f3v)) f3v = GI_INF;
	int c0 = int(b.v[4]), c1 = int(b.v[5]), c2 = int(b.v[6]), c3 = int(b.v[7]);
	// sorting network (0,1)(2,3)(0,2)(1,3)(1,2): ascending entry distance, rejected children (inf) last
#if WF_CSWAP_MINMAX
	// compare-exchange as min / max + two selects (5 ALU instructions); the branchy form compiles to a chain of predicated moves (~9)
#define WF_CSWAP(fa, ca, fb, cb) { const bool sw = fb < fa; const float lo = fminf(fa, fb), hi = fmaxf(fa, fb); const int cl = sw ? cb : ca, ch = sw ? ca : cb; fa = lo; fb = hi; ca = cl; cb = ch; }
#else
#define WF_CSWAP(fa, ca, fb, cb) { if (fb < fa) { const float tf = fa; fa = fb; fb = tf; const int tc = ca; ca = cb; cb = tc; } }
#endif
#if WF_SORT_EXCHANGES >= 5
	WF_CSWAP(f0, c0, f1, c1) WF_CSWAP(f2, c2, f3v, c3) WF_CSWAP(f0, c0, f2, c2) WF_CSWAP(f1, c1, f3v, c3) WF_CSWAP(f1, c1, f2, c2)
#undef WF_CSWAP
	if (f0 < GI_INF)
	{
#if WF_STEP_OVERFLOW_CHECK
		// one capacity check per step instead of one per push: a walk this deep (never seen on a real scene) is handed to the reference walk at once
		if (L.sp > GI_STACK_SIZE - 3) { L.flags |= WFL_OVERFLOW; L.node = WF_NODE_DONE; L.pend = 0; return; }
		if (f3v < GI_INF) wfPushUnchecked(L, stack, c3, f3v);
		if (f2 < GI_INF) wfPushUnchecked(L, stack, c2, f2);
		if (f1 < GI_INF) wfPushUnchecked(L, stack, c1, f1);
#else
		if (f3v < GI_INF) wfPush(L, stack, c3, f3v);
		if (f2 < GI_INF) wfPush(L, stack, c2, f2);
		if (f1 < GI_INF) wfPush(L, stack, c1, f1);
#endif
		L.node = c0;
	}
	else wfPop(L, stack);
#else
	// three exchanges: the nearest child is exact (slot 0), slot 2 holds the loser of the final and is pushed last (popped first);
	// the order of the stack is a heuristic only -- every entry carries its own entry distance for the cull at pop time
	WF_CSWAP(f0, c0, f1, c1) WF_CSWAP(f2, c2, f3v, c3) WF_CSWAP(f0, c0, f2, c2)
#undef WF_CSWAP
	if (f0 < GI_INF)
	{
		if (f3v < GI_INF) wfPush(L, stack, c3, f3v);
		if (f1 < GI_INF) wfPush(L, stack, c1, f1);
		if (f2 < GI_INF) wfPush(L, stack, c2, f2);
		L.node = c0;
	}
	else wfPop(L, stack);
#endif
	wfStash(L, stack);
}
#else
GI_HD bool wfSlabTestChild(const WfLane& L, const uint32_t* node, int child, float& nearLow) { return wfSlabTest(L, node[3 * child], node[3 * child + 1], node[3 * child + 2], nearLow); }

template <class STK, class CNT>
GI_HD void wfInnerStep(WfLane& L, const DScene& s, const STK& stack, CNT& cnt)
{
	// one 32-byte record, one 256-bit load: (c0.x, c0.y, c0.z, c1.x, c1.y, c1.z, child0, child1)
	const U8 n = ld8(s.qNodes + 8 * size_t(L.node));
	cnt.template add<WFC_NODES>(2u); cnt.template add<WFC_INNER_STEPS>(1u);
	float f0, f1;
	const bool h0 = wfSlabTest(L, n.v[0], n.v[1], n.v[2], f0);
	const bool h1 = wfSlabTest(L, n.v[3], n.v[4], n.v[5], f1);
	int c0 = int(n.v[6]), c1 = int(n.v[7]);
	if (h0 && h1)
	{
		if (f1 < f0) { int t = c0; c0 = c1; c1 = t; float tf = f0; f0 = f1; f1 = tf; }
		wfPush(L, stack, c1, f1);
		L.node = c0;
	}
	else if (h0) L.node = c0;
	else if (h1) L.node = c1;
	else wfPop(L, stack);
	wfStash(L, stack);
}

#endif

// what follows a tested leaf: the next stack entry, or (WF_POSTPONE_LEAF) the node the lane already stands at
template <class STK>
GI_HD void wfLeafDone(WfLane& L, const STK& stack)
{
#if WF_POSTPONE_LEAF
	wfStash(L, stack);
#else
	wfPop(L, stack);
#endif
}

template <class STK, class CNT>
GI_HD void wfLeafStep(WfLane& L, const DScene& s, const STK& stack, WfCand* cand, CNT& cnt)
{
#if WF_POSTPONE_LEAF
	const int leaf = ~L.pend;
	L.pend = 0;
#else
	const int leaf = ~L.node;
#endif
	// One 96-byte record: both triangles + the leaf node's own box in the .w components; all six loads are issued first.
	F4 v[6];
	ld8x3(s.leafPos + 6 * size_t(leaf), v);
	cnt.template add<WFC_LEAF_STEPS>(1u);
	// The reference visits this leaf iff its own node box passes the exact-zone test with first <= rayLength
	// (traceray.glsl:64-69); a triangle of an unvisited leaf does not count.  Most leaf visits hit no triangle at all, so the
	// box test (three IEEE divisions + slab test) is made lazily: at the first triangle hit, with the limit the lane had when it
	// reached the leaf (L.limit only changes after the gate).
	float leafFirst = 0.0f, leafLast;
#if WF_LAZY_LEAF_BOX
	bool gateKnown = false;
#else
	cnt.template add<WFC_NODES>(1u);
	if (!(intersectBox(L.o, wfInvDir(L), v[0].w, v[1].w, v[2].w, v[3].w, v[4].w, v[5].w, leafFirst, leafLast) && leafFirst <= L.limit)) { wfLeafDone(L, stack); return; }
#endif
GI_UNROLL
	for (int slot = 0; slot < 2; ++slot)
	{
		const int lt = 2 * leaf + slot;
		const F4 q0 = v[3 * slot], q1 = v[3 * slot + 1], q2 = v[3 * slot + 2];
		if (!(q0.x == q0.x)) break; // unused slot (traceray.glsl:127)
		cnt.template add<WFC_TRIS>(1u);
		float t; F3 nb, nn;
		if (intersectTriangle(L.o, L.d, f3(q0.x, q0.y, q0.z), f3(q1.x, q1.y, q1.z), f3(q2.x, q2.y, q2.z), t, nb, nn) && t < L.tMaxInit)
		{
#if WF_LAZY_LEAF_BOX
			if (!gateKnown)
			{
				gateKnown = true;
				cnt.template add<WFC_NODES>(1u);
				if (!(intersectBox(L.o, wfInvDir(L), v[0].w, v[1].w, v[2].w, v[3].w, v[4].w, v[5].w, leafFirst, leafLast) && leafFirst <= L.limit)) break;
			}
#endif
			if (L.flags & WFL_ANY) { L.flags |= WFL_OCCLUDED; L.node = WF_NODE_DONE; L.pend = 0; return; }
			if (L.flags & WFL_COLLECT)
			{
				// pass 2: L.best is the result of pass 1 and stays fixed
				const float window = L.best * GI_AMBIGUITY_REL + s.ambiguityAbs;
				if (t <= L.best + window)
				{
					if (t > L.best + 0.5f * window) L.flags |= WFL_GUARD;
					else if (L.numCand < WF_MAX_CAND)
					{
						cand[L.numCand].t = t; cand[L.numCand].first = leafFirst; cand[L.numCand].leaf = lt;
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
	wfLeafDone(L, stack);
}

// Called when a lane's walk ran out of nodes.  Returns true when the ray's result (L.best, L.hitLeaf / WFL_OCCLUDED) is
// final and equal to the reference's; false when the lane was restarted for the collecting second pass.
template <class CNT>
GI_HD bool wfEndOfWalk(WfLane& L, const DScene& s, WfCand* cand, CNT& cnt)
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
				cnt.template add<WFC_FALLBACK>(1u);
				if (!walk && s.leafPreorder)
				{
					// pass 2: same lane, culling limit fixed at the window of pass 1's result
					L.flags |= WFL_COLLECT;
					L.numCand = 0; wfStackReset(L);
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
	cnt.template add<WFC_NODES>(st.nodes); cnt.template add<WFC_TRIS>(st.tris);
	return done;
}

// One ray from start to end on one lane (what a warp does for 32 rays at once, interleaved with other rays).
template <class CNT>
GI_HD void wfWalkSingle(WfLane& L, const DScene& s, F3 o, F3 d, float tMax, bool isAny, CNT& cnt)
{
	StackEntry entries[GI_STACK_SIZE];
	WfStackArray stack; stack.e = entries;
	WfCand cand[WF_MAX_CAND];
	wfBegin(L, s, o, d, tMax, isAny, 0u, cnt);
	for (;;)
	{
		wfStash(L, stack);
		if (wfFinished(L)) { if (wfEndOfWalk(L, s, cand, cnt)) return; continue; }
#if WF_POSTPONE_LEAF
		// as late as possible: the stash is only tested when the lane is blocked at a second leaf or has run out of nodes
		if (L.node >= 0) wfInnerStep(L, s, stack, cnt);
		else wfLeafStep(L, s, stack, cand, cnt);
#else
		if (L.node < 0) wfLeafStep(L, s, stack, cand, cnt);
#if WF_ANY_UNSORTED && WF_CENTRE_EXTENT
		else if (L.flags & WFL_ANY) wfInnerStepAny(L, s, stack, cnt);
#endif
		else wfInnerStep(L, s, stack, cnt);
#endif
	}
}

} // namespace gik
