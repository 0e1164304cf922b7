// vrm_flat.cuh -- the traversal of vrm_core.cuh re-expressed as a per-ray STATE MACHINE whose states are code blocks
// that a warp can schedule.
//
// Why: the reference's control flow is four levels of nested data-dependent loops (scene walk > region march > voxel
// steps / cluster jumps, then the same again for the shadow ray).  Executed as written, a warp serialises lanes that are
// in different loops or have different trip counts: the first sm_100a capture showed 10.6 of 32 threads active per
// instruction for VCS + longest axis (profiles/r01a_ncu_render_vcs_longestaxis.json).  Here a ray is always in one of five
// states and each state's work is one block of straight-line code:
//     kStRegion  enter the region under the ray (table entry already read): null-region skip / load the region / leave
//     kStHead    longest axis only: loop condition of Renderer.cuh:787 + the order of this iteration's voxel tests
//     kStMain    [one advance: next voxel edge | cluster edge | cluster jump] + ONE voxel test (space check + lookup)
//     kStHit     shade the hit (applyLighting) and turn the lane into the hit's shadow ray
//     kStDone    pixel resolved (the lane can take a new pixel)
// A warp then runs, per iteration, only the block that most of its lanes are waiting for (vrm_render.cu,
// render_scheduled_kernel: __match_any_sync / __reduce_min_sync majority vote), so every executed instruction has many
// active lanes whatever the individual rays' phases are.  `step()` runs the blocks in program order for single-ray callers.
//
// ARITHMETIC IS UNCHANGED: the same operations in the same order as vrm_core.cuh / the reference -- only the order in
// which different rays' operations are interleaved changes.  tests/hostsim runs this very code on the CPU against the
// oracle (bit-exact RGB, hit maps and event counters).
#pragma once

#include "vrm_core.cuh"

namespace vrm
{

enum FlatState : int
{
	kStMain = 0,
	kStRegion = 1,
	kStHead = 2,
	kStHit = 3,
	kStDone = 4
};

// What the advance of a kStMain step moves to.  All are "t_i = (next_i - o_i) / dir_i, o += (min t + EPSILON) * dir".
enum AdvMode : int
{
	kAdvNone = 0,     // no advance: a longest-axis voxel test on gridValues
	kAdvNext = 1,     // next voxel edge +-EPSILON                      (Renderer.cuh:269-280,320-331)
	kAdvCluster = 2,  // cluster edge of the voxel under the ray        (Renderer.cuh:293-304)
	kAdvJump = 3      // one iteration of performVoxelSpaceJump's loop   (Renderer.cuh:707-721): cluster edge of gridValues, scaled direction
};

template <int ST, int ALGO, bool STATS>
struct FlatRay
{
	using P = typename std::conditional<ALGO == kAlgoOriginal, PermIdentity, PermRuntime>::type;
	static constexpr bool kLA = ALGO != kAlgoOriginal;

	// current ray (primary or shadow), region-local, walk space
	float o[3];  // original algorithm: ray origin; longest axis: oldRay origin (the reference copies between the two only
	             // at points where they are equal, Renderer.cuh:726,768,912)
	RayDir k;    // direction + exact-division constants (vrm_core.cuh)
	RayDir ko;   // the same for the longest-axis-scaled direction (Ray.cuh:69): per-ray constants, not per region
	int reg[3];
	int32_t ri;
	RegionRef<ST> r;
	P p;
	int st;
	bool shadow;    // this is the shadow ray of an already shaded hit
	bool shadowLA;  // ... walked with the longest-axis routines (Renderer.cuh:633-694) rather than the original ones (174-235)
	int mode;       // AdvMode of the next kStMain step
	// tX,tY,tZ,tMin of the last kAdvNext advance (the OUTER values of Renderer.cuh:273-277: a cluster skip leaves them
	// stale) or of the last kAdvJump advance (Renderer.cuh:713-716, tMin including +EPSILON); a ray is in one of the two
	// phases at a time and each phase sets them before it can hit, so one set of registers serves both
	float t0, t1, t2, tMin;
	// longest-axis state (slot 0 = longest axis)
	float ro[3];
	int g[3], ad[3];
	uint32_t seq;
	int nTests;
	bool roundDown;
	// A pending hit (kStHit) re-uses registers that are dead between the hit and the start of the shadow ray: its position
	// lives in ro[], its packed normal / shadow-routine bits in `mode` (bits 0-1 normal axis (world), bit 2 normal sign
	// negative, bit 3 longest-axis shadow routine) and its voxel colour in `result`.
	// `result` = voxel colour (kStHit) -> shaded colour waiting for its shadow ray -> final pixel colour (kStDone).
	uint32_t result;

	VRM_HD bool guardSkip() const { return shadow && !shadowLA; }  // zero-direction guards in the null-region skip (Renderer.cuh:191-193)
	VRM_HD bool guardAdv() const { return shadow; }                // ... and in shadowRayMarchVoxelGrid (Renderer.cuh:113-115)

	VRM_HD void finish(uint32_t colour)
	{
		result = colour;
		st = kStDone;
	}

	// rebase into the neighbouring region after the ray left the current one (Renderer.cuh:421-429) and read its entry
	VRM_HD void change_region(RayCtx<ST, STATS>& c)
	{
		rebase_region(o, reg);
		ri = position_sane(o) ? region_entry(c, p, reg) : -2;  // see position_sane (vrm_core.cuh)
		st = kStRegion;
	}

	// rayMarchVoxelScene / rayMarchVoxelSceneLongestAxis up to the first region (Renderer.cuh:338-378, 917-954)
	VRM_HD void start_primary(RayCtx<ST, STATS>& c, const float* originW, const float* dirW, float scale)
	{
		shadow = false; shadowLA = false; result = 0;
		if constexpr (kLA) p = rank_axes(dirW[0], dirW[1], dirW[2]);
		float sW[3] = {canonical_zero(vmul(scale, vsub(originW[0], c.translation[0]))), canonical_zero(vmul(scale, vsub(originW[1], c.translation[1]))),
		               canonical_zero(vmul(scale, vsub(originW[2], c.translation[2])))};
		float dw[3];
		to_walk(p, sW, o); to_walk(p, dirW, dw);
		k = make_raydir(dw[0], dw[1], dw[2]);
		if constexpr (kLA) ko = scaled_raydir(k);
		reg[0] = (int)floorf(vmul(o[0], 0.015625f)); reg[1] = (int)floorf(vmul(o[1], 0.015625f)); reg[2] = (int)floorf(vmul(o[2], 0.015625f));
		const int minC = c.sv.minCoord;
		const uint32_t D = c.sv.diameter;
		while (reg[0] - minC < 0 || reg[1] - minC < 0 || reg[2] - minC < 0 ||
		       (uint32_t)(reg[0] - minC) > D - 1 || (uint32_t)(reg[1] - minC) > D - 1 || (uint32_t)(reg[2] - minC) > D - 1)
		{
			int far = (int)(D + (uint32_t)minC);
			float a0 = vdiv(vsub((float)((k.d[0] < 0.0f ? far : minC) * kRegion), o[0]), k.d[0]);
			float a1 = vdiv(vsub((float)((k.d[1] < 0.0f ? far : minC) * kRegion), o[1]), k.d[1]);
			float a2 = vdiv(vsub((float)((k.d[2] < 0.0f ? far : minC) * kRegion), o[2]), k.d[2]);
			if (a0 <= 0.0f) a0 = INFINITY;
			if (a1 <= 0.0f) a1 = INFINITY;
			if (a2 <= 0.0f) a2 = INFINITY;
			float m = min3(a0, a1, a2);
			if (m == INFINITY || m != m) { finish(0); return; }  // (a NaN tMin only arises from 0/0: treated as a miss)
			float s = vadd(m, kEps);
			o[0] = along(o[0], s, k.d[0]); o[1] = along(o[1], s, k.d[1]); o[2] = along(o[2], s, k.d[2]);
			reg[0] = (int)floorf(vmul(o[0], 0.015625f)); reg[1] = (int)floorf(vmul(o[1], 0.015625f)); reg[2] = (int)floorf(vmul(o[2], 0.015625f));
		}
		for (int i = 0; i < 3; i++) o[i] = vmul(1.0f, vsub(o[i], (float)(reg[i] * kRegion)));
		ri = region_entry(c, p, reg);
		st = kStRegion;
	}

	// ---- kStHit: applyLighting(...) * !isInShadow...(Ray(hit, LIGHT_DIRECTION), currentRegion)  (Renderer.cuh:314-315,821-822,...)
	VRM_HD void record_hit(uint32_t col, const float* pos, int nAxisW, float nSign, bool laKind)
	{
		if (shadow) { finish(0); return; }  // any voxel on the shadow ray: colour * !inShadow = 0
		result = col;
		ro[0] = pos[0]; ro[1] = pos[1]; ro[2] = pos[2];
		mode = nAxisW | (nSign < 0.0f ? 4 : 0) | (laKind ? 8 : 0);
		st = kStHit;
	}

	VRM_HD void do_hit(RayCtx<ST, STATS>& c)
	{
		const int nAxisW = mode & 3;
		const float nSign = (mode & 4) ? -1.0f : 1.0f;
		const bool laKind = (mode & 8) != 0;
		float hitW[3];
		int regW[3];
		to_world(p, ro, hitW); to_world(p, reg, regW);
		result = apply_lighting(c.light, c.translation, result, nAxisW, nSign, hitW, regW);
		if (!c.light.useShadows) { finish(result); return; }
		shadow = true;
		shadowLA = laKind;
		if constexpr (kLA)
		{
			if (laKind) p = rank_axes(c.light.dir[0], c.light.dir[1], c.light.dir[2]);
			else { p.a0 = 0; p.a1 = 1; p.a2 = 2; }
		}
		float dw[3];
		to_walk(p, hitW, o); to_walk(p, c.light.dir, dw); to_walk(p, regW, reg);
		k = make_raydir(dw[0], dw[1], dw[2]);
		if constexpr (kLA) { if (laKind) ko = scaled_raydir(k); }
		ri = region_entry(c, p, reg);
		st = kStRegion;
	}

	// ---- kStRegion ---------------------------------------------------------------------------------------------
	VRM_HD void do_region(RayCtx<ST, STATS>& c)
	{
		if (ri == -2) { finish(shadow ? result : 0u); return; }  // left the scene: background / not shadowed
		if (ri == -1)
		{
			// null-region skip to the region edge, no +EPSILON (Renderer.cuh:384-410, guarded twin 185-211)
			float n0 = k.d[0] > 0.0f ? vadd((float)kRegion, kEps) : vsub(0.0f, kEps);
			float n1 = k.d[1] > 0.0f ? vadd((float)kRegion, kEps) : vsub(0.0f, kEps);
			float n2 = k.d[2] > 0.0f ? vadd((float)kRegion, kEps) : vsub(0.0f, kEps);
			float a0, a1, a2;
			div3(vsub(n0, o[0]), vsub(n1, o[1]), vsub(n2, o[2]), k, a0, a1, a2);
			if (guardSkip()) { a0 = (k.d[0] != 0.0f) ? a0 : INFINITY; a1 = (k.d[1] != 0.0f) ? a1 : INFINITY; a2 = (k.d[2] != 0.0f) ? a2 : INFINITY; }
			const float m = min3(a0, a1, a2);
			o[0] = along(o[0], m, k.d[0]); o[1] = along(o[1], m, k.d[1]); o[2] = along(o[2], m, k.d[2]);
			change_region(c);
			return;
		}
		r = load_region<ST>(c.sv, ri);
		bool la = false;
		if constexpr (kLA) la = !(shadow && !shadowLA);
		if (la)
		{
			// rayMarchVoxelGridLongestAxis prologue, Renderer.cuh:763-784
			g[0] = (int)o[0]; g[1] = (int)o[1]; g[2] = (int)o[2];
			ad[0] = k.d[0] < 0.0f ? -1 : 1;
			float t = ad[0] > 0 ? vdiv(vsub(vadd(vadd((float)g[0], kEps), 1.0f), o[0]), 1.0f)
			                    : vdiv(vsub(vsub((float)g[0], kEps), o[0]), -1.0f);
			ro[0] = along(o[0], t, ko.d[0]); ro[1] = along(o[1], t, ko.d[1]); ro[2] = along(o[2], t, ko.d[2]);
			ad[1] = (int)ro[1] - g[1];
			ad[2] = (int)ro[2] - g[2];
			roundDown = ko.d[1] < 0.0f;
			st = kStHead;
		}
		else { mode = kAdvNext; st = kStMain; }  // the region march starts with one step before the first test (Renderer.cuh:269-280)
	}

	// ---- kStHead (longest axis): Renderer.cuh:787-805 ------------------------------------------------------------------
	VRM_HD void do_head()
	{
		if (!grid_in_region(g[0] + ad[0], g[1] + ad[1], g[2] + ad[2]))
		{
			// Renderer.cuh:911-914: finish the region with the original algorithm from oldRay's origin (o already is it)
			mode = kAdvNext;
			st = kStMain;
			return;
		}
		if (ad[2] != 0 && ad[1] != 0)
		{
			float rounded = roundDown ? floorf(o[1]) : ceilf(o[1]);
			float tt = div1(vsub(rounded, o[1]), ko, 1);
			float shortestPosition = vadd(o[2], vmul(ko.d[2], tt));
			int shorterDiff = (int)floorf(shortestPosition) - g[2];
			seq = shorterDiff != 0 ? (2u | (1u << 2)) : (1u | (2u << 2));
			nTests = 3;
		}
		else if (ad[1] != 0) { seq = 1u; nTests = 2; }
		else if (ad[2] != 0) { seq = 2u; nTests = 2; }
		else { seq = 0u; nTests = 1; }
		mode = kAdvNone;
		st = kStMain;
	}

	// ---- kStMain: [one advance] + one voxel test ------------------------------------------------------------------------
	VRM_HD void do_main(RayCtx<ST, STATS>& c)
	{
		int c0, c1, c2, slot = 0;
		const bool test = kLA && mode == kAdvNone;
		const bool jump = kLA && mode == kAdvJump;
		if (!test)
		{
			// zero-direction guards exist only in the shadow routines of the ORIGINAL algorithm (Renderer.cuh:113-115,137-139,
			// 160-162); the longest-axis jump guards nothing (Renderer.cuh:457-459)
			const bool gd = jump ? false : guardAdv();
			RayDir e = k;
			if constexpr (kLA) { if (jump) e = ko; }
			const float e0 = e.d[0], e1 = e.d[1], e2 = e.d[2];
			float n0, n1, n2;
			if (mode == kAdvNext)
			{
				n0 = next_edge(e0, o[0]); n1 = next_edge(e1, o[1]); n2 = next_edge(e2, o[2]);
			}
			else
			{
				// cluster edge of the voxel of the failed test: (int)o for the original algorithm, gridValues in a jump
				n0 = (float)cluster_edge(e0, jump ? g[0] : (int)o[0]);
				n1 = (float)cluster_edge(e1, jump ? g[1] : (int)o[1]);
				n2 = (float)cluster_edge(e2, jump ? g[2] : (int)o[2]);
			}
			float a0, a1, a2;
			div3(vsub(n0, o[0]), vsub(n1, o[1]), vsub(n2, o[2]), e, a0, a1, a2);
			if (gd) { a0 = (e0 != 0.0f) ? a0 : INFINITY; a1 = (e1 != 0.0f) ? a1 : INFINITY; a2 = (e2 != 0.0f) ? a2 : INFINITY; }
			float m = min3(a0, a1, a2);
			if (m == 0.0f && mode != kAdvNext)
			{
				// the ray sits on a cluster face it cannot leave: fast-forward the EPSILON crawl (crawl_skip, vrm_core.cuh)
				const int skipped = crawl_skip(o, e, jump ? g[0] : (int)o[0], jump ? g[1] : (int)o[1], jump ? g[2] : (int)o[2]);
				if (skipped > 0)
				{
					if (STATS) { c.st.nExist += skipped; c.st.nExistFalse += skipped; c.st.nCrawlSkipped += skipped; }
					div3(vsub(n0, o[0]), vsub(n1, o[1]), vsub(n2, o[2]), e, a0, a1, a2);  // same cell, same edges; guards are moot on the fast path
					m = min3(a0, a1, a2);
				}
			}
			const float s = vadd(m, kEps);
			if (mode == kAdvNext) { t0 = a0; t1 = a1; t2 = a2; tMin = m; }
			if (jump) { t0 = a0; t1 = a1; t2 = a2; tMin = s; }
			o[0] = along(o[0], s, e0); o[1] = along(o[1], s, e1); o[2] = along(o[2], s, e2);
			// grid_in_region((int)floorf(o)) of the jump (Renderer.cuh:719-723) and isRayInRegion(o) agree for every o
			if (!ray_in_region(o)) { change_region(c); return; }
			c0 = (int)o[0]; c1 = (int)o[1]; c2 = (int)o[2];  // == (int)floorf(o) inside a region
			if (jump) { g[0] = c0; g[1] = c1; g[2] = c2; }
		}
		else
		{
			slot = (int)(seq & 3u);
			seq >>= 2;
			g[0] += slot == 0 ? ad[0] : 0;
			g[1] += slot == 1 ? ad[1] : 0;
			g[2] += slot == 2 ? ad[2] : 0;
			c0 = g[0]; c1 = g[1]; c2 = g[2];
		}

		// the voxel test: doesVoxelSpaceExist + lookupVoxel
		const bool e = space_exists(c, r, p, c0, c1, c2);
		uint32_t col = kEmpty;
		if (e) col = lookup_voxel(c, r, p, reg, c0, c1, c2);

		if (col != kEmpty)
		{
			if (!test)
			{
				// original algorithm: Renderer.cuh:312-315; jump: Renderer.cuh:733-738 (tMin carries +EPSILON there, so the
				// comparison normally falls through to the Z normal)
				int nAxisW = normal_axis_from_t(p, t0, t1, t2, tMin);
				float dn = p.axis(0) == nAxisW ? k.d[0] : (p.axis(1) == nAxisW ? k.d[1] : k.d[2]);  // scaled direction = s * d, s > 0: same sign
				record_hit(col, o, nAxisW, copysignf(1.0f, -dn), jump);
			}
			else if constexpr (kLA)
			{
				float odS = pick3(slot, ko.d[0], ko.d[1], ko.d[2]);
				float pos[3];
				if (slot == 0) { pos[0] = ro[0]; pos[1] = ro[1]; pos[2] = ro[2]; }  // Renderer.cuh:899
				else
				{
					// getLocalHitLocation, Renderer.cuh:753-758
					float ooS = pick3(slot, o[0], o[1], o[2]);
					float tl = odS > 0.0f ? vdiv(vsub(ceilf(ooS), ooS), odS) : vdiv(vsub(floorf(ooS), ooS), odS);
					pos[0] = along(o[0], tl, ko.d[0]); pos[1] = along(o[1], tl, ko.d[1]); pos[2] = along(o[2], tl, ko.d[2]);
				}
				record_hit(col, pos, p.axis(slot), copysignf(1.0f, -odS), true);
			}
			return;
		}

		// no voxel here: decide the next micro-step
		if (!test)
		{
			if (!jump) mode = e ? kAdvNext : kAdvCluster;
			else if (e)
			{
				if constexpr (kLA)
				{
					// the jump reached a cluster that exists but the voxel is empty: re-snap to the longest axis and `continue`
					// the while loop (Renderer.cuh:742-750)
					float tNext = div1(vsub(ko.d[0] > 0.0f ? ceilf(o[0]) : floorf(o[0]), o[0]), ko, 0);
					float tt = vadd(tNext, kEps);
					ro[0] = along(o[0], tt, ko.d[0]); ro[1] = along(o[1], tt, ko.d[1]); ro[2] = along(o[2], tt, ko.d[2]);
					ad[1] = (int)ro[1] - g[1];
					ad[2] = (int)ro[2] - g[2];
					st = kStHead;
				}
			}
			// else: still no voxel space: another jump iteration (mode stays kAdvJump)
			return;
		}
		if constexpr (kLA)
		{
			if (!e)
			{
				// performVoxelSpaceJump (Renderer.cuh:808-810 -> 696): its while condition repeats the exist check of the failed
				// test on the same voxel -- same answer, only the counter sees it
				if (STATS) { c.st.nExist++; c.st.nExistFalse++; }
				mode = kAdvJump;
			}
			else if (--nTests == 0)
			{
				// Renderer.cuh:903-908
				o[0] = ro[0]; o[1] = ro[1]; o[2] = ro[2];
				ro[0] = vadd(ro[0], ko.d[0]); ro[1] = vadd(ro[1], ko.d[1]); ro[2] = vadd(ro[2], ko.d[2]);
				ad[1] = (int)ro[1] - g[1];
				ad[2] = (int)ro[2] - g[2];
				st = kStHead;
			}
		}
	}

	// Run the block of the current state.  Returns true when the pixel is resolved.
	VRM_HD bool step(RayCtx<ST, STATS>& c)
	{
		if (st == kStRegion) do_region(c);
		else if (st == kStHit) do_hit(c);
		else
		{
			if constexpr (kLA) { if (st == kStHead) do_head(); }
			if (st == kStMain) do_main(c);
		}
		return st == kStDone;
	}
};

// Convenience for single-ray callers (trace kernels, host sim): run the state machine to completion.
template <int ST, int ALGO, bool STATS>
VRM_HD uint32_t march_scene_flat(RayCtx<ST, STATS>& c, const float* originW, const float* dirW, float scale)
{
	FlatRay<ST, ALGO, STATS> ray;
	ray.start_primary(c, originW, dirW, scale);
	while (ray.st != kStDone) ray.step(c);
	return ray.result;
}

}  // namespace vrm
