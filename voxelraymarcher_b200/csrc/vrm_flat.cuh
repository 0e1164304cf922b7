// vrm_flat.cuh -- the traversal of vrm_core.cuh re-expressed as ONE flat loop of micro-steps (a per-ray state machine).
//
// Why: the reference's control flow is four levels of nested data-dependent loops (scene walk > region march > voxel
// steps / cluster jumps, then the same again for the shadow ray).  Executed as written, a warp serialises lanes that are
// in different loops: the first sm_100a capture showed 10.6 of 32 threads active per instruction for VCS + longest axis
// (profiles/r01a_ncu_render_vcs_longestaxis.json).  Here every lane performs exactly one "voxel test" per iteration of
// a single warp-uniform loop, whatever phase it is in (primary or shadow ray, original steps, longest-axis tests,
// cluster jumps); region changes are the only other state.  A finished lane can then be refilled with a new pixel
// (persistent-thread ray queue, vrm_render.cu).
//
// ARITHMETIC IS UNCHANGED: the same operations in the same order as vrm_core.cuh / the reference -- only the order in
// which different rays' operations are interleaved changes.  tests/hostsim runs this very stepper on the CPU against the
// oracle (bit-exact RGB, hit maps and event counters).
#pragma once

#include "vrm_core.cuh"

namespace vrm
{

enum FlatState : int
{
	kStRegion = 0,  // (re)entering region `reg`: `ri` holds its table entry; null regions are skipped one per step
	kStAdv = 1,     // "original" stepping (also the longest-axis tail): advance, then test the voxel under the ray
	kStHead = 2,    // longest axis: head of the while loop (Renderer.cuh:787): loop condition + order of this iteration's tests
	kStTest = 3,    // longest axis: next pending voxel test of the iteration
	kStJump = 4,    // longest axis: inside performVoxelSpaceJump's while loop
	kStDone = 5
};

template <int ST, int ALGO, bool STATS>
struct FlatRay
{
	using P = typename std::conditional<ALGO == kAlgoOriginal, PermIdentity, PermRuntime>::type;

	// current ray (primary or shadow), region-local, walk space
	float o[3];  // original algorithm: ray origin; longest axis: oldRay origin (the reference copies between the two only
	             // at points where they are equal, Renderer.cuh:726,768,912)
	float d[3];
	int reg[3];
	int32_t ri;
	RegionRef<ST> r;
	P p;
	int st;
	bool shadow;    // this is the shadow ray of an already shaded hit
	bool shadowLA;  // ... walked with the longest-axis routines (Renderer.cuh:633-694) rather than the original ones (174-235)
	// original-algorithm state
	float t0, t1, t2, tMin;  // OUTER tX,tY,tZ,tMin of Renderer.cuh:273-277 (stale after a cluster skip)
	bool modeNext;           // next advance goes to the next voxel edge (true) or to the cluster edge (false)
	// longest-axis state (slot 0 = longest axis)
	float od[3], ro[3];
	int g[3], ad[3];
	uint32_t seq;
	int nTests;
	bool roundDown;
	float j0, j1, j2, jMin;  // tX,tY,tZ,tMin of the last cluster jump (Renderer.cuh:701-716)
	// result
	uint32_t lit;     // shaded colour waiting for its shadow ray
	uint32_t result;  // final pixel colour once st == kStDone

	VRM_HD bool guardSkip() const { return shadow && !shadowLA; }  // zero-direction guards in the null-region skip (Renderer.cuh:191-193)
	VRM_HD bool guardAdv() const { return shadow; }                // ... and in shadowRayMarchVoxelGrid (Renderer.cuh:113-115)

	VRM_HD float tdiv(float next, float oi, float di, bool guard) const
	{
		float t = vdiv(vsub(next, oi), di);
		if (guard) t = (di != 0.0f) ? t : INFINITY;
		return t;
	}

	VRM_HD void finish(uint32_t colour)
	{
		result = colour;
		st = kStDone;
	}

	// rebase into the neighbouring region after the ray left the current one (Renderer.cuh:421-429) and read its entry
	VRM_HD void change_region(RayCtx<ST, STATS>& c)
	{
		rebase_region(o, reg);
		// float -> int of a NaN is INT_MIN on the reference's host build, i.e. "outside the scene"; CUDA would give 0 and spin
		bool sane = o[0] == o[0] && o[1] == o[1] && o[2] == o[2];
		ri = sane ? region_entry(c, p, reg) : -2;
		st = kStRegion;
	}

	// rayMarchVoxelScene / rayMarchVoxelSceneLongestAxis up to the first region (Renderer.cuh:338-378, 917-954)
	VRM_HD void start_primary(RayCtx<ST, STATS>& c, const float* originW, const float* dirW, float scale)
	{
		shadow = false; shadowLA = false; lit = 0; result = 0;
		if constexpr (ALGO != kAlgoOriginal) p = rank_axes(dirW[0], dirW[1], dirW[2]);
		float sW[3] = {vmul(scale, vsub(originW[0], c.translation[0])), vmul(scale, vsub(originW[1], c.translation[1])), vmul(scale, vsub(originW[2], c.translation[2]))};
		to_walk(p, sW, o); to_walk(p, dirW, d);
		reg[0] = (int)floorf(vdiv(o[0], (float)kRegion)); reg[1] = (int)floorf(vdiv(o[1], (float)kRegion)); reg[2] = (int)floorf(vdiv(o[2], (float)kRegion));
		const int minC = c.sv.minCoord;
		const uint32_t D = c.sv.diameter;
		while (reg[0] - minC < 0 || reg[1] - minC < 0 || reg[2] - minC < 0 ||
		       (uint32_t)(reg[0] - minC) > D - 1 || (uint32_t)(reg[1] - minC) > D - 1 || (uint32_t)(reg[2] - minC) > D - 1)
		{
			int far = (int)(D + (uint32_t)minC);
			float a0 = vdiv(vsub((float)((d[0] < 0.0f ? far : minC) * kRegion), o[0]), d[0]);
			float a1 = vdiv(vsub((float)((d[1] < 0.0f ? far : minC) * kRegion), o[1]), d[1]);
			float a2 = vdiv(vsub((float)((d[2] < 0.0f ? far : minC) * kRegion), o[2]), d[2]);
			if (a0 <= 0.0f) a0 = INFINITY;
			if (a1 <= 0.0f) a1 = INFINITY;
			if (a2 <= 0.0f) a2 = INFINITY;
			float m = min3(a0, a1, a2);
			if (m == INFINITY || m != m) { finish(0); return; }  // (a NaN tMin only arises from 0/0: treated as a miss)
			float s = vadd(m, kEps);
			o[0] = along(o[0], s, d[0]); o[1] = along(o[1], s, d[1]); o[2] = along(o[2], s, d[2]);
			reg[0] = (int)floorf(vdiv(o[0], (float)kRegion)); reg[1] = (int)floorf(vdiv(o[1], (float)kRegion)); reg[2] = (int)floorf(vdiv(o[2], (float)kRegion));
		}
		for (int i = 0; i < 3; i++) o[i] = vmul(1.0f, vsub(o[i], (float)(reg[i] * kRegion)));
		ri = region_entry(c, p, reg);
		st = kStRegion;
	}

	// A hit of the PRIMARY ray: shade it (applyLighting) and turn this lane into the hit's shadow ray.
	// pos = hit position in walk space; laKind = which isInShadow* routine the reference calls at this hit site.
	VRM_HD void on_primary_hit(RayCtx<ST, STATS>& c, uint32_t col, const float* pos, int nAxisW, float nSign, bool laKind)
	{
		float hitW[3];
		int regW[3];
		to_world(p, pos, hitW); to_world(p, reg, regW);
		lit = apply_lighting(c.light, c.translation, col, nAxisW, nSign, hitW, regW);
		if (!c.light.useShadows) { finish(lit); return; }
		shadow = true;
		shadowLA = laKind;
		if constexpr (ALGO != kAlgoOriginal)
		{
			if (laKind) p = rank_axes(c.light.dir[0], c.light.dir[1], c.light.dir[2]);
			else { p.a0 = 0; p.a1 = 1; p.a2 = 2; }
		}
		to_walk(p, hitW, o); to_walk(p, c.light.dir, d); to_walk(p, regW, reg);
		ri = region_entry(c, p, reg);
		st = kStRegion;
	}

	VRM_HD void on_hit(RayCtx<ST, STATS>& c, uint32_t col, const float* pos, int nAxisW, float nSign, bool laKind)
	{
		if (shadow) finish(0);  // colour * !inShadow
		else on_primary_hit(c, col, pos, nAxisW, nSign, laKind);
	}

	// rayMarchVoxelGridLongestAxis prologue, Renderer.cuh:763-784
	VRM_HD void la_setup()
	{
		float k = vdiv(1.0f, fabsf(d[0]));
		od[0] = vmul(k, d[0]); od[1] = vmul(k, d[1]); od[2] = vmul(k, d[2]);
		g[0] = (int)o[0]; g[1] = (int)o[1]; g[2] = (int)o[2];
		ad[0] = d[0] < 0.0f ? -1 : 1;
		float t = ad[0] > 0 ? vdiv(vsub(vadd(vadd((float)g[0], kEps), 1.0f), o[0]), 1.0f)
		                    : vdiv(vsub(vsub((float)g[0], kEps), o[0]), -1.0f);
		ro[0] = along(o[0], t, od[0]); ro[1] = along(o[1], t, od[1]); ro[2] = along(o[2], t, od[2]);
		ad[1] = (int)ro[1] - g[1];
		ad[2] = (int)ro[2] - g[2];
		roundDown = od[1] < 0.0f;
		j0 = j1 = j2 = jMin = 0.0f;
	}

	// One micro-step.  Returns true when the pixel is resolved (st == kStDone, `result` valid).
	VRM_HD bool step(RayCtx<ST, STATS>& c)
	{
		if (st == kStRegion)
		{
			if (ri == -2) { finish(shadow ? lit : 0u); return true; }  // left the scene: background / not shadowed
			if (ri == -1)
			{
				// null-region skip, Renderer.cuh:384-410 (guarded twin 185-211)
				const bool gd = guardSkip();
				float n0 = d[0] > 0.0f ? vadd((float)kRegion, kEps) : vsub(0.0f, kEps);
				float n1 = d[1] > 0.0f ? vadd((float)kRegion, kEps) : vsub(0.0f, kEps);
				float n2 = d[2] > 0.0f ? vadd((float)kRegion, kEps) : vsub(0.0f, kEps);
				float m = min3(tdiv(n0, o[0], d[0], gd), tdiv(n1, o[1], d[1], gd), tdiv(n2, o[2], d[2], gd));
				o[0] = along(o[0], m, d[0]); o[1] = along(o[1], m, d[1]); o[2] = along(o[2], m, d[2]);
				change_region(c);
				return false;
			}
			r = load_region<ST>(c.sv, ri);
			bool la = false;
			if constexpr (ALGO != kAlgoOriginal) la = !(shadow && !shadowLA);
			if (la) { la_setup(); st = kStHead; }
			else { modeNext = true; st = kStAdv; }  // the region march starts with one step before the first test (Renderer.cuh:269-280)
		}

		if constexpr (ALGO != kAlgoOriginal)
		{
			if (st == kStHead)
			{
				if (!grid_in_region(g[0] + ad[0], g[1] + ad[1], g[2] + ad[2]))
				{
					// Renderer.cuh:911-914: finish the region with the original algorithm from oldRay's origin (o already is it)
					modeNext = true;
					st = kStAdv;
				}
				else
				{
					if (ad[2] != 0 && ad[1] != 0)  // Renderer.cuh:792-805
					{
						float rounded = roundDown ? floorf(o[1]) : ceilf(o[1]);
						float tt = vdiv(vsub(rounded, o[1]), od[1]);
						float shortestPosition = vadd(o[2], vmul(od[2], tt));
						int shorterDiff = (int)floorf(shortestPosition) - g[2];
						seq = shorterDiff != 0 ? (2u | (1u << 2)) : (1u | (2u << 2));
						nTests = 3;
					}
					else if (ad[1] != 0) { seq = 1u; nTests = 2; }
					else if (ad[2] != 0) { seq = 2u; nTests = 2; }
					else { seq = 0u; nTests = 1; }
					st = kStTest;
				}
			}
		}

		// ---- position of this step's voxel test -----------------------------------------------------------------
		int c0, c1, c2, slot = 0;
		if (st == kStAdv)
		{
			const bool gd = guardAdv();
			float a0, a1, a2;
			if (modeNext)
			{
				a0 = tdiv(next_edge(d[0], o[0]), o[0], d[0], gd);
				a1 = tdiv(next_edge(d[1], o[1]), o[1], d[1], gd);
				a2 = tdiv(next_edge(d[2], o[2]), o[2], d[2], gd);
			}
			else
			{
				// cluster skip, Renderer.cuh:293-304; the voxel of the failed test is still (int)o
				a0 = tdiv((float)cluster_edge(d[0], (int)o[0]), o[0], d[0], gd);
				a1 = tdiv((float)cluster_edge(d[1], (int)o[1]), o[1], d[1], gd);
				a2 = tdiv((float)cluster_edge(d[2], (int)o[2]), o[2], d[2], gd);
			}
			float m = min3(a0, a1, a2);
			if (modeNext) { t0 = a0; t1 = a1; t2 = a2; tMin = m; }  // the skip's t values shadow the outer ones (Renderer.cuh:297-301)
			float s = vadd(m, kEps);
			o[0] = along(o[0], s, d[0]); o[1] = along(o[1], s, d[1]); o[2] = along(o[2], s, d[2]);
			if (!ray_in_region(o)) { change_region(c); return false; }
			c0 = (int)o[0]; c1 = (int)o[1]; c2 = (int)o[2];
		}
		else
		{
			if (st == kStTest)
			{
				slot = (int)(seq & 3u);
				seq >>= 2;
				g[0] += slot == 0 ? ad[0] : 0;
				g[1] += slot == 1 ? ad[1] : 0;
				g[2] += slot == 2 ? ad[2] : 0;
			}
			c0 = g[0]; c1 = g[1]; c2 = g[2];
		}

		// ---- the voxel test: doesVoxelSpaceExist + lookupVoxel ------------------------------------------------------
		const bool e = space_exists(c, r, p, c0, c1, c2);
		uint32_t col = kEmpty;
		if (e) col = lookup_voxel(c, r, p, reg, c0, c1, c2);

		if (st == kStAdv)
		{
			if (col != kEmpty)
			{
				int nAxisW = normal_axis_from_t(p, t0, t1, t2, tMin);  // Renderer.cuh:312
				float dn = p.axis(0) == nAxisW ? d[0] : (p.axis(1) == nAxisW ? d[1] : d[2]);
				on_hit(c, col, o, nAxisW, copysignf(1.0f, -dn), false);
				return st == kStDone;
			}
			modeNext = e;
			return false;
		}

		if constexpr (ALGO != kAlgoOriginal)
		{
			if (col != kEmpty)
			{
				if (st == kStJump)
				{
					// Renderer.cuh:733-738: tMin carries +EPSILON, so this normally falls through to the Z normal
					int nAxisW = normal_axis_from_t(p, j0, j1, j2, jMin);
					float dn = p.axis(0) == nAxisW ? od[0] : (p.axis(1) == nAxisW ? od[1] : od[2]);
					on_hit(c, col, o, nAxisW, copysignf(1.0f, -dn), true);
				}
				else
				{
					float odS = pick3(slot, od[0], od[1], od[2]);
					float pos[3];
					if (slot == 0) { pos[0] = ro[0]; pos[1] = ro[1]; pos[2] = ro[2]; }  // Renderer.cuh:899
					else
					{
						// getLocalHitLocation, Renderer.cuh:753-758
						float ooS = pick3(slot, o[0], o[1], o[2]);
						float tl = odS > 0.0f ? vdiv(vsub(ceilf(ooS), ooS), odS) : vdiv(vsub(floorf(ooS), ooS), odS);
						pos[0] = along(o[0], tl, od[0]); pos[1] = along(o[1], tl, od[1]); pos[2] = along(o[2], tl, od[2]);
					}
					on_hit(c, col, pos, p.axis(slot), copysignf(1.0f, -odS), true);
				}
				return st == kStDone;
			}
			if (!e)
			{
				// Entering performVoxelSpaceJump from a failed test repeats the exist check in its while condition
				// (Renderer.cuh:808-810 then 705): same voxel, same answer -- only the counter sees it.
				if (STATS && st == kStTest) { c.st.nExist++; c.st.nExistFalse++; }
				// one iteration of the jump loop, Renderer.cuh:707-728
				j0 = vdiv(vsub((float)cluster_edge(od[0], g[0]), o[0]), od[0]);
				j1 = vdiv(vsub((float)cluster_edge(od[1], g[1]), o[1]), od[1]);
				j2 = vdiv(vsub((float)cluster_edge(od[2], g[2]), o[2]), od[2]);
				jMin = vadd(min3(j0, j1, j2), kEps);
				o[0] = along(o[0], jMin, od[0]); o[1] = along(o[1], jMin, od[1]); o[2] = along(o[2], jMin, od[2]);
				g[0] = (int)floorf(o[0]); g[1] = (int)floorf(o[1]); g[2] = (int)floorf(o[2]);
				if (!grid_in_region(g[0], g[1], g[2])) { change_region(c); return false; }  // Renderer.cuh:723-728
				st = kStJump;
				return false;
			}
			// the voxel space exists but holds no voxel here
			if (st == kStJump)
			{
				// re-snap to the longest axis and `continue` the while loop, Renderer.cuh:742-750
				float tNext = od[0] > 0.0f ? vdiv(vsub(ceilf(o[0]), o[0]), od[0]) : vdiv(vsub(floorf(o[0]), o[0]), od[0]);
				float tt = vadd(tNext, kEps);
				ro[0] = along(o[0], tt, od[0]); ro[1] = along(o[1], tt, od[1]); ro[2] = along(o[2], tt, od[2]);
				ad[1] = (int)ro[1] - g[1];
				ad[2] = (int)ro[2] - g[2];
				st = kStHead;
				return false;
			}
			// kStTest
			if (--nTests == 0)
			{
				// Renderer.cuh:903-908
				o[0] = ro[0]; o[1] = ro[1]; o[2] = ro[2];
				ro[0] = vadd(ro[0], od[0]); ro[1] = vadd(ro[1], od[1]); ro[2] = vadd(ro[2], od[2]);
				ad[1] = (int)ro[1] - g[1];
				ad[2] = (int)ro[2] - g[2];
				st = kStHead;
			}
		}
		return false;
	}
};

// Convenience for single-ray callers (trace kernels, host sim): run the stepper to completion.
template <int ST, int ALGO, bool STATS>
VRM_HD uint32_t march_scene_flat(RayCtx<ST, STATS>& c, const float* originW, const float* dirW, float scale)
{
	FlatRay<ST, ALGO, STATS> ray;
	ray.start_primary(c, originW, dirW, scale);
	while (ray.st != kStDone) ray.step(c);
	return ray.result;
}

}  // namespace vrm
