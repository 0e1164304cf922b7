// vrm_lean.cuh -- the state machine of vrm_flat.cuh cut down to what the HOT kernels execute (the headline
// kernel was issue-bound at ~3400 thread-instructions per ray, a third of them control flow and operand selection).
//
// Same rays, same IEEE operations in the same order as vrm_flat.cuh / vrm_core.cuh / the reference; what changes is where the
// per-ray constants live and which code the hot loop carries:
//  * PER-RAY CONSTANTS IN SHARED MEMORY.  The direction, its reciprocal, the longest-axis-scaled direction and its reciprocal, the
//    division threshold and the walk-slot -> storage-code multipliers / masks / region-table strides never change along a ray
//    (they change once, when the lane turns into its hit's shadow ray).  They sit in six 16-byte vectors per thread in shared
//    memory ([vector][thread]: conflict-free LDS.128), so an advance fetches its operand set with two loads at a mode-dependent
//    offset instead of selecting six registers with twelve instructions, and ~20 registers are free for the walk itself.
//  * NO SLOW PATHS IN THE HOT LOOP.  IEEE division for unsafe operands, the crawl fast-forward and the region-face ping-pong
//    probe all start from one condition: min |numerator| < threshold in the advance (an exactly zero numerator -- the crawl --
//    included).  A lane that meets it PARKS: it flags its ray in a bitmap and stops; resume_kernel (vrm_render.cu), which owns all
//    the slow code, re-traces the flagged rays from their start with the generic machine.  A handful of rays per frame, each worth
//    microseconds; a bitmap has no capacity to overflow.
//  * ONE PROGRAM COUNTER (pc) instead of state + advance mode; zero-direction guards dropped (a zero component makes the
//    threshold NaN, so such a ray parks at its first advance); hit normals of cluster jumps computed at the hit, not per jump.
//
// tests/hostsim runs this source on the CPU (mode "lean") against the oracle: RGB, hit voxels and event counters.
// Reference file:line citations are relative to /root/reference/VoxelRaymarcher/src.
#pragma once

#include "vrm_flat.cuh"

namespace vrm
{

struct
#if defined(__CUDACC__)
__align__(16)
#else
alignas(16)
#endif
Vec4
{
	float x, y, z, w;
};

VRM_HD uint32_t fbits(float f) { return float_bits(f); }
VRM_HD float bitsf(uint32_t u) { return bits_float(u); }

// (int)floorf(x): one F2I.FLOOR on the device (same value for every x, NaN and out-of-range included: both forms saturate)
VRM_HD int floor_to_int(float x)
{
#if defined(__CUDA_ARCH__)
	return __float2int_rd(x);
#else
	return (int)floorf(x);
#endif
}

// indices of the constant vectors (element j of a lane's set is kc[j * STRIDE])
constexpr int kKcD = 0;     // d0 d1 d2 thr
constexpr int kKcRD = 1;    // rd0 rd1 rd2 | rs2 (uint bits)
constexpr int kKcSD = 2;    // sd0 sd1 sd2 thr      (longest axis: the scaled direction, Ray.cuh:69)
constexpr int kKcSRD = 3;   // srd0 srd1 srd2 | -
constexpr int kKcMul = 4;   // storage-code multipliers mul0 mul1 mul2 | rs0   (uint bits)
constexpr int kKcMask = 5;  // storage-code masks mask0 mask1 mask2 | rs1      (uint bits; VCS only)
constexpr int kKcVectors = 6;

enum LeanPc : int
{
	kPcTest = 0,        // longest axis: bump gridValues along the next slot of `seq`, test the voxel  (Renderer.cuh:792-901)
	kPcAdvNext = 1,     // advance to the next voxel edge +-EPSILON, test                               (Renderer.cuh:269-280,320-331)
	kPcAdvCluster = 2,  // advance to the cluster edge of the voxel under the ray, test                 (Renderer.cuh:293-304)
	kPcAdvJump = 3,     // one iteration of performVoxelSpaceJump: cluster edge of gridValues, scaled direction (Renderer.cuh:707-721)
	kPcAdvRegion = 4,   // null region: advance to the region edge, no +EPSILON                         (Renderer.cuh:384-410)
	kPcRegion = 5,      // region entry (table entry already read)
	kPcHead = 6,        // longest axis loop head (Renderer.cuh:787-805)
	kPcHit = 7,         // waiting to be shaded
	kPcDone = 8,
	kPcParkBit = 16     // OR-ed onto an advance pc: the step needs a slow path; the ray is re-traced by the generic machine (resume kernel)
};

constexpr uint32_t kFlNeg0 = 32u;  // longest-axis walk: direction component of slot 0 is negative (axisDiff[L] = -1)

template <int ST, int ALGO, bool STATS, int STRIDE>
struct LeanRay
{
	static constexpr bool kLA = ALGO != kAlgoOriginal;
	using Generic = FlatRay<ST, ALGO, STATS>;

	float o[3];        // as FlatRay::o
	float ro[3];       // as FlatRay::ro (longest axis `ray` origin; pending hit: its position)
	int g[3];
	int ad1, ad2;
	uint32_t seq;
	int pc;
	uint32_t fl;       // kFl* bits of vrm_flat.cuh (shadow flags, equality bits, permutation) + kFlNeg0
	uint32_t ur[3];
	int32_t ri;
	RegionRef<ST> r;   // hash table only (the VCS addresses everything from ri)
	uint32_t result;
	int hitMode;       // pending hit: packed normal / shadow-routine bits (FlatRay keeps them in `mode`)

	VRM_HD bool shadow() const { return (fl & kFlShadow) != 0; }
	VRM_HD bool shadowOriginal() const { return (fl & (kFlShadow | kFlShadowLA)) == kFlShadow; }
	VRM_HD void finish(uint32_t colour) { result = colour; pc = kPcDone; }

	// ---- constants --------------------------------------------------------------------------------------------------------
	static VRM_HD Vec4 ld(const Vec4* kc, int j) { return kc[j * STRIDE]; }
	static VRM_HD void st4(Vec4* kc, int j, float x, float y, float z, float w) { Vec4 v; v.x = x; v.y = y; v.z = z; v.w = w; kc[j * STRIDE] = v; }

	// direction constants + permutation-dependent storage constants of the ray that starts now
	VRM_HD void set_constants(const SceneView& sv, Vec4* kc, const PermRuntime& p, const RayDir& k, const RayDir& ko, float thr)
	{
		fl = (fl & ~((63u << kFlPermShift) | kFlNeg0)) | pack_perm(p) | (k.d[0] < 0.0f ? kFlNeg0 : 0u);
		const uint32_t D = sv.diameter;
		const int a[3] = {p.a0, p.a1, p.a2};
		uint32_t mul[3], mask[3], rs[3];
		for (int i = 0; i < 3; i++)
		{
			if (ST == kStorageHash) { mul[i] = 1u << (kHashKeyBits * (2 - a[i])); mask[i] = 0u; }
			else { const uint32_t sh = (uint32_t)(6 - 3 * a[i]); mul[i] = 65u << sh; mask[i] = 0x1E07u << sh; }
			rs[i] = a[i] == 0 ? 1u : (a[i] == 1 ? D : D * D);
		}
		st4(kc, kKcD, k.d[0], k.d[1], k.d[2], thr);
		st4(kc, kKcRD, k.rd[0], k.rd[1], k.rd[2], bitsf(rs[2]));
		if (kLA)
		{
			st4(kc, kKcSD, ko.d[0], ko.d[1], ko.d[2], thr);
			st4(kc, kKcSRD, ko.rd[0], ko.rd[1], ko.rd[2], 0.0f);
		}
		st4(kc, kKcMul, bitsf(mul[0]), bitsf(mul[1]), bitsf(mul[2]), bitsf(rs[0]));
		st4(kc, kKcMask, bitsf(mask[0]), bitsf(mask[1]), bitsf(mask[2]), bitsf(rs[1]));
	}

	// VoxelScene::isRayInScene + getRegionStorageStructure (Renderer.cuh:29-44)
	VRM_HD void read_region_entry(RayCtx<ST, STATS>& c, const Vec4* kc)
	{
		const uint32_t D = c.sv.diameter;
		if (!(ur[0] < D && ur[1] < D && ur[2] < D)) { ri = -2; return; }
		if (STATS) c.st.nRegionReads++;
		const uint32_t rs0 = fbits(ld(kc, kKcMul).w), rs1 = fbits(ld(kc, kKcMask).w), rs2 = fbits(ld(kc, kKcRD).w);
		ri = ldg(c.sv.regionTable + (ur[0] * rs0 + ur[1] * rs1 + ur[2] * rs2));
	}

	// rebase into the neighbouring region after the ray left the current one (Renderer.cuh:421-429) and read its entry
	VRM_HD void change_region(RayCtx<ST, STATS>& c, const Vec4* kc)
	{
		for (int i = 0; i < 3; i++)
		{
			const int diff = floor_to_int(vmul(o[i], 0.015625f));  // o / 64 (exact scaling by a power of two)
			ur[i] += (uint32_t)diff;
			o[i] = vsub(o[i], (float)(diff * kRegion));
		}
		pc = kPcRegion;
		if (!position_sane(o)) { ri = -2; return; }  // see position_sane (vrm_core.cuh)
		read_region_entry(c, kc);
	}

	// rayMarchVoxelScene / rayMarchVoxelSceneLongestAxis up to the first region (Renderer.cuh:338-378, 917-954)
	VRM_HD void start_primary(RayCtx<ST, STATS>& c, Vec4* kc, const float* originW, const float* dirW, float scale)
	{
		fl = 0; result = 0; pc = kPcRegion; ri = -2; seq = 0; ad1 = ad2 = 0; hitMode = 0;
		g[0] = g[1] = g[2] = 0; ro[0] = ro[1] = ro[2] = 0.0f;
		PermRuntime p;
		p.a0 = 0; p.a1 = 1; p.a2 = 2;
		if constexpr (kLA) p = rank_axes(dirW[0], dirW[1], dirW[2]);
		float sW[3] = {canonical_zero(vmul(scale, vsub(originW[0], c.translation[0]))), canonical_zero(vmul(scale, vsub(originW[1], c.translation[1]))),
		               canonical_zero(vmul(scale, vsub(originW[2], c.translation[2])))};
		float dw[3];
		to_walk(p, sW, o); to_walk(p, dirW, dw);
		const RayDir k = make_raydir(dw[0], dw[1], dw[2]);
		float thr = k.thr;
		RayDir ko = k;
		if constexpr (kLA)
		{
			ko = scaled_raydir(k);
			if (!(ko.thr == ko.thr)) thr = NAN;
		}
		set_constants(c.sv, kc, p, k, ko, thr);
		int reg[3] = {floor_to_int(vmul(o[0], 0.015625f)), floor_to_int(vmul(o[1], 0.015625f)), floor_to_int(vmul(o[2], 0.015625f))};
		const int minC = c.sv.minCoord;
		const uint32_t D = c.sv.diameter;
		while (reg[0] - minC < 0 || reg[1] - minC < 0 || reg[2] - minC < 0 ||
		       (uint32_t)(reg[0] - minC) > D - 1 || (uint32_t)(reg[1] - minC) > D - 1 || (uint32_t)(reg[2] - minC) > D - 1)
		{
			// scene-entry loop, Renderer.cuh:349-373
			const int far = (int)(D + (uint32_t)minC);
			float a0, a1, a2;
			div3(vsub((float)((k.d[0] < 0.0f ? far : minC) * kRegion), o[0]), vsub((float)((k.d[1] < 0.0f ? far : minC) * kRegion), o[1]),
			     vsub((float)((k.d[2] < 0.0f ? far : minC) * kRegion), o[2]), k.d[0], k.d[1], k.d[2], k.rd[0], k.rd[1], k.rd[2], thr, a0, a1, a2);
			if (a0 <= 0.0f) a0 = INFINITY;
			if (a1 <= 0.0f) a1 = INFINITY;
			if (a2 <= 0.0f) a2 = INFINITY;
			const float m = min3(a0, a1, a2);
			if (m == INFINITY || m != m) { finish(0); return; }  // (a NaN tMin only arises from 0/0: treated as a miss)
			const float s = vadd(m, kEps);
			o[0] = along(o[0], s, k.d[0]); o[1] = along(o[1], s, k.d[1]); o[2] = along(o[2], s, k.d[2]);
			reg[0] = floor_to_int(vmul(o[0], 0.015625f)); reg[1] = floor_to_int(vmul(o[1], 0.015625f)); reg[2] = floor_to_int(vmul(o[2], 0.015625f));
		}
		for (int i = 0; i < 3; i++)
		{
			o[i] = vsub(o[i], (float)(reg[i] * kRegion));  // Renderer.cuh:376-378
			ur[i] = (uint32_t)(reg[i] - minC);
		}
		read_region_entry(c, kc);
	}

	// ---- kPcHit: applyLighting(...) * !isInShadow...(Ray(hit, LIGHT_DIRECTION), currentRegion)  (Renderer.cuh:314-315,821-822,...)
	VRM_HD void record_hit(RayCtx<ST, STATS>& c, uint32_t col, float p0, float p1, float p2, int nAxisW, float nSign, bool laKind, int v0, int v1, int v2)
	{
		if (shadow()) { finish(0); return; }  // any voxel on the shadow ray: colour * !inShadow = 0
		if (c.hitOut)
		{
			const PermRuntime p = unpack_perm(fl);
			const int minC = c.sv.minCoord;
			const int reg[3] = {(int)ur[0] + minC, (int)ur[1] + minC, (int)ur[2] + minC};
			record_hit_voxel(c, p, reg, v0, v1, v2);
		}
		result = col;
		ro[0] = p0; ro[1] = p1; ro[2] = p2;
		hitMode = nAxisW | (nSign < 0.0f ? 4 : 0) | (laKind ? 8 : 0);
		pc = kPcHit;
	}

	VRM_HD void do_hit(RayCtx<ST, STATS>& c, Vec4* kc)
	{
		const int nAxisW = hitMode & 3;
		const float nSign = (hitMode & 4) ? -1.0f : 1.0f;
		const bool laKind = kLA && (hitMode & 8) != 0;
		float hitW[3];
		int regW[3];
		{
			const PermRuntime p = unpack_perm(fl);
			const int minC = c.sv.minCoord;
			const int reg[3] = {(int)ur[0] + minC, (int)ur[1] + minC, (int)ur[2] + minC};
			to_world(p, ro, hitW); to_world(p, reg, regW);
		}
		result = apply_lighting_flat(c.light, c.translation, result, nAxisW, nSign, hitW, regW);
		if (!c.light.useShadows) { finish(result); return; }
		fl = kFlShadow | (laKind ? kFlShadowLA : 0u);
		PermRuntime p;
		p.a0 = 0; p.a1 = 1; p.a2 = 2;
		if (laKind) p = unpack_perm(c.lw.laPerm);
		to_walk(p, hitW, o);
		{
			int regWalk[3];
			to_walk(p, regW, regWalk);
			const int minC = c.sv.minCoord;
			for (int i = 0; i < 3; i++) ur[i] = (uint32_t)(regWalk[i] - minC);
		}
		RayDir k, ko;
		float thr;
		if (laKind)
		{
			for (int i = 0; i < 3; i++) { k.d[i] = c.lw.laD[i]; k.rd[i] = c.lw.laR[i]; ko.d[i] = c.lw.laSD[i]; ko.rd[i] = c.lw.laSR[i]; }
			thr = c.lw.laThr;
		}
		else
		{
			for (int i = 0; i < 3; i++) { k.d[i] = c.lw.idD[i]; k.rd[i] = c.lw.idR[i]; ko.d[i] = k.d[i]; ko.rd[i] = k.rd[i]; }
			thr = c.lw.idThr;
		}
		set_constants(c.sv, kc, p, k, ko, thr);
		pc = kPcRegion;
		read_region_entry(c, kc);
	}

	// ---- kPcRegion --------------------------------------------------------------------------------------------------------------
	VRM_HD void do_region(RayCtx<ST, STATS>& c, const Vec4* kc)
	{
		if (ri == -2) { finish(shadow() ? result : 0u); return; }  // left the scene: background / not shadowed
		if (ri == -1) { pc = kPcAdvRegion; return; }                // null region: skip to its edge through the advance site
		r = load_region<ST>(c.sv, ri);
		bool la = false;
		if constexpr (kLA) la = !shadowOriginal();
		if (la)
		{
			// rayMarchVoxelGridLongestAxis prologue, Renderer.cuh:763-784.  The reference divides by the scaled longest
			// component +-1.0f: x / 1 = x and x / -1 = -x exactly.
			const Vec4 S = ld(kc, kKcSD);
			g[0] = (int)o[0]; g[1] = (int)o[1]; g[2] = (int)o[2];
			const float t = (fl & kFlNeg0) ? -vsub(vsub((float)g[0], kEps), o[0]) : vsub(vadd(vadd((float)g[0], kEps), 1.0f), o[0]);
			ro[0] = along(o[0], t, S.x); ro[1] = along(o[1], t, S.y); ro[2] = along(o[2], t, S.z);
			ad1 = (int)ro[1] - g[1];
			ad2 = (int)ro[2] - g[2];
			pc = kPcHead;
		}
		else pc = kPcAdvNext;  // the region march starts with one step before the first test (Renderer.cuh:269-280)
	}

	// ---- kPcHead (longest axis): Renderer.cuh:787-805 ------------------------------------------------------------------
	VRM_HD void do_head(const Vec4* kc)
	{
		const int ad0 = (fl & kFlNeg0) ? -1 : 1;
		if (!grid_in_region(g[0] + ad0, g[1] + ad1, g[2] + ad2))
		{
			// Renderer.cuh:911-914: finish the region with the original algorithm from oldRay's origin (o already is it)
			pc = kPcAdvNext;
			return;
		}
		if (ad2 != 0 && ad1 != 0)
		{
			const Vec4 S = ld(kc, kKcSD), Q = ld(kc, kKcSRD);
			const float rounded = S.y < 0.0f ? floorf(o[1]) : ceilf(o[1]);  // Renderer.cuh:784
			const float tt = div1(vsub(rounded, o[1]), S.y, Q.y, S.w);
			const float shortestPosition = vadd(o[2], vmul(S.z, tt));
			const int shorterDiff = floor_to_int(shortestPosition) - g[2];
			seq = shorterDiff != 0 ? (2u | (1u << 2)) : (1u | (2u << 2));
		}
		else seq = ad1 != 0 ? 1u : (ad2 != 0 ? 2u : 0u);
		pc = kPcTest;
	}

	// ---- the advance (pc in kPcAdvNext .. kPcAdvRegion) -----------------------------------------------------------------------
	// Leaves the lane in its advance pc when the ray is still inside the region (the voxel test follows), in kPcRegion when it left,
	// with kPcParkBit set -- state untouched -- when the step needs one of the slow paths.  a0..a2 / s: the t values and the step of THIS
	// advance, for the hit normal of a jump (Renderer.cuh:733-738).
	VRM_HD void do_advance(RayCtx<ST, STATS>& c, const Vec4* kc, float& a0, float& a1, float& a2, float& s)
	{
		const bool jump = kLA && pc == kPcAdvJump;
		const bool skip = pc == kPcAdvRegion;
		const Vec4 E = ld(kc, jump ? kKcSD : kKcD);    // e0 e1 e2 thr
		const Vec4 Q = ld(kc, jump ? kKcSRD : kKcRD);  // q0 q1 q2
		float n0, n1, n2;
		if (pc == kPcAdvNext)
		{
			n0 = next_edge(E.x, o[0]); n1 = next_edge(E.y, o[1]); n2 = next_edge(E.z, o[2]);
		}
		else
		{
			// cluster edge of the voxel of the failed test ((int)o for the original algorithm, gridValues in a jump: g either
			// way), or the region's far face +EPSILON / near face -EPSILON
			const float lo = skip ? vsub(0.0f, kEps) : 0.0f, hi = skip ? vadd((float)kRegion, kEps) : 8.0f;
			const uint32_t keep = skip ? 0u : ~7u;
			n0 = vadd((float)(int)((uint32_t)g[0] & keep), E.x > 0.0f ? hi : lo);
			n1 = vadd((float)(int)((uint32_t)g[1] & keep), E.y > 0.0f ? hi : lo);
			n2 = vadd((float)(int)((uint32_t)g[2] & keep), E.z > 0.0f ? hi : lo);
		}
		const float x0 = vsub(n0, o[0]), x1 = vsub(n1, o[1]), x2 = vsub(n2, o[2]);
		// ONE test sends every slow case away: unsafe direction (thr is NaN), a numerator in the denormal range, and an exactly zero
		// numerator -- the only way min t can be 0, i.e. the crawl / ping-pong situations and their fast-forwards
		const float mabs = fminf(fabsf(x0), fminf(fabsf(x1), fabsf(x2)));
		if (!(mabs >= E.w)) { pc |= kPcParkBit; return; }
		a0 = div_by_const(x0, E.x, Q.x); a1 = div_by_const(x1, E.y, Q.y); a2 = div_by_const(x2, E.z, Q.z);
		// (the zero-direction guards of the original shadow routines, Renderer.cuh:113-115,191-193, only act on a component that is
		// exactly zero; such a ray has thr = NaN and never gets here)
		const float m = min3(a0, a1, a2);
		s = skip ? m : vadd(m, kEps);
		if (pc == kPcAdvNext)
		{
			// tX,tY,tZ,tMin of Renderer.cuh:273-277 (a cluster skip leaves them stale): all the normal ever asks is which equal tMin
			fl = (fl & ~kFlEqMask) | (a0 == m ? kFlEq0 : 0u) | (a1 == m ? kFlEq0 << 1 : 0u) | (a2 == m ? kFlEq0 << 2 : 0u);
		}
		o[0] = along(o[0], s, E.x); o[1] = along(o[1], s, E.y); o[2] = along(o[2], s, E.z);
		// grid_in_region((int)floorf(o)) of the jump (Renderer.cuh:719-723) and isRayInRegion(o) agree for every o
		if (skip || !ray_in_region(o)) { change_region(c, kc); return; }
		g[0] = (int)o[0]; g[1] = (int)o[1]; g[2] = (int)o[2];  // == (int)floorf(o) inside a region
	}

	// ---- the voxel test: doesVoxelSpaceExist + lookupVoxel on region-local walk-space coordinates ---------------------------
	VRM_HD bool voxel_test(RayCtx<ST, STATS>& c, const Vec4* kc, int c0, int c1, int c2, uint32_t& col)
	{
		col = kEmpty;
		const Vec4 M = ld(kc, kKcMul);
		if constexpr (ST == kStorageHash)
		{
			// key = c0 << ks0 | c1 << ks1 | c2 << ks2 with 7-bit fields: the products cannot overlap (a coordinate is at most 64)
			const uint32_t key = (uint32_t)c0 * fbits(M.x) + (uint32_t)c1 * fbits(M.y) + (uint32_t)c2 * fbits(M.z);
#if VRM_HASH_CLUSTER_FILTER
			if (hash_cluster_occupied(c.sv.clusterMask, r.ri, key))  // negative filter, see lookup_voxel (vrm_core.cuh)
#endif
			{
				const unsigned long long e1 = ldg(c.sv.slots + (r.base1 + hash_slot1(key, r.seed1, r.n)));
				const unsigned long long e2 = ldg(c.sv.slots + (r.base2 + hash_slot2(key, r.seed2, r.n)));
				if ((uint32_t)(e1 >> 32) == key) col = (uint32_t)e1;
				else if ((uint32_t)(e2 >> 32) == key) col = (uint32_t)e2;
			}
			if (STATS) { c.st.nExist++; c.st.nLookup++; c.st.nProbe2++; if (col != kEmpty) c.st.nLookupHit++; }
			return true;
		}
		else
		{
			// cc = cluster id << 9 | in-cluster code; a coordinate v contributes ((v * 65) & 0x1E07) << shift = (v * mul) & mask
			// (vrm_flat.cuh voxel_test; bit 12 of the mask only matters for v = 64, the reference's undefined corner)
			const Vec4 K = ld(kc, kKcMask);
			const uint32_t cc = (((uint32_t)c0 * fbits(M.x)) & fbits(K.x)) | (((uint32_t)c1 * fbits(M.y)) & fbits(K.y)) | (((uint32_t)c2 * fbits(M.z)) & fbits(K.z));
			const uint2 h = ldg(c.sv.headers + ((uint32_t)ri * 8192u + (cc >> 5)));
			const bool e = (h.y & kHeaderClusterExists) != 0;
			if (STATS) { c.st.nExist++; if (!e) c.st.nExistFalse++; }
			if (!e) return false;
			const uint32_t bit = cc & 31u;
			if ((h.x >> bit) & 1u)
			{
				col = ldg(c.sv.values + ((h.y & ~kHeaderClusterExists) + (uint32_t)popc32(h.x & ((1u << bit) - 1u))));
#if VRM_COORD64_EMPTY
				if constexpr (kLA) { if (((uint32_t)c0 | (uint32_t)c1 | (uint32_t)c2) & 64u) col = kEmpty; }  // "a coordinate of 64", vrm_core.cuh
#endif
			}
			if (STATS) { c.st.nLookup++; if (col != kEmpty) c.st.nLookupHit++; }
			return true;
		}
	}

	// ---- [bump] + voxel test + what comes next (pc in kPcTest .. kPcAdvJump) -------------------------------------------------
	VRM_HD void do_test(RayCtx<ST, STATS>& c, const Vec4* kc, float a0, float a1, float a2, float s)
	{
		int slot = 0;
		const int was = pc;
		if (kLA && was == kPcTest)
		{
			slot = (int)(seq & 3u);
			seq >>= 2;
			g[0] += slot == 0 ? ((fl & kFlNeg0) ? -1 : 1) : 0;
			g[1] += slot == 1 ? ad1 : 0;
			g[2] += slot == 2 ? ad2 : 0;
		}
		uint32_t col;
		const bool e = voxel_test(c, kc, g[0], g[1], g[2], col);
		if (col != kEmpty)
		{
			const PermRuntime p = unpack_perm(fl);
			if (!(kLA && was == kPcTest))
			{
				// original algorithm: Renderer.cuh:312-315; jump: Renderer.cuh:733-738 (tMin carries +EPSILON there, so the
				// comparison normally falls through to the Z normal).  getNormalFromTValues tests X, then Y, else Z.
				uint32_t eq = fl;
				if (kLA && was == kPcAdvJump) eq = (a0 == s ? kFlEq0 : 0u) | (a1 == s ? kFlEq0 << 1 : 0u) | (a2 == s ? kFlEq0 << 2 : 0u);
				int mW = 0;
				if (eq & kFlEq0) mW |= 1 << p.a0;
				if (eq & (kFlEq0 << 1)) mW |= 1 << p.a1;
				if (eq & (kFlEq0 << 2)) mW |= 1 << p.a2;
				const int nAxisW = (mW & 1) ? 0 : ((mW & 2) ? 1 : 2);
				const Vec4 Dv = ld(kc, kKcD);
				const float dn = p.a0 == nAxisW ? Dv.x : (p.a1 == nAxisW ? Dv.y : Dv.z);  // scaled direction = s * d, s > 0: same sign
				record_hit(c, col, o[0], o[1], o[2], nAxisW, copysignf(1.0f, -dn), kLA && was == kPcAdvJump, g[0], g[1], g[2]);
			}
			else
			{
				const Vec4 S = ld(kc, kKcSD);
				const float odS = pick3(slot, S.x, S.y, S.z);
				float p0 = ro[0], p1 = ro[1], p2 = ro[2];  // Renderer.cuh:899
				if (slot != 0)
				{
					// getLocalHitLocation, Renderer.cuh:753-758
					const float ooS = pick3(slot, o[0], o[1], o[2]);
					const float tl = odS > 0.0f ? vdiv(vsub(ceilf(ooS), ooS), odS) : vdiv(vsub(floorf(ooS), ooS), odS);
					p0 = along(o[0], tl, S.x); p1 = along(o[1], tl, S.y); p2 = along(o[2], tl, S.z);
				}
				record_hit(c, col, p0, p1, p2, p.axis(slot), copysignf(1.0f, -odS), true, g[0], g[1], g[2]);
			}
			return;
		}
		if (kLA && was == kPcTest)
		{
			if (!e)
			{
				// performVoxelSpaceJump (Renderer.cuh:808-810 -> 696): its while condition repeats the exist check of the failed
				// test on the same voxel -- same answer, only the counter sees it
				if (STATS) { c.st.nExist++; c.st.nExistFalse++; }
				pc = kPcAdvJump;
			}
			else if (slot == 0)
			{
				// the longest axis was this iteration's last test: Renderer.cuh:903-908
				const Vec4 S = ld(kc, kKcSD);
				o[0] = ro[0]; o[1] = ro[1]; o[2] = ro[2];
				ro[0] = vadd(ro[0], S.x); ro[1] = vadd(ro[1], S.y); ro[2] = vadd(ro[2], S.z);
				ad1 = (int)ro[1] - g[1];
				ad2 = (int)ro[2] - g[2];
				pc = kPcHead;
			}
		}
		else if (kLA && was == kPcAdvJump)
		{
			if (e)
			{
				// the jump reached a cluster that exists but the voxel is empty: re-snap to the longest axis and `continue`
				// the while loop (Renderer.cuh:742-750)
				const Vec4 S = ld(kc, kKcSD), Q = ld(kc, kKcSRD);
				const float tNext = div1(vsub(S.x > 0.0f ? ceilf(o[0]) : floorf(o[0]), o[0]), S.x, Q.x, S.w);
				const float tt = vadd(tNext, kEps);
				ro[0] = along(o[0], tt, S.x); ro[1] = along(o[1], tt, S.y); ro[2] = along(o[2], tt, S.z);
				ad1 = (int)ro[1] - g[1];
				ad2 = (int)ro[2] - g[2];
				pc = kPcHead;
			}
			// else: still no voxel space: another jump iteration
		}
		else pc = e ? kPcAdvNext : kPcAdvCluster;
	}

	// One pass over the marching blocks in program order (region -> head -> advance -> test).
	VRM_HD void step_marching(RayCtx<ST, STATS>& c, const Vec4* kc)
	{
		if (pc == kPcRegion) do_region(c, kc);
		if constexpr (kLA) { if (pc == kPcHead) do_head(kc); }
		float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, s = 0.0f;
		if (pc >= kPcAdvNext && pc <= kPcAdvRegion) do_advance(c, kc, a0, a1, a2, s);
		if (pc <= kPcAdvJump) do_test(c, kc, a0, a1, a2, s);
	}

};

// Single-ray driver (trace kernels, host sim).  Returns true when the ray finished here (colour in ray.result), false when it
// parked: the caller hands it to the generic machine.
template <int ST, int ALGO, bool STATS, int STRIDE>
VRM_HD bool march_scene_lean(RayCtx<ST, STATS>& c, Vec4* kc, const float* originW, const float* dirW, float scale, LeanRay<ST, ALGO, STATS, STRIDE>& ray)
{
	ray.start_primary(c, kc, originW, dirW, scale);
	for (;;)
	{
		while (ray.pc <= kPcHead) ray.step_marching(c, kc);
		if (ray.pc != kPcHit) break;
		ray.do_hit(c, kc);
	}
	return ray.pc == kPcDone;
}

#if defined(__CUDACC__)
// Warp-cooperative form with the HIT BARRIER of vrm_flat.cuh (march_scene_flat_warp): lanes whose primary ray has hit wait until
// every lane of the warp has hit, finished or parked; the tile then shades and starts its shadow rays in the same iteration.
template <int ST, int ALGO, bool STATS, int STRIDE>
__device__ __forceinline__ bool march_scene_lean_warp(RayCtx<ST, STATS>& c, Vec4* kc, bool active, const float* originW, const float* dirW, float scale,
                                                      LeanRay<ST, ALGO, STATS, STRIDE>& ray)
{
	ray.pc = kPcDone; ray.result = 0;
	if (active) ray.start_primary(c, kc, originW, dirW, scale);
	for (;;)
	{
		if (__any_sync(0xFFFFFFFFu, ray.pc <= kPcHead))
		{
			if (ray.pc <= kPcHead) ray.step_marching(c, kc);
			continue;
		}
		if (!__any_sync(0xFFFFFFFFu, ray.pc == kPcHit)) break;
		if (ray.pc == kPcHit) ray.do_hit(c, kc);
	}
	return ray.pc == kPcDone;
}
#endif

}  // namespace vrm
