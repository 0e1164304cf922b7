// TEST INFRASTRUCTURE ONLY.  Host compile of the product's per-ray traversal core (csrc/vrm_core.cuh, the exact source
// the sm_100a kernels are built from) so that its arithmetic and control flow can be checked against the oracle in
// the CPU-only test tier.  Nothing in the product links or calls this; the storage layouts are filled here by a
// naive sequential builder (the real one is the GPU pipeline in csrc/vrm_build.cu).
//   g++ -std=c++17 -O2 -ffp-contract=off -fPIC -shared -fopenmp
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <map>
#include <tuple>
#include <vector>

struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };

#include "../../voxelraymarcher_b200/csrc/vrm_core.cuh"
#include "../../voxelraymarcher_b200/csrc/vrm_flat.cuh"
#include "../../voxelraymarcher_b200/csrc/vrm_lean.cuh"

using namespace vrm;

namespace {

struct SimScene
{
	std::vector<int32_t> xyz;
	std::vector<uint32_t> rgb;
	int storage = -1;
	int32_t minCoord = 0, maxCoord = 0;
	uint32_t diameter = 0, filled = 0;
	std::vector<int32_t> regionTable;
	std::vector<HashRegionDesc> hashDesc;
	std::vector<unsigned long long> slots;
	std::vector<uint2> headers;
	std::vector<uint32_t> clusterMask;
	std::vector<uint32_t> values;
	Lighting light;
	SceneView view() const
	{
		SceneView v;
		v.regionTable = regionTable.data(); v.diameter = diameter; v.minCoord = minCoord;
		v.hashDesc = hashDesc.data(); v.slots = slots.data();
		v.headers = headers.data(); v.clusterMask = clusterMask.data(); v.values = values.data();
		return v;
	}
};

Lighting gLight = {{0.57735026f, 0.57735026f, 0.57735026f}, {1, 1, 1}, {10, 10, -10}, 0, 1};
unsigned long long gCrawlSkipped = 0;  // cluster-skip iterations fast-forwarded by crawl_skip (render calls only)
int gFlat = 1;  // 1: the flat state machine of vrm_flat.cuh; 0: the nested form of vrm_core.cuh; 2: the lean machine of vrm_lean.cuh; 3: the flat machine with its fast paths (vrm_flat.cuh fast_jump / fast_nullskip) taken per ray
unsigned long long gLeanParked = 0;  // rays the lean machine handed to the generic one (mode 2)

template <int ST, int ALGO>
uint32_t march(RayCtx<ST, true>& c, const float* o, const float* d, float scale)
{
	if (gFlat == 2)
	{
		Vec4 kc[kKcVectors];
		LeanRay<ST, ALGO, true, 1> ray;
		if (march_scene_lean<ST, ALGO, true, 1>(c, kc, o, d, scale, ray)) return ray.result;
		// parked: the ray is re-traced from its start by the generic machine, as resume_kernel does (vrm_render.cu)
		#pragma omp atomic
		gLeanParked++;
		c.reset();
	}
	if (gFlat == 3) return march_scene_flat_fast<ST, ALGO, true>(c, o, d, scale);  // the generic machine with its warp-uniform fast paths taken per ray
	return gFlat ? march_scene_flat<ST, ALGO, true>(c, o, d, scale) : march_scene<ST, ALGO, true>(c, o, d, scale);
}

int floordiv64(int v) { return v >= 0 ? v / 64 : -((-v + 63) / 64); }

template <int ST, int ALGO>
void render_rows(const SimScene& s, const float* cam, const float* tr, float scale, uint32_t W, uint32_t H, uint8_t* rgb, int32_t* hits, uint64_t* counters, uint32_t* lookups, int nThreads)
{
	uint64_t total[5] = {0, 0, 0, 0, 0};
	#pragma omp parallel num_threads(nThreads)
	{
		RayCtx<ST, true> c;
		c.sv = s.view(); c.light = gLight; c.lw = make_light_walk(gLight); c.hitOut = c.hit;
		c.translation[0] = tr[0]; c.translation[1] = tr[1]; c.translation[2] = tr[2];
		uint64_t local[5] = {0, 0, 0, 0, 0};
		unsigned long long crawl = 0;
		#pragma omp for schedule(dynamic, 1)
		for (int64_t y = 0; y < (int64_t)H; y++)
			for (uint32_t x = 0; x < W; x++)
			{
				float o[3], d[3];
				c.reset();
				if (gFlat) primary_ray_flat(cam, x, (uint32_t)y, W, H, 1.0f / (float)W, 1.0f / (float)H, o, d);  // what the state-machine kernels run
				else primary_ray(cam, x, (uint32_t)y, W, H, o, d);
				uint32_t color = march<ST, ALGO>(c, o, d, scale);
				size_t p = (size_t)y * W + x;
				rgb[3 * p] = (uint8_t)(color >> 16); rgb[3 * p + 1] = (uint8_t)((color >> 8) & 0xFF); rgb[3 * p + 2] = (uint8_t)(color & 0xFF);
				if (hits) memcpy(hits + 4 * p, c.hit, 16);
				if (lookups) lookups[p] = getenv("SIM_COUNT_STEPS") ? (uint32_t)(c.st.nExist - c.st.nCrawlSkipped) : (uint32_t)c.st.nLookup;
				local[0] += c.st.nExist; local[1] += c.st.nExistFalse; local[2] += c.st.nLookup; local[3] += c.st.nLookupHit;
				if (c.st.nLookup > local[4]) local[4] = c.st.nLookup;
				crawl += c.st.nCrawlSkipped;
			}
		#pragma omp critical
		{
			for (int i = 0; i < 4; i++) total[i] += local[i];
			if (local[4] > total[4]) total[4] = local[4];
			gCrawlSkipped += crawl;
		}
	}
	if (counters) memcpy(counters, total, sizeof(total));
}

template <int ST, int ALGO>
void trace(const SimScene& s, const float* rays, uint64_t n, const float* tr, float scale, uint32_t* colour, int32_t* hits, uint64_t* counters, int nThreads)
{
	uint64_t total[5] = {0, 0, 0, 0, 0};
	#pragma omp parallel num_threads(nThreads)
	{
		RayCtx<ST, true> c;
		c.sv = s.view(); c.light = gLight; c.lw = make_light_walk(gLight); c.hitOut = c.hit;
		c.translation[0] = tr[0]; c.translation[1] = tr[1]; c.translation[2] = tr[2];
		uint64_t local[5] = {0, 0, 0, 0, 0};
		#pragma omp for schedule(dynamic, 256)
		for (int64_t i = 0; i < (int64_t)n; i++)
		{
			c.reset();
			colour[i] = march<ST, ALGO>(c, rays + 6 * i, rays + 6 * i + 3, scale);
			if (hits) memcpy(hits + 4 * i, c.hit, 16);
			#pragma omp atomic
			gCrawlSkipped += c.st.nCrawlSkipped;
			local[0] += c.st.nExist; local[1] += c.st.nExistFalse; local[2] += c.st.nLookup; local[3] += c.st.nLookupHit;
			if (c.st.nLookup > local[4]) local[4] = c.st.nLookup;
		}
		#pragma omp critical
		{
			for (int i = 0; i < 4; i++) total[i] += local[i];
			if (local[4] > total[4]) total[4] = local[4];
		}
	}
	if (counters) memcpy(counters, total, sizeof(total));
}

}  // namespace

extern "C" {

void sim_set_flat(int flat) { gFlat = flat; }
unsigned long long sim_lean_parked() { unsigned long long v = gLeanParked; gLeanParked = 0; return v; }
int sim_is_flat() { return gFlat; }

// crawl_skip against the literal iterations it replaces: pseudo-random positions (many of them exactly on cluster faces, integers and
// powers of two, or a few ulps away from them) and directions whose EPSILON steps range from "cannot move the coordinate" to
// hundreds of ulps.  For every case in which crawl_skip skips M > 0 iterations: executing o <- RN(o + RN(EPSILON * d)) M times must
// give the same bits, and every intermediate position must still truncate into the cluster cell that is being skipped.
// Returns the number of failing cases; *skippedCases receives how many cases actually skipped something.
uint64_t sim_check_crawl_skip(uint32_t seed, uint64_t count, uint64_t* skippedCases)
{
	uint64_t bad = 0, used = 0, st = seed * 0x9E3779B97F4A7C15ull + 12345;
	auto next = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (uint32_t)(st >> 20); };
	auto unit = [&]() { return (float)(next() & 0xFFFFFF) / 16777216.0f; };
	for (uint64_t n = 0; n < count; n++)
	{
		float o[3], d[3];
		int v[3];
		for (int i = 0; i < 3; i++)
		{
			const int cell = (int)(next() % 8u) * 8;
			float y;
			switch (next() % 6u)
			{
			case 0: y = (float)cell; break;                                            // on the lower cluster face
			case 1: y = (float)(cell + (int)(next() % 8u)); break;                      // on an integer
			case 2: y = (float)(1u << (next() % 6u)); break;                            // a power of two
			case 3: y = bits_float(float_bits((float)(cell + 1 + (int)(next() % 7u))) + (next() % 5u) - 2u); break;  // a few ulps around an integer
			default: y = (float)cell + 8.0f * unit(); break;
			}
			if (!(y >= 0.0f && y < 64.0f)) y = 1.5f;
			o[i] = y;
			v[i] = (int)y;
			const float mag = (next() % 3u == 0u) ? 1.0f : ((next() % 2u) ? 0.05f : 0.004f);
			d[i] = (unit() * 2.0f - 1.0f) * mag;
			if (d[i] == 0.0f) d[i] = 0.001f;
		}
		const RayDir k = make_raydir(d[0], d[1], d[2]);
		float p[3] = {o[0], o[1], o[2]};
		const int m = crawl_skip(p, k, v[0], v[1], v[2]);
		if (m <= 0) continue;
		used++;
		float q[3] = {o[0], o[1], o[2]};
		bool ok = true;
		for (int it = 0; it < m && ok; it++)
			for (int i = 0; i < 3; i++)
			{
				q[i] = vadd(q[i], vmul(kEps, k.d[i]));
				if (it + 1 < m && (((int)q[i]) & ~7) != (v[i] & ~7)) ok = false;
			}
		for (int i = 0; i < 3; i++) if (float_bits(q[i]) != float_bits(p[i])) ok = false;
		if (!ok) bad++;
	}
	if (skippedCases) *skippedCases = used;
	return bad;
}

// primary_ray_flat's image-plane divisions against IEEE division: every pixel coordinate of an image side of n pixels.
// Returns the number of mismatches (0 expected).
uint64_t sim_check_image_division(uint32_t n)
{
	uint64_t bad = 0;
	const float fn = (float)n, inv = 1.0f / fn;
	for (uint32_t x = 0; x <= n; x++)
	{
		const float num = vadd((float)x, 0.5f);
		const float q = div_by_const(num, fn, inv), ref = num / fn;
		if (memcmp(&q, &ref, 4) != 0) bad++;
	}
	return bad;
}

// div3 (reciprocal + FMA residual form with its slow-path guard) against IEEE division for count pseudo-random numerator
// triples divided by one common denominator (the ray-length normalisation of primary_ray_flat).
uint64_t sim_check_common_division(uint32_t seed, uint64_t count)
{
	uint64_t bad = 0, st = seed * 0x9E3779B97F4A7C15ull + 1;
	auto next = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (uint32_t)(st >> 16); };
	for (uint64_t i = 0; i < count; i++)
	{
		float x[3], len;
		for (int k = 0; k < 3; k++) { uint32_t b = (next() & 0x807FFFFFu) | ((100u + next() % 40u) << 23); memcpy(&x[k], &b, 4); }
		{ uint32_t b = (next() & 0x007FFFFFu) | ((110u + next() % 30u) << 23); memcpy(&len, &b, 4); }
		if (i % 7 == 0) x[i % 3] = 0.0f;
		const float rl = vrcp(len);
		const float thr = dir_component_safe(len) ? 7.888609052210118e-31f : NAN;
		float q[3];
		div3(x[0], x[1], x[2], len, len, len, rl, rl, rl, thr, q[0], q[1], q[2]);
		for (int k = 0; k < 3; k++) { const float ref = x[k] / len; if (memcmp(&q[k], &ref, 4) != 0) bad++; }
	}
	return bad;
}
unsigned long long sim_crawl_skipped() { unsigned long long v = gCrawlSkipped; gCrawlSkipped = 0; return v; }
void* sim_scene_create() { return new SimScene(); }
void sim_scene_destroy(void* h) { delete static_cast<SimScene*>(h); }

void sim_scene_add_voxels(void* h, const int32_t* xyz, const uint32_t* rgb, uint64_t n)
{
	SimScene* s = static_cast<SimScene*>(h);
	s->xyz.insert(s->xyz.end(), xyz, xyz + 3 * n);
	s->rgb.insert(s->rgb.end(), rgb, rgb + n);
}

int sim_scene_build(void* h, int storageType)
{
	SimScene* s = static_cast<SimScene*>(h);
	if (s->storage != -1) return 1;
	size_t n = s->rgb.size();
	// region -> (cluster-major 18-bit code -> colour), last write wins
	std::map<std::tuple<int, int, int>, std::map<uint32_t, uint32_t>> regions;
	for (size_t i = 0; i < n; i++)
	{
		int r[3]; uint32_t l[3];
		for (int a = 0; a < 3; a++)
		{
			int v = s->xyz[3 * i + a];
			r[a] = floordiv64(v);
			l[a] = (uint32_t)(v - r[a] * 64);
			s->minCoord = std::min(s->minCoord, r[a]);
			s->maxCoord = std::max(s->maxCoord, r[a]);
		}
		uint32_t cid = ((l[0] >> 3) << 6) | ((l[1] >> 3) << 3) | (l[2] >> 3);
		uint32_t code = ((l[0] & 7) << 6) | ((l[1] & 7) << 3) | (l[2] & 7);
		regions[std::make_tuple(r[2], r[1], r[0])][(cid << 9) | code] = s->rgb[i];
	}
	s->diameter = (uint32_t)(s->maxCoord - s->minCoord + 1);
	uint32_t D = s->diameter;
	s->regionTable.assign((size_t)D * D * D, -1);
	s->filled = (uint32_t)regions.size();
	s->headers.assign((size_t)s->filled * 512 * 16, uint2{0, 0});
	s->clusterMask.assign((size_t)s->filled * 16, 0);
	int32_t ri = 0;
	for (auto& kv : regions)
	{
		int rz = std::get<0>(kv.first), ry = std::get<1>(kv.first), rx = std::get<2>(kv.first);
		s->regionTable[(size_t)(rx - s->minCoord) + (size_t)(ry - s->minCoord) * D + (size_t)(rz - s->minCoord) * D * D] = ri;
		// VCS: colours in (cluster, code) order + per-word {mask, first index}
		uint2* hdr = s->headers.data() + (size_t)ri * 512 * 16;
		for (auto& v : kv.second)
		{
			uint32_t cid = v.first >> 9, code = v.first & 511;
			hdr[cid * 16 + (code >> 5)].x |= 1u << (code & 31);
			s->clusterMask[(size_t)ri * 16 + (cid >> 5)] |= 1u << (cid & 31);
			s->values.push_back(v.second);
		}
		uint32_t running = (uint32_t)(s->values.size() - kv.second.size());
		for (uint32_t w = 0; w < 512 * 16; w++) { hdr[w].y = running; running += (uint32_t)__builtin_popcount(hdr[w].x); }
		for (uint32_t w = 0; w < 512 * 16; w++)
			if ((s->clusterMask[(size_t)ri * 16 + (w >> 9)] >> ((w >> 4) & 31)) & 1u) hdr[w].y |= kHeaderClusterExists;
		// cuckoo: sequential insertion with the product's hash functions
		uint32_t N = (uint32_t)kv.second.size();
		HashRegionDesc d;
		d.n = N + N / 4 + 2;
		d.slotBase = (uint32_t)s->slots.size();
		s->slots.resize(s->slots.size() + 2 * (size_t)d.n, kEmptySlot);
		for (uint32_t attempt = 0;; attempt++)
		{
			d.seed1 = (0x9E3779B9u * (attempt + 1) + (uint32_t)ri * 0x85EBCA77u) | 1u; d.seed2 = ((0x7F4A7C15u * (attempt + 1)) ^ ((uint32_t)ri * 0xC2B2AE3Du) ^ 0x27D4EB2Fu) | 1u;  // odd multipliers
			std::fill(s->slots.begin() + d.slotBase, s->slots.end(), kEmptySlot);
			bool ok = true;
			for (auto& v : kv.second)
			{
				uint32_t cid = v.first >> 9, code = v.first & 511;
				uint32_t x = ((cid >> 6) << 3) | (code >> 6), y = (((cid >> 3) & 7) << 3) | ((code >> 3) & 7), z = ((cid & 7) << 3) | (code & 7);
				unsigned long long e = ((unsigned long long)hash_key(x, y, z) << 32) | v.second;
				int table = 0, it = 0;
				for (; it < 500; it++)
				{
					uint32_t key = (uint32_t)(e >> 32);
					size_t idx = table == 0 ? d.slotBase + hash_slot1(key, d.seed1, d.n) : d.slotBase + d.n + hash_slot2(key, d.seed2, d.n);
					std::swap(e, s->slots[idx]);
					if (e == kEmptySlot) break;
					table ^= 1;
				}
				if (it == 500) { ok = false; break; }
			}
			if (ok) break;
			if (attempt > 64) return 2;
		}
		s->hashDesc.push_back(d);
		ri++;
	}
	// zeroed guard space, as in the product's builder: the reference can test a voxel with a coordinate of exactly 64 (undefined
	// behaviour there) and both forms of the traversal then form cluster ids of up to (8 << 6) | (8 << 3) | 8 = 584
	s->headers.resize(s->headers.size() + 80 * 16, uint2{0, 0});
	s->clusterMask.resize(s->clusterMask.size() + 32, 0u);
	s->storage = storageType;
	return 0;
}

void sim_scene_info(void* h, uint32_t* diameter, int32_t* minCoord, uint32_t* filled)
{
	SimScene* s = static_cast<SimScene*>(h);
	*diameter = s->diameter; *minCoord = s->minCoord; *filled = s->filled;
}

void sim_set_lighting(const float* dir, const float* color, const float* pos, int usePoint, int useShadows)
{
	memcpy(gLight.dir, dir, 12); memcpy(gLight.color, color, 12); memcpy(gLight.pos, pos, 12);
	gLight.usePoint = usePoint != 0; gLight.useShadows = useShadows != 0;
}

void sim_make_unit_vector(const float* v, float* out)
{
	float l = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
	out[0] = v[0] / l; out[1] = v[1] / l; out[2] = v[2] / l;
}

void sim_camera_make(const float*, const float*, const float*, float, float, float*) {}  // cameras come from the oracle / product API

int sim_render(void* h, const float* cam, const float* tr, uint32_t scale, int algorithm, uint32_t W, uint32_t H,
	uint8_t* rgb, int32_t* hits, uint64_t* counters, uint32_t* lookups, int nThreads)
{
	SimScene* s = static_cast<SimScene*>(h);
	if (s->storage == -1) return 1;
	if (nThreads < 1) nThreads = 1;
	float sc = (float)scale;
	if (s->storage == kStorageHash)
	{
		if (algorithm == kAlgoOriginal) render_rows<kStorageHash, kAlgoOriginal>(*s, cam, tr, sc, W, H, rgb, hits, counters, lookups, nThreads);
		else render_rows<kStorageHash, kAlgoLongestAxis>(*s, cam, tr, sc, W, H, rgb, hits, counters, lookups, nThreads);
	}
	else
	{
		if (algorithm == kAlgoOriginal) render_rows<kStorageVcs, kAlgoOriginal>(*s, cam, tr, sc, W, H, rgb, hits, counters, lookups, nThreads);
		else render_rows<kStorageVcs, kAlgoLongestAxis>(*s, cam, tr, sc, W, H, rgb, hits, counters, lookups, nThreads);
	}
	return 0;
}

int sim_trace_rays(void* h, const float* rays, uint64_t n, const float* tr, uint32_t scale, int algorithm,
	uint32_t* colour, int32_t* hits, uint64_t* counters, int nThreads)
{
	SimScene* s = static_cast<SimScene*>(h);
	if (s->storage == -1) return 1;
	if (nThreads < 1) nThreads = 1;
	float sc = (float)scale;
	if (s->storage == kStorageHash)
	{
		if (algorithm == kAlgoOriginal) trace<kStorageHash, kAlgoOriginal>(*s, rays, n, tr, sc, colour, hits, counters, nThreads);
		else trace<kStorageHash, kAlgoLongestAxis>(*s, rays, n, tr, sc, colour, hits, counters, nThreads);
	}
	else
	{
		if (algorithm == kAlgoOriginal) trace<kStorageVcs, kAlgoOriginal>(*s, rays, n, tr, sc, colour, hits, counters, nThreads);
		else trace<kStorageVcs, kAlgoLongestAxis>(*s, rays, n, tr, sc, colour, hits, counters, nThreads);
	}
	return 0;
}

int sim_lookup(void* h, const int32_t* xyz, uint64_t n, uint32_t* out, uint8_t* exists)
{
	SimScene* s = static_cast<SimScene*>(h);
	if (s->storage == -1) return 1;
	PermIdentity p;
	for (uint64_t i = 0; i < n; i++)
	{
		int reg[3], l[3];
		for (int a = 0; a < 3; a++) { reg[a] = floordiv64(xyz[3 * i + a]); l[a] = xyz[3 * i + a] - reg[a] * 64; }
		out[i] = kEmpty;
		if (exists) exists[i] = 0;
		if (s->storage == kStorageHash)
		{
			RayCtx<kStorageHash, false> c; c.sv = s->view(); c.reset();
			int32_t ri = region_entry(c, p, reg);
			if (ri < 0) continue;
			auto r = load_region<kStorageHash>(c.sv, ri);
			if (exists) exists[i] = 1;
			out[i] = lookup_voxel(c, r, p, reg, l[0], l[1], l[2]);
		}
		else
		{
			RayCtx<kStorageVcs, false> c; c.sv = s->view(); c.reset();
			int32_t ri = region_entry(c, p, reg);
			if (ri < 0) continue;
			auto r = load_region<kStorageVcs>(c.sv, ri);
			bool e = space_exists(c, r, p, l[0], l[1], l[2]);
			if (exists) exists[i] = e;
			if (e) out[i] = lookup_voxel(c, r, p, reg, l[0], l[1], l[2]);
		}
	}
	return 0;
}

}  // extern "C"

// Debug aid: step one ray through the state machine and print a window of micro-steps.
extern "C" int sim_debug_ray(void* h, const float* ray, int algorithm, long from, long count)
{
	SimScene* s = static_cast<SimScene*>(h);
	if (s->storage != kStorageVcs) return 1;
	RayCtx<kStorageVcs, true> c;
	c.sv = s->view(); c.light = gLight; c.lw = make_light_walk(gLight); c.hitOut = c.hit; c.translation[0] = c.translation[1] = c.translation[2] = 0.0f; c.reset();
	float tr1 = 1.0f;
	auto run = [&](auto& rayState) {
		rayState.start_primary(c, ray, ray + 3, tr1);
		long n = 0;
		while (rayState.st != kStDone && n < from + count)
		{
			if (n >= from)
				printf("step %ld st %d shadow %d ureg %u %u %u o %.9g %.9g %.9g  d %.9g %.9g %.9g exist %llu/%llu\n", n, rayState.st, (int)rayState.shadow(),
				       rayState.ur[0], rayState.ur[1], rayState.ur[2], rayState.o[0], rayState.o[1], rayState.o[2], rayState.d[0], rayState.d[1], rayState.d[2],
				       c.st.nExist, c.st.nExistFalse);
			rayState.step(c);
			n++;
		}
		return n;
	};
	if (algorithm == kAlgoOriginal) { FlatRay<kStorageVcs, kAlgoOriginal, true> r; run(r); }
	else { FlatRay<kStorageVcs, kAlgoLongestAxis, true> r; run(r); }
	return 0;
}

// ---- development aid (tools/warp_profile.py): what does a WARP of the render kernel execute? -----------------------------------
// A fast VCS-only builder for full-size scenes (the std::map builder above is for the small test scenes) and a lockstep simulation of
// march_scene_flat_warp over 8x4 pixel tiles that records, per pass, which blocks of the state machine have at least one lane.
extern "C" int sim_scene_build_vcs_fast(void* h)
{
	SimScene* s = static_cast<SimScene*>(h);
	if (s->storage != -1) return 1;
	const size_t n = s->rgb.size();
	for (size_t i = 0; i < 3 * n; i++) { const int r = floordiv64(s->xyz[i]); s->minCoord = std::min(s->minCoord, r); s->maxCoord = std::max(s->maxCoord, r); }
	s->diameter = (uint32_t)(s->maxCoord - s->minCoord + 1);
	const uint32_t D = s->diameter;
	std::vector<std::pair<uint64_t, uint32_t>> keyed(n);  // (region z,y,x | cluster-major code, insertion index)
	for (size_t i = 0; i < n; i++)
	{
		uint32_t u[3], l[3];
		for (int a = 0; a < 3; a++) { const int v = s->xyz[3 * i + a], r = floordiv64(v); u[a] = (uint32_t)(r - s->minCoord); l[a] = (uint32_t)(v - r * 64); }
		const uint32_t cid = ((l[0] >> 3) << 6) | ((l[1] >> 3) << 3) | (l[2] >> 3), code = ((l[0] & 7) << 6) | ((l[1] & 7) << 3) | (l[2] & 7);
		keyed[i] = {((uint64_t)(u[0] + u[1] * D + u[2] * D * D) << 18) | (cid << 9) | code, (uint32_t)i};
	}
	std::sort(keyed.begin(), keyed.end());
	s->regionTable.assign((size_t)D * D * D, -1);
	int32_t ri = -1;
	uint64_t lastRegion = ~0ull;
	for (size_t i = 0; i < n; i++)
	{
		if (i + 1 < n && keyed[i + 1].first == keyed[i].first) continue;  // last write wins
		const uint64_t region = keyed[i].first >> 18;
		if (region != lastRegion)
		{
			lastRegion = region; ri++;
			s->regionTable[(size_t)region] = ri;
			s->headers.resize((size_t)(ri + 1) * 512 * 16, uint2{0, 0});
			s->clusterMask.resize((size_t)(ri + 1) * 16, 0u);
		}
		const uint32_t cc = (uint32_t)(keyed[i].first & 0x3FFFFu), cid = cc >> 9;
		uint2& w = s->headers[(size_t)ri * 8192 + (cc >> 5)];
		if (w.x == 0u) w.y = (uint32_t)s->values.size();
		w.x |= 1u << (cc & 31u);
		s->clusterMask[(size_t)ri * 16 + (cid >> 5)] |= 1u << (cid & 31u);
		s->values.push_back(s->rgb[keyed[i].second]);
	}
	s->filled = (uint32_t)(ri + 1);
	for (uint32_t r = 0; r < s->filled; r++)
		for (uint32_t w = 0; w < 8192; w++)
			if ((s->clusterMask[(size_t)r * 16 + (w >> 9)] >> ((w >> 4) & 31)) & 1u) s->headers[(size_t)r * 8192 + w].y |= kHeaderClusterExists;
	s->headers.resize(s->headers.size() + 80 * 16, uint2{0, 0});
	s->clusterMask.resize(s->clusterMask.size() + 32, 0u);
	s->storage = kStorageVcs;
	return 0;
}

// Categories a lane can be in at the top of a pass: 0 region (stored, or leaving the scene), 1 head, 2 main/test, 3 main/jump, 4 main/next, 5 main/cluster, 6 region (null: skip),
// 7 waiting with a hit, 8 done.  hist[signature] += 1 per pass, signature = bit mask of the categories present; lanes[signature][category] += lanes.
// Tiles (8x4 pixels) are sampled every tileStride-th in both directions.  VCS + longest axis only.
extern "C" int sim_warp_profile(void* h, const float* cam, const float* tr, uint32_t scale, uint32_t W, uint32_t H, uint32_t tileStride,
	uint64_t* hist /*512*/, uint64_t* lanes /*512*9*/, uint64_t* totals /*4: tiles, passes, shade passes, lane-passes marching*/, int nThreads)
{
	SimScene* s = static_cast<SimScene*>(h);
	if (s->storage != kStorageVcs) return 1;
	using Ray = FlatRay<kStorageVcs, kAlgoLongestAxis, false>;
	const uint32_t tilesX = (W + 7) / 8, tilesY = (H + 3) / 4;
	std::vector<uint64_t> H0(512, 0), L0(512 * 9, 0);
	uint64_t T[4] = {0, 0, 0, 0};
	#pragma omp parallel num_threads(nThreads)
	{
		std::vector<uint64_t> hl(512, 0), ll(512 * 9, 0);
		uint64_t t[4] = {0, 0, 0, 0};
		RayCtx<kStorageVcs, false> c;
		c.sv = s->view(); c.light = gLight; c.lw = make_light_walk(gLight); c.hitOut = nullptr;
		c.translation[0] = tr[0]; c.translation[1] = tr[1]; c.translation[2] = tr[2];
		c.reset();
		c.skipDead = 1u;
		#pragma omp for schedule(dynamic, 1)
		for (int64_t ty = 0; ty < (int64_t)tilesY; ty += tileStride)
			for (uint32_t tx = 0; tx < tilesX; tx += tileStride)
			{
				Ray ray[32];
				for (int l = 0; l < 32; l++)
				{
					const uint32_t x = tx * 8 + (l & 7), y = (uint32_t)ty * 4 + (l >> 3);
					ray[l].st = kStDone; ray[l].result = 0;
					if (x < W && y < H)
					{
						float o[3], d[3];
						primary_ray_flat(cam, x, y, W, H, 1.0f / (float)W, 1.0f / (float)H, o, d);
						ray[l].start_primary(c, o, d, (float)scale);
					}
				}
				t[0]++;
				for (;;)
				{
					uint32_t sig = 0; int cnt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
					bool marching = false, anyHit = false;
					for (int l = 0; l < 32; l++)
					{
						int cat;
						switch (ray[l].st)
						{
						case kStRegion: cat = ray[l].ri == -1 ? 6 : 0; break;
						case kStHead: cat = 1; break;
						case kAdvNone: cat = 2; break;
						case kAdvJump: cat = 3; break;
						case kAdvNext: cat = 4; break;
						case kAdvCluster: cat = 5; break;
						case kAdvRegion: cat = 6; break;
						case kStHit: cat = 7; anyHit = true; break;
						default: cat = 8; break;
						}
						if (cat <= 6) marching = true;
						if (cat < 8) sig |= 1u << cat;
						cnt[cat]++;
					}
					if (marching)
					{
						bool shadowPass = false;
						for (int l = 0; l < 32; l++) if (ray[l].st <= kStHead && ray[l].shadow()) shadowPass = true;
						if (shadowPass) sig |= 256u;  // (the hit barrier makes a pass all-primary or all-shadow; bit 8 = "done" is not a marching category)
						hl[sig]++; t[1]++;
						for (int k = 0; k < 9; k++) ll[sig * 9 + k] += cnt[k];
						for (int l = 0; l < 32; l++) if (ray[l].st <= kStHead) { ray[l].template step_marching<kPpOff>(c); t[3]++; }
						continue;
					}
					if (!anyHit) break;
					t[2]++;
					for (int l = 0; l < 32; l++) if (ray[l].st == kStHit) ray[l].do_hit(c);
				}
			}
		#pragma omp critical
		{
			for (int i = 0; i < 512; i++) H0[i] += hl[i];
			for (int i = 0; i < 512 * 9; i++) L0[i] += ll[i];
			for (int i = 0; i < 4; i++) T[i] += t[i];
		}
	}
	memcpy(hist, H0.data(), 512 * 8); memcpy(lanes, L0.data(), 512 * 9 * 8); memcpy(totals, T, 32);
	return 0;
}


// Development aid (tools/warp_profile.py --regroup): what would ABANDON + RE-TRACE buy on a frame with long-tailed tiles?  Phase 1: as
// sim_warp_profile, but once a tile has run `budget` passes and at most `maxLanes` lanes are still marching, those lanes are dropped and
// their pixels appended to a list (the lanes waiting with a hit go on).  Phase 2: the listed pixels are traced from scratch, 32 consecutive
// list entries per warp.  out: {tiles, passes phase 1, passes without regrouping, listed pixels, passes phase 2, lane-passes phase 2}
extern "C" int sim_regroup_profile(void* h, const float* cam, const float* tr, uint32_t scale, uint32_t W, uint32_t H, uint32_t tileStride,
	uint32_t budget, uint32_t maxLanes, uint64_t* out, int nThreads)
{
	SimScene* s = static_cast<SimScene*>(h);
	if (s->storage != kStorageVcs) return 1;
	using Ray = FlatRay<kStorageVcs, kAlgoLongestAxis, false>;
	const uint32_t tilesX = (W + 7) / 8, tilesY = (H + 3) / 4;
	uint64_t T[6] = {0, 0, 0, 0, 0, 0};
	std::vector<uint32_t> list;
	auto run_warp = [&](RayCtx<kStorageVcs, false>& c, Ray* ray, uint32_t budgetHere, uint32_t lanesHere, const uint32_t* pix, std::vector<uint32_t>* parked, uint64_t& passes, uint64_t& lanePasses) {
		uint32_t n = 0;
		for (;;)
		{
			int marching = 0; bool anyHit = false;
			for (int l = 0; l < 32; l++) { if (ray[l].st <= kStHead) marching++; if (ray[l].st == kStHit) anyHit = true; }
			if (marching)
			{
				if (parked && n >= budgetHere && (uint32_t)marching <= lanesHere)
				{
					for (int l = 0; l < 32; l++) if (ray[l].st <= kStHead) { parked->push_back(pix[l]); ray[l].st = kStDone; }
					continue;
				}
				n++; passes++;
				for (int l = 0; l < 32; l++) if (ray[l].st <= kStHead) { ray[l].template step_marching<kPpOff>(c); lanePasses++; }
				continue;
			}
			if (!anyHit) break;
			for (int l = 0; l < 32; l++) if (ray[l].st == kStHit) ray[l].do_hit(c);
		}
	};
	#pragma omp parallel num_threads(nThreads)
	{
		uint64_t t[6] = {0, 0, 0, 0, 0, 0};
		std::vector<uint32_t> mine;
		RayCtx<kStorageVcs, false> c;
		c.sv = s->view(); c.light = gLight; c.lw = make_light_walk(gLight); c.hitOut = nullptr;
		c.translation[0] = tr[0]; c.translation[1] = tr[1]; c.translation[2] = tr[2];
		c.reset(); c.skipDead = 1u;
		#pragma omp for schedule(dynamic, 1)
		for (int64_t ty = 0; ty < (int64_t)tilesY; ty += tileStride)
			for (uint32_t tx = 0; tx < tilesX; tx += tileStride)
				for (int variant = 0; variant < 2; variant++)
				{
					Ray ray[32]; uint32_t pix[32];
					for (int l = 0; l < 32; l++)
					{
						const uint32_t x = tx * 8 + (l & 7), y = (uint32_t)ty * 4 + (l >> 3);
						pix[l] = y * W + x;
						ray[l].st = kStDone; ray[l].result = 0;
						if (x < W && y < H) { float o[3], d[3]; primary_ray_flat(cam, x, y, W, H, 1.0f / (float)W, 1.0f / (float)H, o, d); ray[l].start_primary(c, o, d, (float)scale); }
					}
					uint64_t lp = 0;
					if (variant == 0) { t[0]++; run_warp(c, ray, budget, maxLanes, pix, &mine, t[1], lp); }
					else run_warp(c, ray, 0, 0, pix, nullptr, t[2], lp);
				}
		#pragma omp critical
		{
			for (int i = 0; i < 6; i++) T[i] += t[i];
			list.insert(list.end(), mine.begin(), mine.end());
		}
	}
	std::sort(list.begin(), list.end(), [&](uint32_t a, uint32_t b) {  // tile order, as warps would append them
		const uint32_t ax = a % W, ay = a / W, bx = b % W, by = b / W;
		const uint64_t ka = ((uint64_t)(ay / 4) * tilesX + ax / 8) * 32 + (ay % 4) * 8 + ax % 8, kb = ((uint64_t)(by / 4) * tilesX + bx / 8) * 32 + (by % 4) * 8 + bx % 8;
		return ka < kb; });
	T[3] = list.size();
	RayCtx<kStorageVcs, false> c;
	c.sv = s->view(); c.light = gLight; c.lw = make_light_walk(gLight); c.hitOut = nullptr;
	c.translation[0] = tr[0]; c.translation[1] = tr[1]; c.translation[2] = tr[2];
	c.reset(); c.skipDead = 1u;
	for (size_t b = 0; b < list.size(); b += 32)
	{
		Ray ray[32]; uint32_t pix[32];
		for (int l = 0; l < 32; l++)
		{
			ray[l].st = kStDone; ray[l].result = 0; pix[l] = 0;
			if (b + l < list.size()) { const uint32_t x = list[b + l] % W, y = list[b + l] / W; float o[3], d[3]; primary_ray_flat(cam, x, y, W, H, 1.0f / (float)W, 1.0f / (float)H, o, d); ray[l].start_primary(c, o, d, (float)scale); }
		}
		run_warp(c, ray, 0, 0, pix, nullptr, T[4], T[5]);
	}
	memcpy(out, T, sizeof(T));
	return 0;
}


// Development aid (tools/warp_profile.py --compact N): what would packing the SHADOW rays of the N tiles of a CTA into full warps buy (the primary
// phase stays per tile; behind a CTA barrier the live shadow rays of the CTA's tiles are packed 32 per warp in tile order)?
// out: {CTAs, primary passes, shadow passes per tile (today), shadow passes packed, shadow lane-passes, shadow rays, worst-tile primary passes summed over CTAs}
extern "C" int sim_cta_compact_profile(void* h, const float* cam, const float* tr, uint32_t scale, uint32_t W, uint32_t H, uint32_t ctaStride, uint32_t tilesPerCta,
	uint64_t* out, int nThreads)
{
	SimScene* s = static_cast<SimScene*>(h);
	if (s->storage != kStorageVcs || tilesPerCta == 0 || tilesPerCta > 16) return 1;
	using Ray = FlatRay<kStorageVcs, kAlgoLongestAxis, false>;
	const uint32_t ctasX = (W + 8 * tilesPerCta - 1) / (8 * tilesPerCta), ctasY = (H + 3) / 4;
	uint64_t T[7] = {0, 0, 0, 0, 0, 0, 0};
	auto run = [](RayCtx<kStorageVcs, false>& c, Ray* ray, uint64_t& passes, uint64_t& lanePasses) {
		uint64_t n = 0;
		for (;;)
		{
			bool marching = false;
			for (int l = 0; l < 32; l++) if (ray[l].st <= kStHead) marching = true;
			if (!marching) break;
			passes++; n++;
			for (int l = 0; l < 32; l++) if (ray[l].st <= kStHead) { ray[l].template step_marching<kPpOff>(c); lanePasses++; }
		}
		return n;
	};
	#pragma omp parallel num_threads(nThreads)
	{
		uint64_t t[7] = {0, 0, 0, 0, 0, 0, 0};
		RayCtx<kStorageVcs, false> c;
		c.sv = s->view(); c.light = gLight; c.lw = make_light_walk(gLight); c.hitOut = nullptr;
		c.translation[0] = tr[0]; c.translation[1] = tr[1]; c.translation[2] = tr[2];
		c.reset(); c.skipDead = 1u;
		std::vector<Ray> tiles(32 * 16), packed;
		#pragma omp for schedule(dynamic, 1)
		for (int64_t cy = 0; cy < (int64_t)ctasY; cy += ctaStride)
			for (uint32_t cx = 0; cx < ctasX; cx += ctaStride)
			{
				t[0]++;
				uint64_t dummy = 0, worst = 0;
				packed.clear();
				for (uint32_t k = 0; k < tilesPerCta; k++)
				{
					Ray* ray = &tiles[32 * k];
					for (int l = 0; l < 32; l++)
					{
						const uint32_t x = (cx * tilesPerCta + k) * 8 + (l & 7), y = (uint32_t)cy * 4 + (l >> 3);
						ray[l].st = kStDone; ray[l].result = 0;
						if (x < W && y < H) { float o[3], d[3]; primary_ray_flat(cam, x, y, W, H, 1.0f / (float)W, 1.0f / (float)H, o, d); ray[l].start_primary(c, o, d, (float)scale); }
					}
					const uint64_t n = run(c, ray, t[1], dummy);
					if (n > worst) worst = n;
					for (int l = 0; l < 32; l++) if (ray[l].st == kStHit) ray[l].do_hit(c);
					for (int l = 0; l < 32; l++) if (ray[l].st <= kStHead) { packed.push_back(ray[l]); t[5]++; }
					run(c, ray, t[2], t[4]);
				}
				t[6] += worst;
				while (packed.size() % 32) { Ray r; r.st = kStDone; r.result = 0; packed.push_back(r); }
				for (size_t b = 0; b < packed.size(); b += 32) run(c, &packed[b], t[3], dummy);
			}
		#pragma omp critical
		for (int i = 0; i < 7; i++) T[i] += t[i];
	}
	memcpy(out, T, sizeof(T));
	return 0;
}
