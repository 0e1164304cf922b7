/* TEST INFRASTRUCTURE ONLY (oracle/).  Not part of the product; the product never links or calls this.
 *
 * Plain-C restatement ("port") of the reference's hot path: scene partition rules, the two storage
 * structures' OBSERVABLE semantics (key -> colour, cluster occupancy), both traversal algorithms, the
 * shadow rays and the lighting.  Built with  gcc -O2 -ffp-contract=off  (IEEE fp32, no contraction), which
 * SURVEY.md §7 "hard part 1" defines as the canonical arithmetic of the reference.
 *
 * Parity pinning: the reference ships NO golden vectors or tests.  This restatement is pinned
 * differentially against the unmodified reference headers built for the host (oracle/_ref/libvrm_ref_host.so,
 * see oracle/ref_host.cpp) by tests/test_oracle_vs_reference.py, and against the fixtures in tests/golden/
 * that were generated from that reference build (tests/golden/make_golden.py).
 *
 * Every function cites the reference file:line (relative to /root/reference/VoxelRaymarcher/src) it follows.
 * The storage layout here is deliberately naive (sorted key array + bsearch per 64^3 region): only lookup
 * results and cluster occupancy are observable through the StorageStructure seam
 * (storage/StorageStructure.cuh:12-17).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EPSILON 0.0001f               /* geometry/VoxelFunctions.cuh:19 */
#define EMPTY_VAL (1u << 30)          /* geometry/VoxelFunctions.cuh:20-21 */
#define CONTINUE_VAL (EMPTY_VAL + 2u) /* geometry/VoxelFunctions.cuh:23 */
#define BLOCK_SIZE 64                 /* geometry/VoxelFunctions.cuh:24 */
#define PI_F 3.141592f                /* math/MathConstants.cuh:3 */

typedef struct Region
{
	uint32_t n;
	uint32_t* keys; /* sorted */
	uint32_t* vals;
	uint8_t cluster[512];
} Region;

typedef struct Scene
{
	/* staging (insertion order) */
	int32_t* sx; uint32_t* srgb; uint64_t sn, scap;
	/* built */
	Region** table;
	uint32_t diameter;
	int32_t minCoord, maxCoord;
	uint32_t filled;
	int storage; /* -1 = not built, 0 = VCS, 1 = hash table */
} Scene;

/* Main.cu:26-42 / geometry/VoxelFunctions.cuh:27-35: process-wide constants, as in the reference */
static float LIGHT_DIRECTION[3] = {0.57735026f, 0.57735026f, 0.57735026f};
static float LIGHT_COLOR[3] = {1.0f, 1.0f, 1.0f};
static float LIGHT_POSITION[3] = {10.0f, 10.0f, -10.0f};
static const float LIGHT_CONSTANT = 1.0f, LIGHT_LINEAR = 0.045f, LIGHT_QUADRATIC = 0.0075f;
static int USE_POINT_LIGHT = 0;
static int USE_SHADOWS = 1;

typedef struct Ctx
{
	const Scene* scene;
	float translation[3];
	uint32_t scale;
	/* recorder (mirrors oracle/ref_host.cpp Recording) */
	int32_t hit[4];
	uint64_t nExist, nExistFalse, nLookup, nLookupHit;
} Ctx;

/* ---------------------------------------------------------------- vector helpers (math/Vector3.cuh) */

static float v_length(const float* v) { return sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); } /* Vector3.cuh:81 */
static void v_unit(const float* v, float* out)                                                   /* Vector3.cuh:161-165 */
{
	float l = v_length(v);
	out[0] = v[0] / l; out[1] = v[1] / l; out[2] = v[2] / l;
}
static void v_cross(const float* a, const float* b, float* out) /* Vector3.cuh:153-158 */
{
	out[0] = a[1] * b[2] - a[2] * b[1];
	out[1] = -(a[0] * b[2] - a[2] * b[0]);
	out[2] = a[0] * b[1] - a[1] * b[0];
}
/* o += t * d  with the reference's operand order  origin + (t * direction)  (Vector3.cuh:105-109,134-138) */
static void advance(float* o, const float* d, float t)
{
	o[0] = o[0] + t * d[0]; o[1] = o[1] + t * d[1]; o[2] = o[2] + t * d[2];
}

/* ---------------------------------------------------------------- storage semantics */

static uint32_t make_key(uint32_t x, uint32_t y, uint32_t z) { return (x << 20) | (y << 10) | z; } /* VoxelFunctions.cuh:41-46 */
static int cluster_id(uint32_t x, uint32_t y, uint32_t z) { return (int)(((x / 8) << 6) | ((y / 8) << 3) | (z / 8)); } /* VoxelClusterStore.cuh:21-24 */

/* StorageStructure::doesVoxelSpaceExist: always true for the hash table (StorageStructure.cuh:49-52),
 * cluster pointer non-null for the VCS (VoxelClusterStore.cuh:93-99). */
static int space_exists(Ctx* c, const Region* r, int32_t x, int32_t y, int32_t z)
{
	int e = c->scene->storage == 1 ? 1 : r->cluster[cluster_id((uint32_t)x, (uint32_t)y, (uint32_t)z)];
	c->nExist++;
	if (!e) c->nExistFalse++;
	return e;
}

/* StorageStructure::lookupVoxel: key -> colour or EMPTY (CuckooHashTable.cuh:59-76, VoxelClusterStore.cuh:101-135) */
static uint32_t region_find(const Region* r, uint32_t key)
{
	int64_t lo = 0, hi = (int64_t)r->n - 1;
	while (lo <= hi)
	{
		int64_t mid = lo + (hi - lo) / 2;
		if (r->keys[mid] == key) return r->vals[mid];
		if (r->keys[mid] < key) lo = mid + 1; else hi = mid - 1;
	}
	return EMPTY_VAL;
}

static uint32_t lookup_voxel(Ctx* c, const Region* r, const int* creg, int32_t x, int32_t y, int32_t z)
{
	uint32_t v = region_find(r, make_key((uint32_t)x, (uint32_t)y, (uint32_t)z));
	c->nLookup++;
	if (v != EMPTY_VAL)
	{
		c->nLookupHit++;
		if (!c->hit[3])
		{
			c->hit[0] = creg[0] * 64 + x; c->hit[1] = creg[1] * 64 + y; c->hit[2] = creg[2] * 64 + z; c->hit[3] = 1;
		}
	}
	return v;
}

/* renderer/Renderer.cuh:29-44 */
static int in_scene(const Scene* s, const int* r)
{
	uint32_t ux = (uint32_t)(r[0] - s->minCoord), uy = (uint32_t)(r[1] - s->minCoord), uz = (uint32_t)(r[2] - s->minCoord);
	return ux < s->diameter && uy < s->diameter && uz < s->diameter;
}
static const Region* region_at(const Scene* s, const int* r)
{
	uint32_t ux = (uint32_t)(r[0] - s->minCoord), uy = (uint32_t)(r[1] - s->minCoord), uz = (uint32_t)(r[2] - s->minCoord);
	return s->table[ux + uy * s->diameter + uz * s->diameter * s->diameter];
}

/* ---------------------------------------------------------------- lighting (Renderer.cuh:57-86,237-258) */

static uint32_t vec_to_rgb(const float* c) /* VoxelFunctions.cuh:76-82 */
{
	uint32_t r = (uint32_t)(c[0] * 255.0f), g = (uint32_t)(c[1] * 255.0f), b = (uint32_t)(c[2] * 255.0f);
	return (r << 16) | (g << 8) | b;
}
static void rgb_to_vec(uint32_t color, float* out) /* VoxelFunctions.cuh:54-74 */
{
	out[0] = (color >> 16) / 255.0f; out[1] = ((color >> 8) & 0xFF) / 255.0f; out[2] = (color & 0xFF) / 255.0f;
}

static uint32_t apply_lighting(uint32_t voxelColor, const float* normal, const float* regionWorld, const float* rayOrigin)
{
	float color[3], diffuse[3], out[3];
	rgb_to_vec(voxelColor, color);
	if (USE_POINT_LIGHT) /* Renderer.cuh:68-86 */
	{
		float hit[3] = {regionWorld[0] + rayOrigin[0], regionWorld[1] + rayOrigin[1], regionWorld[2] + rayOrigin[2]};
		float toLight[3] = {LIGHT_POSITION[0] - hit[0], LIGHT_POSITION[1] - hit[1], LIGHT_POSITION[2] - hit[2]};
		float distance = v_length(toLight);
		float lightDir[3];
		v_unit(toLight, lightDir);
		float attenuation = 1 / (LIGHT_CONSTANT + LIGHT_LINEAR * distance + LIGHT_QUADRATIC * (distance * distance));
		float diff = fmaxf(normal[0] * lightDir[0] + normal[1] * lightDir[1] + normal[2] * lightDir[2], 0.0f);
		for (int i = 0; i < 3; i++) { diffuse[i] = diff * LIGHT_COLOR[i]; out[i] = (attenuation * diffuse[i]) * color[i]; }
		return vec_to_rgb(out);
	}
	/* Renderer.cuh:57-66 */
	float diff = fmaxf(normal[0] * LIGHT_DIRECTION[0] + normal[1] * LIGHT_DIRECTION[1] + normal[2] * LIGHT_DIRECTION[2], 0.0f);
	for (int i = 0; i < 3; i++) { diffuse[i] = diff * LIGHT_COLOR[i]; out[i] = color[i] * diffuse[i]; }
	return vec_to_rgb(out);
}

static void normal_from_t(float tX, float tY, float tZ, float tMin, const float* d, float* n) /* Renderer.cuh:237-247 */
{
	n[0] = n[1] = n[2] = 0.0f;
	if (tX == tMin) n[0] = copysignf(1.0f, -d[0]);
	else if (tY == tMin) n[1] = copysignf(1.0f, -d[1]);
	else n[2] = copysignf(1.0f, -d[2]);
}

/* ---------------------------------------------------------------- helpers shared by the walks */

static int ray_in_region(const float* o) /* Renderer.cuh:93-98 */
{
	return o[0] >= 0.0f && o[0] < BLOCK_SIZE && o[1] >= 0.0f && o[1] < BLOCK_SIZE && o[2] >= 0.0f && o[2] < BLOCK_SIZE;
}
static int grid_in_region(int32_t x, int32_t y, int32_t z) /* Renderer.cuh:436-439 (uint32 compare) */
{
	return (uint32_t)x < BLOCK_SIZE && (uint32_t)y < BLOCK_SIZE && (uint32_t)z < BLOCK_SIZE;
}
static float next_edge(float dir, float v) /* Renderer.cuh:47-55,263-265 */
{
	return dir > 0.0f ? ceilf(v) + EPSILON : floorf(v) - EPSILON;
}
/* (next - o) / d, optionally with the shadow routine's zero-direction guard (Renderer.cuh:113-115) */
static float t_to(float next, float o, float d, int guard)
{
	if (guard && !(d != 0.0f)) return INFINITY;
	return (next - o) / d;
}
static float min3(float a, float b, float c) { return fminf(a, fminf(b, c)); } /* CUDA min(float,float) == fminf */

/* Move to the next region after a region walk ended (Renderer.cuh:421-429 and its twins) */
static void rebase_region(float* o, int* creg)
{
	int32_t dx = (int32_t)floorf(o[0] / BLOCK_SIZE), dy = (int32_t)floorf(o[1] / BLOCK_SIZE), dz = (int32_t)floorf(o[2] / BLOCK_SIZE);
	creg[0] += dx; creg[1] += dy; creg[2] += dz;
	o[0] = (o[0] - (float)(dx * BLOCK_SIZE)) * 1.0f;
	o[1] = (o[1] - (float)(dy * BLOCK_SIZE)) * 1.0f;
	o[2] = (o[2] - (float)(dz * BLOCK_SIZE)) * 1.0f;
}

/* Null-region skip (Renderer.cuh:384-410; guarded variant 185-211).  Returns 0 when the ray left the scene. */
static const Region* skip_null_regions(const Scene* s, float* o, const float* d, int* creg, const Region* r, int guard, int* left)
{
	*left = 0;
	while (r == NULL)
	{
		float nX = d[0] > 0.0f ? BLOCK_SIZE + EPSILON : 0.0f - EPSILON;
		float nY = d[1] > 0.0f ? BLOCK_SIZE + EPSILON : 0.0f - EPSILON;
		float nZ = d[2] > 0.0f ? BLOCK_SIZE + EPSILON : 0.0f - EPSILON;
		float tMin = min3(t_to(nX, o[0], d[0], guard), t_to(nY, o[1], d[1], guard), t_to(nZ, o[2], d[2], guard));
		advance(o, d, tMin);
		rebase_region(o, creg);
		if (in_scene(s, creg)) r = region_at(s, creg);
		else { *left = 1; return NULL; }
	}
	return r;
}

static void region_world(const Ctx* c, const int* creg, float* out) /* Renderer.cuh:413 */
{
	out[0] = c->translation[0] + (float)(creg[0] * BLOCK_SIZE);
	out[1] = c->translation[1] + (float)(creg[1] * BLOCK_SIZE);
	out[2] = c->translation[2] + (float)(creg[2] * BLOCK_SIZE);
}

/* ---------------------------------------------------------------- "original" traversal */

/* shadowRayMarchVoxelGrid, Renderer.cuh:100-172 (guard = 1) and the step structure shared with
 * rayMarchVoxelGrid, Renderer.cuh:260-336 (guard = 0).  Returns the raw voxel colour or EMPTY_VAL;
 * outT receives the OUTER tX,tY,tZ,tMin (stale after a cluster skip, exactly as in the reference). */
static uint32_t march_grid_steps(Ctx* c, float* o, const float* d, const Region* r, const int* creg, int guard, float* outT)
{
	float tX = t_to(next_edge(d[0], o[0]), o[0], d[0], guard);
	float tY = t_to(next_edge(d[1], o[1]), o[1], d[1], guard);
	float tZ = t_to(next_edge(d[2], o[2]), o[2], d[2], guard);
	float tMin = min3(tX, tY, tZ);
	advance(o, d, tMin + EPSILON);
	while (ray_in_region(o))
	{
		int32_t vx = (int32_t)o[0], vy = (int32_t)o[1], vz = (int32_t)o[2];
		if (!space_exists(c, r, vx, vy, vz))
		{
			int32_t nX = d[0] > 0.0f ? ((vx / 8) + 1) * 8 : (vx / 8) * 8;
			int32_t nY = d[1] > 0.0f ? ((vy / 8) + 1) * 8 : (vy / 8) * 8;
			int32_t nZ = d[2] > 0.0f ? ((vz / 8) + 1) * 8 : (vz / 8) * 8;
			float sMin = min3(t_to((float)nX, o[0], d[0], guard), t_to((float)nY, o[1], d[1], guard), t_to((float)nZ, o[2], d[2], guard));
			advance(o, d, sMin + EPSILON);
			continue;
		}
		uint32_t col = lookup_voxel(c, r, creg, vx, vy, vz);
		if (col != EMPTY_VAL)
		{
			if (outT) { outT[0] = tX; outT[1] = tY; outT[2] = tZ; outT[3] = tMin; }
			return col;
		}
		tX = t_to(next_edge(d[0], o[0]), o[0], d[0], guard);
		tY = t_to(next_edge(d[1], o[1]), o[1], d[1], guard);
		tZ = t_to(next_edge(d[2], o[2]), o[2], d[2], guard);
		tMin = min3(tX, tY, tZ);
		advance(o, d, tMin + EPSILON);
	}
	return EMPTY_VAL;
}

/* isInShadowOriginalRayMarch, Renderer.cuh:174-235 */
static int in_shadow_original(Ctx* c, const float* origin, const int* region)
{
	if (!USE_SHADOWS) return 0;
	const Scene* s = c->scene;
	float o[3] = {origin[0], origin[1], origin[2]};
	const float* d = LIGHT_DIRECTION;
	int creg[3] = {region[0], region[1], region[2]};
	while (in_scene(s, creg))
	{
		int left;
		const Region* r = skip_null_regions(s, o, d, creg, region_at(s, creg), 1, &left);
		if (left) return 0;
		if (march_grid_steps(c, o, d, r, creg, 1, NULL) != EMPTY_VAL) return 1;
		rebase_region(o, creg);
	}
	return 0;
}

/* rayMarchVoxelGrid, Renderer.cuh:260-336 */
static uint32_t march_grid_original(Ctx* c, float* o, const float* d, const Region* r, const int* creg)
{
	float t[4];
	uint32_t col = march_grid_steps(c, o, d, r, creg, 0, t);
	if (col == EMPTY_VAL) return EMPTY_VAL;
	float normal[3], world[3];
	normal_from_t(t[0], t[1], t[2], t[3], d, normal);
	region_world(c, creg, world);
	uint32_t lit = apply_lighting(col, normal, world, o);
	return lit * (uint32_t)!in_shadow_original(c, o, creg);
}

/* ---------------------------------------------------------------- "longest axis" traversal */

/* Ray::convertRayToLongestAxisDirection, rays/Ray.cuh:19-71 (strict > tie rules) */
static void rank_axes(const float* d, int* L, int* M, int* S, float* scaled)
{
	float ax = fabsf(d[0]), ay = fabsf(d[1]), az = fabsf(d[2]), k;
	if (ax > ay && ax > az) { *L = 0; if (ay > az) { *M = 1; *S = 2; } else { *M = 2; *S = 1; } k = 1.0f / ax; }
	else if (ay > az)       { *L = 1; if (ax > az) { *M = 0; *S = 2; } else { *M = 2; *S = 0; } k = 1.0f / ay; }
	else                    { *L = 2; if (ax > ay) { *M = 0; *S = 1; } else { *M = 1; *S = 0; } k = 1.0f / az; }
	scaled[0] = k * d[0]; scaled[1] = k * d[1]; scaled[2] = k * d[2];
}

static void local_hit_location(const float* oo, const float* od, int axis, float* out) /* Renderer.cuh:753-758 */
{
	float t = od[axis] > 0.0f ? (ceilf(oo[axis]) - oo[axis]) / od[axis] : (floorf(oo[axis]) - oo[axis]) / od[axis];
	out[0] = oo[0] + t * od[0]; out[1] = oo[1] + t * od[1]; out[2] = oo[2] + t * od[2];
}

static int in_shadow_longest_axis(Ctx* c, const float* origin, const int* region);

typedef struct LAState
{
	float oo[3], od[3]; /* oldRay: origin, longest-axis-scaled direction */
	float ro[3];        /* ray origin (direction == od) */
	int32_t g[3], ad[3];
	int L, M, S;
} LAState;

/* performVoxelSpaceJump (Renderer.cuh:696-751) and performShadowVoxelSpaceJump (Renderer.cuh:441-492) */
static uint32_t voxel_space_jump(Ctx* c, LAState* st, float* origO, const Region* r, const int* creg, int shadow)
{
	float tX = 0.0f, tY = 0.0f, tZ = 0.0f, tMin = 0.0f;
	while (!space_exists(c, r, st->g[0], st->g[1], st->g[2]))
	{
		int32_t nX = st->od[0] > 0.0f ? ((st->g[0] / 8) + 1) * 8 : (st->g[0] / 8) * 8;
		int32_t nY = st->od[1] > 0.0f ? ((st->g[1] / 8) + 1) * 8 : (st->g[1] / 8) * 8;
		int32_t nZ = st->od[2] > 0.0f ? ((st->g[2] / 8) + 1) * 8 : (st->g[2] / 8) * 8;
		tX = ((float)nX - st->oo[0]) / st->od[0];
		tY = ((float)nY - st->oo[1]) / st->od[1];
		tZ = ((float)nZ - st->oo[2]) / st->od[2];
		tMin = min3(tX, tY, tZ) + EPSILON;
		advance(st->oo, st->od, tMin);
		st->g[0] = (int32_t)floorf(st->oo[0]); st->g[1] = (int32_t)floorf(st->oo[1]); st->g[2] = (int32_t)floorf(st->oo[2]);
		if (!grid_in_region(st->g[0], st->g[1], st->g[2]))
		{
			origO[0] = st->oo[0]; origO[1] = st->oo[1]; origO[2] = st->oo[2];
			return EMPTY_VAL;
		}
	}
	uint32_t col = lookup_voxel(c, r, creg, st->g[0], st->g[1], st->g[2]);
	if (col != EMPTY_VAL)
	{
		if (shadow) return col;
		float normal[3], world[3];
		normal_from_t(tX, tY, tZ, tMin, st->od, normal); /* tMin already carries +EPSILON: Renderer.cuh:716,736 */
		region_world(c, creg, world);
		return apply_lighting(col, normal, world, st->oo) * (uint32_t)!in_shadow_longest_axis(c, st->oo, creg);
	}
	int L = st->L;
	float tNext = st->od[L] > 0.0f ? (ceilf(st->oo[L]) - st->oo[L]) / st->od[L] : (floorf(st->oo[L]) - st->oo[L]) / st->od[L];
	float tt = tNext + EPSILON;
	st->ro[0] = st->oo[0] + tt * st->od[0]; st->ro[1] = st->oo[1] + tt * st->od[1]; st->ro[2] = st->oo[2] + tt * st->od[2];
	st->ad[st->M] = (int32_t)st->ro[st->M] - st->g[st->M];
	st->ad[st->S] = (int32_t)st->ro[st->S] - st->g[st->S];
	return CONTINUE_VAL;
}

/* One "bump grid, test space, look up" unit of Renderer.cuh:807-823 etc.  Returns 1 when *result is final,
 * 2 when the caller must `continue` the while loop, 0 to fall through. */
static int la_test_axis(Ctx* c, LAState* st, float* origO, const Region* r, const int* creg, int shadow, int axis, int isLongest, uint32_t* result)
{
	st->g[axis] += st->ad[axis];
	if (!space_exists(c, r, st->g[0], st->g[1], st->g[2]))
	{
		uint32_t j = voxel_space_jump(c, st, origO, r, creg, shadow);
		if (j != CONTINUE_VAL) { *result = j; return 1; }
		return 2;
	}
	uint32_t col = lookup_voxel(c, r, creg, st->g[0], st->g[1], st->g[2]);
	if (col != EMPTY_VAL)
	{
		if (shadow) { *result = col; return 1; }
		float normal[3] = {0.0f, 0.0f, 0.0f}, world[3], hitLoc[3];
		normal[axis] = copysignf(1.0f, -st->od[axis]);
		if (isLongest) { hitLoc[0] = st->ro[0]; hitLoc[1] = st->ro[1]; hitLoc[2] = st->ro[2]; } /* Renderer.cuh:899 */
		else local_hit_location(st->oo, st->od, axis, hitLoc);                                   /* Renderer.cuh:820 */
		region_world(c, creg, world);
		*result = apply_lighting(col, normal, world, hitLoc) * (uint32_t)!in_shadow_longest_axis(c, hitLoc, creg);
		return 1;
	}
	return 0;
}

/* rayMarchVoxelGridLongestAxis (Renderer.cuh:760-915) / shadowRayMarchVoxelGridLongestAxis (Renderer.cuh:495-631) */
static uint32_t march_grid_longest_axis(Ctx* c, float* o, const float* d, const Region* r, const int* creg, int shadow)
{
	LAState st;
	rank_axes(d, &st.L, &st.M, &st.S, st.od);
	int L = st.L, M = st.M, S = st.S;
	st.oo[0] = o[0]; st.oo[1] = o[1]; st.oo[2] = o[2];
	st.g[0] = (int32_t)o[0]; st.g[1] = (int32_t)o[1]; st.g[2] = (int32_t)o[2];
	st.ad[0] = st.ad[1] = st.ad[2] = 0;
	st.ad[L] = d[L] < 0.0f ? -1 : 1;
	float t = st.ad[L] > 0 ? ((float)st.g[L] + EPSILON + 1 - o[L]) / (float)st.ad[L]
	                       : ((float)st.g[L] - EPSILON - o[L]) / (float)st.ad[L];
	st.ro[0] = st.oo[0] + t * st.od[0]; st.ro[1] = st.oo[1] + t * st.od[1]; st.ro[2] = st.oo[2] + t * st.od[2];
	st.ad[M] = (int32_t)st.ro[M] - st.g[M];
	st.ad[S] = (int32_t)st.ro[S] - st.g[S];
	int roundDown = st.od[M] < 0.0f;

	while (grid_in_region(st.g[L] + st.ad[L], st.g[M] + st.ad[M], st.g[S] + st.ad[S]))
	{
		uint32_t result = 0;
		int rc = 0;
		if (st.ad[S] != 0 && st.ad[M] != 0)
		{
			float rounded = roundDown ? floorf(st.oo[M]) : ceilf(st.oo[M]);
			float t1 = (rounded - st.oo[M]) / st.od[M];
			float shortestPosition = st.oo[S] + st.od[S] * t1;
			int32_t shorterDiff = (int32_t)floorf(shortestPosition) - st.g[S];
			int first = M, second = S;
			if (shorterDiff != 0) { first = S; second = M; }
			rc = la_test_axis(c, &st, o, r, creg, shadow, first, 0, &result);
			if (rc == 1) return result;
			if (rc == 2) continue;
			rc = la_test_axis(c, &st, o, r, creg, shadow, second, 0, &result);
			if (rc == 1) return result;
			if (rc == 2) continue;
		}
		else if (st.ad[M] != 0)
		{
			rc = la_test_axis(c, &st, o, r, creg, shadow, M, 0, &result);
			if (rc == 1) return result;
			if (rc == 2) continue;
		}
		else if (st.ad[S] != 0)
		{
			rc = la_test_axis(c, &st, o, r, creg, shadow, S, 0, &result);
			if (rc == 1) return result;
			if (rc == 2) continue;
		}
		rc = la_test_axis(c, &st, o, r, creg, shadow, L, 1, &result);
		if (rc == 1) return result;
		if (rc == 2) continue;

		st.oo[0] = st.ro[0]; st.oo[1] = st.ro[1]; st.oo[2] = st.ro[2];
		st.ro[0] = st.ro[0] + st.od[0]; st.ro[1] = st.ro[1] + st.od[1]; st.ro[2] = st.ro[2] + st.od[2];
		st.ad[M] = (int32_t)st.ro[M] - st.g[M];
		st.ad[S] = (int32_t)st.ro[S] - st.g[S];
	}
	/* Renderer.cuh:911-914 / 627-630: finish the region with the original algorithm from oldRay's origin */
	o[0] = st.oo[0]; o[1] = st.oo[1]; o[2] = st.oo[2];
	if (shadow) return march_grid_steps(c, o, d, r, creg, 1, NULL);
	return march_grid_original(c, o, d, r, creg);
}

/* isInShadowRayMarchVoxelSceneLongestAxis, Renderer.cuh:633-694 (no zero-direction guards) */
static int in_shadow_longest_axis(Ctx* c, const float* origin, const int* region)
{
	if (!USE_SHADOWS) return 0;
	const Scene* s = c->scene;
	float o[3] = {origin[0], origin[1], origin[2]};
	const float* d = LIGHT_DIRECTION;
	int creg[3] = {region[0], region[1], region[2]};
	while (in_scene(s, creg))
	{
		int left;
		const Region* r = skip_null_regions(s, o, d, creg, region_at(s, creg), 0, &left);
		if (left) return 0;
		if (march_grid_longest_axis(c, o, d, r, creg, 1) != EMPTY_VAL) return 1;
		rebase_region(o, creg);
	}
	return 0;
}

/* ---------------------------------------------------------------- scene walk */

/* rayMarchVoxelScene (Renderer.cuh:338-434) / rayMarchVoxelSceneLongestAxis (Renderer.cuh:917-1010).
 * algorithm: 0 = longest axis, 1 = original (Main.cu:58-68). */
static uint32_t march_scene(Ctx* c, const float* worldO, const float* d, int algorithm)
{
	const Scene* s = c->scene;
	float sc = (float)c->scale;
	float o[3] = {(worldO[0] - c->translation[0]) * sc, (worldO[1] - c->translation[1]) * sc, (worldO[2] - c->translation[2]) * sc}; /* Ray.cuh:14-17 */
	int creg[3] = {(int32_t)floorf(o[0] / BLOCK_SIZE), (int32_t)floorf(o[1] / BLOCK_SIZE), (int32_t)floorf(o[2] / BLOCK_SIZE)};
	int32_t minC = s->minCoord;
	uint32_t D = s->diameter;
	/* scene-entry loop, Renderer.cuh:349-373 (mixed signed/unsigned compares kept) */
	while (creg[0] - minC < 0 || creg[1] - minC < 0 || creg[2] - minC < 0 ||
	       (uint32_t)(creg[0] - minC) > D - 1 || (uint32_t)(creg[1] - minC) > D - 1 || (uint32_t)(creg[2] - minC) > D - 1)
	{
		int32_t nX = d[0] < 0.0f ? (int32_t)(D + (uint32_t)minC) : 0 + minC;
		int32_t nY = d[1] < 0.0f ? (int32_t)(D + (uint32_t)minC) : 0 + minC;
		int32_t nZ = d[2] < 0.0f ? (int32_t)(D + (uint32_t)minC) : 0 + minC;
		float tX = ((float)(nX * BLOCK_SIZE) - o[0]) / d[0];
		float tY = ((float)(nY * BLOCK_SIZE) - o[1]) / d[1];
		float tZ = ((float)(nZ * BLOCK_SIZE) - o[2]) / d[2];
		if (tX <= 0.0f) tX = INFINITY;
		if (tY <= 0.0f) tY = INFINITY;
		if (tZ <= 0.0f) tZ = INFINITY;
		float tMin = min3(tX, tY, tZ);
		if (tMin == INFINITY) return 0;
		advance(o, d, tMin + EPSILON);
		creg[0] = (int32_t)floorf(o[0] / BLOCK_SIZE); creg[1] = (int32_t)floorf(o[1] / BLOCK_SIZE); creg[2] = (int32_t)floorf(o[2] / BLOCK_SIZE);
	}
	/* to region-local coordinates, Renderer.cuh:376-378 */
	o[0] = (o[0] - (float)(creg[0] * BLOCK_SIZE)) * 1.0f;
	o[1] = (o[1] - (float)(creg[1] * BLOCK_SIZE)) * 1.0f;
	o[2] = (o[2] - (float)(creg[2] * BLOCK_SIZE)) * 1.0f;
	while (in_scene(s, creg))
	{
		int left;
		const Region* r = skip_null_regions(s, o, d, creg, region_at(s, creg), 0, &left);
		if (left) return 0;
		uint32_t col = algorithm == 1 ? march_grid_original(c, o, d, r, creg) : march_grid_longest_axis(c, o, d, r, creg, 0);
		if (col != EMPTY_VAL) return col;
		rebase_region(o, creg);
	}
	return 0;
}

/* ---------------------------------------------------------------- scene build */

typedef struct Staged { int32_t reg[3]; uint32_t key; uint32_t val; uint64_t order; } Staged;

static int cmp_staged(const void* a, const void* b)
{
	const Staged* p = (const Staged*)a; const Staged* q = (const Staged*)b;
	for (int i = 2; i >= 0; i--) if (p->reg[i] != q->reg[i]) return p->reg[i] < q->reg[i] ? -1 : 1;
	if (p->key != q->key) return p->key < q->key ? -1 : 1;
	return p->order < q->order ? -1 : (p->order > q->order);
}

static void ctx_init(Ctx* c, const Scene* s, const float* translation, uint32_t scale)
{
	memset(c, 0, sizeof(*c));
	c->scene = s;
	c->translation[0] = translation[0]; c->translation[1] = translation[1]; c->translation[2] = translation[2];
	c->scale = scale;
}

static void ctx_reset(Ctx* c)
{
	c->hit[0] = c->hit[1] = c->hit[2] = c->hit[3] = 0;
	c->nExist = c->nExistFalse = c->nLookup = c->nLookupHit = 0;
}

/* ================================================================ C ABI (same entry points as oracle/ref_host.cpp) */

void* orc_scene_create(void)
{
	Scene* s = (Scene*)calloc(1, sizeof(Scene));
	s->storage = -1;
	return s;
}

void orc_scene_destroy(void* h)
{
	Scene* s = (Scene*)h;
	if (!s) return;
	if (s->table)
	{
		uint32_t size = s->diameter * s->diameter * s->diameter;
		for (uint32_t i = 0; i < size; i++)
			if (s->table[i]) { free(s->table[i]->keys); free(s->table[i]->vals); free(s->table[i]); }
		free(s->table);
	}
	free(s->sx); free(s->srgb); free(s);
}

void orc_scene_add_voxels(void* h, const int32_t* xyz, const uint32_t* rgb, uint64_t n)
{
	Scene* s = (Scene*)h;
	if (s->sn + n > s->scap)
	{
		s->scap = (s->sn + n) * 2;
		s->sx = (int32_t*)realloc(s->sx, s->scap * 3 * sizeof(int32_t));
		s->srgb = (uint32_t*)realloc(s->srgb, s->scap * sizeof(uint32_t));
	}
	memcpy(s->sx + 3 * s->sn, xyz, n * 3 * sizeof(int32_t));
	memcpy(s->srgb + s->sn, rgb, n * sizeof(uint32_t));
	s->sn += n;
}

/* VoxelSceneCPU::insertVoxel (geometry/VoxelSceneCPU.cuh:16-46) + generateVoxelScene (:49-93) */
int orc_scene_build(void* h, int storageType)
{
	Scene* s = (Scene*)h;
	if (s->storage != -1) return 1;
	Staged* st = (Staged*)malloc((s->sn ? s->sn : 1) * sizeof(Staged));
	s->minCoord = 0; s->maxCoord = 0; /* VoxelSceneCPU.cuh:129-130 */
	for (uint64_t i = 0; i < s->sn; i++)
	{
		uint32_t l[3];
		for (int a = 0; a < 3; a++)
		{
			int32_t v = s->sx[3 * i + a];
			st[i].reg[a] = (int32_t)floorf(v / (float)BLOCK_SIZE);       /* VoxelSceneCPU.cuh:19-21 */
			l[a] = (uint32_t)(((v % BLOCK_SIZE) + BLOCK_SIZE) % BLOCK_SIZE); /* VoxelSceneCPU.cuh:24-26 */
			if (st[i].reg[a] < s->minCoord) s->minCoord = st[i].reg[a];  /* VoxelSceneCPU.cuh:28-35 */
			if (st[i].reg[a] > s->maxCoord) s->maxCoord = st[i].reg[a];
		}
		st[i].key = make_key(l[0], l[1], l[2]);
		st[i].val = s->srgb[i];
		st[i].order = i;
	}
	qsort(st, s->sn, sizeof(Staged), cmp_staged);
	s->diameter = (uint32_t)(s->maxCoord - s->minCoord + 1);
	uint32_t size = s->diameter * s->diameter * s->diameter;
	s->table = (Region**)calloc(size, sizeof(Region*));
	s->filled = 0;
	uint64_t i = 0;
	while (i < s->sn)
	{
		uint64_t j = i;
		while (j < s->sn && st[j].reg[0] == st[i].reg[0] && st[j].reg[1] == st[i].reg[1] && st[j].reg[2] == st[i].reg[2]) j++;
		Region* r = (Region*)calloc(1, sizeof(Region));
		r->keys = (uint32_t*)malloc((j - i) * sizeof(uint32_t));
		r->vals = (uint32_t*)malloc((j - i) * sizeof(uint32_t));
		for (uint64_t k = i; k < j; k++)
		{
			/* last write wins (VoxelSceneCPU.cuh:45): equal keys are adjacent, ordered by insertion */
			if (k + 1 < j && st[k + 1].key == st[k].key) continue;
			r->keys[r->n] = st[k].key; r->vals[r->n] = st[k].val; r->n++;
			r->cluster[cluster_id(st[k].key >> 20, (st[k].key >> 10) & 0x3FF, st[k].key & 0x3FF)] = 1; /* VoxelClusterStore.cuh:26-32 */
		}
		uint32_t ux = (uint32_t)(st[i].reg[0] - s->minCoord), uy = (uint32_t)(st[i].reg[1] - s->minCoord), uz = (uint32_t)(st[i].reg[2] - s->minCoord);
		s->table[ux + uy * s->diameter + uz * s->diameter * s->diameter] = r; /* VoxelSceneCPU.cuh:61-62 */
		s->filled++;
		i = j;
	}
	free(st);
	s->storage = storageType;
	return 0;
}

void orc_scene_info(void* h, uint32_t* diameter, int32_t* minCoord, uint32_t* filled)
{
	Scene* s = (Scene*)h;
	*diameter = s->diameter; *minCoord = s->minCoord; *filled = s->filled;
}

void orc_set_lighting(const float* dir, const float* color, const float* pos, int usePoint, int useShadows)
{
	memcpy(LIGHT_DIRECTION, dir, 12); memcpy(LIGHT_COLOR, color, 12); memcpy(LIGHT_POSITION, pos, 12);
	USE_POINT_LIGHT = usePoint != 0; USE_SHADOWS = useShadows != 0;
}

void orc_make_unit_vector(const float* v, float* out) { v_unit(v, out); }

/* Camera::Camera, renderer/camera/Camera.cuh:11-23 */
void orc_camera_make(const float* origin, const float* lookAt, const float* up, float fov, float aspect, float* out)
{
	float halfHeight = tanf((fov * PI_F / 180.f) / 2.0f);
	float halfWidth = halfHeight * aspect;
	float dir[3] = {lookAt[0] - origin[0], lookAt[1] - origin[1], lookAt[2] - origin[2]}, w[3], wxup[3], u[3], v[3];
	v_unit(dir, w);
	v_cross(w, up, wxup);
	v_unit(wxup, u);
	v_cross(u, w, v);
	for (int i = 0; i < 3; i++)
	{
		out[i] = origin[i];
		out[3 + i] = origin[i] - halfWidth * u[i] - halfHeight * v[i] + w[i];
		out[6 + i] = (2 * halfWidth) * u[i];
		out[9 + i] = (2 * halfHeight) * v[i];
		out[12 + i] = w[i];
	}
}

/* calculateWorldRay (Renderer.cuh:1013-1022) + Camera::generateRay (Camera.cuh:25-29) */
static void primary_ray(const float* cam, uint32_t x, uint32_t y, uint32_t W, uint32_t H, float* o, float* d)
{
	float u = ((float)x + 0.5f) / (float)W;
	float v = ((float)(H - y) + 0.5f) / (float)H;
	float rel[3];
	for (int i = 0; i < 3; i++)
	{
		o[i] = cam[3 + i] + u * cam[6 + i] + v * cam[9 + i];
		rel[i] = o[i] - cam[i];
	}
	v_unit(rel, d);
}

static void fold_counters(uint64_t* counters, const Ctx* c)
{
	counters[0] += c->nExist; counters[1] += c->nExistFalse; counters[2] += c->nLookup; counters[3] += c->nLookupHit;
	if (c->nLookup > counters[4]) counters[4] = c->nLookup;
}

int orc_render(void* h, const float* camera15, const float* translation, uint32_t scale, int algorithm,
	uint32_t width, uint32_t height, uint8_t* rgb, int32_t* hits, uint64_t* counters, uint32_t* lookupsPerPixel, int nThreads)
{
	Scene* s = (Scene*)h;
	if (s->storage == -1) return 1;
	if (nThreads < 1) nThreads = 1;
	uint64_t total[5] = {0, 0, 0, 0, 0};
	#pragma omp parallel num_threads(nThreads)
	{
		Ctx c;
		ctx_init(&c, s, translation, scale);
		uint64_t local[5] = {0, 0, 0, 0, 0};
		#pragma omp for schedule(dynamic, 1)
		for (int64_t y = 0; y < (int64_t)height; y++)
		{
			for (uint32_t x = 0; x < width; x++)
			{
				float o[3], d[3];
				ctx_reset(&c);
				primary_ray(camera15, x, (uint32_t)y, width, height, o, d);
				uint32_t color = march_scene(&c, o, d, algorithm);
				size_t p = (size_t)y * width + x;
				rgb[3 * p] = (uint8_t)(color >> 16); rgb[3 * p + 1] = (uint8_t)((color >> 8) & 0xFF); rgb[3 * p + 2] = (uint8_t)(color & 0xFF); /* Renderer.cuh:1024-1031 */
				if (hits) memcpy(hits + 4 * p, c.hit, 16);
				if (lookupsPerPixel) lookupsPerPixel[p] = (uint32_t)c.nLookup;
				fold_counters(local, &c);
			}
		}
		#pragma omp critical
		{
			for (int i = 0; i < 4; i++) total[i] += local[i];
			if (local[4] > total[4]) total[4] = local[4];
		}
	}
	if (counters) memcpy(counters, total, sizeof(total));
	return 0;
}

int orc_trace_rays(void* h, const float* rays, uint64_t n, const float* translation, uint32_t scale, int algorithm,
	uint32_t* colour, int32_t* hits, uint64_t* counters, int nThreads)
{
	Scene* s = (Scene*)h;
	if (s->storage == -1) return 1;
	if (nThreads < 1) nThreads = 1;
	uint64_t total[5] = {0, 0, 0, 0, 0};
	#pragma omp parallel num_threads(nThreads)
	{
		Ctx c;
		ctx_init(&c, s, translation, scale);
		uint64_t local[5] = {0, 0, 0, 0, 0};
		#pragma omp for schedule(dynamic, 256)
		for (int64_t i = 0; i < (int64_t)n; i++)
		{
			ctx_reset(&c);
			colour[i] = march_scene(&c, rays + 6 * i, rays + 6 * i + 3, algorithm);
			if (hits) memcpy(hits + 4 * i, c.hit, 16);
			fold_counters(local, &c);
		}
		#pragma omp critical
		{
			for (int i = 0; i < 4; i++) total[i] += local[i];
			if (local[4] > total[4]) total[4] = local[4];
		}
	}
	if (counters) memcpy(counters, total, sizeof(total));
	return 0;
}

int orc_lookup(void* h, const int32_t* xyz, uint64_t n, uint32_t* out, uint8_t* exists)
{
	Scene* s = (Scene*)h;
	if (s->storage == -1) return 1;
	for (uint64_t i = 0; i < n; i++)
	{
		int r[3]; uint32_t l[3];
		for (int a = 0; a < 3; a++)
		{
			int32_t v = xyz[3 * i + a];
			r[a] = (int32_t)floorf(v / (float)BLOCK_SIZE);
			l[a] = (uint32_t)(((v % BLOCK_SIZE) + BLOCK_SIZE) % BLOCK_SIZE);
		}
		out[i] = EMPTY_VAL;
		if (exists) exists[i] = 0;
		if (!in_scene(s, r)) continue;
		const Region* reg = region_at(s, r);
		if (!reg) continue;
		int e = s->storage == 1 ? 1 : reg->cluster[cluster_id(l[0], l[1], l[2])];
		if (exists) exists[i] = (uint8_t)e;
		if (e) out[i] = region_find(reg, make_key(l[0], l[1], l[2]));
	}
	return 0;
}
