// TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product; the product never links this.
//
// Host build of the UNMODIFIED reference hot path (SURVEY.md §8c "CPU oracle").
// The reference headers are compiled where they lie under /root/reference (never copied):
//   g++ -std=c++17 -O2 -DNDEBUG -ffp-contract=off -Ioracle/fakecuda -I/root/reference/VoxelRaymarcher/src
// and the result goes to oracle/_ref/libvrm_ref_host.so (git-ignored, travels to the GPU box).
//
// This file only adds a C ABI around the reference's own functions:
//   VoxelSceneCPU::insertVoxel / generateVoxelScene   (geometry/VoxelSceneCPU.cuh:16-93)
//   generateVoxelScene "kernel"                       (renderer/Renderer.cuh:1066-1086)
//   rayMarchSceneOriginal / rayMarchSceneJumpAxis     (renderer/Renderer.cuh:1033-1063)
//   rayMarchVoxelScene / rayMarchVoxelSceneLongestAxis (renderer/Renderer.cuh:338, 917)
//   Camera::Camera                                     (renderer/camera/Camera.cuh:11-23)
//   VoxelCube / VoxelSphere generators                 (geometry/VoxelCube.cuh, VoxelSphere.cuh)
//   VoxelFile::readVoxelFile                           (geometry/VoxelFile.cuh:9-35; reads "resources/<name>" under the working directory)
// plus a recording StorageStructure wrapper (the reference's own virtual seam,
// storage/StorageStructure.cuh:12-17) that extracts the per-pixel first-hit voxel, which the
// reference itself never outputs, and counts lookups for the roofline model.
#include <cuda_runtime.h>

thread_local uint3 threadIdx = {0, 0, 0};
thread_local uint3 blockIdx = {0, 0, 0};
thread_local dim3 blockDim;

#include <iostream>
#include <sstream>
#include <thread>
#include <atomic>
#include <mutex>

#include "geometry/VoxelCube.cuh"
#include "geometry/VoxelFile.cuh"
#include "geometry/VoxelFunctions.cuh"
#include "geometry/VoxelSceneCPU.cuh"
#include "geometry/VoxelSphere.cuh"
#include "renderer/Renderer.cuh"
#include "renderer/camera/Camera.cuh"

namespace {

struct PixelRec
{
	int32_t hit[4];      // global voxel x,y,z + hit flag of the FIRST successful lookup of this pixel
	uint64_t nExist, nExistFalse, nLookup, nLookupHit;
};
thread_local PixelRec* tlRec = nullptr;

// Forwards to the reference's adapter object; records what it is asked.
class Recording : public StorageStructure
{
public:
	Recording(StorageStructure* in, int32_t rx, int32_t ry, int32_t rz) : inner(in), regX(rx), regY(ry), regZ(rz) {}
	uint32_t lookupVoxel(int32_t x, int32_t y, int32_t z) const override
	{
		uint32_t r = inner->lookupVoxel(x, y, z);
		PixelRec* rec = tlRec;
		if (rec)
		{
			rec->nLookup++;
			if (r != EMPTY_VAL)
			{
				rec->nLookupHit++;
				if (!rec->hit[3])
				{
					rec->hit[0] = regX * 64 + x;
					rec->hit[1] = regY * 64 + y;
					rec->hit[2] = regZ * 64 + z;
					rec->hit[3] = 1;
				}
			}
		}
		return r;
	}
	bool doesVoxelSpaceExist(int32_t x, int32_t y, int32_t z) const override
	{
		bool e = inner->doesVoxelSpaceExist(x, y, z);
		PixelRec* rec = tlRec;
		if (rec)
		{
			rec->nExist++;
			if (!e) rec->nExistFalse++;
		}
		return e;
	}
	StorageStructure* inner;
	int32_t regX, regY, regZ;
};

struct RefScene
{
	VoxelSceneCPU cpu;
	std::vector<StorageStructure*> plain;      // what the reference's generateVoxelScene produced
	std::vector<StorageStructure*> recording;  // same, wrapped
	uint32_t diameter = 0;
	int32_t minCoord = 0;
	uint32_t filled = 0;
	int storageType = -1;
	size_t nInserted = 0;
};

std::mutex gCoutMutex;

struct CoutSilencer
{
	std::streambuf* old;
	std::ostringstream sink;
	CoutSilencer() { old = std::cout.rdbuf(sink.rdbuf()); }
	~CoutSilencer() { std::cout.rdbuf(old); }
};

Camera cameraFromFloats(const float* c)
{
	Camera cam(Vector3f(0, 0, 0), Vector3f(0, 0, -1), Vector3f(0, 1, 0), 60.0f, 1.0f);
	cam.origin = Vector3f(c[0], c[1], c[2]);
	cam.lowerLeftCorner = Vector3f(c[3], c[4], c[5]);
	cam.horizontalVector = Vector3f(c[6], c[7], c[8]);
	cam.verticalVector = Vector3f(c[9], c[10], c[11]);
	cam.forwardVector = Vector3f(c[12], c[13], c[14]);
	return cam;
}

void accumulate(uint64_t* counters, const PixelRec& rec, uint64_t& maxLookups)
{
	counters[0] += rec.nExist;
	counters[1] += rec.nExistFalse;
	counters[2] += rec.nLookup;
	counters[3] += rec.nLookupHit;
	if (rec.nLookup > maxLookups) maxLookups = rec.nLookup;
}

}  // namespace

extern "C" {

void* refh_scene_create() { return new RefScene(); }

void refh_scene_destroy(void* h) { delete static_cast<RefScene*>(h); }  // per-region stores leak, as in the reference

void refh_scene_add_voxels(void* h, const int32_t* xyz, const uint32_t* rgb, uint64_t n)
{
	RefScene* s = static_cast<RefScene*>(h);
	for (uint64_t i = 0; i < n; i++)
		s->cpu.insertVoxel(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], rgb[i]);
	s->nInserted += n;
}

// The reference's scene reader (Main.cu:99): inserts the voxels of resources/<filename> (relative to the process' working directory).
void refh_scene_load_file(void* h, const char* filename)
{
	VoxelFile::readVoxelFile(static_cast<RefScene*>(h)->cpu, filename);
}

// The reference's own (unused by its main) procedural generators, for generator parity tests.
void refh_scene_add_cube(void* h, int32_t x, int32_t y, int32_t z, int32_t halfWidth)
{
	VoxelCube::generateVoxelCube(static_cast<RefScene*>(h)->cpu, x, y, z, halfWidth);
}
void refh_scene_add_sphere(void* h, uint32_t x, uint32_t y, uint32_t z, uint32_t r, int checkered)
{
	RefScene* s = static_cast<RefScene*>(h);
	if (checkered) VoxelSphere::generateCheckeredVoxelSphere(s->cpu, x, y, z, r);
	else VoxelSphere::generateVoxelSphere(s->cpu, x, y, z, r);
}

// storageType: 0 = VOXEL_CLUSTER_STORE, 1 = HASH_TABLE (geometry/VoxelFunctions.cuh:37)
int refh_scene_build(void* h, int storageType)
{
	RefScene* s = static_cast<RefScene*>(h);
	if (s->storageType != -1) return 1;
	{
		std::lock_guard<std::mutex> lock(gCoutMutex);
		CoutSilencer quiet;
		s->cpu.generateVoxelScene(StorageType(storageType));
	}
	s->storageType = storageType;
	s->diameter = s->cpu.getArrayDiameter();
	s->minCoord = s->cpu.getMinCoord();
	uint32_t size = s->cpu.getArraySize();
	// the reference never zeroes this table; the harness owns it, so it does.
	s->plain.assign(size, nullptr);
	generateVoxelScene(s->plain.data(), s->cpu.deviceVoxelScene, size, StorageType(storageType));
	s->recording.assign(size, nullptr);
	uint32_t d = s->diameter;
	for (uint32_t i = 0; i < size; i++)
	{
		if (!s->plain[i]) continue;
		s->filled++;
		int32_t rx = static_cast<int32_t>(i % d) + s->minCoord;
		int32_t ry = static_cast<int32_t>((i / d) % d) + s->minCoord;
		int32_t rz = static_cast<int32_t>(i / (d * d)) + s->minCoord;
		s->recording[i] = new Recording(s->plain[i], rx, ry, rz);
	}
	return 0;
}

void refh_scene_info(void* h, uint32_t* diameter, int32_t* minCoord, uint32_t* filled)
{
	RefScene* s = static_cast<RefScene*>(h);
	*diameter = s->diameter;
	*minCoord = s->minCoord;
	*filled = s->filled;
}

// Main.cu:26-42 setupConstantValues; the symbols are process-wide, exactly as in the reference.
void refh_set_lighting(const float* dir, const float* color, const float* pos, int usePoint, int useShadows)
{
	Vector3f d(dir[0], dir[1], dir[2]), c(color[0], color[1], color[2]), p(pos[0], pos[1], pos[2]);
	bool up = usePoint != 0, us = useShadows != 0;
	cudaMemcpyToSymbol(LIGHT_DIRECTION, &d, sizeof(Vector3f));
	cudaMemcpyToSymbol(LIGHT_COLOR, &c, sizeof(Vector3f));
	cudaMemcpyToSymbol(LIGHT_POSITION, &p, sizeof(Vector3f));
	cudaMemcpyToSymbol(USE_POINT_LIGHT, &up, sizeof(bool));
	cudaMemcpyToSymbol(USE_SHADOWS, &us, sizeof(bool));
}

// Main.cu:28: makeUnitVector(Vector3f(1,1,1)) and friends, evaluated by the reference's own Vector3.
void refh_make_unit_vector(const float* v, float* out)
{
	Vector3f u = makeUnitVector(Vector3f(v[0], v[1], v[2]));
	out[0] = u.getX(); out[1] = u.getY(); out[2] = u.getZ();
}

// Camera.cuh:11-23; out = origin, lowerLeftCorner, horizontal, vertical, forward (15 floats = the 60-byte struct).
void refh_camera_make(const float* origin, const float* lookAt, const float* up, float fov, float aspect, float* out)
{
	Camera cam(Vector3f(origin[0], origin[1], origin[2]), Vector3f(lookAt[0], lookAt[1], lookAt[2]), Vector3f(up[0], up[1], up[2]), fov, aspect);
	static_assert(sizeof(Camera) == 60, "Camera is 5 x Vector3f");
	memcpy(out, &cam, 60);
}

// algorithm: 0 = longest axis, 1 = original (Main.cu:58-68).  hits = 4 x int32 per pixel (x,y,z,flag), nullable.
// counters (nullable) = {exist checks, exist checks answering false, lookups, lookups that found a voxel, max lookups of any pixel}.
// lookupsPerPixel (nullable) = lookups issued by each pixel (primary + shadow).
int refh_render(void* h, const float* camera15, const float* translation, uint32_t scale, int algorithm,
	uint32_t width, uint32_t height, uint8_t* rgb, int32_t* hits, uint64_t* counters, uint32_t* lookupsPerPixel, int nThreads)
{
	RefScene* s = static_cast<RefScene*>(h);
	if (s->storageType == -1) return 1;
	Camera cam = cameraFromFloats(camera15);
	VoxelSceneInfo info(Vector3f(translation[0], translation[1], translation[2]), scale);
	bool record = hits || counters || lookupsPerPixel;
	StorageStructure** table = record ? s->recording.data() : s->plain.data();
	if (nThreads < 1) nThreads = 1;
	std::atomic<uint32_t> nextRow(0);
	std::vector<std::vector<uint64_t>> perThread(nThreads, std::vector<uint64_t>(5, 0));
	auto worker = [&](int tid)
	{
		blockDim = dim3(1, 1, 1);
		threadIdx = {0, 0, 0};
		uint64_t maxLookups = 0;
		for (;;)
		{
			uint32_t y = nextRow.fetch_add(1);
			if (y >= height) break;
			for (uint32_t x = 0; x < width; x++)
			{
				PixelRec rec = {};
				tlRec = record ? &rec : nullptr;
				blockIdx = {x, y, 0};
				if (algorithm == 1)
					rayMarchSceneOriginal(width, height, &cam, &info, rgb, table, s->diameter, s->minCoord);
				else
					rayMarchSceneJumpAxis(width, height, &cam, &info, rgb, table, s->diameter, s->minCoord);
				tlRec = nullptr;
				if (hits) memcpy(hits + 4 * (static_cast<size_t>(y) * width + x), rec.hit, 16);
				if (lookupsPerPixel) lookupsPerPixel[static_cast<size_t>(y) * width + x] = static_cast<uint32_t>(rec.nLookup);
				accumulate(perThread[tid].data(), rec, maxLookups);
			}
		}
		perThread[tid][4] = maxLookups;
	};
	std::vector<std::thread> pool;
	for (int t = 1; t < nThreads; t++) pool.emplace_back(worker, t);
	worker(0);
	for (auto& t : pool) t.join();
	if (counters)
	{
		for (int i = 0; i < 5; i++) counters[i] = 0;
		for (int t = 0; t < nThreads; t++)
		{
			for (int i = 0; i < 4; i++) counters[i] += perThread[t][i];
			if (perThread[t][4] > counters[4]) counters[4] = perThread[t][4];
		}
	}
	return 0;
}

// Arbitrary world rays (BASELINE.json config 5): rays = 6 floats each (origin, direction); colour = return value of
// rayMarchVoxelScene / rayMarchVoxelSceneLongestAxis (0 = background).
int refh_trace_rays(void* h, const float* rays, uint64_t n, const float* translation, uint32_t scale, int algorithm,
	uint32_t* colour, int32_t* hits, uint64_t* counters, int nThreads)
{
	RefScene* s = static_cast<RefScene*>(h);
	if (s->storageType == -1) return 1;
	VoxelSceneInfo info(Vector3f(translation[0], translation[1], translation[2]), scale);
	bool record = hits || counters;
	StorageStructure** table = record ? s->recording.data() : s->plain.data();
	VoxelScene scene(table, s->diameter, s->minCoord);
	if (nThreads < 1) nThreads = 1;
	std::vector<std::vector<uint64_t>> perThread(nThreads, std::vector<uint64_t>(5, 0));
	auto worker = [&](int tid)
	{
		uint64_t maxLookups = 0;
		uint64_t lo = n * tid / nThreads, hi = n * (tid + 1) / nThreads;
		for (uint64_t i = lo; i < hi; i++)
		{
			PixelRec rec = {};
			tlRec = record ? &rec : nullptr;
			Ray ray(Vector3f(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]), Vector3f(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]));
			colour[i] = algorithm == 1 ? rayMarchVoxelScene(ray, &info, scene) : rayMarchVoxelSceneLongestAxis(ray, &info, scene);
			tlRec = nullptr;
			if (hits) memcpy(hits + 4 * i, rec.hit, 16);
			accumulate(perThread[tid].data(), rec, maxLookups);
		}
		perThread[tid][4] = maxLookups;
	};
	std::vector<std::thread> pool;
	for (int t = 1; t < nThreads; t++) pool.emplace_back(worker, t);
	worker(0);
	for (auto& t : pool) t.join();
	if (counters)
	{
		for (int i = 0; i < 5; i++) counters[i] = 0;
		for (int t = 0; t < nThreads; t++)
		{
			for (int i = 0; i < 4; i++) counters[i] += perThread[t][i];
			if (perThread[t][4] > counters[4]) counters[4] = perThread[t][4];
		}
	}
	return 0;
}

// The StorageStructure seam itself, on GLOBAL voxel coordinates: region = floor(c / 64), local = c mod 64
// (VoxelSceneCPU.cuh:19-26).  out = lookupVoxel result (1<<30 when absent / region absent);
// exists = doesVoxelSpaceExist (0 when the region is absent).
int refh_lookup(void* h, const int32_t* xyz, uint64_t n, uint32_t* out, uint8_t* exists)
{
	RefScene* s = static_cast<RefScene*>(h);
	if (s->storageType == -1) return 1;
	VoxelScene scene(s->plain.data(), s->diameter, s->minCoord);
	for (uint64_t i = 0; i < n; i++)
	{
		int32_t c[3], r[3], l[3];
		for (int a = 0; a < 3; a++)
		{
			c[a] = xyz[3 * i + a];
			r[a] = static_cast<int32_t>(std::floorf(c[a] / 64.0f));
			l[a] = ((c[a] % 64) + 64) % 64;
		}
		out[i] = EMPTY_VAL;
		if (exists) exists[i] = 0;
		if (!scene.isRayInScene(r[0], r[1], r[2])) continue;
		StorageStructure* st = scene.getRegionStorageStructure(r[0], r[1], r[2]);
		if (!st) continue;
		bool e = st->doesVoxelSpaceExist(l[0], l[1], l[2]);
		if (exists) exists[i] = e ? 1 : 0;
		if (e) out[i] = st->lookupVoxel(l[0], l[1], l[2]);  // VCS lookup dereferences the cluster pointer: only legal when it exists
	}
	return 0;
}

}  // extern "C"
