// TEST INFRASTRUCTURE / BASELINE ONLY (oracle/).  Not part of the product.
//
// The reference's OWN CUDA kernels (renderer/Renderer.cuh:1033-1063) and storage structures, compiled unmodified
// from /root/reference for sm_100a as one translation unit (the reference keeps non-inline __device__ functions and
// __constant__ symbols in headers).  Two builds of this file exist: libvrm_ref_cuda.so (nvcc defaults, -fmad=true:
// "the reference rebuilt on the same B200", the speed baseline of BASELINE.md §4) and libvrm_ref_cuda_exact.so
// (-fmad=false: the bit-exact GPU oracle).  This harness replaces main/Main.cu only: it owns (and zeroes, SURVEY.md
// F10) the region table, accepts any resolution / camera, and can wrap the storage objects in a recording
// StorageStructure to extract the per-pixel first-hit voxel.
#include <iostream>
#include <sstream>

#include "geometry/VoxelFunctions.cuh"
#include "geometry/VoxelSceneCPU.cuh"
#include "renderer/Renderer.cuh"
#include "renderer/camera/Camera.cuh"

namespace
{

__device__ int32_t* gHits = nullptr;  // 4 x int32 per pixel / ray
__device__ uint32_t gWidth = 0;       // 0 = linear ray indexing (trace mode)

__device__ __forceinline__ size_t currentPixel()
{
	uint32_t x = threadIdx.x + blockIdx.x * blockDim.x;
	uint32_t y = threadIdx.y + blockIdx.y * blockDim.y;
	return gWidth ? (size_t)y * gWidth + x : (size_t)x;
}

class Recording : public StorageStructure
{
public:
	__device__ Recording(StorageStructure* in, int32_t rx, int32_t ry, int32_t rz) : inner(in), regX(rx), regY(ry), regZ(rz) {}
	__device__ virtual uint32_t lookupVoxel(int32_t x, int32_t y, int32_t z) const override
	{
		uint32_t r = inner->lookupVoxel(x, y, z);
		if (r != EMPTY_VAL && gHits)
		{
			int32_t* h = gHits + 4 * currentPixel();
			if (!h[3]) { h[0] = regX * 64 + x; h[1] = regY * 64 + y; h[2] = regZ * 64 + z; h[3] = 1; }
		}
		return r;
	}
	__device__ virtual bool doesVoxelSpaceExist(int32_t x, int32_t y, int32_t z) const override { return inner->doesVoxelSpaceExist(x, y, z); }
	StorageStructure* inner;
	int32_t regX, regY, regZ;
};

__global__ void wrapRecording(StorageStructure** plain, StorageStructure** rec, uint32_t d, int32_t minCoord)
{
	uint32_t size = d * d * d;
	for (uint32_t i = 0; i < size; i++)
	{
		rec[i] = nullptr;
		if (plain[i]) rec[i] = new Recording(plain[i], (int32_t)(i % d) + minCoord, (int32_t)((i / d) % d) + minCoord, (int32_t)(i / (d * d)) + minCoord);
	}
}

__global__ void setRecorder(int32_t* hits, uint32_t width) { gHits = hits; gWidth = width; }

__global__ void traceRays(const float* rays, unsigned long long n, VoxelSceneInfo* info, StorageStructure** table, uint32_t d, int32_t minCoord, int algorithm, uint32_t* colour)
{
	unsigned long long i = threadIdx.x + (unsigned long long)blockIdx.x * blockDim.x;
	if (i >= n) return;
	Ray ray(Vector3f(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]), Vector3f(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]));
	VoxelScene scene(table, d, minCoord);
	colour[i] = algorithm == 1 ? rayMarchVoxelScene(ray, info, scene) : rayMarchVoxelSceneLongestAxis(ray, info, scene);
}

struct RefScene
{
	VoxelSceneCPU cpu;
	StorageStructure** plain = nullptr;
	StorageStructure** recording = nullptr;
	uint32_t diameter = 0, filled = 0;
	int32_t minCoord = 0;
	int storageType = -1;
	std::vector<std::pair<int32_t, int32_t>> dummy;
};

struct CoutSilencer
{
	std::streambuf* old;
	std::ostringstream sink;
	CoutSilencer() { old = std::cout.rdbuf(sink.rdbuf()); }
	~CoutSilencer() { std::cout.rdbuf(old); }
};

float gLastMs = 0.0f;

void launchRender(RefScene* s, StorageStructure** table, Camera* dCam, VoxelSceneInfo* dInfo, uint8_t* dFb, int algorithm, uint32_t width, uint32_t height)
{
	// main/Main.cu:109-111,121,126
	uint32_t numThreads = 8;
	dim3 blocks(width / numThreads + 1, height / numThreads + 1);
	dim3 threads(numThreads, numThreads);
	if (algorithm == 0) rayMarchSceneJumpAxis<<<blocks, threads>>>(width, height, dCam, dInfo, dFb, table, s->diameter, s->minCoord);
	else rayMarchSceneOriginal<<<blocks, threads>>>(width, height, dCam, dInfo, dFb, table, s->diameter, s->minCoord);
}

Camera cameraFromFloats(const float* c)
{
	Camera cam(Vector3f(0, 0, 0), Vector3f(0, 0, -1), Vector3f(0, 1, 0), 60.0f, 1.0f);
	memcpy(&cam, c, 60);
	return cam;
}

}  // namespace

extern "C" {

void* refg_scene_create() { return new RefScene(); }
void refg_scene_destroy(void* h)
{
	RefScene* s = static_cast<RefScene*>(h);
	if (s->plain) cudaFree(s->plain);
	if (s->recording) cudaFree(s->recording);
	delete s;  // per-region stores leak, as in the reference (VoxelSceneCPU.cuh:95-105)
}

void refg_scene_add_voxels(void* h, const int32_t* xyz, const uint32_t* rgb, uint64_t n)
{
	RefScene* s = static_cast<RefScene*>(h);
	for (uint64_t i = 0; i < n; i++) s->cpu.insertVoxel(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], rgb[i]);
}

int refg_scene_build(void* h, int storageType)
{
	RefScene* s = static_cast<RefScene*>(h);
	if (s->storageType != -1) return 1;
	{
		CoutSilencer quiet;
		s->cpu.generateVoxelScene(StorageType(storageType));  // host build + cudaMemcpy, VoxelSceneCPU.cuh:49-93
	}
	s->storageType = storageType;
	s->diameter = s->cpu.getArrayDiameter();
	s->minCoord = s->cpu.getMinCoord();
	uint32_t size = s->cpu.getArraySize();
	if (cudaMalloc(&s->plain, sizeof(StorageStructure*) * size) != cudaSuccess) return 2;
	if (cudaMalloc(&s->recording, sizeof(StorageStructure*) * size) != cudaSuccess) return 2;
	cudaMemset(s->plain, 0, sizeof(StorageStructure*) * size);
	cudaDeviceSetLimit(cudaLimitMallocHeapSize, 64u << 20);
	generateVoxelScene<<<1, 1>>>(s->plain, s->cpu.deviceVoxelScene, size, StorageType(storageType));  // Main.cu:211
	wrapRecording<<<1, 1>>>(s->plain, s->recording, s->diameter, s->minCoord);
	if (cudaDeviceSynchronize() != cudaSuccess) return 3;
	std::vector<void*> host(size);
	cudaMemcpy(host.data(), s->cpu.deviceVoxelScene, sizeof(void*) * size, cudaMemcpyDeviceToHost);
	for (void* p : host) if (p) s->filled++;
	return 0;
}

void refg_scene_info(void* h, uint32_t* diameter, int32_t* minCoord, uint32_t* filled)
{
	RefScene* s = static_cast<RefScene*>(h);
	*diameter = s->diameter; *minCoord = s->minCoord; *filled = s->filled;
}

void refg_set_lighting(const float* dir, const float* color, const float* pos, int usePoint, int useShadows)
{
	bool up = usePoint != 0, us = useShadows != 0;
	cudaMemcpyToSymbol(LIGHT_DIRECTION, dir, sizeof(Vector3f));
	cudaMemcpyToSymbol(LIGHT_COLOR, color, sizeof(Vector3f));
	cudaMemcpyToSymbol(LIGHT_POSITION, pos, sizeof(Vector3f));
	cudaMemcpyToSymbol(USE_POINT_LIGHT, &up, sizeof(bool));
	cudaMemcpyToSymbol(USE_SHADOWS, &us, sizeof(bool));
}

void refg_make_unit_vector(const float* v, float* out)
{
	Vector3f u = makeUnitVector(Vector3f(v[0], v[1], v[2]));
	out[0] = u.getX(); out[1] = u.getY(); out[2] = u.getZ();
}

void refg_camera_make(const float* origin, const float* lookAt, const float* up, float fov, float aspect, float* out)
{
	Camera cam(Vector3f(origin[0], origin[1], origin[2]), Vector3f(lookAt[0], lookAt[1], lookAt[2]), Vector3f(up[0], up[1], up[2]), fov, aspect);
	memcpy(out, &cam, 60);
}

float refg_last_kernel_ms() { return gLastMs; }

// Same signature as refh_render; counters / lookupsPerPixel / nThreads are ignored on the GPU.
int refg_render(void* h, const float* camera15, const float* translation, uint32_t scale, int algorithm,
	uint32_t width, uint32_t height, uint8_t* rgb, int32_t* hits, uint64_t*, uint32_t*, int)
{
	RefScene* s = static_cast<RefScene*>(h);
	if (s->storageType == -1) return 1;
	Camera cam = cameraFromFloats(camera15);
	VoxelSceneInfo info(Vector3f(translation[0], translation[1], translation[2]), scale);
	Camera* dCam; VoxelSceneInfo* dInfo; uint8_t* dFb; int32_t* dHits = nullptr;
	size_t px = (size_t)width * height;
	cudaMalloc(&dCam, sizeof(Camera)); cudaMalloc(&dInfo, sizeof(VoxelSceneInfo)); cudaMalloc(&dFb, px * 3);
	cudaMemcpy(dCam, &cam, sizeof(Camera), cudaMemcpyHostToDevice);
	cudaMemcpy(dInfo, &info, sizeof(VoxelSceneInfo), cudaMemcpyHostToDevice);
	cudaMemset(dFb, 0, px * 3);
	if (hits) { cudaMalloc(&dHits, px * 16); cudaMemset(dHits, 0, px * 16); }
	setRecorder<<<1, 1>>>(dHits, width);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaEventRecord(e0);
	launchRender(s, hits ? s->recording : s->plain, dCam, dInfo, dFb, algorithm, width, height);
	cudaEventRecord(e1);
	cudaError_t err = cudaDeviceSynchronize();
	cudaEventElapsedTime(&gLastMs, e0, e1);
	cudaMemcpy(rgb, dFb, px * 3, cudaMemcpyDeviceToHost);
	if (hits) cudaMemcpy(hits, dHits, px * 16, cudaMemcpyDeviceToHost);
	setRecorder<<<1, 1>>>(nullptr, 0);
	cudaDeviceSynchronize();
	cudaFree(dCam); cudaFree(dInfo); cudaFree(dFb); if (dHits) cudaFree(dHits);
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	return err == cudaSuccess ? 0 : 2;
}

// Speed baseline: `warmup` untimed + `iters` timed launches of the reference kernel (plain storage objects, no
// recorder), each timed with CUDA events; msOut receives `iters` durations.
int refg_render_timed(void* h, const float* camera15, const float* translation, uint32_t scale, int algorithm,
	uint32_t width, uint32_t height, int warmup, int iters, float* msOut)
{
	RefScene* s = static_cast<RefScene*>(h);
	if (s->storageType == -1) return 1;
	Camera cam = cameraFromFloats(camera15);
	VoxelSceneInfo info(Vector3f(translation[0], translation[1], translation[2]), scale);
	Camera* dCam; VoxelSceneInfo* dInfo; uint8_t* dFb;
	size_t px = (size_t)width * height;
	cudaMalloc(&dCam, sizeof(Camera)); cudaMalloc(&dInfo, sizeof(VoxelSceneInfo)); cudaMalloc(&dFb, px * 3);
	cudaMemcpy(dCam, &cam, sizeof(Camera), cudaMemcpyHostToDevice);
	cudaMemcpy(dInfo, &info, sizeof(VoxelSceneInfo), cudaMemcpyHostToDevice);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaError_t err = cudaSuccess;
	for (int i = 0; i < warmup + iters && err == cudaSuccess; i++)
	{
		cudaEventRecord(e0);
		launchRender(s, s->plain, dCam, dInfo, dFb, algorithm, width, height);
		cudaEventRecord(e1);
		err = cudaDeviceSynchronize();
		if (i >= warmup) cudaEventElapsedTime(msOut + (i - warmup), e0, e1);
	}
	cudaFree(dCam); cudaFree(dInfo); cudaFree(dFb);
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	return err == cudaSuccess ? 0 : 2;
}

int refg_trace_rays(void* h, const float* rays, uint64_t n, const float* translation, uint32_t scale, int algorithm,
	uint32_t* colour, int32_t* hits, uint64_t*, int)
{
	RefScene* s = static_cast<RefScene*>(h);
	if (s->storageType == -1) return 1;
	VoxelSceneInfo info(Vector3f(translation[0], translation[1], translation[2]), scale);
	VoxelSceneInfo* dInfo; float* dRays; uint32_t* dCol; int32_t* dHits = nullptr;
	cudaMalloc(&dInfo, sizeof(VoxelSceneInfo)); cudaMalloc(&dRays, n * 24); cudaMalloc(&dCol, n * 4);
	cudaMemcpy(dInfo, &info, sizeof(VoxelSceneInfo), cudaMemcpyHostToDevice);
	cudaMemcpy(dRays, rays, n * 24, cudaMemcpyHostToDevice);
	if (hits) { cudaMalloc(&dHits, n * 16); cudaMemset(dHits, 0, n * 16); }
	setRecorder<<<1, 1>>>(dHits, 0);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaEventRecord(e0);
	traceRays<<<(unsigned)((n + 63) / 64), 64>>>(dRays, n, dInfo, hits ? s->recording : s->plain, s->diameter, s->minCoord, algorithm, dCol);
	cudaEventRecord(e1);
	cudaError_t err = cudaDeviceSynchronize();
	cudaEventElapsedTime(&gLastMs, e0, e1);
	cudaMemcpy(colour, dCol, n * 4, cudaMemcpyDeviceToHost);
	if (hits) cudaMemcpy(hits, dHits, n * 16, cudaMemcpyDeviceToHost);
	setRecorder<<<1, 1>>>(nullptr, 0);
	cudaDeviceSynchronize();
	cudaFree(dInfo); cudaFree(dRays); cudaFree(dCol); if (dHits) cudaFree(dHits);
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	return err == cudaSuccess ? 0 : 2;
}

int refg_lookup(void*, const int32_t*, uint64_t, uint32_t*, uint8_t*) { return 9; }  // seam parity is covered by the host build

}  // extern "C"
