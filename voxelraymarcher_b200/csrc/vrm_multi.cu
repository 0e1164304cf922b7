// vrm_multi.cu -- the multi-GPU part of the boundary (SURVEY.md 8b `render_views(handles[], cameras[], ...)`, 8e).
//
// The reference pins device 0 and renders one frame per process (main/Main.cu:82-94, 195-229); sharding a view batch is new
// surface.  Pixels and views are independent, so the path shards with no data-path collective: the structure is replicated (one
// handle per device, each built from the same voxel list), views are dealt out to the devices, and the only exchange is the
// finished RGB8 frames arriving in ONE gather buffer -- written there by the render kernels themselves over NVLink peer memory.
//
//  * vrm_render_views_sharded: ONE process, one handle per device.  Views are claimed DYNAMICALLY: every device keeps a small window
//    of single-view launches in flight and takes the next unclaimed view whenever one of its launches has finished (views of an
//    orbit differ in cost by 2-3x, static blocks left devices idle: DESIGN.md 6).  Kernels store straight into the gather
//    buffer on the first handle's device when peer access is available, else into a local frame followed by cudaMemcpyPeerAsync.
//  * completion flags for the one-process-per-GPU form (torch.distributed plumbing, multigpu.py): a rank's render launches publish
//    a sequence number with release semantics at system scope into a word of the gatherer's memory after the frame's stores;
//    the gatherer waits on the words with a one-warp kernel.  This replaces a 4-byte NCCL all-reduce per step.
#include "vrm_internal.h"
#include "../../include/vrm_b200.h"

#include <chrono>
#include <cstring>
#include <thread>
#include <vector>

namespace
{

__global__ void signal_kernel(uint32_t* flag, uint32_t value)
{
	// stream order puts this kernel behind the render kernel, whose stores (to peer memory too) are performed by then; the release
	// at system scope orders them before the flag for an observer on another device
	__threadfence_system();
	asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}

// the counting form: rank-agnostic "frames landed" counter for dynamically claimed views
__global__ void signal_add_kernel(uint32_t* counter, uint32_t frames)
{
	__threadfence_system();
	atomicAdd_system(counter, frames);
}

// dynamic view claiming across processes: the next unclaimed view index, fetched from a counter in the gatherer's memory into
// page-locked host memory of the claiming rank
__global__ void claim_kernel(uint32_t* counter, uint32_t* out)
{
	*out = atomicAdd_system(counter, 1u);
	__threadfence_system();
}

// One lane per flag: spin until flag >= minValue (sequence numbers compare modulo 2^32) or the time-out passes.
__global__ void wait_flags_kernel(const uint32_t* flags, uint32_t n, uint32_t strideWords, uint32_t minValue, unsigned long long timeoutNs, int* status)
{
	unsigned long long t0;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
	for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
	{
		const uint32_t* p = flags + (size_t)i * strideWords;
		for (;;)
		{
			uint32_t v;
			asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
			if ((int32_t)(v - minValue) >= 0) break;
			unsigned long long t;
			asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
			if (t - t0 > timeoutNs) { if (status) atomicExch(status, 1); return; }
			__nanosleep(200);
		}
	}
}

}  // namespace

// called by vrm_launch_render after the frame's kernels (vrm_render.cu)
void vrm_signal_completion(vrm_scene* s, uint32_t frames)
{
	if (s->d_doneFlag) signal_kernel<<<1, 1, 0, s->stream>>>(s->d_doneFlag, ++s->doneSeq);
	if (s->d_doneCounter) signal_add_kernel<<<1, 1, 0, s->stream>>>(s->d_doneCounter, frames);
}

extern "C" {

int vrm_scene_set_completion_flag(vrm_scene* s, uint32_t* d_flag, uint32_t first_value)
{
	if (!s) return VRM_ERR_INVALID;
	s->d_doneFlag = d_flag;
	s->doneSeq = first_value - 1u;
	return VRM_OK;
}

int vrm_scene_set_completion_counter(vrm_scene* s, uint32_t* d_counter)
{
	if (!s) return VRM_ERR_INVALID;
	s->d_doneCounter = d_counter;
	return VRM_OK;
}

int vrm_claim_next(int device, void* cuda_stream, uint32_t* d_counter, uint32_t* h_claimed)
{
	if (!d_counter || !h_claimed) return VRM_ERR_INVALID;
	if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return VRM_ERR_CUDA; }
	cudaPointerAttributes at;
	if (cudaPointerGetAttributes(&at, h_claimed) != cudaSuccess || at.type != cudaMemoryTypeHost || !at.devicePointer) { cudaGetLastError(); return VRM_ERR_INVALID; }
	claim_kernel<<<1, 1, 0, static_cast<cudaStream_t>(cuda_stream)>>>(d_counter, static_cast<uint32_t*>(at.devicePointer));
	if (cudaGetLastError() != cudaSuccess) return VRM_ERR_CUDA;
	return VRM_OK;
}

int vrm_wait_flags_device(int device, void* cuda_stream, const uint32_t* d_flags, uint32_t n_flags, uint32_t stride_words, uint32_t min_value,
                          uint32_t timeout_ms, int* d_status)
{
	if (!d_flags || n_flags == 0 || stride_words == 0) return VRM_ERR_INVALID;
	if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return VRM_ERR_CUDA; }
	wait_flags_kernel<<<1, 32, 0, static_cast<cudaStream_t>(cuda_stream)>>>(d_flags, n_flags, stride_words, min_value, (unsigned long long)timeout_ms * 1000000ull, d_status);
	if (cudaGetLastError() != cudaSuccess) return VRM_ERR_CUDA;
	return VRM_OK;
}

int vrm_render_views_sharded(vrm_scene* const* scenes, uint32_t n_scenes, const float* cameras, uint32_t n_views, const float translation[3], uint32_t scale,
                             int algorithm, uint32_t width, uint32_t height, uint8_t* rgb_out, int out_on_device, uint32_t* views_per_scene_out, float* total_ms)
{
	if (!scenes || n_scenes == 0 || !scenes[0]) return VRM_ERR_INVALID;
	vrm_scene* root = scenes[0];
	if (!cameras || !translation || !rgb_out || n_views == 0 || width == 0 || height == 0 || (algorithm != VRM_ALGO_ORIGINAL && algorithm != VRM_ALGO_LONGEST_AXIS))
	{ root->lastError = "invalid render arguments"; return VRM_ERR_INVALID; }
	for (uint32_t i = 0; i < n_scenes; i++)
	{
		if (!scenes[i]) return VRM_ERR_INVALID;
		if (scenes[i]->storage < 0) { root->lastError = "scene not built"; return VRM_ERR_STATE; }
		for (uint32_t j = 0; j < i; j++) if (scenes[j] == scenes[i]) { root->lastError = "the same handle twice"; return VRM_ERR_INVALID; }
	}
	const size_t frameBytes = (size_t)width * height * 3;
	const auto wall0 = std::chrono::steady_clock::now();

	// ---- the gather buffer: the caller's device buffer, or a buffer of the root handle that is copied to the host at the end ----
	VRM_CUDA(root, cudaSetDevice(root->device));
	uint8_t* gather = nullptr;
	bool gatherIsMappedHost = false;
	if (out_on_device) gather = rgb_out;
	else
	{
		cudaPointerAttributes at;
		if (n_scenes == 1 && cudaPointerGetAttributes(&at, rgb_out) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer)
		{ gather = static_cast<uint8_t*>(at.devicePointer); gatherIsMappedHost = true; }  // one device + pinned frames: stored straight into them, as in vrm_render_views
		cudaGetLastError();
		if (!gather)
		{
			if (root->gatherBytes < (size_t)n_views * frameBytes)
			{
				VRM_CUDA(root, cudaStreamSynchronize(root->stream));
				if (root->d_gather) cudaFree(root->d_gather);
				root->d_gather = nullptr; root->gatherBytes = 0;
				VRM_CUDA(root, cudaMalloc(&root->d_gather, (size_t)n_views * frameBytes));
				root->gatherBytes = (size_t)n_views * frameBytes;
			}
			gather = root->d_gather;
		}
	}

	// ---- per device: can its kernels store into the gather buffer? ------------------------------------------------------------
	struct Dev
	{
		vrm_scene* s;
		bool direct;             // kernels store into the gather buffer (same device, or peer access)
		uint8_t* local[2];       // else: two local frames + a peer copy behind each render
		cudaEvent_t done[2];     // window of two launches in flight
		int32_t view[2];
		float* d_cams;           // all cameras, uploaded once
		uint32_t count;
	};
	std::vector<Dev> devs(n_scenes);
	int rc = VRM_OK;
	auto cleanup = [&]() {
		for (Dev& d : devs)
		{
			if (!d.s) continue;
			cudaSetDevice(d.s->device);
			cudaStreamSynchronize(d.s->stream);
			for (int k = 0; k < 2; k++) { if (d.done[k]) cudaEventDestroy(d.done[k]); if (d.local[k]) cudaFree(d.local[k]); }
			if (d.d_cams) cudaFree(d.d_cams);
		}
		cudaGetLastError();
	};
	for (uint32_t i = 0; i < n_scenes; i++)
	{
		Dev& d = devs[i];
		d.s = scenes[i]; d.local[0] = d.local[1] = nullptr; d.done[0] = d.done[1] = nullptr; d.view[0] = d.view[1] = -1; d.d_cams = nullptr; d.count = 0;
		d.direct = d.s->device == root->device || gatherIsMappedHost;
		if (cudaSetDevice(d.s->device) != cudaSuccess) { rc = VRM_ERR_CUDA; break; }
		if (!d.direct)
		{
			int can = 0;
			cudaDeviceCanAccessPeer(&can, d.s->device, root->device);
			if (can)
			{
				const cudaError_t e = cudaDeviceEnablePeerAccess(root->device, 0);
				if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) d.direct = true;
				cudaGetLastError();
			}
		}
		bool ok = true;
		for (int k = 0; k < 2 && ok; k++)
		{
			ok = cudaEventCreateWithFlags(&d.done[k], cudaEventDisableTiming) == cudaSuccess;
			if (ok && !d.direct) ok = cudaMalloc(&d.local[k], frameBytes) == cudaSuccess;
		}
		const size_t camBytes = (size_t)n_views * VRM_CAMERA_FLOATS * sizeof(float);
		ok = ok && cudaMalloc(&d.d_cams, camBytes) == cudaSuccess;
		ok = ok && cudaMemcpyAsync(d.d_cams, cameras, camBytes, cudaMemcpyHostToDevice, d.s->stream) == cudaSuccess;
		// the root's stream may still be using the gather buffer's previous contents (and vice versa): order every device behind it
		if (!ok) { rc = vrm_fail_cuda(root, cudaGetLastError(), "vrm_render_views_sharded: per-device setup"); break; }
	}
	if (rc) { cleanup(); return rc; }

	// ---- dynamic view claiming: a device takes the next view whenever one of its two launch slots is free ------------------------
	uint32_t next = 0, finished = 0;
	auto launch = [&](Dev& d, int k) -> int {
		const uint32_t v = next++;
		d.view[k] = (int32_t)v;
		d.count++;
		if (cudaSetDevice(d.s->device) != cudaSuccess) return VRM_ERR_CUDA;
		uint8_t* dst = d.direct ? gather + (size_t)v * frameBytes : d.local[k];
		int r = vrm_launch_render(d.s, d.d_cams + (size_t)v * VRM_CAMERA_FLOATS, 1, translation, scale, algorithm, width, height, dst, nullptr);
		if (r) return r;
		if (!d.direct && cudaMemcpyPeerAsync(gather + (size_t)v * frameBytes, root->device, d.local[k], d.s->device, frameBytes, d.s->stream) != cudaSuccess) return VRM_ERR_CUDA;
		if (cudaEventRecord(d.done[k], d.s->stream) != cudaSuccess) return VRM_ERR_CUDA;
		return VRM_OK;
	};
	for (int k = 0; k < 2 && !rc; k++)
		for (uint32_t i = 0; i < n_scenes && !rc; i++)
			if (next < n_views) rc = launch(devs[i], k);
	while (!rc && finished < n_views)
	{
		bool progressed = false;
		for (uint32_t i = 0; i < n_scenes && !rc; i++)
		{
			Dev& d = devs[i];
			for (int k = 0; k < 2 && !rc; k++)
			{
				if (d.view[k] < 0) continue;
				const cudaError_t q = cudaEventQuery(d.done[k]);
				if (q == cudaErrorNotReady) continue;
				if (q != cudaSuccess) { rc = vrm_fail_cuda(root, q, "vrm_render_views_sharded: a render launch failed"); break; }
				finished++; progressed = true;
				d.view[k] = -1;
				if (next < n_views) rc = launch(d, k);
			}
		}
		if (!progressed) std::this_thread::yield();
	}
	cudaGetLastError();
	if (!rc && !out_on_device && !gatherIsMappedHost)
	{
		if (cudaSetDevice(root->device) != cudaSuccess || cudaMemcpyAsync(rgb_out, gather, (size_t)n_views * frameBytes, cudaMemcpyDeviceToHost, root->stream) != cudaSuccess ||
		    cudaStreamSynchronize(root->stream) != cudaSuccess)
			rc = vrm_fail_cuda(root, cudaGetLastError(), "vrm_render_views_sharded: copy of the gathered frames to the host");
	}
	if (views_per_scene_out) for (uint32_t i = 0; i < n_scenes; i++) views_per_scene_out[i] = devs[i].count;
	cleanup();
	if (total_ms) *total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - wall0).count();
	return rc;
}

}  // extern "C"
