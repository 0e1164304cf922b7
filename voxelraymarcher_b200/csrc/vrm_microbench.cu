// vrm_microbench.cu -- the roofs the traversal kernels are measured against that MEASURED_PEAKS.json does not hold (VERDICT r01
// item 3; SURVEY.md 8d: "micro-benchmark L2 GB/s").  The render kernels gather 8-byte cluster headers / hash slots from a
// working set that lives in L2, so the memory roof that can bind them is the L2's rate for 8-byte random gathers, not the HBM copy
// rate.  Two kernels, both timed with CUDA events on their own stream after a warm-up pass that makes the buffer L2-resident:
//   gather8  every thread reads 8-byte words at pseudo-random indices of the working set (one 32-byte sector per load)
//   stream   every thread reads consecutive 16-byte words, the whole buffer over and over (coalesced: the L2's peak read rate)
// Measurement infrastructure: nothing on the product path calls it.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vrm_b200.h"

namespace
{

__global__ void gather8_kernel(const uint2* __restrict__ words, uint32_t mask, uint32_t loadsPerThread, uint32_t* __restrict__ sink)
{
	uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
	uint32_t acc = 0;
#pragma unroll 8
	for (uint32_t i = 0; i < loadsPerThread; i++)
	{
		x = x * 1664525u + 1013904223u;            // independent of the loaded value: the loads of a thread overlap, like the
		const uint2 v = __ldg(words + ((x >> 7) & mask));  // two hash probes / the header loads of neighbouring lanes do
		acc ^= v.x + v.y;
	}
	if (acc == 0x9E3779B9u) sink[0] = acc;  // never true for the zero-filled buffer's complement pattern; keeps the loads alive
}

__global__ void stream_kernel(const uint4* __restrict__ words, uint64_t n, uint32_t passes, uint32_t* __restrict__ sink)
{
	uint32_t acc = 0;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint32_t p = 0; p < passes; p++)
		for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += stride)
		{
			const uint4 v = __ldg(words + i);
			acc ^= v.x + v.y + v.z + v.w;
		}
	if (acc == 0x9E3779B9u) sink[0] = acc;
}

__global__ void fill_kernel(uint32_t* p, uint64_t n)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = (uint32_t)i * 2246822519u + 1u;
}

}  // namespace

extern "C" int vrm_microbench_l2(int device, uint64_t working_set_bytes, float* gather8_loads_per_ns, float* gather8_gbs, float* stream_gbs)
{
	if (working_set_bytes < (1u << 16) || working_set_bytes > (1ull << 32)) return VRM_ERR_INVALID;
	if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return VRM_ERR_CUDA; }
	// power-of-two number of 8-byte words
	uint64_t words = 1;
	while (words * 2 * 8 <= working_set_bytes) words *= 2;
	const uint64_t bytes = words * 8;
	void* buf = nullptr; uint32_t* sink = nullptr;
	cudaStream_t st = nullptr;
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
	int rc = VRM_OK;
	float msGather = 0.0f, msStream = 0.0f;
	const uint32_t loadsPerThread = 256, passes = 64;
	const unsigned blocks = (unsigned)sms * 8u, threads = 256u;
	if (cudaMalloc(&buf, bytes) != cudaSuccess || cudaMalloc(&sink, 4) != cudaSuccess || cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess ||
	    cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) { rc = VRM_ERR_CUDA; goto done; }
	fill_kernel<<<blocks, threads, 0, st>>>(static_cast<uint32_t*>(buf), bytes / 4);
	// warm-up: pulls the buffer into L2 (and the kernels into the instruction cache)
	gather8_kernel<<<blocks, threads, 0, st>>>(static_cast<const uint2*>(buf), (uint32_t)(words - 1), loadsPerThread, sink);
	stream_kernel<<<blocks, threads, 0, st>>>(static_cast<const uint4*>(buf), bytes / 16, 2, sink);
	cudaEventRecord(e0, st);
	gather8_kernel<<<blocks, threads, 0, st>>>(static_cast<const uint2*>(buf), (uint32_t)(words - 1), loadsPerThread, sink);
	cudaEventRecord(e1, st);
	if (cudaStreamSynchronize(st) != cudaSuccess || cudaEventElapsedTime(&msGather, e0, e1) != cudaSuccess) { rc = VRM_ERR_CUDA; goto done; }
	cudaEventRecord(e0, st);
	stream_kernel<<<blocks, threads, 0, st>>>(static_cast<const uint4*>(buf), bytes / 16, passes, sink);
	cudaEventRecord(e1, st);
	if (cudaStreamSynchronize(st) != cudaSuccess || cudaEventElapsedTime(&msStream, e0, e1) != cudaSuccess) { rc = VRM_ERR_CUDA; goto done; }
	{
		const double loads = (double)blocks * threads * loadsPerThread;
		if (gather8_loads_per_ns) *gather8_loads_per_ns = (float)(loads / (msGather * 1e6));
		if (gather8_gbs) *gather8_gbs = (float)(loads * 8.0 / (msGather * 1e6));      // algorithmic bytes (8 per load); the L2 moves a 32-byte sector for each
		if (stream_gbs) *stream_gbs = (float)((double)bytes * passes / (msStream * 1e6));
	}
done:
	if (rc != VRM_OK) cudaGetLastError();
	if (e0) cudaEventDestroy(e0);
	if (e1) cudaEventDestroy(e1);
	if (st) cudaStreamDestroy(st);
	cudaFree(buf); cudaFree(sink);
	return rc;
}
