// vrm_build.cu -- GPU construction of the two voxel storage structures (sm_100a).
//
// Replaces the reference's HOST builders:
//   VoxelSceneCPU::insertVoxel / generateVoxelScene   geometry/VoxelSceneCPU.cuh:16-93   (std::unordered_map per region)
//   CuckooHashTable ctor + createCuckooHashTable      storage/CuckooHashTable.cuh:20-49,97-178 (sequential eviction loop)
//   VoxelClusterStore ctor                            storage/VoxelClusterStore.cuh:37-85 (bucket + std::sort per cluster)
//   generateVoxelScene<<<1,1>>>                       renderer/Renderer.cuh:1066-1086    (serial device-side new)
//
// Pipeline (all on the handle's stream):
//   1. region min/max reduction            -> minCoord, diameter           (VoxelSceneCPU.cuh:28-35,129-130)
//   2. 64-bit key = regionIndex << 18 | clusterId << 9 | inClusterCode     (cluster-major order)
//   3. stable LSD radix sort (8-bit digits, hand-written: histogram / scan / warp-match ranked scatter)
//   4. last-write-wins dedupe (stable sort keeps insertion order inside an equal-key run; keep the run's last)
//   5. region directory: dense region index per non-empty region, cubic table index ux + uy*D + uz*D*D
//   6a. VCS: per cluster 16 x {32-bit occupancy mask, index of the word's first colour}; per region a 512-bit
//       cluster-exists mask; colours stay in sorted order, so rank = popcount prefix (no binary search)
//   6b. cuckoo: parallel insertion with 64-bit atomicExch eviction chains into per-region table pairs; regions whose
//       chain bound is hit are cleared and re-inserted with fresh seeds (CuckooHashTable.cuh:118-129 does the same
//       on the host with a 300 000-eviction bound).
// Only key -> colour (last insert wins) and cluster occupancy are observable through the lookup seam, so the
// layouts are free (SURVEY.md §7 hard part 4).
#include "vrm_internal.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include "../../include/vrm_b200.h"

#include <algorithm>
#include <cstdio>

using namespace vrm;

namespace
{

constexpr int kThreads = 256;
constexpr int kSortItems = 8;                       // keys per thread in the radix passes
constexpr int kSortTile = kThreads * kSortItems;    // 2048 keys per block
constexpr int kScanItems = 8;
constexpr int kScanTile = kThreads * kScanItems;
constexpr int kMaxEvictions = 400;
constexpr int kMaxRebuilds = 24;

struct DeviceBuf
{
	void* p = nullptr;
	size_t bytes = 0;
	~DeviceBuf() { if (p) cudaFree(p); }
	cudaError_t alloc(size_t n) { bytes = n; return cudaMalloc(&p, n ? n : 16); }
	template <class T> T* as() { return static_cast<T*>(p); }
	void* release() { void* q = p; p = nullptr; return q; }
};

// All scratch of one build comes out of ONE device allocation: a dozen cudaMalloc / cudaFree pairs of hundreds of megabytes
// cost several times the 3-4 ms the kernels of a 32 M voxel build take.
struct Arena
{
	char* base = nullptr;
	size_t cap = 0, off = 0;
	~Arena() { if (base) cudaFree(base); }
	static size_t pad(size_t bytes) { return (bytes + 255) & ~size_t(255); }
	cudaError_t reserve(size_t bytes) { cap = bytes; off = 0; return cudaMalloc(&base, bytes ? bytes : 256); }
	template <class T> T* take(size_t count)
	{
		T* p = reinterpret_cast<T*>(base + off);
		off += pad(count * sizeof(T));
		return off <= cap ? p : nullptr;
	}
};

__device__ __forceinline__ int floor_div64(int v) { return v >> 6; }  // == floorf(v / 64.0f) for |v| < 2^24 (VoxelSceneCPU.cuh:19-21)

// ---- 1. region extent ------------------------------------------------------------------------------------------
__global__ void region_minmax_kernel(const int32_t* __restrict__ xyz, uint64_t n, int* __restrict__ minmax)
{
	int lo = 0, hi = 0;  // the reference initialises both to 0 (VoxelSceneCPU.cuh:129-130)
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n * 3; i += (uint64_t)gridDim.x * blockDim.x)
	{
		int r = floor_div64(xyz[i]);
		lo = min(lo, r); hi = max(hi, r);
	}
	for (int o = 16; o; o >>= 1)
	{
		lo = min(lo, __shfl_xor_sync(0xFFFFFFFFu, lo, o));
		hi = max(hi, __shfl_xor_sync(0xFFFFFFFFu, hi, o));
	}
	if ((threadIdx.x & 31) == 0) { atomicMin(minmax, lo); atomicMax(minmax + 1, hi); }
}

__global__ void set_u32_kernel(uint32_t* p, uint32_t v) { *p = v; }

// ---- 2. keys ---------------------------------------------------------------------------------------------------
__global__ void make_keys_kernel(const int32_t* __restrict__ xyz, const uint32_t* __restrict__ rgb, uint64_t n, uint64_t dstOffset,
                                 int minCoord, uint32_t D, unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals)
{
	uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if (i >= n) return;
	int x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
	int rx = floor_div64(x), ry = floor_div64(y), rz = floor_div64(z);
	uint32_t lx = (uint32_t)(x & 63), ly = (uint32_t)(y & 63), lz = (uint32_t)(z & 63);  // ((c % 64) + 64) % 64, VoxelSceneCPU.cuh:24-26
	unsigned long long region = ((unsigned long long)(uint32_t)(rz - minCoord) * D + (uint32_t)(ry - minCoord)) * D + (uint32_t)(rx - minCoord);
	uint32_t cid = ((lx >> 3) << 6) | ((ly >> 3) << 3) | (lz >> 3);  // getVoxelClusterID, VoxelClusterStore.cuh:21-24
	uint32_t code = ((lx & 7) << 6) | ((ly & 7) << 3) | (lz & 7);
	keys[dstOffset + i] = (region << 18) | (cid << 9) | code;
	vals[dstOffset + i] = rgb[i];
}

// ---- exclusive scan (uint32), recursive three-phase -----------------------------------------------------------------
__global__ void scan_tile_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint64_t n, uint32_t* __restrict__ tileSums)
{
	__shared__ uint32_t warpSums[kThreads / 32];
	uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
	uint32_t v[kScanItems];
	uint32_t sum = 0;
#pragma unroll
	for (int k = 0; k < kScanItems; k++) { v[k] = base + k < n ? in[base + k] : 0u; sum += v[k]; }
	uint32_t incl = sum;
	int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
	if (lane == 31) warpSums[warp] = incl;
	__syncthreads();
	if (warp == 0)
	{
		uint32_t w = lane < kThreads / 32 ? warpSums[lane] : 0u;
		uint32_t wi = w;
		for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xFFFFFFFFu, wi, o); if (lane >= o) wi += t; }
		if (lane < kThreads / 32) warpSums[lane] = wi - w;
		if (lane == kThreads / 32 - 1 && tileSums) tileSums[blockIdx.x] = wi;
	}
	__syncthreads();
	uint32_t run = warpSums[warp] + incl - sum;
#pragma unroll
	for (int k = 0; k < kScanItems; k++) { if (base + k < n) out[base + k] = run; run += v[k]; }
}

__global__ void scan_add_kernel(uint32_t* __restrict__ out, uint64_t n, const uint32_t* __restrict__ tileOffsets)
{
	uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
	uint32_t add = tileOffsets[blockIdx.x];
#pragma unroll
	for (int k = 0; k < kScanItems; k++) if (base + k < n) out[base + k] += add;
}

size_t scan_scratch_elems(uint64_t n)
{
	size_t total = 0;
	while (n > (uint64_t)kScanTile) { n = (n + kScanTile - 1) / kScanTile; total += n; }
	return total + 1;
}

// out may alias in.  scratch must hold scan_scratch_elems(n) uint32.
void exclusive_scan(const uint32_t* in, uint32_t* out, uint64_t n, uint32_t* scratch, cudaStream_t st)
{
	if (n == 0) return;
	uint64_t tiles = (n + kScanTile - 1) / kScanTile;
	if (tiles == 1) { scan_tile_kernel<<<1, kThreads, 0, st>>>(in, out, n, nullptr); return; }
	scan_tile_kernel<<<(unsigned)tiles, kThreads, 0, st>>>(in, out, n, scratch);
	exclusive_scan(scratch, scratch, tiles, scratch + tiles, st);
	scan_add_kernel<<<(unsigned)tiles, kThreads, 0, st>>>(out, n, scratch);
}

// ---- 3. radix sort ---------------------------------------------------------------------------------------------
__global__ void radix_hist_kernel(const unsigned long long* __restrict__ keys, uint64_t n, int shift, uint32_t numTiles, uint32_t* __restrict__ hist)
{
	__shared__ uint32_t h[256];
	h[threadIdx.x] = 0;
	__syncthreads();
	uint64_t base = (uint64_t)blockIdx.x * kSortTile;
#pragma unroll
	for (int k = 0; k < kSortItems; k++)
	{
		uint64_t i = base + (uint64_t)k * kThreads + threadIdx.x;
		if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 255u], 1u);
	}
	__syncthreads();
	hist[(uint64_t)threadIdx.x * numTiles + blockIdx.x] = h[threadIdx.x];  // digit-major so that one scan orders (digit, tile)
}

// Stable scatter: warp w of a tile owns keys [w*256, w*256+256) of the tile and walks them in rounds of 32 consecutive
// keys, so (warp, round, lane) is the input order.  __match_any_sync groups equal digits inside a round.
__global__ void radix_scatter_kernel(const unsigned long long* __restrict__ keysIn, const uint32_t* __restrict__ valsIn, uint64_t n, int shift,
                                     uint32_t numTiles, const uint32_t* __restrict__ histScanned,
                                     unsigned long long* __restrict__ keysOut, uint32_t* __restrict__ valsOut)
{
	__shared__ uint32_t cnt[kThreads / 32][256];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (int i = threadIdx.x; i < (kThreads / 32) * 256; i += kThreads) (&cnt[0][0])[i] = 0;
	__syncthreads();
	const uint64_t warpBase = (uint64_t)blockIdx.x * kSortTile + (uint64_t)warp * (32 * kSortItems);
	unsigned long long key[kSortItems];
	uint32_t rank[kSortItems];
#pragma unroll
	for (int k = 0; k < kSortItems; k++)
	{
		uint64_t i = warpBase + (uint64_t)k * 32 + lane;
		bool valid = i < n;
		key[k] = valid ? keysIn[i] : 0ull;
		uint32_t d = valid ? ((uint32_t)(key[k] >> shift) & 255u) : 0xFFFFFFFFu;
		uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);
		int leader = __ffs(peers) - 1;
		uint32_t before = __popc(peers & ((1u << lane) - 1u));
		uint32_t prev = 0;
		if (valid && lane == leader) { prev = cnt[warp][d]; cnt[warp][d] = prev + __popc(peers); }
		prev = __shfl_sync(0xFFFFFFFFu, prev, leader);
		rank[k] = prev + before;
		__syncwarp();
	}
	__syncthreads();
	{
		// exclusive prefix over the warps of this tile, per digit; add the tile's global offset for the digit
		uint32_t d = threadIdx.x;
		uint32_t run = histScanned[(uint64_t)d * numTiles + blockIdx.x];
		for (int w = 0; w < kThreads / 32; w++) { uint32_t c = cnt[w][d]; cnt[w][d] = run; run += c; }
	}
	__syncthreads();
#pragma unroll
	for (int k = 0; k < kSortItems; k++)
	{
		uint64_t i = warpBase + (uint64_t)k * 32 + lane;
		if (i < n)
		{
			uint32_t d = (uint32_t)(key[k] >> shift) & 255u;
			uint32_t dst = cnt[warp][d] + rank[k];
			keysOut[dst] = key[k];
			valsOut[dst] = valsIn[i];
		}
	}
}

// ---- 4. dedupe -------------------------------------------------------------------------------------------------
__global__ void keep_last_kernel(const unsigned long long* __restrict__ keys, uint64_t n, uint32_t* __restrict__ keep)
{
	uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if (i >= n) return;
	keep[i] = (i + 1 == n || keys[i + 1] != keys[i]) ? 1u : 0u;  // last write wins, VoxelSceneCPU.cuh:45
}

__global__ void compact_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ vals, const uint32_t* __restrict__ keep,
                               const uint32_t* __restrict__ pos, uint64_t n, unsigned long long* __restrict__ ukeys, uint32_t* __restrict__ uvals)
{
	uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if (i >= n || !keep[i]) return;
	ukeys[pos[i]] = keys[i];
	uvals[pos[i]] = vals[i];
}

// ---- 5. region directory ---------------------------------------------------------------------------------------
__global__ void region_head_kernel(const unsigned long long* __restrict__ ukeys, uint64_t u, uint32_t* __restrict__ head)
{
	uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if (j >= u) return;
	head[j] = (j == 0 || (ukeys[j] >> 18) != (ukeys[j - 1] >> 18)) ? 1u : 0u;
}

// headScan = exclusive scan of head.  regionOf[j] = dense region index of voxel j.
__global__ void region_dir_kernel(const unsigned long long* __restrict__ ukeys, const uint32_t* __restrict__ head, const uint32_t* __restrict__ headScan,
                                  uint64_t u, int32_t* __restrict__ regionTable, uint32_t* __restrict__ regionStart, uint32_t* __restrict__ regionOf)
{
	uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if (j >= u) return;
	uint32_t ri = headScan[j] + head[j] - 1u;
	regionOf[j] = ri;
	if (head[j])
	{
		regionTable[ukeys[j] >> 18] = (int32_t)ri;  // VoxelSceneCPU.cuh:61-62 (index), Renderer.cuh:1075 (entry)
		regionStart[ri] = (uint32_t)j;
	}
}

// ---- 6a. voxel cluster store -----------------------------------------------------------------------------------
__global__ void vcs_fill_kernel(const unsigned long long* __restrict__ ukeys, const uint32_t* __restrict__ regionOf, uint64_t u,
                                uint2* __restrict__ headers, uint32_t* __restrict__ clusterMask)
{
	uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if (j >= u) return;
	unsigned long long k = ukeys[j];
	// The voxels of one 32-bit occupancy word are contiguous in the sorted array: its first voxel owns the word and
	// gathers the (at most 32) bits itself -- no atomics, no contention on the densely filled words of solid terrain.
	if (j != 0 && (ukeys[j - 1] >> 5) == (k >> 5)) return;
	uint32_t ri = regionOf[j];
	uint32_t cid = (uint32_t)(k >> 9) & 511u, code = (uint32_t)k & 511u;
	uint32_t mask = 1u << (code & 31);
	for (uint64_t t = j + 1; t < u && t < j + 32; t++)
	{
		unsigned long long kt = ukeys[t];
		if ((kt >> 5) != (k >> 5)) break;
		mask |= 1u << ((uint32_t)kt & 31);
	}
	// colours are sorted by (region, cluster, code): rank inside the word = popcount below the bit
	headers[((size_t)ri * 512 + cid) * 16 + (code >> 5)] = make_uint2(mask, (uint32_t)j);
	bool firstOfCluster = j == 0 || (ukeys[j - 1] >> 9) != (k >> 9);
	if (firstOfCluster) atomicOr(clusterMask + (size_t)ri * 16 + (cid >> 5), 1u << (cid & 31));
}

// Every occupancy word of a cluster that holds at least one voxel carries kHeaderClusterExists in its colour index, so that one
// 8-byte header load answers doesClusterExist AND the occupancy test (vrm_flat.cuh voxel_test).  One thread per cluster.
__global__ void vcs_flag_clusters_kernel(uint2* __restrict__ headers, const uint32_t* __restrict__ clusterMask, uint64_t numClusters)
{
	uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if (c >= numClusters) return;
	if (!((clusterMask[c >> 5] >> (c & 31)) & 1u)) return;
	uint4* words = reinterpret_cast<uint4*>(headers + c * 16);
	for (int w = 0; w < 8; w++)
	{
		uint4 v = words[w];
		v.y |= kHeaderClusterExists; v.w |= kHeaderClusterExists;
		words[w] = v;
	}
}

// Hash storage: the same 512-bit cluster-occupancy mask per region, used as a NEGATIVE FILTER in front of the two probes
// (vrm_core.cuh lookup_voxel): a voxel of a cluster without voxels cannot be in the table.
__global__ void cluster_mask_kernel(const unsigned long long* __restrict__ ukeys, const uint32_t* __restrict__ regionOf, uint64_t u, uint32_t* __restrict__ clusterMask)
{
	uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if (j >= u) return;
	unsigned long long k = ukeys[j];
	if (j != 0 && (ukeys[j - 1] >> 9) == (k >> 9)) return;  // one thread per (region, cluster): the sorted keys keep a cluster's voxels together
	uint32_t cid = (uint32_t)(k >> 9) & 511u;
	atomicOr(clusterMask + (size_t)regionOf[j] * 16 + (cid >> 5), 1u << (cid & 31));
}

// ---- 6b. cuckoo hash table -------------------------------------------------------------------------------------
__global__ void fill_slots_kernel(unsigned long long* __restrict__ slots, uint64_t n)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) slots[i] = kEmptySlot;
}

__global__ void clear_failed_regions_kernel(unsigned long long* __restrict__ slots, const HashRegionDesc* __restrict__ desc, const uint32_t* __restrict__ retry, uint32_t numRegions)
{
	// one block per region
	uint32_t ri = blockIdx.x;
	if (ri >= numRegions || !retry[ri]) return;
	HashRegionDesc d = desc[ri];
	for (uint32_t i = threadIdx.x; i < 2 * d.n; i += blockDim.x) slots[(size_t)d.slotBase + i] = kEmptySlot;
}

__global__ void cuckoo_insert_kernel(const unsigned long long* __restrict__ ukeys, const uint32_t* __restrict__ uvals, const uint32_t* __restrict__ regionOf, uint64_t u,
                                     const HashRegionDesc* __restrict__ desc, unsigned long long* __restrict__ slots,
                                     const uint32_t* __restrict__ retry /* nullable: insert everything */, uint32_t* __restrict__ failed)
{
	uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if (j >= u) return;
	uint32_t ri = regionOf[j];
	if (retry && !retry[ri]) return;
	HashRegionDesc d = desc[ri];
	unsigned long long k = ukeys[j];
	uint32_t cid = (uint32_t)(k >> 9) & 511u, code = (uint32_t)k & 511u;
	uint32_t x = ((cid >> 6) << 3) | (code >> 6), y = (((cid >> 3) & 7u) << 3) | ((code >> 3) & 7u), z = ((cid & 7u) << 3) | (code & 7u);
	unsigned long long entry = ((unsigned long long)vrm::hash_key(x, y, z) << 32) | uvals[j];
	unsigned long long* t1 = slots + d.slotBase;
	unsigned long long* t2 = t1 + d.n;
	int table = 0;
	for (int it = 0; it < kMaxEvictions; it++)
	{
		uint32_t key = (uint32_t)(entry >> 32);
		unsigned long long* slot = table == 0 ? t1 + hash_slot1(key, d.seed1, d.n) : t2 + hash_slot2(key, d.seed2, d.n);
		entry = atomicExch(slot, entry);  // evict whoever lives there (CuckooHashTable.cuh:137-151), atomically
		if (entry == kEmptySlot) return;
		table ^= 1;                       // the evicted entry moves to its other table
	}
	failed[ri] = 1u;  // cycle bound hit: the region is rebuilt with new hash seeds (CuckooHashTable.cuh:118-129)
}

uint32_t seed_for(uint32_t ri, uint32_t attempt, uint32_t which)
{
	uint32_t h = ri * 0x9E3779B1u + attempt * 0x85EBCA77u + which * 0xC2B2AE3Du + 0x27D4EB2Fu;
	h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
	return h | 1u;  // multiply-shift hashing needs an odd multiplier (vrm_core.cuh hash_slot1/2)
}

unsigned grid_for(uint64_t n) { return (unsigned)((n + kThreads - 1) / kThreads); }

}  // namespace

void vrm_free_structure(vrm_scene* s)
{
	cudaFree(s->d_regionTable); s->d_regionTable = nullptr;
	cudaFree(s->d_hashDesc); s->d_hashDesc = nullptr;
	cudaFree(s->d_slots); s->d_slots = nullptr;
	cudaFree(s->d_headers); s->d_headers = nullptr;
	cudaFree(s->d_clusterMask); s->d_clusterMask = nullptr;
	cudaFree(s->d_values); s->d_values = nullptr;
	s->storage = -1; s->bytes = 0; s->filled = 0; s->unique = 0;
}

// VRM_BUILD_TRACE=1: host-clock stage marks on stderr (each mark synchronises the stream, so traced builds are slower)
struct BuildTrace
{
	bool on = getenv("VRM_BUILD_TRACE") != nullptr;
	std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
	void mark(cudaStream_t st, const char* what)
	{
		if (!on) return;
		cudaStreamSynchronize(st);
		auto t1 = std::chrono::steady_clock::now();
		fprintf(stderr, "[vrm build] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
		t0 = t1;
	}
};

int vrm_build_structure(vrm_scene* s, int storageType, float* buildMs)
{
	cudaStream_t st = s->stream;
	BuildTrace trace;
	const uint64_t n = s->nStaged;
	if (n >= (1ull << 32) - (uint64_t)kSortTile) { s->lastError = "too many voxels for 32-bit indices"; return VRM_ERR_INVALID; }
	VRM_CUDA(s, cudaEventRecord(s->ev0, st));

	// 1. extent
	DeviceBuf minmax;
	VRM_CUDA(s, minmax.alloc(2 * sizeof(int)));
	VRM_CUDA(s, cudaMemsetAsync(minmax.p, 0, 2 * sizeof(int), st));
	for (const VoxelChunk& c : s->chunks)
		if (c.n) region_minmax_kernel<<<(unsigned)std::min<uint64_t>((c.n * 3 + kThreads - 1) / kThreads, 148 * 16), kThreads, 0, st>>>(c.d_xyz, c.n, minmax.as<int>());
	int hostMinMax[2] = {0, 0};
	VRM_CUDA(s, cudaMemcpyAsync(hostMinMax, minmax.p, sizeof(hostMinMax), cudaMemcpyDeviceToHost, st));
	VRM_CUDA(s, cudaStreamSynchronize(st));
	trace.mark(st, "extent");
	const int minCoord = hostMinMax[0];
	const uint64_t D = (uint64_t)(hostMinMax[1] - hostMinMax[0] + 1);
	if (D > 1024) { s->lastError = "region table diameter > 1024 (scene spans more than 65536 voxels)"; return VRM_ERR_INVALID; }
	const uint64_t tableSize = D * D * D;
	int regionBits = 0;
	while ((1ull << regionBits) < tableSize) regionBits++;
	const int keyBits = 18 + regionBits;

	DeviceBuf regionTable;
	if (regionTable.alloc(tableSize * sizeof(int32_t)) != cudaSuccess) { cudaGetLastError(); s->lastError = "region table allocation failed"; return VRM_ERR_NOMEM; }
	VRM_CUDA(s, cudaMemsetAsync(regionTable.p, 0xFF, tableSize * sizeof(int32_t), st));  // -1 = empty region (the reference forgets to clear this)

	uint64_t unique = 0;
	uint32_t numRegions = 0;
	Arena arena;
	DeviceBuf uvalsBuf, regionStart;
	unsigned long long* ukeys = nullptr;
	uint32_t* uvals = nullptr;
	uint32_t* regionOf = nullptr;
	if (n > 0)
	{
		// 2. keys
		const uint32_t numTiles = (uint32_t)((n + kSortTile - 1) / kSortTile);
		const uint64_t histElems = 256ull * numTiles;
		const size_t scratchElems = scan_scratch_elems(std::max<uint64_t>(histElems, n));
		const size_t arenaBytes = 2 * Arena::pad(n * 8) + 2 * Arena::pad(n * 4) + Arena::pad(histElems * 4) + Arena::pad(scratchElems * 4) +
		                          Arena::pad(n * 4) /* pos */ + Arena::pad(n * 8) /* unique keys */ + Arena::pad(n * 4) /* regionOf */;
		if (arena.reserve(arenaBytes) != cudaSuccess) { cudaGetLastError(); s->lastError = "build scratch allocation failed"; return VRM_ERR_NOMEM; }
		trace.mark(st, "arena + region table alloc");
		unsigned long long* keysA = arena.take<unsigned long long>(n);
		unsigned long long* keysB = arena.take<unsigned long long>(n);
		uint32_t* valsA = arena.take<uint32_t>(n);
		uint32_t* valsB = arena.take<uint32_t>(n);
		uint32_t* hist = arena.take<uint32_t>(histElems);
		uint32_t* scratch = arena.take<uint32_t>(scratchElems);
		uint32_t* pos = arena.take<uint32_t>(n);
		ukeys = arena.take<unsigned long long>(n);
		regionOf = arena.take<uint32_t>(n);
		if (!regionOf) { s->lastError = "build scratch arena too small"; return VRM_ERR_NOMEM; }
		uint64_t off = 0;
		for (const VoxelChunk& c : s->chunks)
		{
			if (!c.n) continue;
			make_keys_kernel<<<grid_for(c.n), kThreads, 0, st>>>(c.d_xyz, c.d_rgb, c.n, off, minCoord, (uint32_t)D, keysA, valsA);
			off += c.n;
		}
		trace.mark(st, "keys");
		// 3. stable LSD radix sort
		unsigned long long* kin = keysA; unsigned long long* kout = keysB;
		uint32_t* vin = valsA; uint32_t* vout = valsB;
		for (int shift = 0; shift < keyBits; shift += 8)
		{
			radix_hist_kernel<<<numTiles, kThreads, 0, st>>>(kin, n, shift, numTiles, hist);
			exclusive_scan(hist, hist, histElems, scratch, st);
			radix_scatter_kernel<<<numTiles, kThreads, 0, st>>>(kin, vin, n, shift, numTiles, hist, kout, vout);
			std::swap(kin, kout); std::swap(vin, vout);
		}
		trace.mark(st, "radix sort");
		// 4. dedupe (kout / vout are free now: vout holds the keep flags)
		uint32_t* keep = vout;
		keep_last_kernel<<<grid_for(n), kThreads, 0, st>>>(kin, n, keep);
		exclusive_scan(keep, pos, n, scratch, st);
		uint32_t lastPos = 0, lastKeep = 0;
		VRM_CUDA(s, cudaMemcpyAsync(&lastPos, pos + (n - 1), 4, cudaMemcpyDeviceToHost, st));
		VRM_CUDA(s, cudaMemcpyAsync(&lastKeep, keep + (n - 1), 4, cudaMemcpyDeviceToHost, st));
		VRM_CUDA(s, cudaStreamSynchronize(st));
		unique = (uint64_t)lastPos + lastKeep;
		// the colours outlive the build (VCS: the values array), so they get their own, exactly sized allocation
		if (uvalsBuf.alloc(unique * 4) != cudaSuccess) { cudaGetLastError(); s->lastError = "unique voxel allocation failed"; return VRM_ERR_NOMEM; }
		uvals = uvalsBuf.as<uint32_t>();
		compact_kernel<<<grid_for(n), kThreads, 0, st>>>(kin, vin, keep, pos, n, ukeys, uvals);
		trace.mark(st, "dedupe + compact");
		// 5. region directory (reuse pos / keep as head-scan / head flags: unique <= n)
		uint32_t* head = keep;
		uint32_t* headScan = pos;
		region_head_kernel<<<grid_for(unique), kThreads, 0, st>>>(ukeys, unique, head);
		exclusive_scan(head, headScan, unique, scratch, st);
		uint32_t lastScan = 0, lastHead = 0;
		VRM_CUDA(s, cudaMemcpyAsync(&lastScan, headScan + (unique - 1), 4, cudaMemcpyDeviceToHost, st));
		VRM_CUDA(s, cudaMemcpyAsync(&lastHead, head + (unique - 1), 4, cudaMemcpyDeviceToHost, st));
		VRM_CUDA(s, cudaStreamSynchronize(st));
		numRegions = lastScan + lastHead;
		if (regionStart.alloc(((size_t)numRegions + 1) * 4) != cudaSuccess) { cudaGetLastError(); s->lastError = "region directory allocation failed"; return VRM_ERR_NOMEM; }
		region_dir_kernel<<<grid_for(unique), kThreads, 0, st>>>(ukeys, head, headScan, unique, regionTable.as<int32_t>(), regionStart.as<uint32_t>(), regionOf);
		// regionStart[numRegions] = unique, written from the device (no host staging variable to keep alive)
		set_u32_kernel<<<1, 1, 0, st>>>(regionStart.as<uint32_t>() + numRegions, (uint32_t)unique);
		VRM_CUDA(s, cudaGetLastError());
	}

	trace.mark(st, "region directory");
	uint64_t bytes = tableSize * sizeof(int32_t);
	DeviceBuf headers, clusterMask, hashDesc, slots;
	if (storageType == VRM_STORAGE_VCS)
	{
		const size_t headerBytes = (size_t)numRegions * 512 * 16 * sizeof(uint2);
		// 31-bit colour indices (the top bit of a header's index word is the cluster-exists flag), 32-bit header word indices
		if (unique >= (1ull << 31) || (uint64_t)numRegions * 8192ull >= (1ull << 32)) { s->lastError = "scene too large for the VCS index widths"; return VRM_ERR_INVALID; }
		// Zeroed guard space behind both arrays: the reference can test a voxel with a coordinate of exactly 64 (a ray rebased onto the
		// far face of a region; undefined behaviour there, VoxelClusterStore.cuh:93-99 indexes past its 512-entry table) and the nested
		// traversal then forms a cluster id of up to 575 -- for the LAST region that reads up to 64 clusters past the end.
		const size_t headerGuard = 64 * 16 * sizeof(uint2), maskGuard = 64;
		if (headers.alloc(headerBytes + headerGuard) != cudaSuccess || clusterMask.alloc((size_t)numRegions * 16 * 4 + maskGuard) != cudaSuccess)
		{ cudaGetLastError(); s->lastError = "VCS allocation failed"; return VRM_ERR_NOMEM; }
		VRM_CUDA(s, cudaMemsetAsync(headers.p, 0, headerBytes + headerGuard, st));
		VRM_CUDA(s, cudaMemsetAsync(clusterMask.p, 0, (size_t)numRegions * 16 * 4 + maskGuard, st));
		if (unique)
			vcs_fill_kernel<<<grid_for(unique), kThreads, 0, st>>>(ukeys, regionOf, unique, headers.as<uint2>(), clusterMask.as<uint32_t>());
		if (unique)
			vcs_flag_clusters_kernel<<<grid_for((uint64_t)numRegions * 512), kThreads, 0, st>>>(headers.as<uint2>(), clusterMask.as<uint32_t>(), (uint64_t)numRegions * 512);
		VRM_CUDA(s, cudaGetLastError());
		bytes += headerBytes + (size_t)numRegions * 64 + unique * 4;
	}
	else
	{
		std::vector<uint32_t> starts((size_t)numRegions + 1, 0);
		if (numRegions) VRM_CUDA(s, cudaMemcpyAsync(starts.data(), regionStart.p, starts.size() * 4, cudaMemcpyDeviceToHost, st));
		VRM_CUDA(s, cudaStreamSynchronize(st));
		std::vector<HashRegionDesc> desc(numRegions);
		uint64_t totalSlots = 0;
		for (uint32_t r = 0; r < numRegions; r++)
		{
			uint32_t cnt = starts[r + 1] - starts[r];
			desc[r].n = cnt + cnt / 4 + 2;  // 1.25 N slots per table as in the reference (CuckooHashTable.cuh:23), +2 so tiny regions can always place
			desc[r].slotBase = (uint32_t)totalSlots;
			desc[r].seed1 = seed_for(r, 0, 1); desc[r].seed2 = seed_for(r, 0, 2);
			totalSlots += 2ull * desc[r].n;
		}
		if (totalSlots >= (1ull << 32)) { s->lastError = "hash table needs more than 2^32 slots"; return VRM_ERR_INVALID; }
		DeviceBuf failed, retry;
		if (clusterMask.alloc((size_t)numRegions * 16 * 4 + 64) != cudaSuccess) { cudaGetLastError(); s->lastError = "hash table allocation failed"; return VRM_ERR_NOMEM; }
		VRM_CUDA(s, cudaMemsetAsync(clusterMask.p, 0, (size_t)numRegions * 16 * 4 + 64, st));
		if (unique) cluster_mask_kernel<<<grid_for(unique), kThreads, 0, st>>>(ukeys, regionOf, unique, clusterMask.as<uint32_t>());
		if (hashDesc.alloc((size_t)numRegions * sizeof(HashRegionDesc)) != cudaSuccess || slots.alloc(totalSlots * 8) != cudaSuccess ||
		    failed.alloc((size_t)numRegions * 4) != cudaSuccess || retry.alloc((size_t)numRegions * 4) != cudaSuccess)
		{ cudaGetLastError(); s->lastError = "hash table allocation failed"; return VRM_ERR_NOMEM; }
		if (numRegions)
		{
			VRM_CUDA(s, cudaMemcpyAsync(hashDesc.p, desc.data(), desc.size() * sizeof(HashRegionDesc), cudaMemcpyHostToDevice, st));
			fill_slots_kernel<<<148 * 8, kThreads, 0, st>>>(slots.as<unsigned long long>(), totalSlots);
			std::vector<uint32_t> hostFailed(numRegions);
			bool all = true;
			int attempt = 0;
			for (;; attempt++)
			{
				VRM_CUDA(s, cudaMemsetAsync(failed.p, 0, (size_t)numRegions * 4, st));
				cuckoo_insert_kernel<<<grid_for(unique), kThreads, 0, st>>>(ukeys, uvals, regionOf, unique,
				                                                           hashDesc.as<HashRegionDesc>(), slots.as<unsigned long long>(), all ? nullptr : retry.as<uint32_t>(), failed.as<uint32_t>());
				VRM_CUDA(s, cudaMemcpyAsync(hostFailed.data(), failed.p, (size_t)numRegions * 4, cudaMemcpyDeviceToHost, st));
				VRM_CUDA(s, cudaStreamSynchronize(st));
				bool any = false;
				for (uint32_t r = 0; r < numRegions; r++)
					if (hostFailed[r]) { any = true; desc[r].seed1 = seed_for(r, attempt + 1, 1); desc[r].seed2 = seed_for(r, attempt + 1, 2); }
				if (trace.on) { uint32_t nf = 0; for (uint32_t r = 0; r < numRegions; r++) nf += hostFailed[r] ? 1u : 0u; fprintf(stderr, "[vrm build] cuckoo attempt %d: %u regions failed\n", attempt, nf); }
				if (!any) break;
				if (attempt + 1 >= kMaxRebuilds) { s->lastError = "cuckoo insertion did not converge"; return VRM_ERR_BUILD; }
				VRM_CUDA(s, cudaMemcpyAsync(hashDesc.p, desc.data(), desc.size() * sizeof(HashRegionDesc), cudaMemcpyHostToDevice, st));
				VRM_CUDA(s, cudaMemcpyAsync(retry.p, hostFailed.data(), (size_t)numRegions * 4, cudaMemcpyHostToDevice, st));
				clear_failed_regions_kernel<<<numRegions, kThreads, 0, st>>>(slots.as<unsigned long long>(), hashDesc.as<HashRegionDesc>(), retry.as<uint32_t>(), numRegions);
				all = false;
			}
			VRM_CUDA(s, cudaGetLastError());
		}
		bytes += (size_t)numRegions * sizeof(HashRegionDesc) + totalSlots * 8 + (size_t)numRegions * 64;
	}
	trace.mark(st, storageType == VRM_STORAGE_VCS ? "vcs tables" : "cuckoo insertion");
	VRM_CUDA(s, cudaEventRecord(s->ev1, st));
	VRM_CUDA(s, cudaStreamSynchronize(st));
	VRM_CUDA(s, cudaGetLastError());
	if (buildMs) VRM_CUDA(s, cudaEventElapsedTime(buildMs, s->ev0, s->ev1));

	s->d_regionTable = static_cast<int32_t*>(regionTable.release());
	if (storageType == VRM_STORAGE_VCS)
	{
		s->d_headers = static_cast<uint2*>(headers.release());
		s->d_clusterMask = static_cast<uint32_t*>(clusterMask.release());
		s->d_values = static_cast<uint32_t*>(uvalsBuf.release());
	}
	else
	{
		s->d_hashDesc = static_cast<HashRegionDesc*>(hashDesc.release());
		s->d_clusterMask = static_cast<uint32_t*>(clusterMask.release());
		s->d_slots = static_cast<unsigned long long*>(slots.release());
	}
	s->storage = storageType;
	s->diameter = (uint32_t)D;
	s->minCoord = minCoord;
	s->filled = numRegions;
	s->unique = unique;
	s->bytes = bytes;
	return VRM_OK;
}

// exclusive scan for the other translation units (vrm_generate.cu)
size_t vrm_scan_scratch_elems(uint64_t n) { return scan_scratch_elems(n); }
void vrm_exclusive_scan_u32(const uint32_t* in, uint32_t* out, uint64_t n, uint32_t* scratch, cudaStream_t st) { exclusive_scan(in, out, n, scratch, st); }
