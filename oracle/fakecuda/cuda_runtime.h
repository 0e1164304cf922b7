// TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product.
//
// Minimal stand-in for <cuda_runtime.h> so that the UNMODIFIED reference headers under
// /root/reference/VoxelRaymarcher/src compile as plain host C++ (g++ -ffp-contract=off).
// Recipe follows SURVEY.md §8c "CPU oracle".  Only what the reference's hot-path headers touch
// is provided: execution-space qualifiers, the built-in index variables (thread_local so that
// rows can be rendered from several std::threads), min/max with CUDA's fminf/fmaxf NaN
// semantics, and malloc-backed cudaMalloc/cudaMemcpy/cudaFree.
#pragma once
#include <cmath>
#include <math.h>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

#define __host__
#define __device__
#define __global__
#define __constant__
#define __forceinline__ inline

struct uint3 { unsigned int x, y, z; };
struct dim3 { unsigned int x = 1, y = 1, z = 1; dim3() {} dim3(unsigned a, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };

extern thread_local uint3 threadIdx;
extern thread_local uint3 blockIdx;
extern thread_local dim3 blockDim;

namespace std {
using ::floorf;
using ::ceilf;
using ::copysignf;
}

// CUDA device min/max on floats are fminf/fmaxf (NaN-suppressing), not std::min/std::max.
static inline float min(float a, float b) { return fminf(a, b); }
static inline float max(float a, float b) { return fmaxf(a, b); }

typedef int cudaError_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
template <typename T> static inline cudaError_t cudaMalloc(T** p, size_t n) { *p = static_cast<T*>(malloc(n ? n : 1)); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
template <typename T> static inline cudaError_t cudaMemcpyToSymbol(T& sym, const void* src, size_t n) { memcpy((void*)&sym, src, n); return 0; }
