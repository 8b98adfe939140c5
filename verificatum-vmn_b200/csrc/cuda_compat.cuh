// cuda_compat.cuh -- one launch macro for the engine, plus a HOST EMULATION of the tiny CUDA
// subset the engine uses.
//
// Product build (nvcc, sm_100a): VMX_LAUNCH expands to a plain <<<grid, block, smem, stream>>>
// launch and counts it.
//
// Test build (g++ -DVMX_HOST_EMUL, tests/host_emul only): there is no GPU in the development
// container, so the host-side orchestration (bucket sort, segmented products, scans, table
// construction, byte codecs) is exercised by running every kernel body sequentially on the
// CPU -- one call per (block, thread).  This exists to debug HOST LOGIC without spending
// GPU time; it is never built into, linked with or loaded by the shipped library, and the
// shipped library has no CPU path at all.  Kernels that need warp intrinsics or barriers
// provide a sequential stand-in under #ifdef VMX_HOST_EMUL.
#pragma once
#include <cstdint>
#include <cstddef>

#ifndef VMX_HOST_EMUL
#include <cuda_runtime.h>
#define VMX_DYN_SMEM(type, name) extern __shared__ type name[]
#define VMX_LAUNCH(ctx, kernel, grid, block, smem, ...)                                   \
  do {                                                                                    \
    kernel<<<(unsigned)(grid), (unsigned)(block), (smem), (ctx)->stream>>>(__VA_ARGS__);  \
    (ctx)->launches.fetch_add(1, std::memory_order_relaxed);                              \
  } while (0)
#else
#include <algorithm>
#include <cstdlib>
#include <cstring>
#define __global__
#define __device__
#define __host__
#define __constant__ const
#define __forceinline__ inline
#define __restrict__
#define __grid_constant__
#define __launch_bounds__(...)
struct uint2 { uint32_t x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
struct dim3 { unsigned x = 1, y = 1, z = 1; };
namespace vmx_emul {
extern thread_local dim3 threadIdx_, blockIdx_, blockDim_, gridDim_;
extern thread_local unsigned char* dyn_smem;
}
#define threadIdx (vmx_emul::threadIdx_)
#define blockIdx (vmx_emul::blockIdx_)
#define blockDim (vmx_emul::blockDim_)
#define gridDim (vmx_emul::gridDim_)
#define VMX_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(vmx_emul::dyn_smem)
using std::max;
using std::min;
template <typename T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
static inline int __clz(uint32_t v) { return v ? __builtin_clz(v) : 32; }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, int s) {
  s &= 31; return s ? (lo >> s) | (hi << (32 - s)) : lo;
}
// minimal runtime shims
typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
static inline const char* cudaGetErrorString(cudaError_t) { return "emul"; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaMallocAsync(void** p, size_t n, cudaStream_t) { *p = std::malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { std::free(p); return 0; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = std::malloc(n); return 0; }
static inline cudaError_t cudaFreeHost(void* p) { std::free(p); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return 0; }
#define cudaStreamNonBlocking 1
#define VMX_LAUNCH(ctx, kernel, grid, block, smem, ...)                                   \
  do {                                                                                    \
    std::vector<unsigned char> _sm((smem) + 16);                                          \
    vmx_emul::dyn_smem = _sm.data();                                                      \
    vmx_emul::gridDim_.x = (unsigned)(grid); vmx_emul::blockDim_.x = (unsigned)(block);   \
    for (unsigned _b = 0; _b < (unsigned)(grid); _b++)                                    \
      for (unsigned _t = 0; _t < (unsigned)(block); _t++) {                               \
        vmx_emul::blockIdx_.x = _b; vmx_emul::threadIdx_.x = _t;                          \
        kernel(__VA_ARGS__);                                                              \
      }                                                                                   \
    (ctx)->launches.fetch_add(1, std::memory_order_relaxed);                              \
  } while (0)
#endif
