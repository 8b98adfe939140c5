// scan.cuh -- exclusive prefix sum of uint32 arrays (bucket offsets, chunk offsets).
// Three small kernels (block scan, scan of block sums, add-back); sizes here are <= a few
// million entries, so this is launch-latency bound and deliberately simple.
#pragma once
#include "cuda_compat.cuh"

namespace vmx {

constexpr int kScanBlock = 1024;

#ifndef VMX_HOST_EMUL
// in-place exclusive scan of each block of 1024 entries; block totals to sums[blockIdx]
__global__ void k_scan_block(uint32_t* __restrict__ d, size_t n, uint32_t* __restrict__ sums) {
  __shared__ uint32_t warp_tot[32];
  const size_t i = (size_t)blockIdx.x * kScanBlock + threadIdx.x;
  const uint32_t v = i < n ? d[i] : 0;
  uint32_t x = v;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_tot[wid] = x;
  __syncthreads();
  if (wid == 0) {
    uint32_t w = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    warp_tot[lane] = w;
  }
  __syncthreads();
  const uint32_t base = wid ? warp_tot[wid - 1] : 0;
  if (i < n) d[i] = base + x - v;
  if (threadIdx.x == kScanBlock - 1 && sums) sums[blockIdx.x] = base + x;
}

#else
// sequential stand-in (tests/host_emul only): thread 0 of each block scans its 1024 entries
inline void k_scan_block(uint32_t* d, size_t n, uint32_t* sums) {
  if (threadIdx.x != 0) return;
  const size_t b0 = (size_t)blockIdx.x * kScanBlock;
  uint32_t acc = 0;
  for (size_t i = b0; i < b0 + kScanBlock && i < n; i++) { const uint32_t v = d[i]; d[i] = acc; acc += v; }
  if (sums) sums[blockIdx.x] = acc;
}
#endif

__global__ void k_scan_add(uint32_t* __restrict__ d, size_t n, const uint32_t* __restrict__ sums) {
  const size_t i = (size_t)blockIdx.x * kScanBlock + threadIdx.x;
  if (i < n && blockIdx.x) d[i] += sums[blockIdx.x];
}

}  // namespace vmx
