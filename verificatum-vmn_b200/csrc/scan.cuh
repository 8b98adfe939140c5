// scan.cuh -- exclusive prefix sum of uint32 arrays (bucket offsets, chunk offsets).
// Three small kernels (block scan, scan of block sums, add-back); sizes here are <= a few
// million entries, so this is launch-latency bound and deliberately simple.
#pragma once
#include <cstdint>

namespace vmx {

constexpr int kScanBlock = 1024;

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

__global__ void k_scan_add(uint32_t* __restrict__ d, size_t n, const uint32_t* __restrict__ sums) {
  const size_t i = (size_t)blockIdx.x * kScanBlock + threadIdx.x;
  if (i < n && blockIdx.x) d[i] += sums[blockIdx.x];
}

}  // namespace vmx
