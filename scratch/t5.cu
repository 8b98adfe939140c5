#include "../verificatum-vmn_b200/csrc/mont.cuh"
using namespace vmx;
template<int N>
__global__ void __launch_bounds__(128, 2) k_mul(uint32_t* out, const uint32_t* ain, const uint32_t* bin, int n, int iters, const __grid_constant__ MontParams<N> M) {
  int tid = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t a[N];
  #pragma unroll
  for (int i = 0; i < N; i++) a[i] = ain[(size_t)i * n + tid];
  const uint32_t* bp = bin + tid;
  for (int it = 0; it < iters; it++) {
    mont_mul<N>(a, [&](int i) { return bp[(size_t)i * n]; }, M);
  }
  #pragma unroll
  for (int i = 0; i < N; i++) out[(size_t)i * n + tid] = a[i];
}
template __global__ void k_mul<96>(uint32_t*, const uint32_t*, const uint32_t*, int, int, const __grid_constant__ MontParams<96>);
