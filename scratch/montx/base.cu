#include "../../verificatum-vmn_b200/csrc/layout.cuh"
using namespace vmx;
template <int N>
__global__ void __launch_bounds__(128, 2) k_mul_iter(const uint32_t* __restrict__ a_, const uint32_t* __restrict__ b_, uint32_t* __restrict__ out,
                         size_t cap, size_t n, int iters, const __grid_constant__ MontParams<N> M) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t a[N];
  load_elem<N>(a, a_, cap, i);
  const GlobalLoader B(b_, cap, i);
  for (int it = 0; it < iters; it++) mont_mul<N>(a, B, M);
  store_elem<N>(a, out, cap, i);
}
void* force_inst() { return (void*)&k_mul_iter<96>; }
