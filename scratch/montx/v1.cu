#include "../../verificatum-vmn_b200/csrc/layout.cuh"
using namespace vmx;
__device__ __forceinline__ uint64_t pack64(uint32_t lo, uint32_t hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
__device__ __forceinline__ void unpack64(uint32_t& lo, uint32_t& hi, uint64_t v) { asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); }
template <int N, typename Loader>
__device__ __forceinline__ void mont_mul_v1(uint32_t (&a)[N], Loader ld2, const MontParams<N>& M) {
  uint64_t T[N / 2 + 1];
#pragma unroll
  for (int i = 0; i < N / 2 + 1; i++) T[i] = 0;
  Word2 b = ld2(0);
#pragma unroll 1
  for (int i = 0; i < N; i += 2) {
    const Word2 nb = ld2(i + 2 < N ? i + 2 : i);
    uint32_t t[N + 2];
#pragma unroll
    for (int k = 0; k < N / 2 + 1; k++) unpack64(t[2 * k], t[2 * k + 1], T[k]);
    mont_rowpair<N>(t, a, b.x, b.y, M);
#pragma unroll
    for (int k = 0; k < N / 2 + 1; k++) T[k] = pack64(t[2 * k], t[2 * k + 1]);
    b = nb;
  }
  uint32_t t[N + 2];
#pragma unroll
  for (int k = 0; k < N / 2 + 1; k++) unpack64(t[2 * k], t[2 * k + 1], T[k]);
  mont_final_sub<N>(a, t, M);
}
template <int N>
__global__ void __launch_bounds__(128, 2) k_mul_iter(const uint32_t* __restrict__ a_, const uint32_t* __restrict__ b_, uint32_t* __restrict__ out,
                         size_t cap, size_t n, int iters, const __grid_constant__ MontParams<N> M) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t a[N];
  load_elem<N>(a, a_, cap, i);
  const GlobalLoader B(b_, cap, i);
  for (int it = 0; it < iters; it++) mont_mul_v1<N>(a, B, M);
  store_elem<N>(a, out, cap, i);
}
void* force_inst() { return (void*)&k_mul_iter<96>; }
