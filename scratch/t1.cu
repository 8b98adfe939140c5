#include <cstdint>
// experiment: even/odd wide-chain row
template<int N>
__device__ __forceinline__ void row_even(uint32_t* acc, const uint32_t* a, uint32_t b) {
  // acc[j],acc[j+1] += a[j]*b for even j, single carry chain
  asm("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[0]), "+r"(acc[1]) : "r"(a[0]), "r"(b));
  #pragma unroll
  for (int j = 2; j < N; j += 2)
    asm("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[j]), "+r"(acc[j+1]) : "r"(a[j]), "r"(b));
}
extern "C" __global__ void k(uint32_t* out, const uint32_t* in, int n) {
  constexpr int N = 16;
  uint32_t a[N], acc[N+2];
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = 0; i < N; i++) a[i] = in[i * n + t];
  for (int i = 0; i < N+2; i++) acc[i] = 0;
  for (int r = 0; r < 4; r++) {
    uint32_t b = in[(N + r) * n + t];
    row_even<N>(acc, a, b);
    asm("addc.u32 %0, %0, 0;" : "+r"(acc[N]));
  }
  for (int i = 0; i < N+2; i++) out[i * n + t] = acc[i];
}
