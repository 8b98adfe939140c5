#include <cstdint>
template<int N>
__device__ __forceinline__ void row_lohi(uint32_t* acc, const uint32_t* a, uint32_t b) {
  asm("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(acc[0]) : "r"(a[0]), "r"(b));
  #pragma unroll
  for (int j = 1; j < N; j++)
    asm("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(acc[j]) : "r"(a[j]), "r"(b));
  asm("addc.u32 %0, %0, 0;" : "+r"(acc[N]));
  asm("mad.hi.cc.u32 %0, %1, %2, %0;" : "+r"(acc[1]) : "r"(a[0]), "r"(b));
  #pragma unroll
  for (int j = 1; j < N; j++)
    asm("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(acc[j+1]) : "r"(a[j]), "r"(b));
  asm("addc.u32 %0, %0, 0;" : "+r"(acc[N+1]));
}
extern "C" __global__ void k(uint32_t* out, const uint32_t* in, int n) {
  constexpr int N = 16;
  uint32_t a[N], acc[N+2];
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = 0; i < N; i++) a[i] = in[i * n + t];
  for (int i = 0; i < N+2; i++) acc[i] = 0;
  for (int r = 0; r < 4; r++) {
    uint32_t b = in[(N + r) * n + t];
    row_lohi<N>(acc, a, b);
  }
  for (int i = 0; i < N+2; i++) out[i * n + t] = acc[i];
}
