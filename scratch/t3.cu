#include <cstdint>
extern "C" __global__ void k(uint64_t* out, const uint64_t* in, int n) {
  constexpr int N = 8;
  uint64_t a[N], b[N];
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = 0; i < N; i++) { a[i] = in[i * n + t]; b[i] = in[(i+N)*n+t]; }
  asm("add.cc.u64 %0, %0, %1;" : "+l"(a[0]) : "l"(b[0]));
  #pragma unroll
  for (int i = 1; i < N; i++) asm("addc.cc.u64 %0, %0, %1;" : "+l"(a[i]) : "l"(b[i]));
  for (int i = 0; i < N; i++) out[i * n + t] = a[i];
}
