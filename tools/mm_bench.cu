#include <cstdint>
template<int N> struct ModParam { uint32_t n[N]; uint32_t n0inv; };

// t += E/O handling. Two rows per call, relative columns 0..N+3 ; result shifted down by 2.
template<int N>
__device__ __forceinline__ void mont_rowpair(uint32_t (&t)[N+2], const uint32_t (&a)[N], uint32_t b0, uint32_t b1,
                                             const ModParam<N>& M) {
  // relative columns: t[0..N+1] live on entry (t[N+1] may be nonzero? we keep N+2 words: cols 0..N+1)
  uint32_t top2 = 0, top3 = 0;  // columns N+2, N+3
  // ---- row 0 (even): E = even j (pairs (j, j+1)), O = odd j (cols j, j+1)
  // E-ab
  asm("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(t[0]), "+r"(t[1]) : "r"(a[0]), "r"(b0));
  #pragma unroll
  for (int j = 2; j < N; j += 2)
    asm("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(t[j]), "+r"(t[j+1]) : "r"(a[j]), "r"(b0));
  asm("addc.cc.u32 %0, %0, 0; addc.u32 %1, %1, 0;" : "+r"(t[N]), "+r"(t[N+1]));
  // O-ab: odd j at cols (j, j+1)
  {
    uint32_t lo, hi;
    asm("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(lo), "=r"(hi) : "r"(a[1]), "r"(b0));
    asm("add.cc.u32 %0, %0, %2; addc.cc.u32 %1, %1, %3;" : "+r"(t[1]), "+r"(t[2]) : "r"(lo), "r"(hi));
    #pragma unroll
    for (int j = 3; j < N; j += 2) {
      asm("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(lo), "=r"(hi) : "r"(a[j]), "r"(b0));
      asm("addc.cc.u32 %0, %0, %2; addc.cc.u32 %1, %1, %3;" : "+r"(t[j]), "+r"(t[j+1]) : "r"(lo), "r"(hi));
    }
    asm("addc.u32 %0, %0, 0;" : "+r"(t[N+1]));
  }
  uint32_t m = t[0] * M.n0inv;
  // O-mn first
  {
    uint32_t lo, hi;
    asm("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(lo), "=r"(hi) : "r"(M.n[1]), "r"(m));
    asm("add.cc.u32 %0, %0, %2; addc.cc.u32 %1, %1, %3;" : "+r"(t[1]), "+r"(t[2]) : "r"(lo), "r"(hi));
    #pragma unroll
    for (int j = 3; j < N; j += 2) {
      asm("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(lo), "=r"(hi) : "r"(M.n[j]), "r"(m));
      asm("addc.cc.u32 %0, %0, %2; addc.cc.u32 %1, %1, %3;" : "+r"(t[j]), "+r"(t[j+1]) : "r"(lo), "r"(hi));
    }
    asm("addc.u32 %0, %0, 0;" : "+r"(t[N+1]));
  }
  // E-mn
  asm("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(t[0]), "+r"(t[1]) : "r"(M.n[0]), "r"(m));
  #pragma unroll
  for (int j = 2; j < N; j += 2)
    asm("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(t[j]), "+r"(t[j+1]) : "r"(M.n[j]), "r"(m));
  asm("addc.cc.u32 %0, %0, 0; addc.u32 %1, %1, 0;" : "+r"(t[N]), "+r"(t[N+1]));
  // now t[0]==0.  ---- row 1 (odd): products at cols 1+j. E = odd j: pairs (1+j, 2+j); O = even j: cols (1+j, 2+j)
  // E-ab: j=1: (2,3) ... j=N-1: (N, N+1), carry -> top2
  asm("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(t[2]), "+r"(t[3]) : "r"(a[1]), "r"(b1));
  #pragma unroll
  for (int j = 3; j < N; j += 2)
    asm("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(t[j+1]), "+r"(t[j+2]) : "r"(a[j]), "r"(b1));
  asm("addc.u32 %0, %0, 0;" : "+r"(top2));
  // O-ab: even j: cols (1+j, 2+j): 1..N, carry -> N+1, top2
  {
    uint32_t lo, hi;
    asm("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(lo), "=r"(hi) : "r"(a[0]), "r"(b1));
    asm("add.cc.u32 %0, %0, %2; addc.cc.u32 %1, %1, %3;" : "+r"(t[1]), "+r"(t[2]) : "r"(lo), "r"(hi));
    #pragma unroll
    for (int j = 2; j < N; j += 2) {
      asm("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(lo), "=r"(hi) : "r"(a[j]), "r"(b1));
      asm("addc.cc.u32 %0, %0, %2; addc.cc.u32 %1, %1, %3;" : "+r"(t[j+1]), "+r"(t[j+2]) : "r"(lo), "r"(hi));
    }
    asm("addc.cc.u32 %0, %0, 0; addc.u32 %1, %1, 0;" : "+r"(t[N+1]), "+r"(top2));
  }
  m = t[1] * M.n0inv;
  // O-mn
  {
    uint32_t lo, hi;
    asm("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(lo), "=r"(hi) : "r"(M.n[0]), "r"(m));
    asm("add.cc.u32 %0, %0, %2; addc.cc.u32 %1, %1, %3;" : "+r"(t[1]), "+r"(t[2]) : "r"(lo), "r"(hi));
    #pragma unroll
    for (int j = 2; j < N; j += 2) {
      asm("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(lo), "=r"(hi) : "r"(M.n[j]), "r"(m));
      asm("addc.cc.u32 %0, %0, %2; addc.cc.u32 %1, %1, %3;" : "+r"(t[j+1]), "+r"(t[j+2]) : "r"(lo), "r"(hi));
    }
    asm("addc.cc.u32 %0, %0, 0; addc.u32 %1, %1, 0;" : "+r"(t[N+1]), "+r"(top2));
  }
  // E-mn with shift by 2: pairs (1+j, 2+j), j odd -> write to (j-1, j)
  asm("mad.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;" : "=r"(t[0]), "=r"(t[1]) : "r"(M.n[1]), "r"(m), "r"(t[2]), "r"(t[3]));
  #pragma unroll
  for (int j = 3; j < N; j += 2)
    asm("madc.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;" : "=r"(t[j-1]), "=r"(t[j]) : "r"(M.n[j]), "r"(m), "r"(t[j+1]), "r"(t[j+2]));
  asm("addc.cc.u32 %0, %1, 0; addc.u32 %2, 0, 0;" : "=r"(t[N]), "+r"(top2), "=r"(t[N+1]));
}

template<int N>
__global__ void __launch_bounds__(128, 2) k_mul(uint32_t* out, const uint32_t* ain, const uint32_t* bin, int n, int iters, const __grid_constant__ ModParam<N> M) {
  int tid = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t a[N], t[N+2];
  #pragma unroll
  for (int i = 0; i < N; i++) a[i] = ain[(size_t)i * n + tid];
  for (int it = 0; it < iters; it++) {
    #pragma unroll
    for (int i = 0; i < N+2; i++) t[i] = 0;
    #pragma unroll 1
    for (int i = 0; i < N; i += 2) {
      uint32_t b0 = bin[(size_t)i * n + tid], b1 = bin[(size_t)(i+1) * n + tid];
      mont_rowpair<N>(t, a, b0, b1, M);
    }
    #pragma unroll
    for (int i = 0; i < N; i++) a[i] = t[i];
  }
  #pragma unroll
  for (int i = 0; i < N; i++) out[(size_t)i * n + tid] = a[i];
  out[(size_t)N * n + tid] = t[N];
}

#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);}}while(0)
int main(int argc, char** argv) {
  constexpr int N = 96;
  const char* pfile = argv[1]; const char* ofile = argv[2];
  ModParam<N> M;
  FILE* f = fopen(pfile, "r"); for (int i = 0; i < N; i++) fscanf(f, "%x", &M.n[i]); fscanf(f, "%x", &M.n0inv); fclose(f);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int n = sms * 256;
  std::vector<uint32_t> ha((size_t)N * n), hb((size_t)N * n), ho((size_t)(N + 1) * n);
  uint64_t s = 88172645463325252ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 16); };
  for (auto& x : ha) x = rnd();
  for (auto& x : hb) x = rnd();
  for (int t = 0; t < n; t++) { ha[(size_t)(N - 1) * n + t] >>= 1; hb[(size_t)(N - 1) * n + t] >>= 1; }  // < p
  uint32_t *da, *db, *dout;
  CK(cudaMalloc(&da, ha.size() * 4)); CK(cudaMalloc(&db, hb.size() * 4)); CK(cudaMalloc(&dout, ho.size() * 4));
  CK(cudaMemcpy(da, ha.data(), ha.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice));
  k_mul<N><<<n / 128, 128>>>(dout, da, db, n, 1, M);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost));
  f = fopen(ofile, "w");
  for (int t : {0, 1, 77, n - 1}) {
    for (int i = 0; i < N; i++) fprintf(f, "%08x ", ha[(size_t)i * n + t]); fprintf(f, "\n");
    for (int i = 0; i < N; i++) fprintf(f, "%08x ", hb[(size_t)i * n + t]); fprintf(f, "\n");
    for (int i = 0; i <= N; i++) fprintf(f, "%08x ", ho[(size_t)i * n + t]); fprintf(f, "\n");
  }
  fclose(f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int iters : {16, 64}) {
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
      cudaEventRecord(e0);
      k_mul<N><<<n / 128, 128>>>(dout, da, db, n, iters, M);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double mm = (double)n * iters / (best * 1e-3);
    printf("t4 TPE N=%d: n=%d iters=%d %.3f ms -> %.3e modmul/s, %.3e MAC/s\n", N, n, iters, best, mm, mm * (2.0 * N * N + N));
  }
  return 0;
}
