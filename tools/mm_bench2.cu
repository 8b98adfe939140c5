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

#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);}}while(0)
template<int N> int run(const char* pfile, const char* ofile) {
  MontParams<N> M;
  FILE* f = fopen(pfile, "r"); for (int i = 0; i < N; i++) fscanf(f, "%x", &M.n[i]); fscanf(f, "%x", &M.n0inv); fclose(f);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int n = sms * 256;
  std::vector<uint32_t> ha((size_t)N * n), hb((size_t)N * n), ho((size_t)(N) * n);
  uint64_t s = 88172645463325252ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 16); };
  for (auto& x : ha) x = rnd();
  for (auto& x : hb) x = rnd();
  for (int t = 0; t < n; t++) { ha[(size_t)(N - 1) * n + t] >>= 1; hb[(size_t)(N - 1) * n + t] >>= 1; }  // < p
  uint32_t *da, *db, *dout;
  CK(cudaMalloc(&da, ha.size() * 4)); CK(cudaMalloc(&db, hb.size() * 4)); CK(cudaMalloc(&dout, ho.size() * 4));
  CK(cudaMemcpy(da, ha.data(), ha.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice));
  k_mul<N><<<n / 128, 128>>>(dout, da, db, n, 3, M);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost));
  f = fopen(ofile, "w");
  for (int t : {0, 1, 77, n - 1}) {
    for (int i = 0; i < N; i++) fprintf(f, "%08x ", ha[(size_t)i * n + t]); fprintf(f, "\n");
    for (int i = 0; i < N; i++) fprintf(f, "%08x ", hb[(size_t)i * n + t]); fprintf(f, "\n");
    for (int i = 0; i < N; i++) fprintf(f, "%08x ", ho[(size_t)i * n + t]); fprintf(f, "\n");
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
    printf("t5 TPE N=%d: n=%d iters=%d %.3f ms -> %.3e modmul/s, %.3e MAC/s\n", N, n, iters, best, mm, mm * (2.0 * N * N + N));
  }
  return 0;
}
int main(int argc, char** argv) {
  run<96>(argv[1], argv[2]);
  run<64>(argv[3], argv[4]);
  return 0;
}
