// Microbenchmark: integer pipe throughput on sm_100a. Prints warp-instr/clk/SM for several instruction mixes.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1;}}while(0)

constexpr int ITERS = 4096;
constexpr int UNR = 16;   // instrs per chain-set per iteration

// mode 0: IMAD.WIDE.U32 independent accumulators (ILP=8), no carry
// mode 1: IMAD.WIDE.U32.X carry chain (dependent through CC), one chain of 8 pairs per step, 2 chains interleaved by compiler
// mode 2: IMAD (32-bit lo) ILP 8
// mode 3: IMAD.HI.U32 ILP 8
// mode 4: IADD3 ILP 8
// mode 5: IMAD.WIDE (ILP8) + IADD3 (ILP8) interleaved 1:1
// mode 6: single dependent IMAD.WIDE chain (latency)
// mode 7: single dependent carry chain IMAD.WIDE.X (latency through predicate)
template<int MODE>
__global__ void kern(uint32_t* out, uint32_t x, uint32_t y, int iters) {
  uint32_t a = threadIdx.x * 2654435761u + x, b = blockIdx.x * 40503u + y;
  uint64_t acc[8]; uint32_t r[16];
  #pragma unroll
  for (int i = 0; i < 8; i++) { acc[i] = a * (i + 1); }
  #pragma unroll
  for (int i = 0; i < 16; i++) { r[i] = (b + i) * a; }
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) {
      #pragma unroll
      for (int u = 0; u < UNR; u++) {
        #pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"((uint32_t)acc[(i + 4) & 7]), "r"(r[i]));
      }
    } else if (MODE == 1) {
      #pragma unroll
      for (int u = 0; u < UNR; u++) {
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(r[0]), "+r"(r[1]) : "r"(r[8]), "r"(b));
        #pragma unroll
        for (int i = 2; i < 16; i += 2)
          asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(r[i]), "+r"(r[i+1]) : "r"(r[(i + 8) & 15]), "r"(b));
      }
    } else if (MODE == 2) {
      #pragma unroll
      for (int u = 0; u < UNR; u++) {
        #pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(r[(i + 4) & 7]), "r"(r[i + 8]));
      }
    } else if (MODE == 3) {
      #pragma unroll
      for (int u = 0; u < UNR; u++) {
        #pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(r[(i + 4) & 7]), "r"(r[i + 8]));
      }
    } else if (MODE == 4) {
      #pragma unroll
      for (int u = 0; u < UNR; u++) {
        #pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(r[i + 8]));
      }
    } else if (MODE == 5) {
      #pragma unroll
      for (int u = 0; u < UNR; u++) {
        #pragma unroll
        for (int i = 0; i < 8; i++) {
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"((uint32_t)acc[(i + 4) & 7]), "r"(b));
          asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(r[i + 8]));
        }
      }
    } else if (MODE == 6) {
      #pragma unroll
      for (int u = 0; u < UNR * 8; u++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[0]) : "r"((uint32_t)acc[0]), "r"(b));
    } else if (MODE == 8) {
      uint32_t q[16];
      #pragma unroll
      for (int i = 0; i < 16; i++) q[i] = r[i] ^ it;
      #pragma unroll
      for (int u = 0; u < UNR; u++) {
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(r[0]), "+r"(r[1]) : "r"(r[8]), "r"(b));
        asm volatile("add.u32 %0, %0, %1;" : "+r"(q[0]) : "r"(q[8]));
        #pragma unroll
        for (int i = 2; i < 16; i += 2) {
          asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(r[i]), "+r"(r[i+1]) : "r"(r[(i + 8) & 15]), "r"(b));
          asm volatile("add.u32 %0, %0, %1;" : "+r"(q[i/2]) : "r"(q[i/2 + 8]));
        }
      }
      #pragma unroll
      for (int i = 0; i < 8; i++) acc[i] += q[i];
    } else if (MODE == 9) {
      #pragma unroll
      for (int u = 0; u < UNR; u++) {
        #pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(acc[i]) : "r"((uint32_t)(acc[(i + 4) & 7] >> 32)), "r"(r[i]));
      }
    } else if (MODE == 7) {
      #pragma unroll
      for (int u = 0; u < UNR * 8; u++)
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1; addc.u32 %0, %0, 0;" : "+r"(r[0]), "+r"(r[1]) : "r"(r[1]), "r"(b));
    }
  }
  uint32_t s = 0;
  #pragma unroll
  for (int i = 0; i < 8; i++) s += (uint32_t)acc[i] + (uint32_t)(acc[i] >> 32);
  #pragma unroll
  for (int i = 0; i < 16; i++) s += r[i];
  if (s == 0x12345678) out[0] = s;
}

template<int MODE> int run(const char* name, int instr_per_iter, int warps_per_sm, uint32_t* d) {
  int dev_sms; cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
  int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int threads = 128, blocks_per_sm = warps_per_sm / 4; if (blocks_per_sm < 1) { blocks_per_sm = 1; threads = warps_per_sm * 32; }
  kern<MODE><<<dev_sms * blocks_per_sm, threads>>>(d, 1, 2, 64);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    kern<MODE><<<dev_sms * blocks_per_sm, threads>>>(d, 1, 2, ITERS);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  double winstr = (double)ITERS * instr_per_iter * warps_per_sm;   // per SM
  double per_s = winstr * dev_sms / (best * 1e-3);
  printf("%-44s warps/SM=%2d  %8.3f ms  %.3e warp-instr/s  (%.2f warp-instr/clk/SM @maxclk %d MHz)\n", name, warps_per_sm, best, per_s,
         winstr / (best * 1e-3 * khz * 1e3), khz / 1000);
  return 0;
}

int main() {
  uint32_t* d; CK(cudaMalloc(&d, 1024));
  for (int w : {4, 8, 16, 32}) {
    run<0>("IMAD.WIDE.U32 ilp8", UNR * 8, w, d);
    run<1>("IMAD.WIDE.U32.X carry chain(8)", UNR * 8, w, d);
    run<2>("IMAD lo ilp8", UNR * 8, w, d);
    run<3>("IMAD.HI ilp8", UNR * 8, w, d);
    run<4>("IADD ilp8", UNR * 8, w, d);
    run<5>("IMAD.WIDE + IADD 1:1 (counting both)", UNR * 16, w, d);
    run<6>("IMAD.WIDE dependent (latency)", UNR * 8, w, d);
    run<7>("IMAD.WIDE.X + addc dependent", UNR * 16, w, d);
    run<9>("IMAD.WIDE.U32 mul-only ilp8", UNR * 8, w, d);
    run<8>("IMAD.WIDE.X chain + IADD 1:1 (count WIDE only)", UNR * 8, w, d);
  }
  return 0;
}
