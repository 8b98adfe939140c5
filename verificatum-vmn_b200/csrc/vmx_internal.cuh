// vmx_internal.cuh -- host-side structures behind the opaque handles of include/vmx.h.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/vmx.h"
#include "kernels_elem.cuh"
#include "kernels_mexp.cuh"
#include "kernels_prg.cuh"
#include "kernels_ring.cuh"
#include "scan.cuh"

namespace vmx {

constexpr int kMaxLimbs = 96;

void set_error(const char* fmt, ...);

#define VMX_CU(expr)                                                                       \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ::vmx::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return VMX_ECUDA;                                                                    \
    }                                                                                      \
  } while (0)

#define VMX_TRY(expr)            \
  do {                           \
    int _s = (expr);             \
    if (_s != VMX_OK) return _s; \
  } while (0)

struct FixedTable {
  uint32_t* d = nullptr;  // nwin * 2^w entries, limb-major, Montgomery form
  size_t cap = 0;
  int w = 0, nwin = 0;
};

struct Modulus {
  uint32_t n[kMaxLimbs];
  uint32_t n0inv;
  int bits;
  uint32_t* consts = nullptr;  // device, cap = 4: [0] = R^2 mod n, [1] = R mod n (Montgomery one), [2] = 1
  template <int N> MontParams<N> params() const {
    MontParams<N> M;
    for (int i = 0; i < N; i++) M.n[i] = n[i];
    M.n0inv = n0inv;
    return M;
  }
};

}  // namespace vmx

struct vmx_ctx {
  int device = 0;
  int nl = 0;  // limbs per residue: 64 or 96
  size_t eb = 0, rb = 0;
  cudaStream_t stream = nullptr;
  vmx::Modulus P, Q;
  std::vector<uint8_t> g_be;
  std::mutex mu;
  std::map<std::string, vmx::FixedTable> tables;  // key = base bytes
  int fixed_window = 0;                           // 0 = choose from n
  int* d_flag = nullptr;                          // device int flag
  unsigned* d_bits = nullptr;
  int* h_flag = nullptr;                          // pinned
  std::atomic<uint64_t> launches{0}, modmuls{0};
  int sm_count = 148;
};

struct vmx_garr {
  vmx_ctx* ctx;
  size_t n, cap;
  uint32_t* d;
};

struct vmx_rarr {
  vmx_ctx* ctx;
  size_t n, cap;
  uint32_t* d;
  mutable int bits;  // cached max bit length, -1 = unknown
};
