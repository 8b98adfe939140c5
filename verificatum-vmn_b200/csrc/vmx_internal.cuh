// vmx_internal.cuh -- host-side structures behind the opaque handles of include/vmx.h.
#pragma once
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/vmx.h"
#include "cuda_compat.cuh"
#include "kernels_elem.cuh"
#include "kernels_ec.cuh"
#include "kernels_member.cuh"
#include "kernels_mexp.cuh"
#include "kernels_prg.cuh"
#include "kernels_perm.cuh"
#include "kernels_ring.cuh"
#include "scan.cuh"

namespace vmx {

constexpr int kMaxLimbs = 96;

void set_error(const char* fmt, ...);

#define VMX_CU(expr)                                                                                 \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess) {                                                                         \
      ::vmx::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);  \
      return VMX_ECUDA;                                                                              \
    }                                                                                                \
  } while (0)

#define VMX_TRY(expr)            \
  do {                           \
    int _s = (expr);             \
    if (_s != VMX_OK) return _s; \
  } while (0)

// Fixed-base window table: entry (k, d) = base^(d * 2^(w*k)), Montgomery form, limb-major,
// element index (k << w) + d.
struct FixedTable {
  uint32_t* d = nullptr;
  size_t cap = 0;
  int w = 0, nwin = 0;
  size_t bytes = 0;       // device memory of the table
  uint64_t last_use = 0;  // tick of the context's table clock (least recently used tables are evicted first)
};

struct Modulus {
  uint32_t n[kMaxLimbs];
  uint32_t n0inv = 0;
  int bits = 0;
  // device, limb-major with cap = 4: [0] = R^2 mod n, [1] = R mod n (Montgomery one), [2] = 1
  uint32_t* consts = nullptr;
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
  int kind = 0;            // 0: ModPGroup, 1: ECqPGroup (256-bit prime curve)
  int nl = 0;              // limbs per residue: 16, 32, 64 or 96 (ModP); 8 (curve: coordinate field and Z_q)
  int gl = 0;              // limbs per stored group element: nl (ModP), 16 (affine curve point)
  size_t cb = 0;           // curve: bytes of one serialised coordinate
  vmx::EcCurve ecc;        // curve: field, coefficients and constants handed to the kernels
  bool ec_sqrt_ok = false; // curve: p = 3 mod 4 (square roots by one exponentiation)
  uint64_t prg_consumed = 0;  // stream bytes the last *_prg_sha256 call consumed
  size_t eb = 0, rb = 0;   // bytes of a serialised group / ring element
  cudaStream_t stream = nullptr;
  vmx::Modulus P, Q;
  std::vector<uint32_t> pm2;  // p - 2 (inversion exponent)
  std::mutex mu;
  std::recursive_mutex api;  // serialises API calls of host threads on this context
  std::map<std::string, vmx::FixedTable> tables;  // key = canonical base bytes
  int fixed_window = 0;                           // 0 = choose from n
  // tuning knobs (vmx_ctx_set_tuning): production defaults; the parity tests lower them so that small
  // oracle-sized arrays run through the thread-per-element kernels and their multi-chunk loops
  size_t coop_max = 8192;                         // arrays up to this size use the warp-per-element kernels
  size_t var_chunk = 0;                           // elements per launch of k_exp_var / k_exp_var2 (0 = from memory)
  int mexp_window = 0;                            // Pippenger window c (0 = choose from n and the exponent length)
  size_t table_max_bytes = (size_t)20e9;          // largest single fixed-base table (w = 18 at 3072 bits: 17.2 GB)
  size_t table_budget = (size_t)100e9;            // all cached tables together; beyond it the least recently used go
  size_t table_bytes = 0;                         // currently cached
  uint64_t table_clock = 0;
  // Large device blocks (arrays and scratch of 32 MB and more) are recycled here instead of going back to the
  // stream-ordered pool: a step allocates and frees the same ~40 sizes (384 MB arrays, GBs of Pippenger and
  // window-table scratch) in an order that fragments the pool, and a fragmented pool answers with fresh physical
  // memory -- one expProd in three took 1.46 s instead of 0.13 s at N = 10^6 (gpurun_out/s4_trace_device.json).
  // All work of a context is on one stream, so a block freed by the host may be handed out again at once.
  struct BigBlock { void* p; size_t bytes; };
  std::vector<BigBlock> big_free;
  std::mutex big_mu;
  size_t big_free_bytes = 0;
  size_t big_cache_max = (size_t)48e9;
  size_t big_block_min = (size_t)32 << 20;        // blocks of at least this size are recycled (tests lower it)
  int* d_flag = nullptr;                          // device scratch: 4 ints
  int* h_flag = nullptr;                          // pinned scratch: 4 ints
  std::atomic<uint64_t> launches{0}, modmuls{0};
  int sm_count = 148;
  bool safe_prime = false;  // p = 2q + 1: membership = Legendre symbol (k_jacobi)
};

struct vmx_garr {
  vmx_ctx* ctx;
  size_t n, cap;
  uint32_t* d;
  size_t granted = 0;  // bytes behind d as the allocator counted them (dev_alloc / dev_free)
};

struct vmx_rarr {
  vmx_ctx* ctx;
  size_t n, cap;
  uint32_t* d;
  mutable int bits;  // cached max bit length, -1 = unknown
  size_t granted = 0;
};
