// kernels_ring.cuh -- Z_q array kernels (exponent ring of the group).
//
// Ring elements are stored as canonical residues (so that exponent digits can be read
// directly).  Products use the same Montgomery core with the modulus q: for x canonical and
// yM = y*R mod q,  mont_mul(x, yM) = x*y mod q  (canonical), so one operand of every product is
// first lifted to Montgomery form.  Reference call sites: hvzk/PoSBasicTW.java:553-604
// (permute, recLin, prods), :637-653 (shiftPush, mul, add), :861-878 (innerProduct, sum, mulAdd).
#pragma once
#include "kernels_elem.cuh"

namespace vmx {

// a <- a + b mod n, b streamed from memory (element i of a limb-major array).
template <int N>
VMX_DEV void mod_add_stream(uint32_t (&a)[N], const uint32_t* b_, size_t bcap, size_t i, const MontParams<N>& M) {
  const uint4* p = reinterpret_cast<const uint4*>(b_) + i;
  uint32_t c;
  {
    const uint4 v = p[0];
    add_cc(a[0], a[0], v.x); addc_cc(a[1], a[1], v.y); addc_cc(a[2], a[2], v.z); addc_cc(a[3], a[3], v.w);
  }
#pragma unroll
  for (int g = 1; g < N / 4; g++) {
    const uint4 v = p[(size_t)g * bcap];
    addc_cc(a[4 * g], a[4 * g], v.x); addc_cc(a[4 * g + 1], a[4 * g + 1], v.y);
    addc_cc(a[4 * g + 2], a[4 * g + 2], v.z); addc_cc(a[4 * g + 3], a[4 * g + 3], v.w);
  }
  addc(c, 0, 0);
  uint32_t d[N], brw;
  sub_cc(d[0], a[0], M.n[0]);
#pragma unroll
  for (int j = 1; j < N; j++) subc_cc(d[j], a[j], M.n[j]);
  subc(brw, c, 0);
  const bool keep = (brw != 0);
#pragma unroll
  for (int j = 0; j < N; j++) a[j] = keep ? a[j] : d[j];
}

// op 0: out = a + b ; op 1: out = -a ; op 2: out = a - b
template <int N>
VMX_KERNEL(N) k_ring_addsub(const uint32_t* __restrict__ a_, size_t acap, const uint32_t* __restrict__ b_, size_t bcap,
                            uint32_t* __restrict__ out, size_t ocap, size_t n, int op,
                            const __grid_constant__ MontParams<N> M) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t a[N];
  if (op == 0) {
    load_elem<N>(a, a_, acap, i);
    mod_add_stream<N>(a, b_, bcap, i, M);
  } else {
    // x = (op==1 ? a : b);  r = q - x (or 0), then op 2 adds a
    load_elem<N>(a, op == 1 ? a_ : b_, op == 1 ? acap : bcap, i);
    uint32_t nz = 0;
#pragma unroll
    for (int j = 0; j < N; j++) nz |= a[j];
    if (nz) {
      sub_cc(a[0], M.n[0], a[0]);
#pragma unroll
      for (int j = 1; j < N; j++) subc_cc(a[j], M.n[j], a[j]);
    }
    if (op == 2) mod_add_stream<N>(a, a_, acap, i, M);
  }
  store_elem<N>(a, out, ocap, i);
}

// out[i] = a[i] * c  where c is element `cidx` of array c_ (Montgomery-form constant) -- used for
// to-Montgomery (c = R^2), scalar multiples, etc.  If addend != null: out[i] += addend[i].
template <int N>
VMX_KERNEL(N) k_mul_const(const uint32_t* __restrict__ a_, size_t acap, const uint32_t* __restrict__ c_, size_t ccap,
                          size_t cidx, const uint32_t* __restrict__ addend, size_t dcap, uint32_t* __restrict__ out,
                          size_t ocap, size_t n, const __grid_constant__ MontParams<N> M) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t a[N];
  load_elem<N>(a, a_, acap, i);
  mont_mul<N>(a, GlobalLoader(c_, ccap, cidx), M);
  if (addend) mod_add_stream<N>(a, addend, dcap, i, M);
  store_elem<N>(a, out, ocap, i);
}

// out[i] = a[i] * 1 * R^-1 (Montgomery -> canonical)
template <int N>
VMX_KERNEL(N) k_from_mont(const uint32_t* __restrict__ a_, size_t acap, uint32_t* __restrict__ out, size_t ocap,
                          size_t n, const __grid_constant__ MontParams<N> M) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t a[N];
  load_elem<N>(a, a_, acap, i);
  mont_mul<N>(a, OneLoader{}, M);
  store_elem<N>(a, out, ocap, i);
}

// Chunked sums: out[c] = sum_{k in chunk c} x_k  where x_k = a[k] (b == null) or
// mont_mul(a[k], b[k]) (inner product; the caller multiplies the final sum by R^2).
template <int N>
VMX_KERNEL(N) k_chunk_sum(const uint32_t* __restrict__ a_, size_t acap, const uint32_t* __restrict__ b_, size_t bcap,
                          size_t n, int K, uint32_t* __restrict__ out, size_t ocap, uint32_t* __restrict__ tmp,
                          size_t tcap, const __grid_constant__ MontParams<N> M) {
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nch = (n + K - 1) / K;
  if (c >= nch) return;
  const size_t b0 = c * K, b1 = min(n, b0 + (size_t)K);
  uint32_t a[N];
  if (!b_) {
    load_elem<N>(a, a_, acap, b0);
    for (size_t k = b0 + 1; k < b1; k++) mod_add_stream<N>(a, a_, acap, k, M);
  } else {
    // acc kept in tmp[c]; a holds the current product
    load_elem<N>(a, a_, acap, b0);
    mont_mul<N>(a, GlobalLoader(b_, bcap, b0), M);
    for (size_t k = b0 + 1; k < b1; k++) {
      store_elem<N>(a, tmp, tcap, c);
      load_elem<N>(a, a_, acap, k);
      mont_mul<N>(a, GlobalLoader(b_, bcap, k), M);
      mod_add_stream<N>(a, tmp, tcap, c, M);
    }
  }
  store_elem<N>(a, out, ocap, c);
}

// ------------------------------------------------------------------ affine scan
// x_i = x_{i-1} * e_i + b_i  (x_{-1} = 0)   and   y_i = y_{i-1} * e_i  (y_{-1} = 1)
// eM = e in Montgomery form.  Phase A: per chunk totals  A_c = prod e_i,  B_c = value of the
// recurrence over the chunk started from 0.  Phase B: per chunk walk from the incoming state.
template <int N>
VMX_KERNEL(N) k_scan_phaseA(const uint32_t* __restrict__ eM, size_t ecap, const uint32_t* __restrict__ b_, size_t bcap,
                            size_t n, int K, uint32_t* __restrict__ A, size_t Acap, uint32_t* __restrict__ B,
                            size_t Bcap, const uint32_t* __restrict__ consts,
                            const __grid_constant__ MontParams<N> M) {
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nch = (n + K - 1) / K;
  if (c >= nch) return;
  const size_t b0 = c * K, b1 = min(n, b0 + (size_t)K);
  uint32_t a[N];
  // A_c (canonical): start from e_b0 canonical = mont_mul(eM, 1)
  load_elem<N>(a, eM, ecap, b0);
  mont_mul<N>(a, OneLoader{}, M);
  for (size_t k = b0 + 1; k < b1; k++) mont_mul<N>(a, GlobalLoader(eM, ecap, k), M);
  store_elem<N>(a, A, Acap, c);
  if (b_) {
    load_elem<N>(a, b_, bcap, b0);
    for (size_t k = b0 + 1; k < b1; k++) {
      mont_mul<N>(a, GlobalLoader(eM, ecap, k), M);
      mod_add_stream<N>(a, b_, bcap, k, M);
    }
    store_elem<N>(a, B, Bcap, c);
  }
  (void)consts;
}

// Phase B.  Incoming state of chunk c: (IA[c-1], IB[c-1]) (inclusive scans of the chunk totals);
// chunk 0 starts from (1, 0).  want_y: write y (prods) else write x (recLin).
template <int N>
VMX_KERNEL(N) k_scan_phaseB(const uint32_t* __restrict__ eM, size_t ecap, const uint32_t* __restrict__ b_, size_t bcap,
                            size_t n, int K, const uint32_t* __restrict__ IA, size_t IAcap,
                            const uint32_t* __restrict__ IB, size_t IBcap, int want_y, uint32_t* __restrict__ out,
                            size_t ocap, const __grid_constant__ MontParams<N> M) {
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nch = (n + K - 1) / K;
  if (c >= nch) return;
  const size_t b0 = c * K, b1 = min(n, b0 + (size_t)K);
  uint32_t a[N];
  if (want_y) {
    if (c == 0) {
      load_elem<N>(a, eM, ecap, b0);
      mont_mul<N>(a, OneLoader{}, M);
    } else {
      load_elem<N>(a, IA, IAcap, c - 1);
      mont_mul<N>(a, GlobalLoader(eM, ecap, b0), M);
    }
    store_elem<N>(a, out, ocap, b0);
    for (size_t k = b0 + 1; k < b1; k++) {
      mont_mul<N>(a, GlobalLoader(eM, ecap, k), M);
      store_elem<N>(a, out, ocap, k);
    }
  } else {
    if (c == 0) {
      load_elem<N>(a, b_, bcap, b0);
    } else {
      load_elem<N>(a, IB, IBcap, c - 1);
      mont_mul<N>(a, GlobalLoader(eM, ecap, b0), M);
      mod_add_stream<N>(a, b_, bcap, b0, M);
    }
    store_elem<N>(a, out, ocap, b0);
    for (size_t k = b0 + 1; k < b1; k++) {
      mont_mul<N>(a, GlobalLoader(eM, ecap, k), M);
      mod_add_stream<N>(a, b_, bcap, k, M);
      store_elem<N>(a, out, ocap, k);
    }
  }
}

}  // namespace vmx
