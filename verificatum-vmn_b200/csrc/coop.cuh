// coop.cuh -- warp-cooperative Montgomery multiplication: ONE WARP per residue.
//
// The thread-per-element kernels (mont.cuh) reach the IMAD roofline only when ~38,000 residues
// are in flight, and one thread needs ~45 us per 3072-bit modmul (it is bounded by the issue
// rate of ONE scheduler).  Everything sequential in the hot path -- the L squarings of a
// Horner evaluation, the squaring chain under a fixed-base table, exponentiations of the O(1)
// single elements of a proof, arrays of a few thousand elements -- instead spreads one residue
// over the 32 lanes of a warp (N/32 limbs per lane; 16 lanes x 1 limb for N = 16):
//
//   for every word b_i (broadcast with SHFL):            acc += a * b_i        (per lane, local carries)
//                                                         m = acc_0 * n0inv     (lane 0, broadcast)
//                                                         acc += n * m
//                                                         acc >>= 32            (lane l takes word 0 of lane l+1)
//
// Carries never cross lanes inside the loop: the carry word that leaves a lane at position L
// lands, after the one-word right shift, on the lane's own top position, so each lane keeps
// L words plus a small overflow word (< 4, by the fixed point V' <= V/2^32 + 3*2^(32L)); the word at
// position L itself is accumulated in 64 bits before the shift.
// Cross-lane carries are resolved once per multiplication (ripple rounds until no lane has a
// pending carry: 1 round in practice), followed by the conditional subtraction of n.  The
// algorithm was validated against Python bigints in a lane-level model before it was written
// in CUDA; tests/test_gpu_parity.py compares it with the thread-per-element path and the
// oracle.  No host emulation exists for this file (warp shuffles): under VMX_HOST_EMUL the host
// code routes everything through the thread-per-element kernels.
#pragma once
#include "layout.cuh"

#ifndef VMX_HOST_EMUL
namespace vmx {

template <int N>
struct Coop {
  static constexpr int LANES = N >= 32 ? 32 : N;
  static constexpr int L = N / LANES;
  static constexpr unsigned MASK = LANES == 32 ? 0xffffffffu : ((1u << LANES) - 1u);
};

constexpr int kCoopWarps = 4;  // warps (residues) per block

template <int N>
__device__ __forceinline__ void coop_load(uint32_t (&x)[Coop<N>::L], const uint32_t* __restrict__ d, size_t cap,
                                          size_t idx, int lane) {
#pragma unroll
  for (int k = 0; k < Coop<N>::L; k++) {
    const int j = lane * Coop<N>::L + k;
    x[k] = d[((size_t)(j >> 2) * cap + idx) * 4 + (j & 3)];
  }
}

template <int N>
__device__ __forceinline__ void coop_store(const uint32_t (&x)[Coop<N>::L], uint32_t* __restrict__ d, size_t cap,
                                           size_t idx, int lane) {
#pragma unroll
  for (int k = 0; k < Coop<N>::L; k++) {
    const int j = lane * Coop<N>::L + k;
    d[((size_t)(j >> 2) * cap + idx) * 4 + (j & 3)] = x[k];
  }
}

// r = a * b * R^-1 mod n, fully reduced.  r may alias a or b.
template <int N>
__device__ __forceinline__ void coop_mul(uint32_t (&r)[Coop<N>::L], const uint32_t (&a)[Coop<N>::L],
                                         const uint32_t (&b)[Coop<N>::L], const uint32_t (&n)[Coop<N>::L],
                                         uint32_t n0inv, int lane) {
  constexpr int L = Coop<N>::L, LANES = Coop<N>::LANES;
  constexpr unsigned MASK = Coop<N>::MASK;
  uint32_t acc[L];
  uint32_t ov = 0;
#pragma unroll
  for (int k = 0; k < L; k++) acc[k] = 0;
#pragma unroll 1
  for (int jl = 0; jl < LANES; jl++) {
#pragma unroll
    for (int kb = 0; kb < L; kb++) {
      const uint32_t bi = __shfl_sync(MASK, b[kb], jl, LANES);
      uint32_t c = 0;
#pragma unroll
      for (int k = 0; k < L; k++) {
        const uint64_t t = (uint64_t)a[k] * bi + acc[k] + c;
        acc[k] = (uint32_t)t;
        c = (uint32_t)(t >> 32);
      }
      // position L of the lane collects full carry words: accumulate it in 64 bits
      uint64_t top64 = (uint64_t)ov + c;
      const uint32_t m = __shfl_sync(MASK, acc[0] * n0inv, 0, LANES);
      c = 0;
#pragma unroll
      for (int k = 0; k < L; k++) {
        const uint64_t t = (uint64_t)n[k] * m + acc[k] + c;
        acc[k] = (uint32_t)t;
        c = (uint32_t)(t >> 32);
      }
      top64 += c;
      uint32_t down = __shfl_down_sync(MASK, acc[0], 1, LANES);
      if (lane == LANES - 1) down = 0;
#pragma unroll
      for (int k = 0; k + 1 < L; k++) acc[k] = acc[k + 1];
      top64 += down;
      acc[L - 1] = (uint32_t)top64;
      ov = (uint32_t)(top64 >> 32);
    }
  }
  // resolve the overflow words: lane l hands its pending carry to lane l+1; the last lane
  // keeps the top word (limb N of the CIOS result, 0 or 1)
  uint32_t top = 0, carry = ov;
  if (lane == LANES - 1) { top = ov; carry = 0; }
  while (__any_sync(MASK, carry != 0)) {
    uint32_t inc = __shfl_up_sync(MASK, carry, 1, LANES);
    if (lane == 0) inc = 0;
#pragma unroll
    for (int k = 0; k < L; k++) {
      const uint64_t s = (uint64_t)acc[k] + inc;
      acc[k] = (uint32_t)s;
      inc = (uint32_t)(s >> 32);
    }
    if (lane == LANES - 1) { top += inc; carry = 0; } else carry = inc;
  }
  // d = acc - n with cross-lane borrows; result = (top:acc) >= n ? d : acc
  uint32_t d[L];
  uint32_t bw = 0, tb = 0;
#pragma unroll
  for (int k = 0; k < L; k++) {
    const uint64_t s = (uint64_t)acc[k] - n[k] - bw;
    d[k] = (uint32_t)s;
    bw = (uint32_t)(s >> 63);
  }
  if (lane == LANES - 1) { tb = bw; bw = 0; }
  while (__any_sync(MASK, bw != 0)) {
    uint32_t inc = __shfl_up_sync(MASK, bw, 1, LANES);
    if (lane == 0) inc = 0;
#pragma unroll
    for (int k = 0; k < L; k++) {
      const uint64_t s = (uint64_t)d[k] - inc;
      d[k] = (uint32_t)s;
      inc = (uint32_t)(s >> 63);
    }
    if (lane == LANES - 1) { tb += inc; bw = 0; } else bw = inc;
  }
  const int use_d = __shfl_sync(MASK, (int)(top >= tb), LANES - 1, LANES);
#pragma unroll
  for (int k = 0; k < L; k++) r[k] = use_d ? d[k] : acc[k];
}

// common prologue: lane / warp ids, modulus limbs of this lane (consts idx 3 = n)
#define VMX_COOP_PROLOGUE(N)                                                          \
  constexpr int L = Coop<N>::L;                                                       \
  const int lane = threadIdx.x & 31;                                                  \
  const int wib = threadIdx.x >> 5;                                                   \
  if (lane >= Coop<N>::LANES) return;                                                 \
  uint32_t nmod[L];                                                                   \
  coop_load<N>(nmod, consts, 4, 3, lane)

// out[i] = a[i]^E, E = e[i] (escalar = 0) or e[0] (escalar = 1); 4-bit windows, top-down; the
// table of 16 powers lives in shared memory; zero digits skip their multiplication (the branch
// is uniform across the warp because the warp owns one exponent).
template <int N>
__global__ void __launch_bounds__(32 * kCoopWarps)
k_coop_exp(const uint32_t* __restrict__ a_, size_t acap, const uint32_t* __restrict__ e_, size_t ecap, int escalar,
           int ebits, size_t n, const uint32_t* __restrict__ consts, uint32_t n0inv, uint32_t* __restrict__ out,
           size_t ocap) {
  __shared__ uint32_t tab[kCoopWarps][16][N];
  VMX_COOP_PROLOGUE(N);
  const size_t i = (size_t)blockIdx.x * kCoopWarps + wib;
  if (i >= n) return;
  const size_t ei = escalar ? 0 : i;
  uint32_t x[L], t[L];
  coop_load<N>(x, a_, acap, i, lane);
#pragma unroll
  for (int k = 0; k < L; k++) { t[k] = x[k]; tab[wib][1][lane * L + k] = x[k]; }
  for (int d = 2; d < 16; d++) {
    coop_mul<N>(t, t, x, nmod, n0inv, lane);
#pragma unroll
    for (int k = 0; k < L; k++) tab[wib][d][lane * L + k] = t[k];
  }
  const int nwin = (ebits + 3) / 4;
  bool started = false;
  for (int w = nwin - 1; w >= 0; w--) {
    if (started) {
#pragma unroll 1
      for (int s = 0; s < 4; s++) coop_mul<N>(t, t, t, nmod, n0inv, lane);
    }
    const uint32_t d = window_bits<N>(e_, ecap, ei, 4 * w, 4);
    if (d) {
      if (!started) {
#pragma unroll
        for (int k = 0; k < L; k++) t[k] = tab[wib][d][lane * L + k];
        started = true;
      } else {
#pragma unroll
        for (int k = 0; k < L; k++) x[k] = tab[wib][d][lane * L + k];
        coop_mul<N>(t, t, x, nmod, n0inv, lane);
      }
    }
  }
  if (!started) coop_load<N>(t, consts, 4, 1, lane);  // exponent 0 -> one
  coop_store<N>(t, out, ocap, i, lane);
}

// Q[m] = base^(2^m), m = 0..len-1 (one warp)
template <int N>
__global__ void __launch_bounds__(32)
k_coop_sqr_chain(const uint32_t* __restrict__ base, size_t bcap, size_t bidx, uint32_t* __restrict__ Q, size_t qcap,
                 int len, const uint32_t* __restrict__ consts, uint32_t n0inv) {
  VMX_COOP_PROLOGUE(N);
  (void)wib;
  uint32_t x[L];
  coop_load<N>(x, base, bcap, bidx, lane);
  coop_store<N>(x, Q, qcap, 0, lane);
  for (int m = 1; m < len; m++) {
    coop_mul<N>(x, x, x, nmod, n0inv, lane);
    coop_store<N>(x, Q, qcap, (size_t)m, lane);
  }
}

// out[oidx + c] = prod_m Y[c * Mcount + m]^(16^m), m = 0..Mcount-1, Horner from the top: one warp (block) per
// column c -- the components of a product-group expProd share the exponents, so their Horner chains (L
// sequential squarings each) run side by side instead of one after the other.
template <int N>
__global__ void __launch_bounds__(32)
k_coop_horner(const uint32_t* __restrict__ Y, size_t ycap, int Mcount, uint32_t* __restrict__ out, size_t ocap,
              size_t oidx, const uint32_t* __restrict__ consts, uint32_t n0inv) {
  VMX_COOP_PROLOGUE(N);
  (void)wib;
  const size_t base = (size_t)blockIdx.x * Mcount;
  uint32_t a[L], y[L];
  coop_load<N>(a, Y, ycap, base + (size_t)Mcount - 1, lane);
  for (int m = Mcount - 2; m >= 0; m--) {
#pragma unroll 1
    for (int q = 0; q < 4; q++) coop_mul<N>(a, a, a, nmod, n0inv, lane);
    coop_load<N>(y, Y, ycap, base + (size_t)m, lane);
    coop_mul<N>(a, a, y, nmod, n0inv, lane);
  }
  coop_store<N>(a, out, ocap, oidx + blockIdx.x, lane);
}

// Y[g] = prod_{v=1}^{15} X[15g + v-1]^v (running products), one warp per group
template <int N>
__global__ void __launch_bounds__(32 * kCoopWarps)
k_coop_weighted_small(const uint32_t* __restrict__ X, size_t xcap, size_t ngroups, uint32_t* __restrict__ Y,
                      size_t ycap, const uint32_t* __restrict__ consts, uint32_t n0inv) {
  VMX_COOP_PROLOGUE(N);
  const size_t g = (size_t)blockIdx.x * kCoopWarps + wib;
  if (g >= ngroups) return;
  uint32_t run[L], tot[L], x[L];
  coop_load<N>(run, X, xcap, g * 15 + 14, lane);
#pragma unroll
  for (int k = 0; k < L; k++) tot[k] = run[k];
  for (int v = 14; v >= 1; v--) {
    coop_load<N>(x, X, xcap, g * 15 + v - 1, lane);
    coop_mul<N>(run, run, x, nmod, n0inv, lane);
    coop_mul<N>(tot, tot, run, nmod, n0inv, lane);
  }
  coop_store<N>(tot, Y, ycap, g, lane);
}

// out[i] = prod_{p < parts} in[p*n + i]  and generic short products: one warp per output
template <int N>
__global__ void __launch_bounds__(32 * kCoopWarps)
k_coop_combine_parts(const uint32_t* __restrict__ in, size_t icap, size_t n, int parts, uint32_t* __restrict__ out,
                     size_t ocap, const uint32_t* __restrict__ consts, uint32_t n0inv) {
  VMX_COOP_PROLOGUE(N);
  const size_t i = (size_t)blockIdx.x * kCoopWarps + wib;
  if (i >= n) return;
  uint32_t a[L], x[L];
  coop_load<N>(a, in, icap, i, lane);
  for (int p = 1; p < parts; p++) {
    coop_load<N>(x, in, icap, (size_t)p * n + i, lane);
    coop_mul<N>(a, a, x, nmod, n0inv, lane);
  }
  coop_store<N>(a, out, ocap, i, lane);
}

// out[i] = a[i] * b[i]  (self-test of the cooperative multiplier against k_mul)
template <int N>
__global__ void __launch_bounds__(32 * kCoopWarps)
k_coop_mul(const uint32_t* __restrict__ a_, size_t acap, const uint32_t* __restrict__ b_, size_t bcap, size_t n,
           uint32_t* __restrict__ out, size_t ocap, const uint32_t* __restrict__ consts, uint32_t n0inv) {
  VMX_COOP_PROLOGUE(N);
  const size_t i = (size_t)blockIdx.x * kCoopWarps + wib;
  if (i >= n) return;
  uint32_t a[L], b[L];
  coop_load<N>(a, a_, acap, i, lane);
  coop_load<N>(b, b_, bcap, i, lane);
  coop_mul<N>(a, a, b, nmod, n0inv, lane);
  coop_store<N>(a, out, ocap, i, lane);
}

}  // namespace vmx
#endif  // !VMX_HOST_EMUL
