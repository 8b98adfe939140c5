// kernels_member.cuh -- group-membership check on import for safe-prime groups.
//
// For p = 2q + 1 the order-q subgroup of Z_p^* is the set of quadratic residues, so
// PGroup.toElementArray's membership test (hvzk/PoSBasicTW.java:507,787-792,
// mixnet/ShufflerElGamalSession.java:205) is a Legendre symbol.  Euler's criterion x^q costs
// VAR(L_q) ~ 3,600 modmuls per element -- more than a whole proof-of-shuffle verification -- so the
// symbol is computed with the binary Jacobi algorithm instead: ~1.4 * bits subtract-and-shift
// steps of O(N) word operations, about the price of 50-100 modmuls, on the ALU pipe (no IMAD).
//
// One thread per element; both running values (a, m) stay in registers (2N words).  Per step:
//     strip the factors of two of a        (2|m) = -1 iff m = 3,5 (mod 8), applied for odd counts
//     if a < m: swap                       reciprocity: sign flips iff a = m = 3 (mod 4)
//     a <- a - m                           now even (or zero: finished, gcd = m)
// The residues are in Montgomery form x*R mod p; (R|p) = (2|p)^(32N) = 1, so the symbol of the
// stored value IS the symbol of x -- no conversion.
#pragma once
#include "layout.cuh"

namespace vmx {

VMX_DEV int ctz32(uint32_t v) {
#ifdef VMX_HOST_EMUL
  return __builtin_ctz(v);
#else
  return __ffs((int)v) - 1;
#endif
}

// returns +1, -1 or 0 (gcd != 1 or input zero)
template <int N>
VMX_DEV int jacobi_binary(uint32_t (&a)[N], uint32_t (&m)[N]) {
  uint32_t t = 0;  // sign bit
  for (int guard = 0; guard < 64 * N + 64; guard++) {
    // a == 0 mod 2^32: whole-limb shifts (32 factors of two: even count, no sign change)
    while (a[0] == 0) {
      uint32_t nz = 0;
#pragma unroll
      for (int j = 0; j < N; j++) nz |= a[j];
      if (nz == 0) {  // a == 0: gcd = m
        uint32_t rest = m[0] ^ 1u;
#pragma unroll
        for (int j = 1; j < N; j++) rest |= m[j];
        return rest == 0 ? (t ? -1 : 1) : 0;
      }
#pragma unroll
      for (int j = 0; j + 1 < N; j++) a[j] = a[j + 1];
      a[N - 1] = 0;
    }
    const int tz = ctz32(a[0]);
    if (tz) {
#pragma unroll
      for (int j = 0; j + 1 < N; j++) a[j] = __funnelshift_r(a[j], a[j + 1], tz);
      a[N - 1] >>= tz;
      const uint32_t m8 = m[0] & 7u;
      if ((tz & 1) && (m8 == 3u || m8 == 5u)) t ^= 1u;
    }
    // a odd, m odd: order them
    uint32_t d, brw;
    sub_cc(d, a[0], m[0]);
#pragma unroll
    for (int j = 1; j < N; j++) subc_cc(d, a[j], m[j]);
    subc(brw, 0, 0);
    const bool lt = brw != 0;
    if (lt && (a[0] & 3u) == 3u && (m[0] & 3u) == 3u) t ^= 1u;
#pragma unroll
    for (int j = 0; j < N; j++) {
      const uint32_t x = a[j], y = m[j];
      a[j] = lt ? y : x;
      m[j] = lt ? x : y;
    }
    sub_cc(a[0], a[0], m[0]);
#pragma unroll
    for (int j = 1; j < N; j++) subc_cc(a[j], a[j], m[j]);
  }
  return 0;
}

// *err |= kErrMember if any element has Legendre symbol != +1 modulo the (prime) modulus
template <int N>
VMX_KERNEL(N) k_jacobi(const uint32_t* __restrict__ a_, size_t acap, size_t n, int* __restrict__ err,
                       const __grid_constant__ MontParams<N> M) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t a[N], m[N];
  load_elem<N>(a, a_, acap, i);
#pragma unroll
  for (int j = 0; j < N; j++) m[j] = M.n[j];
  if (jacobi_binary<N>(a, m) != 1) atomicOr(err, (int)kErrMember);
}

}  // namespace vmx
