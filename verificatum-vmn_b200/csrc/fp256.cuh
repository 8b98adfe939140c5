// fp256.cuh -- 8-limb (256-bit) Montgomery arithmetic, one residue per thread, all in registers.
//
// Used for the coordinate field of the 256-bit prime curves behind ECqPGroup (P-256 is the
// reference's benchmark curve, demo/mixnet/benchmarks/bench_config:33-50) and for their exponent
// ring Z_q.  The thread-per-element CIOS of mont.cuh needs N to be a multiple of 16; at N = 8 a
// product is so short that the separated form is the better fit:
//
//   1. 16-word product by rows, E/O split: a_j * b_i lands on columns (i+j, i+j+1); products on
//      an even column accumulate in `e`, the others in `o` (o[k] = column k+1), so that every
//      multiply-accumulate is ONE IMAD.WIDE on an aligned register pair with a carry chain along
//      the row: 64 IMAD.WIDE + 14 carry words + one 15-word merge.
//   2. Montgomery reduction, word serial.  Generic modulus: 64 more multiply-accumulates.
//      P-256 (p = 2^256 - 2^224 + 2^192 + 2^96 - 1, -1/p = 1 mod 2^32): m_i = t_i and
//      m_i * p is four shifted copies of m_i, so a row is a 6-word add chain and no multiply.
//      Both produce the same residues (the special path is the generic recurrence with the
//      products written out), so arrays, tables and tests are independent of the path.
//   3. One conditional subtraction (p and q of P-256 use all 256 bits: the carry word counts).
//
// Unit of work: one field multiplication = 2*8^2 + 8 = 136 word MACs nominal; the
// P-256 path executes 64 of them plus ~90 adds.
#pragma once
#include "cuda_compat.cuh"
#include "ptx_arith.cuh"

namespace vmx {

struct Fp256 {
  uint32_t n[8];    // modulus, little-endian limbs
  uint32_t n0inv;   // -n^{-1} mod 2^32
  uint32_t solinas; // 1: n is the P-256 prime (shift-and-add reduction)
};

// t = a * b (16 words)
VMX_DEV void fp_mul_wide(uint32_t (&t)[16], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
  uint32_t e[16], o[16];
#pragma unroll
  for (int k = 0; k < 16; k++) { e[k] = 0; o[k] = 0; }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    if ((i & 1) == 0) {
      // even j -> e pairs (i+j, i+j+1); odd j -> o pairs at o[i+j-1], o[i+j]
      mad_wide_cc(e[i], e[i + 1], a[0], b[i]);
#pragma unroll
      for (int j = 2; j < 8; j += 2) madc_wide_cc(e[i + j], e[i + j + 1], a[j], b[i]);
      if (i + 8 < 16) addc(e[i + 8], e[i + 8], 0);
      mad_wide_cc(o[i], o[i + 1], a[1], b[i]);
#pragma unroll
      for (int j = 3; j < 8; j += 2) madc_wide_cc(o[i + j - 1], o[i + j], a[j], b[i]);
      if (i + 8 < 15) addc(o[i + 8], o[i + 8], 0);
    } else {
      // odd j -> column i+j even -> e pairs; even j -> o pairs at o[i+j-1], o[i+j]
      mad_wide_cc(e[i + 1], e[i + 2], a[1], b[i]);
#pragma unroll
      for (int j = 3; j < 8; j += 2) madc_wide_cc(e[i + j], e[i + j + 1], a[j], b[i]);
      if (i + 9 < 16) addc(e[i + 9], e[i + 9], 0);
      mad_wide_cc(o[i - 1], o[i], a[0], b[i]);
#pragma unroll
      for (int j = 2; j < 8; j += 2) madc_wide_cc(o[i + j - 1], o[i + j], a[j], b[i]);
      if (i + 7 < 15) addc(o[i + 7], o[i + 7], 0);
    }
  }
  t[0] = e[0];
  add_cc(t[1], e[1], o[0]);
#pragma unroll
  for (int k = 2; k < 15; k++) addc_cc(t[k], e[k], o[k - 1]);
  addc(t[15], e[15], o[14]);
}

// a^2 as 16 words: the 28 products a_i a_j (i < j) once, doubled, plus the 8 squares on the (aligned) even
// pairs: 36 IMAD.WIDE instead of 64.
VMX_DEV void fp_sqr_wide(uint32_t (&t)[16], const uint32_t (&a)[8]) {
  uint32_t e[16], o[16];
#pragma unroll
  for (int k = 0; k < 16; k++) { e[k] = 0; o[k] = 0; }
  // row i holds a_j * a_i for j > i; column i+j even -> e pair (i+j, i+j+1), odd -> o[i+j-1], o[i+j]
#pragma unroll
  for (int i = 0; i < 7; i++) {
    // same-parity partners j = i+2, i+4, ... -> even columns
    if (i + 2 < 8) {
      mad_wide_cc(e[2 * i + 2], e[2 * i + 3], a[i + 2], a[i]);
#pragma unroll
      for (int j = i + 4; j < 8; j += 2) madc_wide_cc(e[i + j], e[i + j + 1], a[j], a[i]);
      // last pair written: j = i + 2 * ((7 - i) / 2); carry word lands two columns above it
      const int jl = i + 2 * ((7 - i) / 2);
      if (i + jl + 2 < 16) addc(e[i + jl + 2], e[i + jl + 2], 0);
    }
    // opposite-parity partners j = i+1, i+3, ... -> odd columns
    mad_wide_cc(o[2 * i], o[2 * i + 1], a[i + 1], a[i]);
#pragma unroll
    for (int j = i + 3; j < 8; j += 2) madc_wide_cc(o[i + j - 1], o[i + j], a[j], a[i]);
    const int jo = i + 1 + 2 * ((7 - i - 1) / 2);
    if (i + jo + 1 < 15) addc(o[i + jo + 1], o[i + jo + 1], 0);
  }
  // s = e + (o << 32)
  uint32_t s[16];
  s[0] = e[0];
  add_cc(s[1], e[1], o[0]);
#pragma unroll
  for (int k = 2; k < 15; k++) addc_cc(s[k], e[k], o[k - 1]);
  addc(s[15], e[15], o[14]);
  // t = 2 s + sum_i a_i^2 2^(64 i)
  add_cc(s[0], s[0], s[0]);
#pragma unroll
  for (int k = 1; k < 15; k++) addc_cc(s[k], s[k], s[k]);
  addc(s[15], s[15], s[15]);
  mad_wide_cc(s[0], s[1], a[0], a[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) madc_wide_cc(s[2 * i], s[2 * i + 1], a[i], a[i]);
#pragma unroll
  for (int k = 0; k < 16; k++) t[k] = s[k];
}

// r = t / 2^256 mod n (Montgomery reduction of a 16-word value < n * 2^256), fully reduced.
// SOL (compile time): n is the P-256 prime.  Kernels are instantiated for both so that a curve context
// carries only the reduction it uses (the straight-line code of one point addition is ~50 KB of SASS).
// The P-256 prime, for the instantiations that know it at compile time (immediates instead of loads).
VMX_DEV uint32_t p256_word(int j) { return j < 3 ? 0xffffffffu : j < 6 ? 0u : j == 6 ? 1u : 0xffffffffu; }

template <bool SOL>
VMX_DEV void fp_redc(uint32_t (&r)[8], uint32_t (&t)[16], const Fp256& F) {
  uint32_t extra = 0;  // pending carry into column i+9
  if (SOL) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const uint32_t m = t[i];
      const uint32_t negm = 0u - m;
      const uint32_t hi = m - (m != 0 ? 1u : 0u) + extra;  // <= 2^32 - 1: no overflow
      add_cc(t[i + 3], t[i + 3], m);
      addc_cc(t[i + 4], t[i + 4], 0);
      addc_cc(t[i + 5], t[i + 5], 0);
      addc_cc(t[i + 6], t[i + 6], m);
      addc_cc(t[i + 7], t[i + 7], negm);
      addc_cc(t[i + 8], t[i + 8], hi);
      addc(extra, 0, 0);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const uint32_t m = t[i] * F.n0inv;
      uint32_t c = 0;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const uint64_t s = (uint64_t)m * F.n[j] + t[i + j] + c;
        t[i + j] = (uint32_t)s;
        c = (uint32_t)(s >> 32);
      }
      const uint64_t s = (uint64_t)t[i + 8] + c + extra;
      t[i + 8] = (uint32_t)s;
      extra = (uint32_t)(s >> 32);
    }
  }
  // value = t[8..15] + extra * 2^256 < 2n
  uint32_t d[8], brw;
  sub_cc(d[0], t[8], SOL ? p256_word(0) : F.n[0]);
#pragma unroll
  for (int j = 1; j < 8; j++) subc_cc(d[j], t[8 + j], SOL ? p256_word(j) : F.n[j]);
  subc(brw, extra, 0);
  const bool keep = (brw != 0);  // borrow out of the 9-word subtraction: value < n
#pragma unroll
  for (int j = 0; j < 8; j++) r[j] = keep ? t[8 + j] : d[j];
}

// r = a * b * 2^-256 mod n, expanded in place (the exponent-ring kernels, a few multiplications per thread).
template <bool SOL>
VMX_DEV void fp_mul_inline(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8], const Fp256& F) {
  uint32_t t[16];
  fp_mul_wide(t, a, b);
  fp_redc<SOL>(r, t, F);
}

// The curve kernels CALL their field multiplication: one point addition is 16 multiplications, i.e. ~55 KB
// of SASS when expanded in place, and the loop of a scalar multiplication then runs out of the 32 KB
// instruction cache (measured: 2.9 "no instruction" stall cycles per issued instruction, profiles/).  Operands
// and result travel in registers (a 32-byte struct by value: R4.. in the device ABI; no local memory); the
// 3.5 KB body stays resident in the instruction caches.
struct F8 { uint32_t v[8]; };
#ifndef VMX_HOST_EMUL
#define VMX_FN __device__ __noinline__
#else
#define VMX_FN inline
#endif
template <bool SOL>
VMX_FN F8 fp_mul_fn(F8 a, F8 b, const Fp256* F) {
  uint32_t t[16];
  fp_mul_wide(t, a.v, b.v);
  F8 r;
  fp_redc<SOL>(r.v, t, *F);
  return r;
}
template <bool SOL>
VMX_FN F8 fp_sqr_fn(F8 a, const Fp256* F) {
  uint32_t t[16];
  fp_sqr_wide(t, a.v);
  F8 r;
  fp_redc<SOL>(r.v, t, *F);
  return r;
}
// r = a * b * 2^-256 mod n.  r may alias a or b.
template <bool SOL>
VMX_DEV void fp_mul(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8], const Fp256& F) {
  F8 A, B;
#pragma unroll
  for (int j = 0; j < 8; j++) { A.v[j] = a[j]; B.v[j] = b[j]; }
  const F8 R = fp_mul_fn<SOL>(A, B, &F);
#pragma unroll
  for (int j = 0; j < 8; j++) r[j] = R.v[j];
}
template <bool SOL>
VMX_DEV void fp_sqr(uint32_t (&r)[8], const uint32_t (&a)[8], const Fp256& F) {
  F8 A;
#pragma unroll
  for (int j = 0; j < 8; j++) A.v[j] = a[j];
  const F8 R = fp_sqr_fn<SOL>(A, &F);
#pragma unroll
  for (int j = 0; j < 8; j++) r[j] = R.v[j];
}

// r = a + b mod n (a, b < n)
VMX_DEV void fp_add(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8], const Fp256& F) {
  uint32_t s[8], c;
  add_cc(s[0], a[0], b[0]);
#pragma unroll
  for (int j = 1; j < 8; j++) addc_cc(s[j], a[j], b[j]);
  addc(c, 0, 0);
  uint32_t d[8], brw;
  sub_cc(d[0], s[0], F.n[0]);
#pragma unroll
  for (int j = 1; j < 8; j++) subc_cc(d[j], s[j], F.n[j]);
  subc(brw, c, 0);
  const bool keep = (brw != 0);
#pragma unroll
  for (int j = 0; j < 8; j++) r[j] = keep ? s[j] : d[j];
}

// r = a - b mod n (a, b < n)
VMX_DEV void fp_sub(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8], const Fp256& F) {
  uint32_t d[8], brw;
  sub_cc(d[0], a[0], b[0]);
#pragma unroll
  for (int j = 1; j < 8; j++) subc_cc(d[j], a[j], b[j]);
  subc(brw, 0, 0);
  const uint32_t mask = brw;  // 0 or 0xffffffff
  add_cc(r[0], d[0], F.n[0] & mask);
#pragma unroll
  for (int j = 1; j < 7; j++) addc_cc(r[j], d[j], F.n[j] & mask);
  addc(r[7], d[7], F.n[7] & mask);
}

VMX_DEV void fp_dbl(uint32_t (&r)[8], const uint32_t (&a)[8], const Fp256& F) { fp_add(r, a, a, F); }

VMX_DEV void fp_neg(uint32_t (&r)[8], const uint32_t (&a)[8], const Fp256& F) {
  uint32_t nz = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) nz |= a[j];
  uint32_t d[8];
  sub_cc(d[0], F.n[0], a[0]);
#pragma unroll
  for (int j = 1; j < 7; j++) subc_cc(d[j], F.n[j], a[j]);
  subc(d[7], F.n[7], a[7]);
#pragma unroll
  for (int j = 0; j < 8; j++) r[j] = nz ? d[j] : 0u;
}

VMX_DEV bool fp_is_zero(const uint32_t (&a)[8]) {
  uint32_t nz = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) nz |= a[j];
  return nz == 0;
}
VMX_DEV bool fp_eq(const uint32_t (&a)[8], const uint32_t (&b)[8]) {
  uint32_t x = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) x |= a[j] ^ b[j];
  return x == 0;
}
// a < b as 256-bit integers
VMX_DEV bool fp_lt(const uint32_t (&a)[8], const uint32_t (&b)[8]) {
  uint32_t d, brw;
  sub_cc(d, a[0], b[0]);
#pragma unroll
  for (int j = 1; j < 8; j++) subc_cc(d, a[j], b[j]);
  subc(brw, 0, 0);
  (void)d;
  return brw != 0;
}
VMX_DEV void fp_copy(uint32_t (&r)[8], const uint32_t (&a)[8]) {
#pragma unroll
  for (int j = 0; j < 8; j++) r[j] = a[j];
}

// r = a^e mod n for a 256-bit exponent given as plain limbs (uniform across threads: the bits
// come from the parameter bank).  `one` = R mod n.  Used for inversion (e = n - 2) and square
// roots (e = (n + 1) / 4).
template <bool SOL>
VMX_DEV void fp_pow(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&e)[8], const uint32_t (&one)[8],
                    const Fp256& F) {
  uint32_t acc[8];
  fp_copy(acc, one);
  bool started = false;
#pragma unroll 1
  for (int bit = 255; bit >= 0; bit--) {
    if (started) fp_sqr<SOL>(acc, acc, F);
    if ((e[bit >> 5] >> (bit & 31)) & 1u) {
      if (started) fp_mul<SOL>(acc, acc, a, F); else { fp_copy(acc, a); started = true; }
    }
  }
  fp_copy(r, acc);
}

}  // namespace vmx
