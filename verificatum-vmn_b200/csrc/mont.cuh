// mont.cuh -- thread-per-element Montgomery multiplication (CIOS, two rows per loop trip).
//
// One thread owns one N-limb residue: `a` (N registers) and the running CIOS accumulator
// `t` (N+2 registers) stay in registers; the second operand is streamed one 32-bit word per
// row; the modulus is read straight from the kernel-parameter constant bank (no registers).
//
// Why two rows per trip and an "odd block": IMAD.WIDE needs an even-aligned (lo,hi) register
// pair.  A product a[j]*b[i] lands on columns (i+j, i+j+1); half of the products of every row
// are therefore mis-aligned with respect to a fixed accumulator.  We call products on an even
// absolute column E-class (fused straight into `t`) and the others O-class.  O-class products
// of a row PAIR (i, i+1) all fall on the same pairs (1,2),(3,4),..; they are accumulated --
// fused, with carry chains -- into a small separate block `o` of 2*PB registers that is
// aligned on its own, and folded into `t` with one add chain per block.  The last E chain of
// the trip writes its results two registers lower, which implements the CIOS right shift by
// two words for free (no MOVs).  Result: 4*N IMAD.WIDE + ~1.1*N IADD3 per trip.
//
// Algorithmic unit: 1 modmul = 2N^2+N word MACs; this kernel executes exactly
// 2N^2 IMAD.WIDE + N mul.lo for it.
#pragma once
#include "cuda_compat.cuh"
#include "ptx_arith.cuh"
#include "fp256.cuh"

namespace vmx {

template <int N>
struct MontParams {
  uint32_t n[N];    // modulus, little-endian limbs
  uint32_t n0inv;   // -n^{-1} mod 2^32
};

constexpr int kPairBlock = 8;  // O-class pairs per block (16 columns)

// One trip = rows (i, i+1).  t holds relative columns 0..N+1 on entry and on exit (exit
// columns are entry columns 2..N+3).
template <int N>
VMX_DEV void mont_rowpair(uint32_t (&t)[N + 2], const uint32_t (&a)[N], uint32_t b0, uint32_t b1,
                          const MontParams<N>& M) {
  static_assert(N % (2 * kPairBlock) == 0, "N must be a multiple of 16");
  constexpr int PB = kPairBlock;
  uint32_t top2 = 0;  // column N+2

  // E-class, row 0: a[even j] * b0 on pairs (j, j+1)
  mad_wide_cc(t[0], t[1], a[0], b0);
#pragma unroll
  for (int j = 2; j < N; j += 2) madc_wide_cc(t[j], t[j + 1], a[j], b0);
  addc_cc(t[N], t[N], 0);
  addc(t[N + 1], t[N + 1], 0);

  const uint32_t m0 = t[0] * M.n0inv;

  // E-class, row 0 reduction: n[even j] * m0
  mad_wide_cc(t[0], t[1], M.n[0], m0);
#pragma unroll
  for (int j = 2; j < N; j += 2) madc_wide_cc(t[j], t[j + 1], M.n[j], m0);
  addc_cc(t[N], t[N], 0);
  addc(t[N + 1], t[N + 1], 0);

  // E-class, row 1: a[odd j] * b1 on pairs (1+j, 2+j)
  mad_wide_cc(t[2], t[3], a[1], b1);
#pragma unroll
  for (int j = 3; j < N; j += 2) madc_wide_cc(t[j + 1], t[j + 2], a[j], b1);
  addc(top2, top2, 0);

  // O-class: pair q sits on columns (c, c+1), c = 2q+1, and receives
  //   a[c]*b0 + n[c]*m0 + a[c-1]*b1 + n[c-1]*m1
  uint32_t m1 = 0;
  uint32_t cw = 0;  // carry word handed from one block to the next
#pragma unroll
  for (int kb = 0; kb < N / 2 / PB; kb++) {
    uint32_t o[2 * PB];
    uint32_t oc;
    const int c0 = 2 * kb * PB + 1;
    // chain A (no carries between pairs; the first pair absorbs the incoming carry word)
    mad_wide_cc3(o[0], o[1], a[c0], b0, cw, 0);
#pragma unroll
    for (int r = 1; r < PB; r++) mul_wide(o[2 * r], o[2 * r + 1], a[c0 + 2 * r], b0);
    // chain B
    mad_wide_cc(o[0], o[1], M.n[c0], m0);
#pragma unroll
    for (int r = 1; r < PB; r++) madc_wide_cc(o[2 * r], o[2 * r + 1], M.n[c0 + 2 * r], m0);
    addc(oc, 0, 0);
    // chain C
    mad_wide_cc(o[0], o[1], a[c0 - 1], b1);
#pragma unroll
    for (int r = 1; r < PB; r++) madc_wide_cc(o[2 * r], o[2 * r + 1], a[c0 - 1 + 2 * r], b1);
    addc(oc, oc, 0);
    if (kb == 0) m1 = (t[1] + o[0]) * M.n0inv;
    // chain D
    mad_wide_cc(o[0], o[1], M.n[c0 - 1], m1);
#pragma unroll
    for (int r = 1; r < PB; r++) madc_wide_cc(o[2 * r], o[2 * r + 1], M.n[c0 - 1 + 2 * r], m1);
    addc(oc, oc, 0);
    // fold the block into t
    add_cc(t[c0], t[c0], o[0]);
#pragma unroll
    for (int r = 1; r < 2 * PB; r++) addc_cc(t[c0 + r], t[c0 + r], o[r]);
    addc(cw, oc, 0);
  }
  add_cc(t[N + 1], t[N + 1], cw);
  addc(top2, top2, 0);

  // E-class, row 1 reduction: n[odd j] * m1 on pairs (1+j, 2+j), written two columns lower.
  mad_wide_cc3(t[0], t[1], M.n[1], m1, t[2], t[3]);
#pragma unroll
  for (int j = 3; j < N; j += 2) madc_wide_cc3(t[j - 1], t[j], M.n[j], m1, t[j + 1], t[j + 2]);
  addc_cc(t[N], top2, 0);
  addc(t[N + 1], 0, 0);
}

// r = t - n if t >= n else t, where t has N+1 significant words (t < 2n).
template <int N>
VMX_DEV void mont_final_sub(uint32_t (&r)[N], const uint32_t (&t)[N + 2], const MontParams<N>& M) {
  uint32_t d[N];
  uint32_t top;
  sub_cc(d[0], t[0], M.n[0]);
#pragma unroll
  for (int j = 1; j < N; j++) subc_cc(d[j], t[j], M.n[j]);
  subc(top, t[N], 0);
  // top == 0  -> t >= n (no borrow out of the N+1 word subtraction): take d; else keep t
  const bool keep = (top != 0);
#pragma unroll
  for (int j = 0; j < N; j++) r[j] = keep ? t[j] : d[j];
}

struct Word2 { uint32_t x, y; };

// The accumulator crosses the loop back edge as 64-bit values: ptxas then keeps every (lo, hi) word
// pair in an aligned register pair for the whole loop.  With 32-bit loop-carried words it picked
// an unaligned assignment and re-paired them with ~150 IMAD.MOV/MOV per trip (against 384
// IMAD.WIDE), which cost ~12 % of the integer pipe (profiles/r01_ncu_full_exp_fixed_exp_var.txt).
#ifndef VMX_HOST_EMUL
VMX_DEV uint64_t pack_pair(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
VMX_DEV void unpack_pair(uint32_t& lo, uint32_t& hi, uint64_t v) { asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); }
#else
VMX_DEV uint64_t pack_pair(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }
VMX_DEV void unpack_pair(uint32_t& lo, uint32_t& hi, uint64_t v) { lo = (uint32_t)v; hi = (uint32_t)(v >> 32); }
#endif

// a <- a * b * R^{-1} mod n, with b streamed through `ld2(i)` (any callable returning the
// word pair (b[i], b[i+1]) for even i).  Result fully reduced to [0, n).
template <int N, typename Loader>
VMX_DEV void mont_mul(uint32_t (&a)[N], Loader ld2, const MontParams<N>& M) {
  if constexpr (N == 8) {  // 256-bit residues (curve groups): separated product + reduction, fp256.cuh
    Fp256 F;
    uint32_t b[8];
#pragma unroll
    for (int i = 0; i < 8; i += 2) { const Word2 w = ld2(i); b[i] = w.x; b[i + 1] = w.y; }
#pragma unroll
    for (int i = 0; i < 8; i++) F.n[i] = M.n[i];
    F.n0inv = M.n0inv;
    F.solinas = 0;
    fp_mul_inline<false>(a, a, b, F);
    return;
  } else {
  uint64_t T[N / 2 + 1];
#pragma unroll
  for (int k = 0; k < N / 2 + 1; k++) T[k] = 0;
  uint32_t t[N + 2];
  // the operand words of trip i+1 are requested before trip i is computed (the loads are
  // gathers from L2/HBM tables in the exponentiation kernels; one trip hides their latency)
  Word2 b = ld2(0);
#pragma unroll 1
  for (int i = 0; i < N; i += 2) {
    const Word2 nb = ld2(i + 2 < N ? i + 2 : i);
#pragma unroll
    for (int k = 0; k < N / 2 + 1; k++) unpack_pair(t[2 * k], t[2 * k + 1], T[k]);
    mont_rowpair<N>(t, a, b.x, b.y, M);
#pragma unroll
    for (int k = 0; k < N / 2 + 1; k++) T[k] = pack_pair(t[2 * k], t[2 * k + 1]);
    b = nb;
  }
#pragma unroll
  for (int k = 0; k < N / 2 + 1; k++) unpack_pair(t[2 * k], t[2 * k + 1], T[k]);
  mont_final_sub<N>(a, t, M);
  }
}

// ------------------------------------------------------------------------------------------------ squaring
// a^2 with every cross product a[i]*a[j], i != j, of DIFFERENT 16-word blocks computed once.
//
// Write a = sum_I A_I B^(16 I) (B = 2^32, A_I = 16 words).  Then
//     a^2 = sum_I A_I^2 B^(32 I)  +  sum_I (2 A_I) B^(16 I) * A_{>I},      A_{>I} = sum_{j >= 16(I+1)} a[j] B^j.
// Row i of block I (i = 16 I + r) therefore multiplies
//     the diagonal block   a[j], 16 I <= j < 16(I+1),  by x = a[i]               (the block squared in full)
//     the columns beyond   a[j], j >= 16(I+1),          by y = word r of 2 A_I   (each cross block once, doubled)
// and nothing below the diagonal block.  2 A_I has 16 words d'[r] = (a[i] << 1) | (r ? a[i-1] >> 31 : 0) and one
// bit c_I = a[16 I + 15] >> 31 on top; the top bit contributes c_I * B^(16(I+1)) * A_{>I}, which sits exactly on the
// accumulator's own columns once the block's 16 rows are done (relative column j <-> a[j]): one masked add chain per
// block boundary.  Column index ranges are static per block (the row-wise CIOS keeps columns in registers, so a
// row cannot skip a dynamic number of them): one loop per block, N/16 loops.  The reduction rows are those of the
// multiplication.  Work: N^2 (reduction) + 16 * sum_I (N - 16 I) = N^2 + 8 N (N/16 + 1) IMAD.WIDE:
// 14,592 at N = 96 (multiplication 18,432: 0.79), 6,656 at N = 64 (8,192: 0.81).  The order of the additions differs
// from mont_mul(a, a), the sum does not: the accumulator ends at (a^2 + m n) / R < 2n with the same m.
//
// Structure of a trip as in mont_rowpair (E/O column classes, O-class products in an aligned side block).
template <int N, int D0, int J0>
VMX_DEV void mont_sqr_rowpair(uint32_t (&t)[N + 2], const uint32_t (&a)[N], uint32_t x0, uint32_t x1, uint32_t y0,
                              uint32_t y1, const MontParams<N>& M) {
  static_assert(N % (2 * kPairBlock) == 0 && D0 % (2 * kPairBlock) == 0 && J0 % (2 * kPairBlock) == 0 && J0 > D0 && J0 <= N,
                "blocks are multiples of 16 words");
  constexpr int PB = kPairBlock;
  uint32_t top2 = 0;  // column N+2

  // E-class, row 0: a[even j >= D0] * (x0 | y0) on pairs (j, j+1)
  mad_wide_cc(t[D0], t[D0 + 1], a[D0], x0);
#pragma unroll
  for (int j = D0 + 2; j < N; j += 2) madc_wide_cc(t[j], t[j + 1], a[j], j < J0 ? x0 : y0);
  addc_cc(t[N], t[N], 0);
  addc(t[N + 1], t[N + 1], 0);

  const uint32_t m0 = t[0] * M.n0inv;

  // E-class, row 0 reduction: n[even j] * m0
  mad_wide_cc(t[0], t[1], M.n[0], m0);
#pragma unroll
  for (int j = 2; j < N; j += 2) madc_wide_cc(t[j], t[j + 1], M.n[j], m0);
  addc_cc(t[N], t[N], 0);
  addc(t[N + 1], t[N + 1], 0);

  // E-class, row 1: a[odd j >= D0] * (x1 | y1) on pairs (1+j, 2+j)
  mad_wide_cc(t[D0 + 2], t[D0 + 3], a[D0 + 1], x1);
#pragma unroll
  for (int j = D0 + 3; j < N; j += 2) madc_wide_cc(t[j + 1], t[j + 2], a[j], j < J0 ? x1 : y1);
  addc(top2, top2, 0);

  // O-class: pair q on columns (c, c+1), c = 2q+1, receives a[c]*s0 + n[c]*m0 + a[c-1]*s1 + n[c-1]*m1, the
  // a-terms only for columns of the diagonal block and beyond
  uint32_t m1 = 0;
  uint32_t cw = 0;
#pragma unroll
  for (int kb = 0; kb < N / 2 / PB; kb++) {
    uint32_t o[2 * PB];
    uint32_t oc;
    const int c0 = 2 * kb * PB + 1;
    const bool has_a = (c0 - 1 >= D0);
    const uint32_t s0 = (c0 - 1 < J0) ? x0 : y0, s1 = (c0 - 1 < J0) ? x1 : y1;
    if (has_a) {
      mad_wide_cc3(o[0], o[1], a[c0], s0, cw, 0);
#pragma unroll
      for (int r = 1; r < PB; r++) mul_wide(o[2 * r], o[2 * r + 1], a[c0 + 2 * r], s0);
      mad_wide_cc(o[0], o[1], M.n[c0], m0);
#pragma unroll
      for (int r = 1; r < PB; r++) madc_wide_cc(o[2 * r], o[2 * r + 1], M.n[c0 + 2 * r], m0);
      addc(oc, 0, 0);
      mad_wide_cc(o[0], o[1], a[c0 - 1], s1);
#pragma unroll
      for (int r = 1; r < PB; r++) madc_wide_cc(o[2 * r], o[2 * r + 1], a[c0 - 1 + 2 * r], s1);
      addc(oc, oc, 0);
    } else {
      mad_wide_cc3(o[0], o[1], M.n[c0], m0, cw, 0);
#pragma unroll
      for (int r = 1; r < PB; r++) mul_wide(o[2 * r], o[2 * r + 1], M.n[c0 + 2 * r], m0);
      oc = 0;
    }
    if (kb == 0) m1 = (t[1] + o[0]) * M.n0inv;
    mad_wide_cc(o[0], o[1], M.n[c0 - 1], m1);
#pragma unroll
    for (int r = 1; r < PB; r++) madc_wide_cc(o[2 * r], o[2 * r + 1], M.n[c0 - 1 + 2 * r], m1);
    addc(oc, oc, 0);
    add_cc(t[c0], t[c0], o[0]);
#pragma unroll
    for (int r = 1; r < 2 * PB; r++) addc_cc(t[c0 + r], t[c0 + r], o[r]);
    addc(cw, oc, 0);
  }
  add_cc(t[N + 1], t[N + 1], cw);
  addc(top2, top2, 0);

  // E-class, row 1 reduction: n[odd j] * m1 on pairs (1+j, 2+j), written two columns lower.
  mad_wide_cc3(t[0], t[1], M.n[1], m1, t[2], t[3]);
#pragma unroll
  for (int j = 3; j < N; j += 2) madc_wide_cc3(t[j - 1], t[j], M.n[j], m1, t[j + 1], t[j + 2]);
  addc_cc(t[N], top2, 0);
  addc(t[N + 1], 0, 0);
}

// The BS rows of block I (BS / 2 trips), the correction row of its boundary, then block I + 1.  `x` holds the words
// (a[BS I], a[BS I + 1]) on entry; `ld2(i)` returns (a[i], a[i+1]) from the copy of a in shared memory.
// BS (a multiple of 16 dividing N) trades multiplications against code: every block is its own unrolled loop body
// (BS = 16 at N = 96: 6 bodies, 45 KB of SASS, 0.79 of a multiplication; 32: 3 bodies, 24 KB, 0.83; 48: 2 bodies, 17 KB,
// 0.875) and the loop of an exponentiation kernel holds two or three multiplications besides.
template <int N, int BS, int I, typename Loader>
VMX_DEV void mont_sqr_blocks(uint64_t (&T)[N / 2 + 1], const uint32_t (&a)[N], Loader ld2, Word2 x,
                             const MontParams<N>& M) {
  constexpr int D0 = BS * I, J0 = BS * (I + 1);
  uint32_t t[N + 2];
  uint32_t prevtop = 0;
#pragma unroll 1
  for (int r = 0; r < BS; r += 2) {
    const int i = D0 + r;
    const Word2 nx = ld2(i + 2 < N ? i + 2 : i);
    const uint32_t y0 = (x.x << 1) | prevtop;
    const uint32_t y1 = (x.y << 1) | (x.x >> 31);
    prevtop = x.y >> 31;
#pragma unroll
    for (int k = 0; k < N / 2 + 1; k++) unpack_pair(t[2 * k], t[2 * k + 1], T[k]);
    mont_sqr_rowpair<N, D0, J0>(t, a, x.x, x.y, y0, y1, M);
#pragma unroll
    for (int k = 0; k < N / 2 + 1; k++) T[k] = pack_pair(t[2 * k], t[2 * k + 1]);
    x = nx;
  }
  if constexpr (J0 < N) {
    // the top bit of 2 A_I times A_{>I}: relative column j of the accumulator is now absolute column 16(I+1) + j
    const uint32_t mask = 0u - prevtop;
#pragma unroll
    for (int k = 0; k < N / 2 + 1; k++) unpack_pair(t[2 * k], t[2 * k + 1], T[k]);
    add_cc(t[J0], t[J0], a[J0] & mask);
#pragma unroll
    for (int j = J0 + 1; j < N; j++) addc_cc(t[j], t[j], a[j] & mask);
    addc_cc(t[N], t[N], 0);
    addc(t[N + 1], t[N + 1], 0);
#pragma unroll
    for (int k = 0; k < N / 2 + 1; k++) T[k] = pack_pair(t[2 * k], t[2 * k + 1]);
    mont_sqr_blocks<N, BS, I + 1>(T, a, ld2, x, M);
  }
}

// a <- a * a * R^{-1} mod n; `ld2` streams a copy of a (shared memory).  Result fully reduced.
template <int N, int BS, typename Loader>
VMX_DEV void mont_sqr_tri(uint32_t (&a)[N], Loader ld2, const MontParams<N>& M) {
  static_assert(BS % 16 == 0 && N % BS == 0 && N >= 2 * BS, "block-triangular squaring needs at least two blocks");
  uint64_t T[N / 2 + 1];
#pragma unroll
  for (int k = 0; k < N / 2 + 1; k++) T[k] = 0;
  mont_sqr_blocks<N, BS, 0>(T, a, ld2, ld2(0), M);
  uint32_t t[N + 2];
#pragma unroll
  for (int k = 0; k < N / 2 + 1; k++) unpack_pair(t[2 * k], t[2 * k + 1], T[k]);
  mont_final_sub<N>(a, t, M);
}

}  // namespace vmx
