// kernels_ec.cuh -- thread-per-point kernels of the ECqPGroup engine (256-bit prime curves).
//
// Same array operations as the ModPGroup kernels (kernels_elem.cuh, kernels_mexp.cuh), written in
// the additive notation of the curve:  exp -> scalar multiple, mul -> point addition, prod -> sum.
//   k_ec_from_bytes / k_ec_to_bytes   fixed-width two's-complement coordinates, range + on-curve check
//   k_ec_exp_fixed                    sum_k T[k][digit_k(e_i)]  (affine table entries, mixed additions, no doublings)
//   k_ec_exp_var                      4-bit windows over a per-point table of 15 Jacobian multiples
//   k_ec_seg_sum                      one thread per chunk of a segment (Pippenger buckets, prod)
//   k_ec_weighted_small, k_ec_horner  bucket reduction (hex sub-digits) and the final Horner
//   k_fp_inv_up / _down / _block      batched field inversion (Montgomery's trick, block scans)
//   k_ec_finish                       Jacobian scratch + inverted Z -> affine array
// Every producer writes Jacobian points into a 24-limb scratch array; the host then runs the
// batched inversion.  Exponents are ring elements (canonical, 8 limbs).
#pragma once
#include "ec.cuh"
#include "kernels_mexp.cuh"

namespace vmx {

constexpr int kEcThreads = 128;
constexpr int kInvK = 4;  // residues per thread in the batched inversion
#define VMX_EC_KERNEL template <bool SOL> __global__ void __launch_bounds__(kEcThreads)

// ------------------------------------------------------------------ byte codec
// Element i: x at rawx + i*stride, y at rawy + i*stride, `cb` big-endian bytes each (two's complement:
// non-negative values have a zero top byte; the unit element is x = y = -1, every byte 0xff).
// hdr != 0: each coordinate is preceded by the 5-byte leaf header 0x01 || be32(cb), validated here.
VMX_DEV uint32_t ec_be_word(const uint8_t* src, int cb, int j) {
  uint32_t v = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int b = 4 * j + k;
    if (b < cb) v |= (uint32_t)src[cb - 1 - b] << (8 * k);
  }
  return v;
}
VMX_DEV int ec_check_hdr(const uint8_t* h, int cb) {
  return (h[0] != 1 || h[1] != (uint8_t)(cb >> 24) || h[2] != (uint8_t)(cb >> 16) || h[3] != (uint8_t)(cb >> 8) ||
          h[4] != (uint8_t)cb) ? kErrPad : 0;
}

VMX_EC_KERNEL k_ec_from_bytes(const uint8_t* __restrict__ rawx, const uint8_t* __restrict__ rawy, size_t stride,
                              size_t n, int cb, int hdr, uint32_t* __restrict__ out, size_t cap,
                              int* __restrict__ err, const __grid_constant__ EcCurve C) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* sx = rawx + i * stride;
  const uint8_t* sy = rawy + i * stride;
  int bad = 0;
  if (hdr) bad |= ec_check_hdr(sx - 5, cb) | ec_check_hdr(sy - 5, cb);
  bool allff = true;
  for (int k = 0; k < cb; k++) allff = allff && sx[k] == 0xff && sy[k] == 0xff;
  if (allff) {
    if (bad) atomicOr(err, bad);
    ec_store_affine_inf(out, cap, i);
    return;
  }
  for (int k = 0; k < cb - 32; k++) if (sx[k] != 0 || sy[k] != 0) bad |= kErrRange;
  uint32_t x[8], y[8];
#pragma unroll
  for (int j = 0; j < 8; j++) { x[j] = ec_be_word(sx, cb, j); y[j] = ec_be_word(sy, cb, j); }
  if (!fp_lt(x, C.F.n) || !fp_lt(y, C.F.n)) bad |= kErrRange;
  if (!bad) {
    fp_mul<SOL>(x, x, C.r2, C.F);
    fp_mul<SOL>(y, y, C.r2, C.F);
    if (!ec_on_curve<SOL>(x, y, C)) bad |= kErrMember;
  }
  if (bad) atomicOr(err, bad);
  ec_store_affine(x, y, out, cap, i);
}

VMX_EC_KERNEL k_ec_to_bytes(const uint32_t* __restrict__ in, size_t cap, size_t n, int cb, int hdr,
                            uint8_t* __restrict__ rawx, uint8_t* __restrict__ rawy, size_t stride,
                            const __grid_constant__ EcCurve C) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t x[8], y[8];
  ec_load_affine(x, y, in, cap, i);
  uint8_t* dx = rawx + i * stride;
  uint8_t* dy = rawy + i * stride;
  if (hdr) {
    dx[-5] = 1; dx[-4] = (uint8_t)(cb >> 24); dx[-3] = (uint8_t)(cb >> 16); dx[-2] = (uint8_t)(cb >> 8); dx[-1] = (uint8_t)cb;
    dy[-5] = 1; dy[-4] = (uint8_t)(cb >> 24); dy[-3] = (uint8_t)(cb >> 16); dy[-2] = (uint8_t)(cb >> 8); dy[-1] = (uint8_t)cb;
  }
  if (aff_is_inf(x)) {
    for (int k = 0; k < cb; k++) { dx[k] = 0xff; dy[k] = 0xff; }
    return;
  }
  uint32_t one[8];
#pragma unroll
  for (int j = 0; j < 8; j++) one[j] = j == 0 ? 1u : 0u;
  fp_mul<SOL>(x, x, one, C.F);
  fp_mul<SOL>(y, y, one, C.F);
  for (int k = 0; k < cb - 32; k++) { dx[k] = 0; dy[k] = 0; }
#pragma unroll
  for (int j = 0; j < 8; j++) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int b = 4 * j + k;
      if (b < cb) { dx[cb - 1 - b] = (uint8_t)(x[j] >> (8 * k)); dy[cb - 1 - b] = (uint8_t)(y[j] >> (8 * k)); }
    }
  }
}

// ------------------------------------------------------------------ batched inversion
// out[i] = in[i]^-1 over an 8-limb limb-major array; zero entries are treated as one (their
// inverse is never used: they mark the unit element).
//
//   up:    thread t of a block owns kInvK residues (strided by the block size: coalesced), forms
//          their product T_t, the block runs an inclusive prefix and suffix product scan over the T_t
//          in shared memory, and writes  L_t = (prod of the other threads' T)  and the block total.
//   (the host inverts the array of block totals recursively; <= one block: k_fp_inv_block)
//   down:  1/T_t = L_t / total, then Montgomery's trick backwards inside the thread.
// Field multiplications per residue: (4 (K-1) + 16) / K = 7 at K = 4.
struct InvLoad {
  uint32_t z[kInvK][8];
  uint32_t pre[kInvK][8];  // pre[k] = z[0] * ... * z[k]
};
template <bool SOL>
VMX_DEV void inv_load(InvLoad& L, const uint32_t* __restrict__ in, size_t icap, size_t n, size_t base, int tid,
                      int bsize, const EcCurve& C) {
#pragma unroll
  for (int k = 0; k < kInvK; k++) {
    const size_t i = base + (size_t)k * bsize + tid;
    if (i < n) fp_load(L.z[k], in, icap, i); else fp_copy(L.z[k], C.one);
    if (fp_is_zero(L.z[k])) fp_copy(L.z[k], C.one);
    if (k == 0) fp_copy(L.pre[0], L.z[0]); else fp_mul<SOL>(L.pre[k], L.pre[k - 1], L.z[k], C.F);
  }
}
// given inv = 1 / pre[K-1], write the inverses of the thread's residues
template <bool SOL>
VMX_DEV void inv_finish(const InvLoad& L, uint32_t (&inv)[8], uint32_t* __restrict__ out, size_t ocap, size_t n,
                        size_t base, int tid, int bsize, const EcCurve& C) {
#pragma unroll
  for (int k = kInvK - 1; k >= 0; k--) {
    const size_t i = base + (size_t)k * bsize + tid;
    uint32_t zi[8];
    if (k > 0) { fp_mul<SOL>(zi, inv, L.pre[k - 1], C.F); fp_mul<SOL>(inv, inv, L.z[k], C.F); } else fp_copy(zi, inv);
    if (i < n) fp_store(zi, out, ocap, i);
  }
}

#ifndef VMX_HOST_EMUL
// block-wide products: on return `excl` = product of the T of all OTHER threads, `total` = product of all.
// sh: 2 * 8 * blockDim words.
template <bool SOL>
__device__ __forceinline__ void block_products(const uint32_t (&T)[8], uint32_t (&excl)[8], uint32_t (&total)[8],
                                               uint32_t* sh, const EcCurve& C) {
  const int tid = threadIdx.x, bs = blockDim.x;
  uint32_t* shp = sh;            // prefix, word j of thread t at [j * bs + t]
  uint32_t* shs = sh + 8 * bs;   // suffix
  uint32_t p[8], s[8], o[8];
  fp_copy(p, T); fp_copy(s, T);
  for (int d = 1; d < bs; d <<= 1) {
#pragma unroll
    for (int j = 0; j < 8; j++) { shp[j * bs + tid] = p[j]; shs[j * bs + tid] = s[j]; }
    __syncthreads();
    if (tid >= d) {
#pragma unroll
      for (int j = 0; j < 8; j++) o[j] = shp[j * bs + tid - d];
      fp_mul<SOL>(p, p, o, C.F);
    }
    if (tid + d < bs) {
#pragma unroll
      for (int j = 0; j < 8; j++) o[j] = shs[j * bs + tid + d];
      fp_mul<SOL>(s, s, o, C.F);
    }
    __syncthreads();
  }
  // p = T_0..T_t, s = T_t..T_{bs-1}
#pragma unroll
  for (int j = 0; j < 8; j++) { shp[j * bs + tid] = p[j]; shs[j * bs + tid] = s[j]; }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; j++) total[j] = shs[j * bs];  // suffix of thread 0
  if (tid > 0) {
#pragma unroll
    for (int j = 0; j < 8; j++) excl[j] = shp[j * bs + tid - 1];
  } else {
    fp_copy(excl, C.one);
  }
  if (tid + 1 < bs) {
#pragma unroll
    for (int j = 0; j < 8; j++) o[j] = shs[j * bs + tid + 1];
    fp_mul<SOL>(excl, excl, o, C.F);
  }
  __syncthreads();
}

VMX_EC_KERNEL k_fp_inv_up(const uint32_t* __restrict__ in, size_t icap, size_t n, uint32_t* __restrict__ excl_out,
                          size_t ecap, uint32_t* __restrict__ totals, size_t tcap, const __grid_constant__ EcCurve C) {
  __shared__ uint32_t sh[2 * 8 * kEcThreads];
  const size_t base = (size_t)blockIdx.x * (kEcThreads * kInvK);
  InvLoad L;
  inv_load<SOL>(L, in, icap, n, base, threadIdx.x, kEcThreads, C);
  uint32_t excl[8], total[8];
  block_products<SOL>(L.pre[kInvK - 1], excl, total, sh, C);
  fp_store(excl, excl_out, ecap, (size_t)blockIdx.x * kEcThreads + threadIdx.x);
  if (threadIdx.x == 0) fp_store(total, totals, tcap, blockIdx.x);
}

VMX_EC_KERNEL k_fp_inv_down(const uint32_t* __restrict__ in, size_t icap, size_t n, const uint32_t* __restrict__ excl_in,
                            size_t ecap, const uint32_t* __restrict__ totals_inv, size_t tcap,
                            uint32_t* __restrict__ out, size_t ocap, const __grid_constant__ EcCurve C) {
  const size_t base = (size_t)blockIdx.x * (kEcThreads * kInvK);
  InvLoad L;
  inv_load<SOL>(L, in, icap, n, base, threadIdx.x, kEcThreads, C);
  uint32_t inv[8], ti[8];
  fp_load(inv, excl_in, ecap, (size_t)blockIdx.x * kEcThreads + threadIdx.x);
  fp_load(ti, totals_inv, tcap, blockIdx.x);
  fp_mul<SOL>(inv, inv, ti, C.F);
  inv_finish<SOL>(L, inv, out, ocap, n, base, threadIdx.x, kEcThreads, C);
}

// n <= kEcThreads * kInvK: one block, one Fermat inversion (thread 0).
VMX_EC_KERNEL k_fp_inv_block(const uint32_t* __restrict__ in, size_t icap, size_t n, uint32_t* __restrict__ out,
                             size_t ocap, const __grid_constant__ EcCurve C) {
  __shared__ uint32_t sh[2 * 8 * kEcThreads];
  __shared__ uint32_t sh_inv[8];
  InvLoad L;
  inv_load<SOL>(L, in, icap, n, 0, threadIdx.x, kEcThreads, C);
  uint32_t excl[8], total[8];
  block_products<SOL>(L.pre[kInvK - 1], excl, total, sh, C);
  if (threadIdx.x == 0) {
    uint32_t ti[8];
    fp_pow<SOL>(ti, total, C.pm2, C.one, C.F);
#pragma unroll
    for (int j = 0; j < 8; j++) sh_inv[j] = ti[j];
  }
  __syncthreads();
  uint32_t ti[8];
#pragma unroll
  for (int j = 0; j < 8; j++) ti[j] = sh_inv[j];
  fp_mul<SOL>(excl, excl, ti, C.F);
  inv_finish<SOL>(L, excl, out, ocap, n, 0, threadIdx.x, kEcThreads, C);
}
#endif

// One thread per residue, one Fermat inversion each (host emulation of the three kernels above; on
// the device: arrays too small to be worth a scan).
VMX_EC_KERNEL k_fp_inv_each(const uint32_t* __restrict__ in, size_t icap, size_t n, uint32_t* __restrict__ out,
                            size_t ocap, const __grid_constant__ EcCurve C) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t z[8];
  fp_load(z, in, icap, i);
  if (fp_is_zero(z)) fp_copy(z, C.one);
  fp_pow<SOL>(z, z, C.pm2, C.one, C.F);
  fp_store(z, out, ocap, i);
}

// Jacobian scratch + 1/Z -> affine.  Destination index: tw = 0: dst_off + i; tw > 0 (table level tj of a
// fixed-base table of window width tw): item i = (k, r), r < 2^tj, goes to (k << tw) + 2^tj + r.
VMX_EC_KERNEL k_ec_finish(const uint32_t* __restrict__ jac, size_t jcap, const uint32_t* __restrict__ zinv, size_t zcap,
                          size_t n, uint32_t* __restrict__ out, size_t ocap, size_t dst_off, int tw, int tj,
                          const __grid_constant__ EcCurve C) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  size_t dst = dst_off + i;
  if (tw > 0) dst = ((i >> tj) << tw) + ((size_t)1 << tj) + (i & (((size_t)1 << tj) - 1));
  Jac P;
  ec_load_jac(P, jac, jcap, i);
  if (jac_is_inf(P)) { ec_store_affine_inf(out, ocap, dst); return; }
  uint32_t zi[8], zi2[8];
  fp_load(zi, zinv, zcap, i);
  fp_sqr<SOL>(zi2, zi, C.F);
  fp_mul<SOL>(P.X, P.X, zi2, C.F);
  fp_mul<SOL>(zi2, zi2, zi, C.F);
  fp_mul<SOL>(P.Y, P.Y, zi2, C.F);
  ec_store_affine(P.X, P.Y, out, ocap, dst);
}

// ------------------------------------------------------------------ element-wise
// jac[i] = a[i] + b[i]
VMX_EC_KERNEL k_ec_add(const uint32_t* __restrict__ a_, size_t acap, const uint32_t* __restrict__ b_, size_t bcap,
                       size_t n, uint32_t* __restrict__ jac, size_t jcap, const __grid_constant__ EcCurve C) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t x[8], y[8];
  Jac P;
  ec_load_affine(x, y, a_, acap, i);
  if (aff_is_inf(x)) jac_set_inf(P, C); else jac_from_affine(P, x, y, C);
  ec_load_affine(x, y, b_, bcap, i);
  if (!aff_is_inf(x)) jac_madd<SOL>(P, x, y, C);
  ec_store_jac(P, jac, jcap, i);
}

// out[i] = -a[i]
VMX_EC_KERNEL k_ec_neg(const uint32_t* __restrict__ a_, size_t acap, size_t n, uint32_t* __restrict__ out, size_t ocap,
                       const __grid_constant__ EcCurve C) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t x[8], y[8];
  ec_load_affine(x, y, a_, acap, i);
  if (!aff_is_inf(x)) fp_neg(y, y, C.F);
  ec_store_affine(x, y, out, ocap, i);
}

// jac[i] = sum_j ints[j] * bases[j][i]  (t <= 8 columns, small signed integers; PGroup.expProd of
// elgamal/DistrElGamalSessionBasic.java:502)
struct EcCols {
  const uint32_t* d[8];
  size_t cap[8];
  long long k[8];
  int t;
};
VMX_EC_KERNEL k_ec_cols(const __grid_constant__ EcCols A, size_t n, uint32_t* __restrict__ jac, size_t jcap,
                        const __grid_constant__ EcCurve C) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Jac R;
  jac_set_inf(R, C);
  for (int j = 0; j < A.t; j++) {
    long long k = A.k[j];
    if (k == 0) continue;
    uint32_t x[8], y[8];
    ec_load_affine(x, y, A.d[j], A.cap[j], i);
    if (aff_is_inf(x)) continue;
    if (k < 0) { fp_neg(y, y, C.F); k = -k; }
    const unsigned long long m = (unsigned long long)k;
    Jac T;
    jac_set_inf(T, C);
    int top = 63;
    while (!((m >> top) & 1ull)) top--;
#pragma unroll 1
    for (int b = top; b >= 0; b--) {
      jac_dbl<SOL>(T, C);
      if ((m >> b) & 1ull) jac_madd<SOL>(T, x, y, C);
    }
    jac_add<SOL>(R, T, C);
  }
  ec_store_jac(R, jac, jcap, i);
}

// ------------------------------------------------------------------ fixed base
// table entry (k, d), d >= 1: the affine point d * 2^(w k) * B at index (k << w) + d.
VMX_EC_KERNEL k_ec_exp_fixed(const uint32_t* __restrict__ table, size_t tcap, int w, int nwin,
                             const uint32_t* __restrict__ e_, size_t ecap, size_t n, uint32_t* __restrict__ jac,
                             size_t jcap, const __grid_constant__ EcCurve C) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Jac P;
  jac_set_inf(P, C);
  uint32_t x[8], y[8], nx[8], ny[8];
  uint32_t d = window_bits<8>(e_, ecap, i, 0, w);
  if (d) ec_load_affine(nx, ny, table, tcap, d);
  for (int k = 0; k < nwin; k++) {
    const uint32_t dc = d;
    fp_copy(x, nx); fp_copy(y, ny);
    if (k + 1 < nwin) {  // request the next entry before the addition (gather from L2 / HBM)
      d = window_bits<8>(e_, ecap, i, (k + 1) * w, w);
      if (d) ec_load_affine(nx, ny, table, tcap, ((size_t)(k + 1) << w) + d);
    }
    if (dc && !aff_is_inf(x)) jac_madd<SOL>(P, x, y, C);
  }
  ec_store_jac(P, jac, jcap, i);
}

// Doubling chain Q[m] = 2^m * B, m < len (one thread; once per base).  Jacobian out.
VMX_EC_KERNEL k_ec_dbl_chain(const uint32_t* __restrict__ base, size_t bcap, uint32_t* __restrict__ jac, size_t jcap,
                             int len, const __grid_constant__ EcCurve C) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  uint32_t x[8], y[8];
  Jac P;
  ec_load_affine(x, y, base, bcap, 0);
  if (aff_is_inf(x)) jac_set_inf(P, C); else jac_from_affine(P, x, y, C);
  for (int m = 0; m < len; m++) {
    ec_store_jac(P, jac, jcap, m);
    jac_dbl<SOL>(P, C);
  }
}

// Table level j: item (k, r), r < 2^j:  T[k][2^j + r] = Q[w k + j] + T[k][r]  (T[k][0] = unit).
// Jacobian out at index item; k_ec_finish scatters to the table.
VMX_EC_KERNEL k_ec_table_level(const uint32_t* __restrict__ table, size_t tcap, int w, int nwin, int j, int qlen,
                               const uint32_t* __restrict__ Q, size_t qcap, uint32_t* __restrict__ jac, size_t jcap,
                               const __grid_constant__ EcCurve C) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t per = (size_t)1 << j;
  if (t >= per * nwin) return;
  const int k = (int)(t >> j);
  const size_t r = t & (per - 1);
  const int m = w * k + j;
  Jac P;
  jac_set_inf(P, C);
  if (m < qlen) {
    uint32_t x[8], y[8];
    ec_load_affine(x, y, Q, qcap, m);
    if (!aff_is_inf(x)) jac_from_affine(P, x, y, C);
    if (r) {
      ec_load_affine(x, y, table, tcap, ((size_t)k << w) + r);
      if (!aff_is_inf(x)) jac_madd<SOL>(P, x, y, C);
    }
  }
  ec_store_jac(P, jac, jcap, t);
}

// ------------------------------------------------------------------ variable base
// jac[i] = e[i or 0] * a[i]; 4-bit windows top-down; tab = 15 Jacobian multiples per point
// (entry d of point i at (d-1)*n + i of a 24-limb scratch array).
VMX_EC_KERNEL k_ec_exp_var(const uint32_t* __restrict__ a_, size_t acap, const uint32_t* __restrict__ e_, size_t ecap,
                           int escalar, int ebits, size_t n, uint32_t* __restrict__ tab, size_t tabcap,
                           uint32_t* __restrict__ jac, size_t jcap, const __grid_constant__ EcCurve C) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t ei = escalar ? 0 : i;
  uint32_t x[8], y[8];
  Jac P;
  ec_load_affine(x, y, a_, acap, i);
  if (aff_is_inf(x) || ebits == 0) {
    jac_set_inf(P, C);
    ec_store_jac(P, jac, jcap, i);
    return;
  }
  jac_from_affine(P, x, y, C);
  ec_store_jac(P, tab, tabcap, i);
#pragma unroll 1
  for (int d = 2; d < 16; d++) {
    jac_madd<SOL>(P, x, y, C);
    ec_store_jac(P, tab, tabcap, (size_t)(d - 1) * n + i);
  }
  const int nwin = (ebits + 3) / 4;
  jac_set_inf(P, C);
#pragma unroll 1
  for (int k = nwin - 1; k >= 0; k--) {
    if (k != nwin - 1) {
#pragma unroll 1
      for (int s = 0; s < 4; s++) jac_dbl<SOL>(P, C);
    }
    const uint32_t d = window_bits<8>(e_, ecap, ei, 4 * k, 4);
    if (d) {
      Jac T;
      ec_load_jac(T, tab, tabcap, (size_t)(d - 1) * n + i);
      jac_add<SOL>(P, T, C);
    }
  }
  ec_store_jac(P, jac, jcap, i);
}

// jac[i] = x * a[i] + y[i] * b[i] (x = element 0 of x_): both scalar multiples on ONE chain of doublings, the curve
// form of k_exp_var2 (the verifier's v * B_i - k_E,i * B_{i-1}, hvzk/PoSBasicTW.java:1028-1035).  Two tables of 15
// Jacobian multiples per point; the two additions of a window go through one addition site (code size).
VMX_EC_KERNEL k_ec_exp_var2(const uint32_t* __restrict__ a_, size_t acap, const uint32_t* __restrict__ x_, size_t xcap,
                            int xbits, const uint32_t* __restrict__ b_, size_t bcap, const uint32_t* __restrict__ y_,
                            size_t ycap, int ybits, size_t n, uint32_t* __restrict__ tabA, uint32_t* __restrict__ tabB,
                            size_t tabcap, uint32_t* __restrict__ jac, size_t jcap, const __grid_constant__ EcCurve C) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t x[8], y[8];
  Jac P;
  bool use[2];
#pragma unroll 1
  for (int side = 0; side < 2; side++) {
    uint32_t* tab = side ? tabB : tabA;
    ec_load_affine(x, y, side ? b_ : a_, side ? bcap : acap, i);
    use[side] = !aff_is_inf(x) && (side ? ybits : xbits) != 0;
    if (!use[side]) continue;
    jac_from_affine(P, x, y, C);
    ec_store_jac(P, tab, tabcap, i);
#pragma unroll 1
    for (int d = 2; d < 16; d++) {
      jac_madd<SOL>(P, x, y, C);
      ec_store_jac(P, tab, tabcap, (size_t)(d - 1) * n + i);
    }
  }
  const int nwx = (xbits + 3) / 4, nwy = (ybits + 3) / 4;
  const int nwin = nwx > nwy ? nwx : nwy;
  jac_set_inf(P, C);
#pragma unroll 1
  for (int k = nwin - 1; k >= 0; k--) {
    if (k != nwin - 1) {
#pragma unroll 1
      for (int s = 0; s < 4; s++) jac_dbl<SOL>(P, C);
    }
#pragma unroll 1
    for (int side = 0; side < 2; side++) {
      if (!use[side]) continue;
      const uint32_t d = side ? window_bits<8>(y_, ycap, i, 4 * k, 4) : window_bits<8>(x_, xcap, 0, 4 * k, 4);
      if (d) {
        Jac T;
        ec_load_jac(T, side ? tabB : tabA, tabcap, (size_t)(d - 1) * n + i);
        jac_add<SOL>(P, T, C);
      }
    }
  }
  ec_store_jac(P, jac, jcap, i);
}

// ------------------------------------------------------------------ segmented sums (Pippenger buckets, prod)
// out[c] = sum_{k < len} V[idx[start + k]]; V affine (vjac = 0) or Jacobian (vjac = 1); out Jacobian.
VMX_EC_KERNEL k_ec_seg_sum(const uint32_t* __restrict__ V, size_t vcap, int vjac, const uint32_t* __restrict__ idx,
                           const Chunk* __restrict__ chunks, const uint32_t* __restrict__ nchunks_dev,
                           uint32_t* __restrict__ out, size_t ocap, const __grid_constant__ EcCurve C) {
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= *nchunks_dev) return;
  const Chunk ch = chunks[c];
  Jac P;
  jac_set_inf(P, C);
  if (vjac) {
    for (uint32_t k = 0; k < ch.len; k++) {
      const size_t ik = idx ? idx[ch.start + k] : ch.start + k;
      Jac T;
      ec_load_jac(T, V, vcap, ik);
      jac_add<SOL>(P, T, C);
    }
  } else {
    uint32_t x[8], y[8], nx[8], ny[8];
    if (ch.len) ec_load_affine(nx, ny, V, vcap, idx ? idx[ch.start] : ch.start);
    for (uint32_t k = 0; k < ch.len; k++) {
      fp_copy(x, nx); fp_copy(y, ny);
      if (k + 1 < ch.len) ec_load_affine(nx, ny, V, vcap, idx ? idx[ch.start + k + 1] : ch.start + k + 1);
      if (!aff_is_inf(x)) jac_madd<SOL>(P, x, y, C);
    }
  }
  ec_store_jac(P, out, ocap, c);
}

// Y[g] = sum_{v=1}^{15} v * X[15 g + v - 1]  (running-sum trick); one thread per group; Jacobian in/out.
VMX_EC_KERNEL k_ec_weighted_small(const uint32_t* __restrict__ X, size_t xcap, size_t ngroups, uint32_t* __restrict__ Y,
                                  size_t ycap, const __grid_constant__ EcCurve C) {
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ngroups) return;
  const size_t x0 = g * kSubVals;
  // run += X_v; tot += run, written as ONE addition site "A += Q" with the roles swapped after every step
  Jac A, B, Q;  // (A, B) = (run, tot) before even steps, (tot, run) before odd steps
  jac_set_inf(A, C);
  jac_set_inf(B, C);
#pragma unroll 1
  for (int it = 0; it < 2 * kSubVals; it++) {
    if ((it & 1) == 0) ec_load_jac(Q, X, xcap, x0 + (kSubVals - it / 2) - 1); else Q = B;
    jac_add<SOL>(A, Q, C);
    Q = A; A = B; B = Q;
  }
  ec_store_jac(B, Y, ycap, g);
}

// One thread per column col < ncols: out[col] = sum_m 16^m * Y[col * Mcount + m] by Horner from the top.
VMX_EC_KERNEL k_ec_horner(const uint32_t* __restrict__ Y, size_t ycap, int Mcount, int ncols, uint32_t* __restrict__ out,
                          size_t ocap, const __grid_constant__ EcCurve C) {
  const int col = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (col >= ncols) return;
  Jac P, T;
  jac_set_inf(P, C);
#pragma unroll 1
  for (int m = Mcount - 1; m >= 0; m--) {
    if (m != Mcount - 1) {
#pragma unroll 1
      for (int s = 0; s < 4; s++) jac_dbl<SOL>(P, C);
    }
    ec_load_jac(T, Y, ycap, (size_t)col * Mcount + m);
    jac_add<SOL>(P, T, C);
  }
  ec_store_jac(P, out, ocap, col);
}

// ------------------------------------------------------------------ random elements (generators)
// Candidate j: x = (the j-th `width`-byte big-endian integer masked to `bitlen` bits) mod p, given
// already reduced in xs (canonical); ok[j] = 1 and cand[j] = (x, min(y, p - y)) if x^3 + a x + b is a
// square (p = 3 mod 4: y = rhs^((p+1)/4)), else ok[j] = 0.
VMX_EC_KERNEL k_ec_candidates(const uint32_t* __restrict__ xs, size_t xcap, size_t m, uint32_t* __restrict__ cand,
                              size_t ccap, uint32_t* __restrict__ ok, const __grid_constant__ EcCurve C) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const Fp256& F = C.F;
  uint32_t x[8], rhs[8], y[8], t[8];
  fp_load(x, xs, xcap, j);
  fp_mul<SOL>(x, x, C.r2, F);
  fp_sqr<SOL>(rhs, x, F);
  fp_add(rhs, rhs, C.a, F);
  fp_mul<SOL>(rhs, rhs, x, F);
  fp_add(rhs, rhs, C.b, F);
  fp_pow<SOL>(y, rhs, C.sqe, C.one, F);
  fp_sqr<SOL>(t, y, F);
  if (!fp_eq(t, rhs)) { ok[j] = 0; ec_store_affine_inf(cand, ccap, j); return; }
  // the smaller root as an integer: compare canonical forms
  uint32_t one[8], yc[8], ync[8], yn[8];
#pragma unroll
  for (int k = 0; k < 8; k++) one[k] = k == 0 ? 1u : 0u;
  fp_neg(yn, y, F);
  fp_mul<SOL>(yc, y, one, F);
  fp_mul<SOL>(ync, yn, one, F);
  if (fp_lt(ync, yc)) fp_copy(y, yn);
  ok[j] = 1;
  ec_store_affine(x, y, cand, ccap, j);
}

// dst[pos[j]] = cand[j] for accepted candidates with pos[j] + have < want (pos = exclusive scan of ok)
__global__ void k_ec_compact(const uint4* __restrict__ cand, size_t ccap, const uint32_t* __restrict__ ok,
                             const uint32_t* __restrict__ pos, size_t m, size_t have, size_t want,
                             uint4* __restrict__ out, size_t ocap, uint32_t* __restrict__ last_used) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m || !ok[j]) return;
  const size_t d = have + pos[j];
  if (d >= want) return;
  for (int g = 0; g < 4; g++) out[(size_t)g * ocap + d] = cand[(size_t)g * ccap + j];
  if (d + 1 == want) *last_used = (uint32_t)j + 1;
}

}  // namespace vmx
