// kernels_elem.cuh -- element-wise (thread-per-element) kernels: codec, mul, exponentiations.
//
// Every kernel keeps one residue per thread in registers (see mont.cuh) and streams the other
// operand.  Grids are sized in 128-thread blocks; with 254 registers (N=96) two blocks are
// resident per SM (8 warps, 2 per scheduler), which saturates the half-rate IMAD.WIDE pipe.
#pragma once
#include "layout.cuh"

namespace vmx {

constexpr int kThreads = 128;
template <int N> struct Occ { static constexpr int kMinBlocks = (N > 64) ? 2 : 3; };
#define VMX_KERNEL(N) __global__ void __launch_bounds__(kThreads, Occ<N>::kMinBlocks)

// error flag bits
enum { kErrRange = 1, kErrZero = 2, kErrPad = 4, kErrMember = 8 };

// ------------------------------------------------------------------ byte codec
// raw: n elements of `eb` big-endian bytes each (the fixed-width two's-complement leaf payload
// of the byte tree; values are non-negative).  Bytes beyond 4N must be zero.
// mode 0: group element -> range check 0 < x < mod, to Montgomery form (x * R).
// mode 1: ring element  -> range check 0 <= x < mod, stays canonical.
template <int N>
VMX_DEV uint32_t be_word(const uint8_t* src, int eb, int j) {  // little-endian word j of a big-endian string
  uint32_t v = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int b = 4 * j + k;
    if (b < eb) v |= (uint32_t)src[eb - 1 - b] << (8 * k);
  }
  return v;
}

// Both codec kernels move their block's records between HBM and shared memory as one contiguous run of 16-byte
// vectors (a block of kCodecThreads elements starts at a multiple of 16 bytes whatever the record length, since
// kCodecThreads is a multiple of 16), and the threads pick their bytes out of shared memory.  Reading a 390-byte
// record byte by byte straight from HBM (round 1) touched 13 sectors per warp instruction and ran at 3.8 GB/s:
// 100 ms per 10^6 elements on import and again on export, ~1.7 s of an end-to-end step at N = 10^6.
constexpr int kCodecThreads = 64;

#ifndef VMX_HOST_EMUL
// The `bytes` bytes at g -> shared memory, as aligned 16-byte vectors whatever the alignment of g (PRG output
// starts anywhere in its first 32-byte block): returns where g[0] sits in `stage`.  The vectors cover
// [g - mis, g + bytes) rounded up to 16; the last one stays inside the 16-byte granule that holds g[bytes - 1].
// Needs (bytes + 31) bytes of shared memory (codec_smem in vmx.cu); ends with a block barrier.
VMX_DEV const uint8_t* stage_in(uint4* stage, const uint8_t* g, size_t bytes) {
  const size_t mis = reinterpret_cast<uintptr_t>(g) & 15;
  const uint4* g16 = reinterpret_cast<const uint4*>(g - mis);
  const size_t vecs = (mis + bytes + 15) >> 4;
  for (size_t v = threadIdx.x; v < vecs; v += blockDim.x) stage[v] = g16[v];
  __syncthreads();
  return reinterpret_cast<const uint8_t*>(stage) + mis;
}
#endif

template <int N>
VMX_KERNEL(N) k_from_bytes(const uint8_t* __restrict__ raw, size_t n, int eb, int hdr, int mode,
                           uint32_t* __restrict__ out, size_t cap, const uint32_t* __restrict__ r2,
                           int* __restrict__ err, const __grid_constant__ MontParams<N> M) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t rec = (size_t)(eb + hdr);
#ifndef VMX_HOST_EMUL
  VMX_DYN_SMEM(uint4, stage);
  const size_t e0 = (size_t)blockIdx.x * blockDim.x;
  const uint8_t* sbase = stage_in(stage, raw + e0 * rec, (n - e0 < blockDim.x ? n - e0 : blockDim.x) * rec);
  if (i >= n) return;
  // hdr = 5: the elements are the leaves of a byte tree, 0x01 || be32(eb) || payload each
  const uint8_t* src = sbase + threadIdx.x * rec + hdr;
#else
  if (i >= n) return;
  const uint8_t* src = raw + i * rec + hdr;
#endif
  uint32_t a[N];
  int bad = 0;
  if (hdr) {
    const uint8_t* h = src - 5;
    if (h[0] != 1 || h[1] != (uint8_t)(eb >> 24) || h[2] != (uint8_t)(eb >> 16) || h[3] != (uint8_t)(eb >> 8) ||
        h[4] != (uint8_t)eb) bad |= kErrPad;
  }
  for (int k = 0; k < eb - 4 * N; k++) if (src[k] != 0) bad |= kErrPad;
#pragma unroll
  for (int j = 0; j < N; j++) a[j] = be_word<N>(src, eb, j);
  // x < mod ?
  uint32_t d, brw;
  sub_cc(d, a[0], M.n[0]);
#pragma unroll
  for (int j = 1; j < N; j++) subc_cc(d, a[j], M.n[j]);
  subc(brw, 0, 0);
  if (brw == 0) bad |= kErrRange;
  if (mode == 0) {
    uint32_t nz = 0;
#pragma unroll
    for (int j = 0; j < N; j++) nz |= a[j];
    if (nz == 0) bad |= kErrZero;
    if (!bad) mont_mul<N>(a, GlobalLoader(r2, 4, 0), M);
  }
  if (bad) atomicOr(err, bad);
  store_elem<N>(a, out, cap, i);
}

// mode 0: group element (Montgomery -> canonical first); mode 1: ring element.
template <int N>
VMX_KERNEL(N) k_to_bytes(const uint32_t* __restrict__ in, size_t cap, size_t n, int eb, int hdr, int mode,
                         uint8_t* __restrict__ raw, const __grid_constant__ MontParams<N> M) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t rec = (size_t)(eb + hdr);
#ifndef VMX_HOST_EMUL
  VMX_DYN_SMEM(uint4, stage);
  uint8_t* dst = reinterpret_cast<uint8_t*>(stage) + threadIdx.x * rec + hdr;
  const bool live = i < n;
#else
  if (i >= n) return;
  uint8_t* dst = raw + i * rec + hdr;
  const bool live = true;
#endif
  if (live) {
    uint32_t a[N];
    load_elem<N>(a, in, cap, i);
    if (mode == 0) mont_mul<N>(a, OneLoader{}, M);
    if (hdr) {  // byte-tree leaf header
      dst[-5] = 1; dst[-4] = (uint8_t)(eb >> 24); dst[-3] = (uint8_t)(eb >> 16); dst[-2] = (uint8_t)(eb >> 8); dst[-1] = (uint8_t)eb;
    }
    for (int k = 0; k < eb - 4 * N; k++) dst[k] = 0;
#pragma unroll
    for (int j = 0; j < N; j++) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int b = 4 * j + k;
        if (b < eb) dst[eb - 1 - b] = (uint8_t)(a[j] >> (8 * k));
      }
    }
  }
#ifndef VMX_HOST_EMUL
  __syncthreads();
  {
    const size_t e0 = (size_t)blockIdx.x * blockDim.x;
    if (e0 < n) {
      const size_t cnt = n - e0 < blockDim.x ? n - e0 : blockDim.x;
      uint8_t* g = raw + e0 * rec;
      const size_t bytes = cnt * rec, vecs = bytes >> 4;
      for (size_t v = threadIdx.x; v < vecs; v += blockDim.x) reinterpret_cast<uint4*>(g)[v] = stage[v];
      for (size_t b = (vecs << 4) + threadIdx.x; b < bytes; b += blockDim.x) g[b] = reinterpret_cast<const uint8_t*>(stage)[b];
    }
  }
#endif
}

// raw unsigned integers of `width` bytes (big-endian), masked to `bitlen` bits (0 = all),
// reduced mod q.  The integer is read as chunks of 32N bits, x = sum_k c_k 2^(32N k), and folded
// from the top: acc <- acc * 2^(32N) + c_k (mod q).  Up to 4 chunks (an n_e + n_v + n_r = 612-bit
// integer is 3 chunks over the 256-bit order of a curve group, hvzk/PoSBasicTW.java:470-474).
template <int N>
VMX_KERNEL(N) k_ring_from_raw(const uint8_t* __restrict__ raw, size_t n, int width, int bitlen,
                              uint32_t* __restrict__ out, size_t cap, const uint32_t* __restrict__ r2,
                              int need_reduce, const __grid_constant__ MontParams<N> M) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
#ifndef VMX_HOST_EMUL
  VMX_DYN_SMEM(uint4, stage);
  const size_t e0 = (size_t)blockIdx.x * blockDim.x;
  const uint8_t* sbase = stage_in(stage, raw + e0 * (size_t)width, (n - e0 < blockDim.x ? n - e0 : blockDim.x) * (size_t)width);
  if (i >= n) return;
  const uint8_t* src = sbase + threadIdx.x * (size_t)width;
#else
  if (i >= n) return;
  const uint8_t* src = raw + i * (size_t)width;
#endif
  const int totalbits = bitlen ? bitlen : 8 * width;
  auto word = [&](int j) -> uint32_t {  // little-endian word j of the masked integer
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int byte = 4 * j + k;  // little-endian byte index
      if (byte < width) {
        uint32_t b = src[width - 1 - byte];
        const int hb = totalbits - 8 * byte;  // valid bits in this byte
        if (hb <= 0) b = 0; else if (hb < 8) b &= (1u << hb) - 1u;
        v |= b << (8 * k);
      }
    }
    return v;
  };
  uint32_t acc[N];
  if (!need_reduce) {
#pragma unroll
    for (int j = 0; j < N; j++) acc[j] = word(j);
    store_elem<N>(acc, out, cap, i);
    return;
  }
  const GlobalLoader R2(r2, 4, 0);
  const int nchunks = (8 * width + 32 * N - 1) / (32 * N);
  for (int c = nchunks - 1; c >= 0; c--) {
    uint32_t ch[N];
#pragma unroll
    for (int j = 0; j < N; j++) ch[j] = word(c * N + j);
    mont_mul<N>(ch, R2, M);          // c_k * R
    mont_mul<N>(ch, OneLoader{}, M); // c_k mod q
    if (c == nchunks - 1) {
#pragma unroll
      for (int j = 0; j < N; j++) acc[j] = ch[j];
      continue;
    }
    mont_mul<N>(acc, R2, M);         // acc * 2^(32N) mod q
    uint32_t cy;
    add_cc(acc[0], acc[0], ch[0]);
#pragma unroll
    for (int j = 1; j < N; j++) addc_cc(acc[j], acc[j], ch[j]);
    addc(cy, 0, 0);
    uint32_t d[N], brw;
    sub_cc(d[0], acc[0], M.n[0]);
#pragma unroll
    for (int j = 1; j < N; j++) subc_cc(d[j], acc[j], M.n[j]);
    subc(brw, cy, 0);
    const bool keep = (brw != 0);
#pragma unroll
    for (int j = 0; j < N; j++) acc[j] = keep ? acc[j] : d[j];
  }
  store_elem<N>(acc, out, cap, i);
}

// ------------------------------------------------------------------ out[i] = a[i] * b[i]
template <int N>
VMX_KERNEL(N) k_mul(const uint32_t* __restrict__ a_, size_t acap, const uint32_t* __restrict__ b_, size_t bcap,
                    uint32_t* __restrict__ out, size_t ocap, size_t n, const __grid_constant__ MontParams<N> M) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t a[N];
  load_elem<N>(a, a_, acap, i);
  mont_mul<N>(a, GlobalLoader(b_, bcap, i), M);
  store_elem<N>(a, out, ocap, i);
}

// ------------------------------------------------------------------ batched inversion (Montgomery's trick)
// One level of a product tree with interleaved chunks: thread t of T owns the elements t, t + T, t + 2T, ...
// (consecutive threads touch consecutive elements: every access is a coalesced 16-byte vector per lane).
//   up:   pre[i] = a[t] * a[t + T] * ... * a[i] for the elements after the first, part[t] = the whole product;
//   down: given inv[t] = part[t]^-1 (consumed as scratch), out[i] = a[i]^-1 -- `out` is the `pre` of the up pass.
// 3 multiplications per element and one inversion at the root, against ~1.2 * |p| for a Fermat inversion of
// every element.  Only one residue lives in registers at a time (two would not fit at 96 limbs): the running
// inverse stays in its slot of `inv` and is streamed as the second operand.
template <int N>
VMX_KERNEL(N) k_inv_up(const uint32_t* __restrict__ a_, size_t acap, size_t n, size_t T, uint32_t* __restrict__ pre,
                       size_t pcap, uint32_t* __restrict__ part, size_t partcap, const __grid_constant__ MontParams<N> M) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  uint32_t x[N];
  load_elem<N>(x, a_, acap, t);
  for (size_t i = t + T; i < n; i += T) {
    mont_mul<N>(x, GlobalLoader(a_, acap, i), M);
    store_elem<N>(x, pre, pcap, i);
  }
  store_elem<N>(x, part, partcap, t);
}

template <int N>
VMX_KERNEL(N) k_inv_down(const uint32_t* __restrict__ a_, size_t acap, size_t n, size_t T, uint32_t* inv, size_t icap,
                         uint32_t* out, size_t ocap, const __grid_constant__ MontParams<N> M) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  uint32_t x[N];
  size_t i = t + ((n - 1 - t) / T) * T;  // last element of this thread
  for (; i > t; i -= T) {
    // a[i]^-1 = (running inverse of the prefix up to i) * (prefix up to i - T)
    if (i - T == t) load_elem<N>(x, a_, acap, t); else load_elem<N>(x, out, ocap, i - T);
    mont_mul<N>(x, GlobalLoader(inv, icap, t), M);
    store_elem<N>(x, out, ocap, i);
    // running inverse <- running inverse * a[i]
    load_elem<N>(x, a_, acap, i);
    mont_mul<N>(x, GlobalLoader(inv, icap, t), M);
    store_elem<N>(x, inv, icap, t);
  }
  load_elem<N>(x, inv, icap, t);
  store_elem<N>(x, out, ocap, t);
}

// out[i] = a[i] * b[i]^iters  (benchmark of the raw modmul path)
template <int N>
VMX_KERNEL(N) k_mul_iter(const uint32_t* __restrict__ a_, const uint32_t* __restrict__ b_, uint32_t* __restrict__ out,
                         size_t cap, size_t n, int iters, const __grid_constant__ MontParams<N> M) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t a[N];
  load_elem<N>(a, a_, cap, i);
  const GlobalLoader B(b_, cap, i);
  for (int it = 0; it < iters; it++) mont_mul<N>(a, B, M);
  store_elem<N>(a, out, cap, i);
}

// out[i] = a[i]^(2^iters): the dedicated squaring (mont_sqr) on its own -- self test against k_mul and benchmark
template <int N>
VMX_KERNEL(N) k_sqr_iter(const uint32_t* __restrict__ a_, size_t acap, uint32_t* __restrict__ out, size_t ocap, size_t n,
                         int iters, const __grid_constant__ MontParams<N> M) {
  VMX_DYN_SMEM(uint2, smem);
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t a[N];
  load_elem<N>(a, a_, acap, i);
  for (int it = 0; it < iters; it++) mont_sqr<N>(a, smem + threadIdx.x, blockDim.x, M);
  store_elem<N>(a, out, ocap, i);
}

// ------------------------------------------------------------------ fixed-base exponentiation
// table: nwin * 2^w entries, entry (k, d) = base^(d * 2^(w*k)) in Montgomery form (d = 0 -> one), stored
// ENTRY-MAJOR: the N limbs of entry e are the N consecutive words at table + e * N (a limb-major "array" of
// capacity 1, which is how load_elem / GlobalLoader are pointed at it).  A thread walks its entry 8 bytes per CIOS
// trip, so the 384 bytes of a 3072-bit entry are three 128-byte lines read once each (L1 keeps the line for the
// 16 trips that use it); in the limb-major layout of the arrays the same entry is 24 half-used 32-byte sectors in
// 24 planes (2.07x the algorithmic DRAM bytes, profiles/r02_ncu_exp_fixed.txt, and 24 pages per entry).
// Work item = (element i, window range part): part p of `parts` multiplies windows
// [p*nwin/parts, (p+1)*nwin/parts) and writes to out element p*n + i (parts > 1 -> combined later).
template <int N>
VMX_KERNEL(N) k_exp_fixed(const uint32_t* __restrict__ table, size_t tcap, int w, int nwin,
                          const uint32_t* __restrict__ e_, size_t ecap, size_t n, int parts,
                          uint32_t* __restrict__ out, size_t ocap, const __grid_constant__ MontParams<N> M) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * (size_t)parts) return;
  const size_t i = t % n;
  const int p = (int)(t / n);
  const int k0 = (int)((long long)nwin * p / parts), k1 = (int)((long long)nwin * (p + 1) / parts);
  uint32_t a[N];
  {
    const uint32_t d = window_bits<N>(e_, ecap, i, k0 * w, w);
    load_elem<N>(a, table + (((size_t)k0 << w) + d) * N, 1, 0);
  }
  uint32_t dn = (k0 + 1 < k1) ? window_bits<N>(e_, ecap, i, (k0 + 1) * w, w) : 0;
  for (int k = k0 + 1; k < k1; k++) {
    const GlobalLoader B(table + (((size_t)k << w) + dn) * N, 1, 0);
    if (k + 1 < k1) dn = window_bits<N>(e_, ecap, i, (k + 1) * w, w);
    mont_mul<N>(a, B, M);
  }
  store_elem<N>(a, out, ocap, t);
}

// out[i] = prod_{p < parts} in[p*n + i]
template <int N>
VMX_KERNEL(N) k_combine_parts(const uint32_t* __restrict__ in, size_t icap, size_t n, int parts,
                              uint32_t* __restrict__ out, size_t ocap, const __grid_constant__ MontParams<N> M) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t a[N];
  load_elem<N>(a, in, icap, i);
  for (int p = 1; p < parts; p++) mont_mul<N>(a, GlobalLoader(in, icap, (size_t)p * n + i), M);
  store_elem<N>(a, out, ocap, i);
}

// Squaring chain Q[m] = base^(2^m), m = 0..len-1 (single thread; latency-bound, run once per base).
template <int N>
VMX_KERNEL(N) k_sqr_chain(const uint32_t* __restrict__ base, size_t bcap, size_t bidx, uint32_t* __restrict__ Q,
                          size_t qcap, int len, const __grid_constant__ MontParams<N> M) {
  VMX_DYN_SMEM(uint2, smem);
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  uint32_t a[N];
  load_elem<N>(a, base, bcap, bidx);
  store_elem<N>(a, Q, qcap, 0);
  for (int m = 1; m < len; m++) {
    mont_sqr<N>(a, smem, 1, M);
    store_elem<N>(a, Q, qcap, m);
  }
}

// Table level j: T[k][2^j + r] = Q[w*k + j] * T[k][r], r in [0, 2^j); T[k][0] = one.
// Work item t -> (k, r).
template <int N>
VMX_KERNEL(N) k_table_level(uint32_t* __restrict__ table, size_t tcap, int w, int nwin, int j, int qlen,
                            const uint32_t* __restrict__ Q, size_t qcap, const uint32_t* __restrict__ one,
                            const __grid_constant__ MontParams<N> M) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t per = (size_t)1 << j;
  if (t >= per * nwin) return;
  const int k = (int)(t >> j);
  const size_t r = t & (per - 1);
  const int m = w * k + j;
  uint32_t a[N];
  const size_t dst = ((size_t)k << w) + per + r;
  if (m >= qlen) {  // beyond the exponent range: never addressed, keep defined
    load_elem<N>(a, one, 4, 1);
    store_elem<N>(a, table + dst * N, 1, 0);
    return;
  }
  if (r == 0) {
    load_elem<N>(a, Q, qcap, m);
    if (j == 0) {  // also write entry 0 = one
      uint32_t o[N];
      load_elem<N>(o, one, 4, 1);
      store_elem<N>(o, table + ((size_t)k << w) * N, 1, 0);
    }
  } else {
    load_elem<N>(a, table + (((size_t)k << w) + r) * N, 1, 0);
    mont_mul<N>(a, GlobalLoader(Q, qcap, m), M);
  }
  store_elem<N>(a, table + dst * N, 1, 0);
}

// ------------------------------------------------------------------ variable-base exponentiation
// out[i] = a[i]^{e[i]} (escalar = 0) or a[i]^{e[0]} (escalar = 1); fixed w-bit windows, top-down.
// tab: scratch for per-thread tables, (2^w) * n entries, entry-major as the fixed-base tables: entry d of element i
// is the N words at tab + ((i << w) + d) * N.  Written once (two 16-byte stores fill a sector), read at random d.
template <int N>
VMX_KERNEL(N) k_exp_var(const uint32_t* __restrict__ a_, size_t acap, const uint32_t* __restrict__ e_, size_t ecap,
                        int escalar, int ebits, int w, size_t n, uint32_t* __restrict__ tab, size_t tabcap,
                        const uint32_t* __restrict__ one, uint32_t* __restrict__ out, size_t ocap,
                        const __grid_constant__ MontParams<N> M) {
  VMX_DYN_SMEM(uint2, smem);
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint2* sc = smem + threadIdx.x;
  const unsigned ss = blockDim.x;
  const size_t ei = escalar ? 0 : i;
  uint32_t a[N];
  // table: tab[0] = 1, tab[1] = base, tab[d] = tab[d-1] * base
  load_elem<N>(a, one, 4, 1);
  store_elem<N>(a, tab + ((size_t)i << w) * N, 1, 0);
  load_elem<N>(a, a_, acap, i);
  store_elem<N>(a, tab + (((size_t)i << w) + 1) * N, 1, 0);
  const GlobalLoader Bse(a_, acap, i);
  for (int d = 2; d < (1 << w); d++) {
    mont_mul<N>(a, Bse, M);
    store_elem<N>(a, tab + (((size_t)i << w) + d) * N, 1, 0);
  }
  const int nwin = (ebits + w - 1) / w;
  {
    const uint32_t d = window_bits<N>(e_, ecap, ei, (nwin - 1) * w, w);
    load_elem<N>(a, tab + (((size_t)i << w) + d) * N, 1, 0);
  }
  for (int k = nwin - 2; k >= 0; k--) {
    const uint32_t d = window_bits<N>(e_, ecap, ei, k * w, w);
    for (int s = 0; s < w; s++) mont_sqr<N>(a, sc, ss, M);
    mont_mul<N>(a, GlobalLoader(tab + (((size_t)i << w) + d) * N, 1, 0), M);
  }
  store_elem<N>(a, out, ocap, i);
}

// out[i] = a[i]^{x} * b[i]^{y[i]}: simultaneous exponentiation, one chain of squarings for both bases (the
// verifier's B_i^v * B_{i-1}^{-k_E,i}, hvzk/PoSBasicTW.java:1028-1035, costs max(|v|, |k_E|) squarings instead of
// |v| + |k_E|).  x is element 0 of x_ (one exponent for all), y per element; both use w-bit windows at the same
// positions.  tabA / tabB: 2^w entries per element each, entry-major (entry d of element i at ((i << w) + d) * N).
template <int N>
VMX_KERNEL(N) k_exp_var2(const uint32_t* __restrict__ a_, size_t acap, const uint32_t* __restrict__ x_, size_t xcap,
                         int xbits, const uint32_t* __restrict__ b_, size_t bcap, const uint32_t* __restrict__ y_,
                         size_t ycap, int ybits, int w, size_t n, uint32_t* __restrict__ tabA,
                         uint32_t* __restrict__ tabB, size_t tabcap, const uint32_t* __restrict__ one,
                         uint32_t* __restrict__ out, size_t ocap, const __grid_constant__ MontParams<N> M) {
  VMX_DYN_SMEM(uint2, smem);
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint2* sc = smem + threadIdx.x;
  const unsigned ss = blockDim.x;
  uint32_t a[N];
  for (int side = 0; side < 2; side++) {
    const uint32_t* src = side ? b_ : a_;
    const size_t scap = side ? bcap : acap;
    uint32_t* tab = side ? tabB : tabA;
    load_elem<N>(a, one, 4, 1);
    store_elem<N>(a, tab + ((size_t)i << w) * N, 1, 0);
    load_elem<N>(a, src, scap, i);
    store_elem<N>(a, tab + (((size_t)i << w) + 1) * N, 1, 0);
    const GlobalLoader Bse(src, scap, i);
    for (int d = 2; d < (1 << w); d++) {
      mont_mul<N>(a, Bse, M);
      store_elem<N>(a, tab + (((size_t)i << w) + d) * N, 1, 0);
    }
  }
  const int nwy = (ybits + w - 1) / w, nwx = (xbits + w - 1) / w;
  const int nwin = nwy > nwx ? nwy : nwx;
  // ONE multiplication site and one squaring site in the loop (both bases go through the same inlined
  // mont_mul, selected by pointer): the loop body is ~40 KB of SASS with 32-word squaring blocks and stays in
  // the instruction cache; with a site per base and 16-word blocks it was 73 KB and ran 20 % slower than the
  // plain multiplication it replaced (gpurun_out/s1_launches_100k.csv).
  {
    const uint32_t dy = window_bits<N>(y_, ycap, i, (nwin - 1) * w, w);
    load_elem<N>(a, tabB + (((size_t)i << w) + dy) * N, 1, 0);
  }
  for (int k = nwin - 1; k >= 0; k--) {
    if (k < nwin - 1)
      for (int s = 0; s < w; s++) mont_sqr<N>(a, sc, ss, M);
#pragma unroll 1
    for (int side = (k < nwin - 1 ? 0 : 1); side < 2; side++) {
      uint32_t d;
      const uint32_t* tab;
      if (side == 0) {
        d = window_bits<N>(y_, ycap, i, k * w, w);
        tab = tabB;
      } else {
        if (k >= nwx) continue;  // uniform over the grid: x is one exponent
        d = window_bits<N>(x_, xcap, 0, k * w, w);
        if (d == 0) continue;
        tab = tabA;
      }
      mont_mul<N>(a, GlobalLoader(tab + (((size_t)i << w) + d) * N, 1, 0), M);
    }
  }
  store_elem<N>(a, out, ocap, i);
}

// ------------------------------------------------------------------ data movement (uint4 granularity)
// out[dst(i)] = in[src(i)] for plane-wise copies; one thread per (plane, element).
// bcast != 0: every destination receives source element `bidx` (fill).
__global__ void k_gather(const uint4* __restrict__ in, size_t icap, uint4* __restrict__ out, size_t ocap, size_t n,
                         int planes, const uint32_t* __restrict__ src_idx, const uint32_t* __restrict__ dst_idx,
                         long long src_off, long long dst_off, int bcast, size_t bidx) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * (size_t)planes) return;
  const size_t i = t % n, g = t / n;
  const size_t s = bcast ? bidx : (src_idx ? src_idx[i] : (size_t)((long long)i + src_off));
  const size_t d = dst_idx ? dst_idx[i] : (size_t)((long long)i + dst_off);
  out[g * ocap + d] = in[g * icap + s];
}

// Element-major packing for the exchange between GPUs (permute across shards): row r of `rows` holds the
// `planes` uint4 groups of element idx[r] (idx == null: element r) back to back.  Writes are
// coalesced (consecutive threads -> consecutive 16-byte groups of one row); reads are 16-byte
// gathers.  k_unpack_rows is the inverse scatter into a limb-major array.
__global__ void k_pack_rows(const uint4* __restrict__ in, size_t icap, const uint32_t* __restrict__ idx, size_t count,
                            int planes, uint4* __restrict__ rows) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count * (size_t)planes) return;
  const size_t r = t / planes, g = t % planes;
  rows[t] = in[g * icap + (idx ? idx[r] : r)];
}
__global__ void k_unpack_rows(const uint4* __restrict__ rows, const uint32_t* __restrict__ idx, size_t count, int planes,
                              uint4* __restrict__ out, size_t ocap) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count * (size_t)planes) return;
  const size_t r = t / planes, g = t % planes;
  out[g * ocap + (idx ? idx[r] : r)] = rows[t];
}

// *diff |= any word differs
__global__ void k_equal(const uint4* __restrict__ a, size_t acap, const uint4* __restrict__ b, size_t bcap, size_t n,
                        int planes, int* __restrict__ diff) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * (size_t)planes) return;
  const size_t i = t % n, g = t / n;
  const uint4 x = a[g * acap + i], y = b[g * bcap + i];
  if (x.x != y.x || x.y != y.y || x.z != y.z || x.w != y.w) atomicOr(diff, 1);
}

// max bit length over the array -> atomicMax(bits)
template <int N>
__global__ void k_bitlen(const uint32_t* __restrict__ d, size_t cap, size_t n, unsigned* __restrict__ bits) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned b = 0;
  if (i < n) {
    for (int j = N - 1; j >= 0; j--) {
      const uint32_t v = d[((size_t)(j >> 2) * cap + i) * 4 + (j & 3)];
      if (v) { b = 32u * j + (32u - __clz(v)); break; }
    }
  }
#ifndef VMX_HOST_EMUL
  b = __reduce_max_sync(0xffffffffu, b);
  if ((threadIdx.x & 31) == 0 && b) atomicMax(bits, b);
#else
  if (b) atomicMax(bits, b);
#endif
}

// *diff |= 1 if any of the n elements differs from element `cidx` of c_ (e.g. the Montgomery one)
template <int N>
__global__ void k_differs_from_const(const uint32_t* __restrict__ a_, size_t acap, size_t n,
                                     const uint32_t* __restrict__ c_, size_t ccap, size_t cidx, int* __restrict__ diff) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t x = 0;
  for (int j = 0; j < N; j++)
    x |= a_[((size_t)(j >> 2) * acap + i) * 4 + (j & 3)] ^ c_[((size_t)(j >> 2) * ccap + cidx) * 4 + (j & 3)];
  if (x) atomicOr(diff, 1);
}

}  // namespace vmx
