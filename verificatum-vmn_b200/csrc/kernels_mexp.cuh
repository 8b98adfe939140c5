// kernels_mexp.cuh -- bucketed (Pippenger) multi-exponentiation  prod_i b_i^{e_i}  and the
// segmented-product machinery it shares with prod().
//
// Pipeline for window width c (a multiple of 4), W = ceil(L/c) windows, all at once:
//   1. k_digit_hist     histogram of digits: segment id = k*2^c + digit_k(e_i), digit != 0
//   2. exclusive scan   segment offsets                              (scan.cuh)
//   3. k_digit_scatter  term indices sorted by segment
//   4. segmented product rounds (k_chunk_count / k_chunk_fill / k_seg_prod): every bucket is
//      the product of its terms; a segment of length len is cut in ceil(len/K) balanced
//      chunks, one thread per chunk; rounds repeat on the partial products until one value
//      per segment is left.
//   5. bucket reduction by radix-16 sub-digits: X[k][j][v] = product of the buckets of window
//      k whose j-th hex digit is v (segmented products again, over static index lists);
//      Y[k][j] = prod_v X[k][j][v]^v  (k_weighted_small, running-product trick).
//   6. k_horner         result = prod_m Y[m]^(16^m), m = k*(c/4) + j  (4 squarings per step)
// The group is commutative and all arithmetic is exact, so the result does not depend on the
// order in which terms are multiplied: bit-exact with any CPU evaluation.
//
// Reference call sites served: hvzk/PoSBasicTW.java:408-409,481,690,1021,1063,
// hvzk/CCPoSBasicW.java:381,394,499-504, elgamal/DistrElGamalSessionBasic.java:524-526.
#pragma once
#include "kernels_elem.cuh"

namespace vmx {

constexpr int kSubDigit = 4;
constexpr int kSubVals = 15;  // non-zero values of a hex digit

// 1. histogram.  One thread per (term, window).
template <int N>
__global__ void k_digit_hist(const uint32_t* __restrict__ e_, size_t ecap, size_t n, int c, int W,
                             uint32_t* __restrict__ hist) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * (size_t)W) return;
  const size_t i = t % n;
  const int k = (int)(t / n);
  const uint32_t d = window_bits<N>(e_, ecap, i, k * c, c);
  if (d) atomicAdd(&hist[((size_t)k << c) + d], 1u);
}

// 3. scatter term indices into segment order (cursor = copy of the exclusive offsets).
template <int N>
__global__ void k_digit_scatter(const uint32_t* __restrict__ e_, size_t ecap, size_t n, int c, int W,
                                uint32_t* __restrict__ cursor, uint32_t* __restrict__ idx) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * (size_t)W) return;
  const size_t i = t % n;
  const int k = (int)(t / n);
  const uint32_t d = window_bits<N>(e_, ecap, i, k * c, c);
  if (d) {
    const uint32_t pos = atomicAdd(&cursor[((size_t)k << c) + d], 1u);
    idx[pos] = (uint32_t)i;
  }
}

// 4a. chunks per segment: max(1, ceil(len / K)); entry nseg is set to 0 so that the exclusive
// scan over nseg+1 entries leaves the total in entry nseg.  stats[0] = max segment length.
__global__ void k_chunk_count(const uint32_t* __restrict__ seg_off, size_t nseg, int K, uint32_t* __restrict__ nchunks,
                              uint32_t* __restrict__ stats) {
  const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s > nseg) return;
  if (s == nseg) { nchunks[s] = 0; return; }
  const uint32_t len = seg_off[s + 1] - seg_off[s];
  nchunks[s] = len == 0 ? 1u : (len + K - 1) / K;
  if (len > (uint32_t)K) atomicMax(&stats[0], len);
}

// 4b. balanced chunk descriptors, one thread per CHUNK (a single huge segment -- prod() -- must
// not serialise on one thread): the owning segment is found by binary search in chunk_off
// (exclusive scan of nchunks, nseg + 1 entries).
struct Chunk { uint32_t start, len; };
__global__ void k_chunk_fill(const uint32_t* __restrict__ seg_off, const uint32_t* __restrict__ chunk_off, size_t nseg,
                             Chunk* __restrict__ chunks) {
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= chunk_off[nseg]) return;
  size_t lo = 0, hi = nseg;  // largest s with chunk_off[s] <= c
  while (hi - lo > 1) {
    const size_t mid = (lo + hi) >> 1;
    if (chunk_off[mid] <= c) lo = mid; else hi = mid;
  }
  const size_t s = lo;
  const uint32_t b = seg_off[s], len = seg_off[s + 1] - b;
  const uint32_t c0 = chunk_off[s], m = chunk_off[s + 1] - c0, j = (uint32_t)c - c0;
  const uint32_t l0 = (uint32_t)((uint64_t)len * j / m), l1 = (uint32_t)((uint64_t)len * (j + 1) / m);
  chunks[c] = Chunk{b + l0, l1 - l0};
}

// 4b'. chunks ordered by length, longest first (counting sort over the lengths 0..K): the lanes of a warp of
// k_seg_prod then multiply the same number of terms and finish together -- unsorted, a warp waits for its longest
// chunk (a bucket of 25 terms is cut in 4 x 6.25, its neighbour of 24 in 3 x 8).  bins: K + 2 counters, zeroed;
// after k_chunk_len_hist + k_chunk_len_offsets, bins[l] = first position of the chunks of length l.
__global__ void k_chunk_len_hist(const Chunk* __restrict__ chunks, const uint32_t* __restrict__ nchunks_dev, int K,
                                 uint32_t* __restrict__ bins) {
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = c < *nchunks_dev;
  const uint32_t len = live ? min(chunks[c].len, (uint32_t)K) : 0;
#ifndef VMX_HOST_EMUL
  // one atomic per distinct length in the warp
  const unsigned peers = __match_any_sync(0xffffffffu, live ? len : 0xffffffffu);
  if (live && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&bins[len], (uint32_t)__popc(peers));
#else
  if (live) atomicAdd(&bins[len], 1u);
#endif
}
__global__ void k_chunk_len_offsets(int K, uint32_t* __restrict__ bins) {  // one thread: descending lengths
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  uint32_t run = 0;
  for (int l = K; l >= 0; l--) { const uint32_t cnt = bins[l]; bins[l] = run; run += cnt; }
}
__global__ void k_chunk_len_scatter(const Chunk* __restrict__ chunks, const uint32_t* __restrict__ nchunks_dev, int K,
                                    uint32_t* __restrict__ bins, uint32_t* __restrict__ order) {
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = c < *nchunks_dev;
  const uint32_t len = live ? min(chunks[c].len, (uint32_t)K) : 0;
#ifndef VMX_HOST_EMUL
  const unsigned peers = __match_any_sync(0xffffffffu, live ? len : 0xffffffffu);
  const int leader = __ffs(peers) - 1, lane = threadIdx.x & 31;
  uint32_t base = 0;
  if (live && leader == lane) base = atomicAdd(&bins[len], (uint32_t)__popc(peers));
  base = __shfl_sync(peers, base, leader);
  if (live) order[base + __popc(peers & ((1u << lane) - 1u))] = (uint32_t)c;
#else
  if (live) order[atomicAdd(&bins[len], 1u)] = (uint32_t)c;
#endif
}

// 4c. one thread per chunk: out[c] = prod_{k < len} V[idx[start + k]]   (len = 0 -> one).
// The number of chunks lives on the device (*nchunks_dev): the grid is sized from a host bound.
// order != null: thread t takes chunk order[t] (chunks of equal length side by side).
template <int N>
VMX_KERNEL(N) k_seg_prod(const uint32_t* __restrict__ V, size_t vcap, const uint32_t* __restrict__ idx,
                         const Chunk* __restrict__ chunks, const uint32_t* __restrict__ nchunks_dev,
                         const uint32_t* __restrict__ order, uint32_t* __restrict__ out, size_t ocap,
                         const uint32_t* __restrict__ one, const __grid_constant__ MontParams<N> M) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= *nchunks_dev) return;
  const size_t c = order ? order[t] : t;
  const Chunk ch = chunks[c];
  uint32_t a[N];
  if (ch.len == 0) {
    load_elem<N>(a, one, 4, 1);
  } else {
    const size_t i0 = idx ? idx[ch.start] : ch.start;
    load_elem<N>(a, V, vcap, i0);
    for (uint32_t k = 1; k < ch.len; k++) {
      const size_t ik = idx ? idx[ch.start + k] : ch.start + k;
      mont_mul<N>(a, GlobalLoader(V, vcap, ik), M);
    }
  }
  store_elem<N>(a, out, ocap, c);
}

// 5a. static index lists of the sub-digit products.  Segment id = (k*J + j)*15 + (v-1); it
// lists the 2^(c-4) buckets k*2^c + d whose hex digit j equals v.  One block per segment.
__global__ void k_subdigit_lists(int c, int J, uint32_t* __restrict__ idx) {
  const size_t seg = (size_t)blockIdx.x;
  const uint32_t v = (uint32_t)(seg % kSubVals) + 1;
  const int j = (int)((seg / kSubVals) % J);
  const size_t k = seg / ((size_t)kSubVals * J);
  const size_t cnt = (size_t)1 << (c - kSubDigit);
  const size_t base = seg * cnt;
  const int sh = kSubDigit * j;
  const size_t lowmask = ((size_t)1 << sh) - 1;
  for (size_t t = threadIdx.x; t < cnt; t += blockDim.x) {
    const size_t d = ((t >> sh) << (sh + kSubDigit)) | ((size_t)v << sh) | (t & lowmask);
    idx[base + t] = (uint32_t)((k << c) + d);
  }
}

// seg_off[s] = s * cnt  (s = 0..nseg)
__global__ void k_uniform_offsets(uint32_t* __restrict__ seg_off, size_t nseg, uint32_t cnt) {
  const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s <= nseg) seg_off[s] = (uint32_t)(s * cnt);
}

// 5b. Y[g] = prod_{v=1}^{15} X[15g + v-1]^v  (running-product trick, 28 modmuls:
// run = X_15; tot = X_15; for v = 14..1: run *= X_v; tot *= run).  One thread per group; only
// one residue fits in registers, so `run` lives in the scratch array R and `tot` in Y.
template <int N>
VMX_KERNEL(N) k_weighted_small(const uint32_t* __restrict__ X, size_t xcap, size_t ngroups,
                               uint32_t* __restrict__ Y, size_t ycap, uint32_t* __restrict__ R, size_t rcap,
                               const __grid_constant__ MontParams<N> M) {
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ngroups) return;
  const size_t x0 = g * kSubVals;
  uint32_t a[N];
  load_elem<N>(a, X, xcap, x0 + kSubVals - 1);
  store_elem<N>(a, R, rcap, g);
  store_elem<N>(a, Y, ycap, g);
  for (int v = kSubVals - 1; v >= 1; v--) {
    load_elem<N>(a, R, rcap, g);
    mont_mul<N>(a, GlobalLoader(X, xcap, x0 + v - 1), M);
    store_elem<N>(a, R, rcap, g);
    mont_mul<N>(a, GlobalLoader(Y, ycap, g), M);
    store_elem<N>(a, Y, ycap, g);
  }
}

// 6. out[oidx] = prod_m Y[m]^(16^m), m = 0..Mcount-1, by Horner from the top.  One thread.
template <int N>
VMX_KERNEL(N) k_horner(const uint32_t* __restrict__ Y, size_t ycap, int Mcount, uint32_t* __restrict__ out,
                       size_t ocap, size_t oidx, const __grid_constant__ MontParams<N> M) {
  VMX_DYN_SMEM(uint2, smem);
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  uint32_t a[N];
  load_elem<N>(a, Y, ycap, (size_t)Mcount - 1);
  for (int m = Mcount - 2; m >= 0; m--) {
    for (int q = 0; q < kSubDigit; q++) mont_sqr<N>(a, smem, 1, M);
    mont_mul<N>(a, GlobalLoader(Y, ycap, (size_t)m), M);
  }
  store_elem<N>(a, out, ocap, oidx);
}

}  // namespace vmx
