// kernels_mexp.cuh -- bucketed (Pippenger) multi-exponentiation: prod_i b_i^{e_i}.
//
// Pipeline for window width c, W = ceil(L/c) windows (all windows processed at once):
//   1. k_digit_hist     histogram of non-zero digits: bucket id = k*2^c + digit_k(e_i)
//   2. exclusive scan   bucket offsets                              (scan.cuh)
//   3. k_digit_scatter  term indices sorted by bucket
//   4. segmented product rounds (k_chunk_plan / k_seg_prod): every bucket = product of its
//      terms; segments are cut in chunks of <= K terms, one thread per chunk, rounds repeat on
//      the partial products until one value per segment is left.  Chunks are ordered by length
//      so that the 32 threads of a warp run the same number of modmuls.
//   5. bucket reduction by radix-2^s sub-digits (s = 4): X[k][j][v] = prod of buckets whose
//      j-th sub-digit is v  (again segmented products, static index lists),
//      Y[k][j] = prod_v X[k][j][v]^v  (k_weighted_small)
//   6. k_horner         result = prod_{k,j} Y[k][j]^(2^(c*k + s*j))   (one thread; L squarings)
// The group is commutative and all arithmetic is exact, so the result is independent of the
// order in which terms are multiplied: bit-exact with any CPU evaluation.
#pragma once
#include "kernels_elem.cuh"

namespace vmx {

constexpr int kSubDigit = 4;  // s

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

// 3. scatter term indices into bucket order (cursor = copy of the exclusive offsets).
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

// 4a. number of chunks per segment: max(1, ceil(len / K)).
__global__ void k_chunk_count(const uint32_t* __restrict__ seg_off, size_t nseg, int K, uint32_t* __restrict__ nchunks) {
  const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseg) return;
  const uint32_t len = seg_off[s + 1] - seg_off[s];
  nchunks[s] = len == 0 ? 1u : (len + K - 1) / K;
}

// 4b. chunk descriptors + histogram of chunk lengths (for ordering).  chunk_off = exclusive scan of nchunks.
struct Chunk { uint32_t start, len; };
__global__ void k_chunk_fill(const uint32_t* __restrict__ seg_off, const uint32_t* __restrict__ chunk_off, size_t nseg,
                             int K, Chunk* __restrict__ chunks, uint32_t* __restrict__ len_hist) {
  const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseg) return;
  uint32_t b = seg_off[s];
  const uint32_t e = seg_off[s + 1];
  uint32_t c = chunk_off[s];
  if (b == e) { chunks[c] = Chunk{b, 0}; atomicAdd(&len_hist[0], 1u); return; }
  while (b < e) {
    const uint32_t l = min((uint32_t)K, e - b);
    chunks[c++] = Chunk{b, l};
    atomicAdd(&len_hist[K - l], 1u);  // slot 0 <-> longest
    b += l;
  }
}

// 4c. order: position of every chunk in a longest-first ordering (len_cursor = exclusive scan of len_hist).
__global__ void k_chunk_order(const Chunk* __restrict__ chunks, size_t nchunks, int K, uint32_t* __restrict__ len_cursor,
                              uint32_t* __restrict__ order) {
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nchunks) return;
  const uint32_t l = chunks[c].len;
  const uint32_t slot = l == 0 ? 0 : (uint32_t)K - l;
  // empty chunks share slot 0 with the longest; they are rare (empty segments only)
  order[atomicAdd(&len_cursor[slot], 1u)] = (uint32_t)c;
}

// 4d. one thread per chunk: out[c] = prod_{k < len} V[idx[start + k]]   (len = 0 -> one).
// `lanes` independent value planes (components of a product group) share the index lists:
// lane l reads V + l*vlane words and writes out element l*olane + c.
template <int N>
VMX_KERNEL(N) k_seg_prod(const uint32_t* __restrict__ V, size_t vcap, size_t vlane, const uint32_t* __restrict__ idx,
                         const Chunk* __restrict__ chunks, const uint32_t* __restrict__ order, size_t nchunks,
                         int lanes, uint32_t* __restrict__ out, size_t ocap, size_t olane,
                         const uint32_t* __restrict__ one, const __grid_constant__ MontParams<N> M) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nchunks * (size_t)lanes) return;
  const size_t oc = t % nchunks;
  const int lane = (int)(t / nchunks);
  const uint32_t c = order ? order[oc] : (uint32_t)oc;
  const Chunk ch = chunks[c];
  const size_t voff = (size_t)lane * vlane;
  uint32_t a[N];
  if (ch.len == 0) {
    load_elem<N>(a, one, 4, 1);
  } else {
    const size_t i0 = idx ? idx[ch.start] : ch.start;
    load_elem<N>(a, V, vcap, voff + i0);
    for (uint32_t k = 1; k < ch.len; k++) {
      const size_t ik = idx ? idx[ch.start + k] : ch.start + k;
      mont_mul<N>(a, GlobalLoader(V, vcap, voff + ik), M);
    }
  }
  store_elem<N>(a, out, ocap, (size_t)lane * olane + c);
}

// 5a. static index lists of the sub-digit products.  Segment id = (k*J + j)*V + (v-1), V = 2^s - 1;
// it lists the buckets k*2^c + d whose sub-digit j (width wj = min(s, c - s*j)) equals v:
// 2^(c-wj) entries if v < 2^wj, none otherwise.  seg_off is computed on the host (closed form).
__global__ void k_subdigit_lists(int c, int W, int J, const uint32_t* __restrict__ seg_off, uint32_t* __restrict__ idx) {
  const int s = kSubDigit;
  const int V = (1 << s) - 1;
  const size_t seg = (size_t)blockIdx.x;  // one block per segment
  const int v = (int)(seg % V) + 1;
  const int j = (int)((seg / V) % J);
  const int k = (int)(seg / ((size_t)V * J));
  const int wj = min(s, c - s * j);
  if (v >= (1 << wj)) return;
  const size_t cnt = (size_t)1 << (c - wj);
  const size_t base = seg_off[seg];
  const size_t lowmask = ((size_t)1 << (s * j)) - 1;
  for (size_t t = threadIdx.x; t < cnt; t += blockDim.x) {
    const size_t d = ((t >> (s * j)) << (s * j + wj)) | ((size_t)v << (s * j)) | (t & lowmask);
    idx[base + t] = (uint32_t)(((size_t)k << c) + d);
  }
}

// 5b. Y = prod_{v=1}^{V} X[v]^v for every group of V consecutive X values (running-product trick,
// 2V modmuls: run = X[V]; tot = X[V]; for v = V-1..1: run *= X[v]; tot *= run).
// One thread per (group, lane); X element index = lane*xlane + g*V + (v-1).  Only one residue
// fits in registers, so `run` and `tot` live in the scratch arrays R and Y (element t / slot).
template <int N>
VMX_KERNEL(N) k_weighted_small(const uint32_t* __restrict__ X, size_t xcap, size_t xlane, size_t ngroups, int V,
                               int lanes, uint32_t* __restrict__ Y, size_t ycap, size_t ylane,
                               uint32_t* __restrict__ R, size_t rcap, const __grid_constant__ MontParams<N> M) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ngroups * (size_t)lanes) return;
  const size_t g = t % ngroups;
  const int lane = (int)(t / ngroups);
  const size_t x0 = (size_t)lane * xlane + g * V;
  const size_t yi = (size_t)lane * ylane + g;
  uint32_t a[N];
  load_elem<N>(a, X, xcap, x0 + V - 1);
  store_elem<N>(a, R, rcap, t);
  store_elem<N>(a, Y, ycap, yi);
  for (int v = V - 1; v >= 1; v--) {
    load_elem<N>(a, R, rcap, t);
    mont_mul<N>(a, GlobalLoader(X, xcap, x0 + v - 1), M);
    store_elem<N>(a, R, rcap, t);
    load_elem<N>(a, Y, ycap, yi);
    mont_mul<N>(a, GlobalLoader(R, rcap, t), M);
    store_elem<N>(a, Y, ycap, yi);
  }
}

// 6. result = prod_{k,j} Y[k][j]^(2^(c*k + s*j)) by Horner from the top.  One thread per lane.
template <int N>
VMX_KERNEL(N) k_horner(const uint32_t* __restrict__ Y, size_t ycap, size_t ylane, int c, int W, int J, int lanes,
                       uint32_t* __restrict__ out, size_t ocap, const __grid_constant__ MontParams<N> M) {
  extern __shared__ uint2 smem[];
  const int lane = blockIdx.x * blockDim.x + threadIdx.x;
  if (lane >= lanes) return;
  uint2* sc = smem + threadIdx.x;
  const unsigned ss = blockDim.x;
  const int s = kSubDigit;
  uint32_t a[N];
  const size_t y0 = (size_t)lane * ylane;
  load_elem<N>(a, Y, ycap, y0 + (size_t)(W - 1) * J + (J - 1));
  for (int k = W - 1; k >= 0; k--) {
    for (int j = J - 1; j >= 0; j--) {
      if (k == W - 1 && j == J - 1) continue;
      // squarings between position (k,j) and the previous (higher) one
      const int sq = (j == J - 1) ? (c - s * (J - 1)) : s;  // moving down from (k+1,0) to (k,J-1): width of top digit
      for (int q = 0; q < sq; q++) mont_sqr<N>(a, sc, ss, M);
      mont_mul<N>(a, GlobalLoader(Y, ycap, y0 + (size_t)k * J + j), M);
    }
  }
  store_elem<N>(a, out, ocap, lane);
}

}  // namespace vmx
