// kernels_perm.cuh -- Permutation.random on the device (mixnet/ShufflerElGamalSession.java:408-409,
// mixnet/PermutationCommitment.java:211): `n` keys of `nbytes` random bytes masked to `bits` bits are
// ranked; table[i] = rank of key i (ties by index: a stable sort).
//
// The keys are PRG output, i.e. uniform: one counting pass on their top B bits spreads them over 2^B
// buckets of a handful of keys each, one thread then orders its bucket by insertion and writes the
// ranks.  Sorting compares the leading 64 bits; keys longer than that which tie on them, and buckets
// that are too full (a degenerate random source), raise a flag and the caller ranks on the host.
#pragma once
#include "cuda_compat.cuh"

namespace vmx {

constexpr int kPermMaxBucket = 48;

struct PermKey { unsigned long long key; uint32_t idx; uint32_t pad; };

// leading 64 bits of key i (whole key, right aligned, if it has fewer than 8 bytes)
__device__ __host__ inline unsigned long long perm_lead(const uint8_t* raw, size_t i, int nbytes, int bits) {
  const uint8_t* s = raw + i * (size_t)nbytes;
  unsigned long long v = 0;
  const int take = nbytes < 8 ? nbytes : 8;
  for (int k = 0; k < take; k++) {
    unsigned b = s[k];
    if (k == 0) b &= 0xFFu >> ((8 - bits % 8) % 8);
    v = (v << 8) | b;
  }
  return v;
}
// bucket of a key: its top B bits (of the `lbits` significant bits of the lead word)
__device__ __host__ inline uint32_t perm_bucket(unsigned long long lead, int lbits, int B) {
  return lbits > B ? (uint32_t)(lead >> (lbits - B)) : (uint32_t)lead;
}

__global__ void k_perm_hist(const uint8_t* __restrict__ raw, size_t n, int nbytes, int bits, int lbits, int B,
                            uint32_t* __restrict__ hist) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  atomicAdd(&hist[perm_bucket(perm_lead(raw, i, nbytes, bits), lbits, B)], 1u);
}

__global__ void k_perm_scatter(const uint8_t* __restrict__ raw, size_t n, int nbytes, int bits, int lbits, int B,
                               uint32_t* __restrict__ cursor, PermKey* __restrict__ keys) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long lead = perm_lead(raw, i, nbytes, bits);
  const uint32_t pos = atomicAdd(&cursor[perm_bucket(lead, lbits, B)], 1u);
  keys[pos] = PermKey{lead, (uint32_t)i, 0u};
}

// one thread per bucket: order by (key, idx), write ranks; flag[0] |= 1 on an over-full bucket, |= 2 on
// two long keys with equal leading words
__global__ void k_perm_rank(const uint32_t* __restrict__ off, size_t nbuckets, PermKey* __restrict__ keys, int longkeys,
                            uint32_t* __restrict__ table, int* __restrict__ flag) {
  const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbuckets) return;
  const uint32_t lo = off[b], hi = off[b + 1];
  const uint32_t len = hi - lo;
  if (len > (uint32_t)kPermMaxBucket) { atomicOr(flag, 1); return; }
  PermKey* k = keys + lo;
  for (uint32_t a = 1; a < len; a++) {
    const PermKey x = k[a];
    uint32_t j = a;
    while (j > 0 && (k[j - 1].key > x.key || (k[j - 1].key == x.key && k[j - 1].idx > x.idx))) { k[j] = k[j - 1]; j--; }
    k[j] = x;
  }
  for (uint32_t a = 0; a < len; a++) {
    if (longkeys && a + 1 < len && k[a].key == k[a + 1].key) atomicOr(flag, 2);
    table[k[a].idx] = lo + a;
  }
}

}  // namespace vmx
