// layout.cuh -- HBM layout of element arrays and register <-> memory movers.
//
// "Vectorised limb-major": an array of `cap` residues of N 32-bit limbs is stored as N/4
// planes of uint4; plane g holds limbs 4g..4g+3 of every element, element index fastest:
//
//     word(j, i) = d[ ((j >> 2) * cap + i) * 4 + (j & 3) ]
//
// Consecutive threads (elements) therefore read consecutive 16-byte vectors: one LDG.128 per
// warp covers 512 contiguous bytes (coalesced, vectorised).  A gather of one element touches
// N/4 half-used 32-byte sectors.  Tables of precomputed powers use the same layout.
#pragma once
#include "mont.cuh"

namespace vmx {

template <int N>
VMX_DEV void load_elem(uint32_t (&a)[N], const uint32_t* d, size_t cap, size_t i) {
  const uint4* p = reinterpret_cast<const uint4*>(d) + i;
#pragma unroll
  for (int g = 0; g < N / 4; g++) {
    const uint4 v = p[(size_t)g * cap];
    a[4 * g] = v.x; a[4 * g + 1] = v.y; a[4 * g + 2] = v.z; a[4 * g + 3] = v.w;
  }
}

template <int N>
VMX_DEV void store_elem(const uint32_t (&a)[N], uint32_t* d, size_t cap, size_t i) {
  uint4* p = reinterpret_cast<uint4*>(d) + i;
#pragma unroll
  for (int g = 0; g < N / 4; g++) p[(size_t)g * cap] = make_uint4(a[4 * g], a[4 * g + 1], a[4 * g + 2], a[4 * g + 3]);
}

// Streams one element out of a limb-major array, two words per CIOS trip.
struct GlobalLoader {
  const uint32_t* p;  // d + 4*i
  size_t gs;          // words between planes = 4*cap
  VMX_DEV GlobalLoader(const uint32_t* d, size_t cap, size_t i) : p(d + 4 * i), gs(4 * cap) {}
  VMX_DEV Word2 operator()(int i) const {
    const uint2 v = *reinterpret_cast<const uint2*>(p + (size_t)(i >> 2) * gs + (i & 3));
    return Word2{v.x, v.y};
  }
};

// Per-thread scratch copy of a residue in shared memory, word pair k of thread t at
// s[k * blockDim.x + t] (conflict-free LDS.64/STS.64).  Used as the streamed operand of a
// squaring.
struct SharedLoader {
  const uint2* s;
  unsigned stride;
  VMX_DEV Word2 operator()(int i) const {
    const uint2 v = s[(unsigned)(i >> 1) * stride];
    return Word2{v.x, v.y};
  }
};

template <int N>
VMX_DEV void stash_shared(const uint32_t (&a)[N], uint2* s, unsigned stride) {
#pragma unroll
  for (int k = 0; k < N / 2; k++) s[(unsigned)k * stride] = make_uint2(a[2 * k], a[2 * k + 1]);
}

// The constant 1 as a streamed operand (Montgomery -> canonical conversion).
struct OneLoader {
  VMX_DEV Word2 operator()(int i) const { return Word2{i == 0 ? 1u : 0u, 0u}; }
};

// a <- a^2 (Montgomery), via a shared-memory copy of a: the block-triangular squaring of mont.cuh with blocks of
// VMX_SQR_BLOCK words where the residue has at least two of them, a plain multiplication otherwise.  32-word
// blocks: 0.83 of a multiplication's IMAD.WIDE at 96 limbs (3 blocks), 0.875 at 64 (2 blocks).  16-word blocks
// would be 0.79 but their six loop bodies (45 KB of SASS) push the loop of an exponentiation kernel out of the
// instruction cache: measured on B200 at 3072 bits, n = 10^5 (profiles/r04_squaring_block_sizes.txt),
// x^v * y^k: plain 216 ms, 48-word blocks 194, 32-word 197, 16-word 227; x^e, |e| = 3071: 922 / 824 / 820 / 850.
#ifndef VMX_SQR_BLOCK
#define VMX_SQR_BLOCK 32
#endif
template <int N, int BS = VMX_SQR_BLOCK>
VMX_DEV void mont_sqr(uint32_t (&a)[N], uint2* s, unsigned stride, const MontParams<N>& M) {
  stash_shared<N>(a, s, stride);
  if constexpr (BS % 16 == 0 && N % BS == 0 && N >= 2 * BS) mont_sqr_tri<N, BS>(a, SharedLoader{s, stride}, M);
  else mont_mul<N>(a, SharedLoader{s, stride}, M);
}

// bit window [pos, pos+w) of the little-endian limb string of element i in a limb-major array.
template <int N>
VMX_DEV uint32_t window_bits(const uint32_t* d, size_t cap, size_t i, int pos, int w) {
  const int j = pos >> 5, sh = pos & 31;
  if (j >= N) return 0;
  uint32_t lo = d[((size_t)(j >> 2) * cap + i) * 4 + (j & 3)];
  uint32_t v = lo >> sh;
  if (sh + w > 32 && j + 1 < N) {
    const int j1 = j + 1;
    uint32_t hi = d[((size_t)(j1 >> 2) * cap + i) * 4 + (j1 & 3)];
    v |= hi << (32 - sh);
  }
  return v & ((w >= 32) ? 0xffffffffu : ((1u << w) - 1u));
}

}  // namespace vmx
