// kernels_prg.cuh -- PRGHeuristic(SHA-256) expansion on the device.
//
// The reference derives the batching vector e with
//     prg.setSeed(seed); LargeIntegerArray.random(size, ebitlen, prg)
// (hvzk/PoSBasicTW.java:533-538).  PRGHeuristic(H) output is H(seed||be32(0)) || H(seed||be32(1)) || ...
// and element i of the array is the next ceil(ebitlen/8) bytes, big-endian, reduced mod 2^ebitlen
// (verificatum-vcr 3.1.0 semantics, restated in oracle/crypto.py and oracle/arithm.py).  The
// stream is counter mode, so every element is computed independently by its own thread.
#pragma once
#include <cstdint>
#include "layout.cuh"

namespace vmx {

__device__ __constant__ uint32_t kSha256K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

struct PrgSeed {
  uint8_t bytes[48];  // seed (<= 48 bytes so that seed || counter || padding is one SHA-256 block)
  int len;
};

__device__ __forceinline__ uint32_t rotr32(uint32_t x, int r) { return __funnelshift_r(x, x, r); }

// digest words (big-endian words) of SHA-256(seed || be32(counter)); single block.
__device__ inline void sha256_seed_ctr(const PrgSeed& s, uint32_t counter, uint32_t (&h)[8]) {
  uint32_t w[64];
  uint8_t blk[64];
#pragma unroll
  for (int i = 0; i < 64; i++) blk[i] = 0;
  for (int i = 0; i < s.len; i++) blk[i] = s.bytes[i];
  blk[s.len] = (uint8_t)(counter >> 24); blk[s.len + 1] = (uint8_t)(counter >> 16);
  blk[s.len + 2] = (uint8_t)(counter >> 8); blk[s.len + 3] = (uint8_t)counter;
  blk[s.len + 4] = 0x80;
  const uint32_t bitlen = 8u * (s.len + 4);
  blk[62] = (uint8_t)(bitlen >> 8); blk[63] = (uint8_t)bitlen;
#pragma unroll
  for (int i = 0; i < 16; i++)
    w[i] = ((uint32_t)blk[4 * i] << 24) | ((uint32_t)blk[4 * i + 1] << 16) | ((uint32_t)blk[4 * i + 2] << 8) | blk[4 * i + 3];
#pragma unroll
  for (int i = 16; i < 64; i++) {
    const uint32_t s0 = rotr32(w[i - 15], 7) ^ rotr32(w[i - 15], 18) ^ (w[i - 15] >> 3);
    const uint32_t s1 = rotr32(w[i - 2], 17) ^ rotr32(w[i - 2], 19) ^ (w[i - 2] >> 10);
    w[i] = w[i - 16] + s0 + w[i - 7] + s1;
  }
  uint32_t a = 0x6a09e667, b = 0xbb67ae85, c = 0x3c6ef372, d = 0xa54ff53a, e = 0x510e527f, f = 0x9b05688c,
           g = 0x1f83d9ab, hh = 0x5be0cd19;
#pragma unroll
  for (int i = 0; i < 64; i++) {
    const uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
    const uint32_t ch = (e & f) ^ (~e & g);
    const uint32_t t1 = hh + S1 + ch + kSha256K[i] + w[i];
    const uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
    const uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
    const uint32_t t2 = S0 + mj;
    hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  h[0] = 0x6a09e667 + a; h[1] = 0xbb67ae85 + b; h[2] = 0x3c6ef372 + c; h[3] = 0xa54ff53a + d;
  h[4] = 0x510e527f + e; h[5] = 0x9b05688c + f; h[6] = 0x1f83d9ab + g; h[7] = 0x5be0cd19 + hh;
}

// element i = stream bytes [offset + i*w, offset + (i+1)*w) as a big-endian integer mod 2^bitlen, w = ceil(bitlen/8).
template <int N>
__global__ void k_prg_expand(const __grid_constant__ PrgSeed seed, size_t offset, size_t n, int bitlen,
                             uint32_t* __restrict__ out, size_t cap) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int w = (bitlen + 7) / 8;
  const size_t first = offset + i * (size_t)w;  // first stream byte (most significant)
  uint32_t limb[N];
#pragma unroll
  for (int j = 0; j < N; j++) limb[j] = 0;
  uint32_t h[8];
  uint32_t cur = 0xffffffffu;
  for (int k = 0; k < w; k++) {               // k-th byte from the most significant end
    const size_t pos = first + k;
    const uint32_t blk = (uint32_t)(pos >> 5);
    if (blk != cur) { sha256_seed_ctr(seed, blk, h); cur = blk; }
    const int o = (int)(pos & 31);
    uint32_t byte = (h[o >> 2] >> (24 - 8 * (o & 3))) & 0xff;
    const int le = w - 1 - k;                 // little-endian byte index
    const int hb = bitlen - 8 * le;
    if (hb < 8) byte &= (1u << hb) - 1u;
    // dynamic limb index: write through a small switch-free path (local array, N is small here)
    limb[le >> 2] |= byte << (8 * (le & 3));
  }
  store_elem<N>(limb, out, cap, i);
}

// raw PRG stream: block first_block + t = SHA-256(seed || be32(first_block + t)) -> out[32t .. 32t+32)
__global__ void k_prg_bytes(const __grid_constant__ PrgSeed seed, size_t first_block, size_t nblocks,
                            uint8_t* __restrict__ out) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nblocks) return;
  uint32_t h[8];
  sha256_seed_ctr(seed, (uint32_t)(first_block + t), h);
  uint32_t* o = reinterpret_cast<uint32_t*>(out + 32 * t);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint32_t v = h[i];
    o[i] = (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24);
  }
}

}  // namespace vmx
