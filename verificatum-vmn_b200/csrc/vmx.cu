// vmx.cu -- host side of the C ABI declared in include/vmx.h: handle management, kernel
// orchestration (window tables, Pippenger plan, segmented products, scans) and accounting.
// All arithmetic runs in the sm_100a kernels of kernels_*.cuh; there is no CPU path.
#include "vmx_internal.cuh"
#include "coop.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <memory>

#ifdef VMX_HOST_EMUL
namespace vmx_emul {
thread_local dim3 threadIdx_, blockIdx_, blockDim_, gridDim_;
thread_local unsigned char* dyn_smem = nullptr;
}  // namespace vmx_emul
#endif

namespace vmx {

// ------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

#define VMX_DISPATCH(nl, ...)                                                  \
  switch (nl) {                                                                \
    case 8: { constexpr int N = 8; __VA_ARGS__; } break;                       \
    case 16: { constexpr int N = 16; __VA_ARGS__; } break;                     \
    case 32: { constexpr int N = 32; __VA_ARGS__; } break;                     \
    case 64: { constexpr int N = 64; __VA_ARGS__; } break;                     \
    case 96: { constexpr int N = 96; __VA_ARGS__; } break;                     \
    default: set_error("unsupported limb count %d", (int)(nl)); return VMX_EARG; \
  }

#define VMX_CHECK_LAUNCH() VMX_CU(cudaGetLastError())

static inline unsigned nblocks(size_t n, int per = kThreads) { return (unsigned)((n + per - 1) / per); }
// dynamic shared memory of the byte codec kernels: the records of one block (kCodecThreads elements)
// threads per block and dynamic shared memory of the byte codec kernels: the records of one block plus the slack
// of stage_in; long records (wide random integers) take smaller blocks to stay below the 48 KB default limit
static inline int codec_threads(size_t rec) {
  int t = kCodecThreads;
  while (t > 8 && (size_t)t * rec + 48 > 48 * 1024) t >>= 1;
  return t;
}
static inline size_t codec_smem(size_t rec) { return ((size_t)codec_threads(rec) * rec + 47) & ~(size_t)15; }
static inline size_t cap_for(size_t n) { return std::max<size_t>(8, (n + 7) & ~(size_t)7); }

// ------------------------------------------------------------------ tiny host bignum (setup only)
// little-endian 32-bit limbs; used for R mod n, R^2 mod n, n0inv, p-2: O(bits) shifts at
// context creation.  No group arithmetic is done on the host, with one O(1) exception per
// proof: the inversion of a single element (limbs_inv_mod, vmx_elem_inv).
static bool be_to_limbs(const uint8_t* be, size_t nbytes, uint32_t* out, int N) {
  for (int j = 0; j < N; j++) out[j] = 0;
  for (size_t b = 0; b < nbytes; b++) {
    const uint8_t v = be[nbytes - 1 - b];
    if (b / 4 >= (size_t)N) { if (v) return false; continue; }
    out[b / 4] |= (uint32_t)v << (8 * (b % 4));
  }
  return true;
}
static int limbs_bits(const uint32_t* a, int N) {
  for (int j = N - 1; j >= 0; j--) if (a[j]) return 32 * j + (32 - __builtin_clz(a[j]));
  return 0;
}
static int limbs_cmp(const uint32_t* a, const uint32_t* b, int N) {
  for (int j = N - 1; j >= 0; j--) if (a[j] != b[j]) return a[j] < b[j] ? -1 : 1;
  return 0;
}
static void limbs_sub(uint32_t* a, const uint32_t* b, int N) {
  uint64_t brw = 0;
  for (int j = 0; j < N; j++) { const uint64_t d = (uint64_t)a[j] - b[j] - brw; a[j] = (uint32_t)d; brw = (d >> 32) & 1; }
}
// r = 2^k mod n
static void pow2_mod(uint32_t* r, int k, const uint32_t* n, int N) {
  for (int j = 0; j < N; j++) r[j] = 0;
  r[0] = 1;
  for (int s = 0; s < k; s++) {
    uint32_t c = 0;
    for (int j = 0; j < N; j++) { const uint32_t nc = r[j] >> 31; r[j] = (r[j] << 1) | c; c = nc; }
    if (c || limbs_cmp(r, n, N) >= 0) limbs_sub(r, n, N);
  }
}
static uint32_t neg_inv32(uint32_t n0) {
  uint32_t x = n0;  // n0 * x = 1 mod 2^3
  for (int i = 0; i < 5; i++) x *= 2u - n0 * x;
  return 0u - x;
}
static void limbs_add(uint32_t* a, const uint32_t* b, int N, uint32_t* carry_out) {
  uint64_t c = 0;
  for (int j = 0; j < N; j++) { const uint64_t t = (uint64_t)a[j] + b[j] + c; a[j] = (uint32_t)t; c = t >> 32; }
  *carry_out = (uint32_t)c;
}
// x <- x / 2 mod n (n odd), x < n
static void limbs_half_mod(uint32_t* x, const uint32_t* n, int N) {
  uint32_t top = 0;
  if (x[0] & 1) limbs_add(x, n, N, &top);
  for (int j = 0; j < N; j++) x[j] = (x[j] >> 1) | ((j + 1 < N ? x[j + 1] : top) << 31);
}
// r = a^{-1} mod n for odd n and gcd(a, n) = 1 (binary extended Euclid).  The ONE piece of
// residue arithmetic that runs on the host: PGroupElement.inv()/div() of a SINGLE element
// (hvzk/PoSBasicTW.java:1013-1014 computes two of them per proof; the reference does the same on
// a host BigInteger).  Array inversion is vmx_inv and runs on the device.
static bool limbs_inv_mod(uint32_t* r, const uint32_t* a, const uint32_t* n, int N) {
  std::vector<uint32_t> u(a, a + N), v(n, n + N), x1(N, 0), x2(N, 0);
  x1[0] = 1;
  auto is_one = [&](const std::vector<uint32_t>& t) { if (t[0] != 1) return false; for (int j = 1; j < N; j++) if (t[j]) return false; return true; };
  auto is_zero = [&](const std::vector<uint32_t>& t) { for (int j = 0; j < N; j++) if (t[j]) return false; return true; };
  if (is_zero(u)) return false;
  auto sub_mod = [&](std::vector<uint32_t>& x, const std::vector<uint32_t>& y) {  // x = x - y mod n
    if (limbs_cmp(x.data(), y.data(), N) < 0) { uint32_t c; limbs_add(x.data(), n, N, &c); }
    limbs_sub(x.data(), y.data(), N);
  };
  for (int guard = 0; guard < 4 * 32 * N + 8; guard++) {
    if (is_one(u)) { std::memcpy(r, x1.data(), 4 * (size_t)N); return true; }
    if (is_one(v)) { std::memcpy(r, x2.data(), 4 * (size_t)N); return true; }
    while (!(u[0] & 1)) {
      for (int j = 0; j < N; j++) u[j] = (u[j] >> 1) | ((j + 1 < N ? u[j + 1] : 0u) << 31);
      limbs_half_mod(x1.data(), n, N);
    }
    while (!(v[0] & 1)) {
      for (int j = 0; j < N; j++) v[j] = (v[j] >> 1) | ((j + 1 < N ? v[j + 1] : 0u) << 31);
      limbs_half_mod(x2.data(), n, N);
    }
    if (limbs_cmp(u.data(), v.data(), N) >= 0) { limbs_sub(u.data(), v.data(), N); sub_mod(x1, x2); if (is_zero(u)) return false; }
    else { limbs_sub(v.data(), u.data(), N); sub_mod(x2, x1); }
  }
  return false;
}
// element `idx` of a limb-major host image with capacity `cap`
static void image_put(std::vector<uint32_t>& img, size_t cap, size_t idx, const uint32_t* limbs, int N) {
  for (int j = 0; j < N; j++) img[((size_t)(j >> 2) * cap + idx) * 4 + (j & 3)] = limbs[j];
}

// ------------------------------------------------------------------ device memory helpers
// Device memory of a context: small blocks from the stream-ordered pool, large ones through the context's
// recycling list (vmx_ctx::big_free).  *granted = the size to hand back to dev_free.
static cudaError_t dev_alloc(vmx_ctx* c, size_t bytes, void** p, size_t* granted) {
  if (bytes < c->big_block_min) {
    *granted = bytes;
    return cudaMallocAsync(p, bytes ? bytes : 16, c->stream);
  }
  {
    std::lock_guard<std::mutex> lk(c->big_mu);
    size_t best = (size_t)-1;
    for (size_t i = 0; i < c->big_free.size(); i++) {
      const size_t b = c->big_free[i].bytes;
      if (b >= bytes && b <= bytes + bytes / 4 && (best == (size_t)-1 || b < c->big_free[best].bytes)) best = i;
    }
    if (best != (size_t)-1) {
      *p = c->big_free[best].p;
      *granted = c->big_free[best].bytes;
      c->big_free_bytes -= *granted;
      c->big_free.erase(c->big_free.begin() + (long)best);
      return cudaSuccess;
    }
  }
  const size_t rounded = (bytes + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
  *granted = rounded;
  cudaError_t e = cudaMallocAsync(p, rounded, c->stream);
  if (e != cudaSuccess) {  // out of memory with blocks parked in the list: give them back and try once more
    (void)cudaGetLastError();
    std::lock_guard<std::mutex> lk(c->big_mu);
    for (auto& b : c->big_free) cudaFreeAsync(b.p, c->stream);
    c->big_free.clear();
    c->big_free_bytes = 0;
    e = cudaMallocAsync(p, rounded, c->stream);
  }
  return e;
}
static void dev_free(vmx_ctx* c, void* p, size_t granted) {
  if (!p) return;
  if (granted < c->big_block_min) { cudaFreeAsync(p, c->stream); return; }
  std::lock_guard<std::mutex> lk(c->big_mu);
  c->big_free.push_back({p, granted});
  c->big_free_bytes += granted;
  while (c->big_free_bytes > c->big_cache_max && !c->big_free.empty()) {  // oldest first
    cudaFreeAsync(c->big_free.front().p, c->stream);
    c->big_free_bytes -= c->big_free.front().bytes;
    c->big_free.erase(c->big_free.begin());
  }
}

struct DevBuf {  // stream-ordered temporary
  vmx_ctx* c = nullptr;
  void* p = nullptr;
  size_t granted = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { reset(); }
  void reset() { if (p) dev_free(c, p, granted); p = nullptr; }
  int alloc(vmx_ctx* ctx, size_t bytes) {
    reset();
    c = ctx;
    if (dev_alloc(ctx, bytes, &p, &granted) != cudaSuccess) {
      p = nullptr;
      (void)cudaGetLastError();
      set_error("device allocation of %zu bytes failed", bytes);
      return VMX_ENOMEM;
    }
    return VMX_OK;
  }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// a limb-major element array used as a temporary
struct ElemBuf : DevBuf {
  size_t cap = 0;
  int alloc_elems(vmx_ctx* ctx, size_t n) { cap = cap_for(n); return alloc(ctx, cap * ctx->nl * 4); }
  int alloc_limbs(vmx_ctx* ctx, size_t n, int limbs) { cap = cap_for(n); return alloc(ctx, cap * (size_t)limbs * 4); }
  int alloc_gelems(vmx_ctx* ctx, size_t n) { return alloc_limbs(ctx, n, ctx->gl); }  // group elements
  uint32_t* d() const { return as<uint32_t>(); }
};

// Element counts are bounded well below where cap * limbs * 4 leaves 64 bits (indices of permutations and the launch
// geometry are 32-bit); the reference's arrays are Java arrays, int-indexed.
static constexpr size_t kMaxElems = (size_t)1 << 31;
static int new_garr(vmx_ctx* c, size_t n, vmx_garr** out) {
  *out = nullptr;
  if (n > kMaxElems) { set_error("array of %zu elements is beyond the engine's limit of 2^31", n); return VMX_ESIZE; }
  auto* a = new (std::nothrow) vmx_garr{c, n, cap_for(n), nullptr};
  if (!a) return VMX_ENOMEM;
  void* p = nullptr;
  if (dev_alloc(c, a->cap * c->gl * 4, &p, &a->granted) != cudaSuccess) {
    (void)cudaGetLastError();
    delete a;
    set_error("device allocation of %zu bytes failed", a->cap * c->gl * 4);
    return VMX_ENOMEM;
  }
  a->d = (uint32_t*)p;
  *out = a;
  return VMX_OK;
}
static int new_rarr(vmx_ctx* c, size_t n, vmx_rarr** out) {
  *out = nullptr;
  if (n > kMaxElems) { set_error("array of %zu elements is beyond the engine's limit of 2^31", n); return VMX_ESIZE; }
  auto* a = new (std::nothrow) vmx_rarr{c, n, cap_for(n), nullptr, -1};
  if (!a) return VMX_ENOMEM;
  void* p = nullptr;
  if (dev_alloc(c, a->cap * c->nl * 4, &p, &a->granted) != cudaSuccess) {
    (void)cudaGetLastError();
    delete a;
    set_error("device allocation of %zu bytes failed", a->cap * c->nl * 4);
    return VMX_ENOMEM;
  }
  a->d = (uint32_t*)p;
  *out = a;
  return VMX_OK;
}

static int enter_(const vmx_ctx* c) {
  if (!c) { set_error("null context"); return VMX_EARG; }
  VMX_CU(cudaSetDevice(c->device));
  return VMX_OK;
}
// API entry: select the device and serialise host threads on the context (all work of a
// context is queued on its one stream anyway; the flag scratch is shared).
#define VMX_ENTER(c)    \
  VMX_TRY(enter_(c)); \
  std::unique_lock<std::recursive_mutex> _api_lock(const_cast<vmx_ctx*>(c)->api)

// read back ctx->d_flag[0..k) (synchronises the stream)
static int read_flags(vmx_ctx* c, int k) {
  VMX_CU(cudaMemcpyAsync(c->h_flag, c->d_flag, sizeof(int) * k, cudaMemcpyDeviceToHost, c->stream));
  VMX_CU(cudaStreamSynchronize(c->stream));
  return VMX_OK;
}

// ------------------------------------------------------------------ exclusive scan (uint32)
static int exclusive_scan(vmx_ctx* c, uint32_t* d, size_t n) {
  if (n == 0) return VMX_OK;
  const size_t nb = (n + kScanBlock - 1) / kScanBlock;
  if (nb == 1) {
    VMX_LAUNCH(c, k_scan_block, 1, kScanBlock, 0, d, n, (uint32_t*)nullptr);
    VMX_CHECK_LAUNCH();
    return VMX_OK;
  }
  DevBuf sums;
  VMX_TRY(sums.alloc(c, nb * 4));
  VMX_LAUNCH(c, k_scan_block, nb, kScanBlock, 0, d, n, sums.as<uint32_t>());
  VMX_CHECK_LAUNCH();
  VMX_TRY(exclusive_scan(c, sums.as<uint32_t>(), nb));
  VMX_LAUNCH(c, k_scan_add, nb, kScanBlock, 0, d, n, sums.as<uint32_t>());
  VMX_CHECK_LAUNCH();
  return VMX_OK;
}

static size_t wave_threads(const vmx_ctx* c);
// chunks per round from which k_seg_prod takes them sorted by length (env VMX_MEXP_SORT_MIN: tests force it)
static size_t mexp_sort_min(const vmx_ctx* c) {
  static const long v = [] { const char* e = std::getenv("VMX_MEXP_SORT_MIN"); return e ? std::atol(e) : -1L; }();
  return v >= 0 ? (size_t)v : 4 * wave_threads(c);
}

// ------------------------------------------------------------------ segmented products
// out[s] = prod_{k in [seg_off[s], seg_off[s+1])} V[idx ? idx[k] : k]  (empty -> one) for
// s < nseg.  `total_bound` >= seg_off[nseg].  Runs chunked rounds (K terms per thread).
template <int N>
static int seg_product(vmx_ctx* c, const Modulus& Mod, const uint32_t* V, size_t vcap, const uint32_t* idx,
                       const uint32_t* seg_off, size_t nseg, size_t total_bound, int K, uint32_t* out, size_t ocap) {
  const MontParams<N> M = Mod.params<N>();
  DevBuf off_keep;   // seg_off of the current round when it is one of ours
  ElemBuf val_keep;  // partial products feeding the current round
  const uint32_t* cur_V = V;
  size_t cur_vcap = vcap;
  const uint32_t* cur_idx = idx;
  const uint32_t* cur_off = seg_off;
  size_t cur_total = total_bound;
  for (int round = 0; round < 64; round++) {
    const size_t nch_bound = nseg + cur_total / K + 1;
    DevBuf chunk_off, chunks;
    VMX_TRY(chunk_off.alloc(c, (nseg + 1) * 4));
    VMX_TRY(chunks.alloc(c, nch_bound * sizeof(Chunk)));
    VMX_CU(cudaMemsetAsync(c->d_flag, 0, sizeof(int) * 4, c->stream));
    VMX_LAUNCH(c, k_chunk_count, nblocks(nseg + 1, 256), 256, 0, cur_off, nseg, K, chunk_off.as<uint32_t>(),
               reinterpret_cast<uint32_t*>(c->d_flag));
    VMX_CHECK_LAUNCH();
    VMX_TRY(exclusive_scan(c, chunk_off.as<uint32_t>(), nseg + 1));
    VMX_LAUNCH(c, k_chunk_fill, nblocks(nch_bound, 256), 256, 0, cur_off, chunk_off.as<uint32_t>(), nseg,
               chunks.as<Chunk>());
    VMX_CHECK_LAUNCH();
    VMX_TRY(read_flags(c, 1));
    const bool last = (c->h_flag[0] == 0);  // every segment fits one chunk: chunk id == segment id
    const uint32_t* nch_dev = chunk_off.as<uint32_t>() + nseg;
    // chunks of equal length side by side (worth it once there are several waves of them)
    DevBuf order_buf, bins;
    const uint32_t* order = nullptr;
    if (nch_bound >= mexp_sort_min(c)) {
      VMX_TRY(order_buf.alloc(c, nch_bound * 4));
      VMX_TRY(bins.alloc(c, (K + 2) * 4));
      VMX_CU(cudaMemsetAsync(bins.p, 0, (K + 2) * 4, c->stream));
      VMX_LAUNCH(c, k_chunk_len_hist, nblocks(nch_bound, 256), 256, 0, chunks.as<Chunk>(), nch_dev, K, bins.as<uint32_t>());
      VMX_CHECK_LAUNCH();
      VMX_LAUNCH(c, k_chunk_len_offsets, 1, 32, 0, K, bins.as<uint32_t>());
      VMX_CHECK_LAUNCH();
      VMX_LAUNCH(c, k_chunk_len_scatter, nblocks(nch_bound, 256), 256, 0, chunks.as<Chunk>(), nch_dev, K,
                 bins.as<uint32_t>(), order_buf.as<uint32_t>());
      VMX_CHECK_LAUNCH();
      order = order_buf.as<uint32_t>();
    }
    if (last) {
      VMX_LAUNCH(c, k_seg_prod<N>, nblocks(nseg), kThreads, 0, cur_V, cur_vcap, cur_idx, chunks.as<Chunk>(), nch_dev,
                 order, out, ocap, Mod.consts, M);
      VMX_CHECK_LAUNCH();
      c->modmuls += cur_total;
      return VMX_OK;
    }
    ElemBuf part;
    VMX_TRY(part.alloc_elems(c, nch_bound));
    VMX_LAUNCH(c, k_seg_prod<N>, nblocks(nch_bound), kThreads, 0, cur_V, cur_vcap, cur_idx, chunks.as<Chunk>(),
               nch_dev, order, part.d(), part.cap, Mod.consts, M);
    VMX_CHECK_LAUNCH();
    c->modmuls += cur_total;
    // next round: values = partial products, segments = chunk ranges
    std::swap(val_keep.p, part.p); std::swap(val_keep.c, part.c); std::swap(val_keep.cap, part.cap);
    std::swap(val_keep.granted, part.granted);
    std::swap(off_keep.p, chunk_off.p); std::swap(off_keep.c, chunk_off.c); std::swap(off_keep.granted, chunk_off.granted);
    cur_V = val_keep.d();
    cur_vcap = val_keep.cap;
    cur_idx = nullptr;
    cur_off = off_keep.as<uint32_t>();
    cur_total = nch_bound;
  }
  set_error("segmented product did not converge");
  return VMX_ECUDA;
}

// ------------------------------------------------------------------ constants on the device
static int upload_consts(vmx_ctx* c, Modulus& Mod) {
  const int N = c->nl;
  std::vector<uint32_t> img((size_t)4 * N, 0), r(N), r2(N), one(N, 0);
  pow2_mod(r.data(), 32 * N, Mod.n, N);
  pow2_mod(r2.data(), 64 * N, Mod.n, N);
  one[0] = 1;
  image_put(img, 4, 0, r2.data(), N);
  image_put(img, 4, 1, r.data(), N);
  image_put(img, 4, 2, one.data(), N);
  image_put(img, 4, 3, Mod.n, N);  // the modulus itself (cooperative kernels read their lane's limbs)
  void* p = nullptr;
  VMX_CU(cudaMallocAsync(&p, img.size() * 4, c->stream));
  Mod.consts = (uint32_t*)p;
  VMX_CU(cudaMemcpyAsync(p, img.data(), img.size() * 4, cudaMemcpyHostToDevice, c->stream));
  VMX_CU(cudaStreamSynchronize(c->stream));
  return VMX_OK;
}

// upload one element given as big-endian bytes into a fresh 1-element (cap 8) temporary,
// group: to Montgomery form with range check; ring: canonical with range check.
static int ec_upload_one(vmx_ctx* c, const uint8_t* be, ElemBuf& buf);
static int ec_download_one(vmx_ctx* c, const uint32_t* d, size_t cap, size_t idx, uint8_t* out_be);
static int upload_one(vmx_ctx* c, const uint8_t* be, bool group, ElemBuf& buf) {
  if (group && c->kind == 1) return ec_upload_one(c, be, buf);
  const size_t eb = group ? c->eb : c->rb;
  DevBuf raw;
  VMX_TRY(raw.alloc(c, eb));
  VMX_CU(cudaMemcpyAsync(raw.p, be, eb, cudaMemcpyHostToDevice, c->stream));
  VMX_TRY(buf.alloc_elems(c, 1));
  VMX_CU(cudaMemsetAsync(c->d_flag, 0, sizeof(int) * 4, c->stream));
  const Modulus& Mod = group ? c->P : c->Q;
  VMX_DISPATCH(c->nl, VMX_LAUNCH(c, k_from_bytes<N>, 1, codec_threads(eb), codec_smem(eb), raw.as<uint8_t>(), (size_t)1, (int)eb, 0,
                                 group ? 0 : 1, buf.d(), buf.cap, Mod.consts, c->d_flag, Mod.params<N>()));
  VMX_CHECK_LAUNCH();
  VMX_TRY(read_flags(c, 1));
  if (c->h_flag[0]) { set_error("element out of range (flags %d)", c->h_flag[0]); return VMX_EFORMAT; }
  return VMX_OK;
}

// download element `idx` of a limb-major array as big-endian bytes (synchronises)
static int download_one(vmx_ctx* c, const uint32_t* d, size_t cap, size_t idx, bool group, uint8_t* out_be) {
  if (group && c->kind == 1) return ec_download_one(c, d, cap, idx, out_be);
  const size_t eb = group ? c->eb : c->rb;
  DevBuf raw;
  VMX_TRY(raw.alloc(c, eb));
  const Modulus& Mod = group ? c->P : c->Q;
  // k_to_bytes addresses element i = thread index: shift the base so that thread 0 -> idx
  VMX_DISPATCH(c->nl, VMX_LAUNCH(c, k_to_bytes<N>, 1, codec_threads(eb), codec_smem(eb), d + 4 * idx, cap, (size_t)1, (int)eb, 0,
                                 group ? 0 : 1, raw.as<uint8_t>(), Mod.params<N>()));
  VMX_CHECK_LAUNCH();
  c->modmuls += group ? 1 : 0;
  VMX_CU(cudaMemcpyAsync(out_be, raw.p, eb, cudaMemcpyDeviceToHost, c->stream));
  VMX_CU(cudaStreamSynchronize(c->stream));
  return VMX_OK;
}

static int rarr_bitlen(const vmx_rarr* a, int* bits) {
  if (a->bits < 0) {
    vmx_ctx* c = a->ctx;
    VMX_CU(cudaMemsetAsync(c->d_flag, 0, sizeof(int) * 4, c->stream));
    if (a->n) {
      VMX_DISPATCH(c->nl, VMX_LAUNCH(c, k_bitlen<N>, nblocks(a->n, 256), 256, 0, a->d, a->cap, a->n,
                                     reinterpret_cast<unsigned*>(c->d_flag)));
      VMX_CHECK_LAUNCH();
    }
    VMX_TRY(read_flags(c, 1));
    a->bits = c->h_flag[0];
  }
  *bits = a->bits;
  return VMX_OK;
}

// number of resident threads one "wave" of the thread-per-element kernels fills
static size_t wave_threads(const vmx_ctx* c) {
  return (size_t)c->sm_count * kThreads * (c->nl > 64 ? 2 : 3);
}

// ------------------------------------------------------------------ fixed-base tables
// Window width of a fixed-base table for arrays of n exponents.  A table is built once per base and serves
// every later array (the bases of a mix-net session are g, h0 and the public key components: 8 fixed-base
// arrays per shuffle, mixnet/ShufflerElGamalSession.java:407, hvzk/PoSBasicTW.java:447,606-646,1030), so
// the 2^w entries per window are weighed against kReuse arrays of n products.  Memory bound per table:
// ctx->table_max_bytes (20 GB: w = 18 at 3072 bits is 171 windows x 262,144 x 384 B = 17.2 GB of the 180 GB; four
// long-lived bases); all cached tables together stay below ctx->table_budget (least recently used evicted).
constexpr double kFixedReuse = 8.0;
static double fixed_cost(int w, size_t n, int ebits) {
  const double nwin = (ebits + w - 1) / w;
  return nwin * (kFixedReuse * (double)n + (double)(1u << w));
}
static int choose_fixed_window(const vmx_ctx* c, size_t n, int ebits) {
  if (c->fixed_window) return c->fixed_window;
  double best = 1e300;
  int bw = 4;
  for (int w = 4; w <= 18; w++) {
    const double nwin = (ebits + w - 1) / w;
    const double entries = nwin * (double)(1u << w);
    if (entries * c->nl * 4 > (double)c->table_max_bytes) break;
    const double cost = fixed_cost(w, n, ebits);
    if (cost < best) { best = cost; bw = w; }
  }
  return bw;
}

template <int N>
static int build_table(vmx_ctx* c, const uint32_t* base, size_t bcap, int w, FixedTable& T) {
  const MontParams<N> M = c->P.params<N>();
  const int ebits = c->Q.bits;  // exponents are ring elements < q
  T.w = w;
  T.nwin = (ebits + w - 1) / w;
  const size_t entries = (size_t)T.nwin << w;
  T.cap = cap_for(entries);
  void* p = nullptr;
  if (cudaMallocAsync(&p, T.cap * N * 4, c->stream) != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("fixed-base table allocation failed (%zu bytes)", T.cap * N * 4);
    return VMX_ENOMEM;
  }
  T.d = (uint32_t*)p;
  const int qlen = T.nwin * w;
  ElemBuf Q;
  VMX_TRY(Q.alloc_elems(c, qlen));
#ifndef VMX_HOST_EMUL
  VMX_LAUNCH(c, k_coop_sqr_chain<N>, 1, 32, 0, base, bcap, (size_t)0, Q.d(), Q.cap, qlen, c->P.consts, M.n0inv);
#else
  VMX_LAUNCH(c, k_sqr_chain<N>, 1, 32, N * 4, base, bcap, (size_t)0, Q.d(), Q.cap, qlen, M);
#endif
  VMX_CHECK_LAUNCH();
  c->modmuls += qlen;
  for (int j = 0; j < w; j++) {
    const size_t items = (size_t)T.nwin << j;
    VMX_LAUNCH(c, k_table_level<N>, nblocks(items), kThreads, 0, T.d, T.cap, w, T.nwin, j, qlen, Q.d(), Q.cap,
               c->P.consts, M);
    VMX_CHECK_LAUNCH();
    c->modmuls += items;
  }
  return VMX_OK;
}

// find or build the table for `base_be`, sized for n exponents
static int ec_choose_fixed_window(const vmx_ctx* c, size_t n);
static double ec_fixed_cost(int w, size_t n, int ebits);
static int ec_build_table(vmx_ctx* c, const uint32_t* base, size_t bcap, int w, FixedTable& T);
static int get_table(vmx_ctx* c, const uint8_t* base_be, size_t n, FixedTable* out) {
  const std::string key(reinterpret_cast<const char*>(base_be), c->eb);
  const int ebits = c->Q.bits;
  const bool ec = c->kind == 1;
  const int wbest = ec ? ec_choose_fixed_window(c, n) : choose_fixed_window(c, n, ebits);
  std::lock_guard<std::mutex> lk(c->mu);
  auto it = c->tables.find(key);
  if (it != c->tables.end()) {
    const double have = ec ? ec_fixed_cost(it->second.w, n, ebits) : fixed_cost(it->second.w, n, ebits);
    const double want = ec ? ec_fixed_cost(wbest, n, ebits) : fixed_cost(wbest, n, ebits);
    // a wider table than this array size asks for is never slower to USE: keep it; rebuild only to widen
    if (it->second.w >= wbest || have <= 1.3 * want) {
      it->second.last_use = ++c->table_clock;
      *out = it->second;
      return VMX_OK;
    }
    cudaFreeAsync(it->second.d, c->stream);
    c->table_bytes -= it->second.bytes;
    c->tables.erase(it);
  }
  ElemBuf base;
  VMX_TRY(upload_one(c, base_be, true, base));
  FixedTable T;
  if (ec) VMX_TRY(ec_build_table(c, base.d(), base.cap, wbest, T));
  else VMX_DISPATCH(c->nl, VMX_TRY(build_table<N>(c, base.d(), base.cap, wbest, T)));
  T.bytes = T.cap * (size_t)c->gl * 4;
  T.last_use = ++c->table_clock;
  c->tables[key] = T;
  c->table_bytes += T.bytes;
  // a long-lived context that meets many bases (a key per election, ad-hoc g.exp(array) bases) must not grow
  // without bound: evict the least recently used tables (stream-ordered free: kernels already queued finish first)
  while (c->table_bytes > c->table_budget && c->tables.size() > 1) {
    auto victim = c->tables.end();
    for (auto jt = c->tables.begin(); jt != c->tables.end(); ++jt)
      if (jt->first != key && (victim == c->tables.end() || jt->second.last_use < victim->second.last_use)) victim = jt;
    if (victim == c->tables.end()) break;
    cudaFreeAsync(victim->second.d, c->stream);
    c->table_bytes -= victim->second.bytes;
    c->tables.erase(victim);
  }
  *out = T;
  return VMX_OK;
}

template <int N>
static int exp_fixed_run(vmx_ctx* c, const FixedTable& T, const vmx_rarr* e, int ebits, uint32_t* out, size_t ocap) {
  const MontParams<N> M = c->P.params<N>();
  const size_t n = e->n;
  int nwin = std::min(T.nwin, std::max(1, (ebits + T.w - 1) / T.w));
  // split the windows of one exponent over `parts` threads when that fills the waves better
  const size_t wave = wave_threads(c);
  int parts = 1;
  if (nwin >= 8) {
    double best_eff = 0;
    for (int p = 1; p <= 64 && p * 4 <= nwin; p++) {
      const double waves = (double)n * p / wave;
      // the combine runs on the warp-cooperative multiplier (~10x lower latency, ~7000 warps in flight)
      const double time = std::ceil(waves) * ((double)nwin / p) + (p > 1 ? std::ceil((double)n / 7000.0) * 0.1 * p : 0);
      const double eff = ((double)n * nwin / wave) / time;                 // useful / spent
      if (eff > best_eff * 1.02) { best_eff = eff; parts = p; }
    }
  }
  if (parts == 1) {
    VMX_LAUNCH(c, k_exp_fixed<N>, nblocks(n), kThreads, 0, T.d, T.cap, T.w, nwin, e->d, e->cap, n, 1, out, ocap, M);
    VMX_CHECK_LAUNCH();
    c->modmuls += (uint64_t)n * (nwin - 1);
    return VMX_OK;
  }
  ElemBuf tmp;
  VMX_TRY(tmp.alloc_elems(c, n * parts));
  VMX_LAUNCH(c, k_exp_fixed<N>, nblocks(n * parts), kThreads, 0, T.d, T.cap, T.w, nwin, e->d, e->cap, n, parts,
             tmp.d(), tmp.cap, M);
  VMX_CHECK_LAUNCH();
#ifndef VMX_HOST_EMUL
  VMX_LAUNCH(c, k_coop_combine_parts<N>, nblocks(n, kCoopWarps), 32 * kCoopWarps, 0, tmp.d(), tmp.cap, n, parts, out,
             ocap, c->P.consts, M.n0inv);
#else
  VMX_LAUNCH(c, k_combine_parts<N>, nblocks(n), kThreads, 0, tmp.d(), tmp.cap, n, parts, out, ocap, M);
#endif
  VMX_CHECK_LAUNCH();
  c->modmuls += (uint64_t)n * (nwin - 1);
  return VMX_OK;
}

// ------------------------------------------------------------------ variable-base
static int choose_var_window(int ebits) {
  int bw = 1;
  double best = 1e300;
  for (int w = 1; w <= 6; w++) {
    const double cost = (double)(1 << w) + ebits + (double)ebits / w;
    if (cost < best) { best = cost; bw = w; }
  }
  return bw;
}

// out[i] = a[i]^{E[i or 0]} where E is a limb-major exponent array (capacity ecap)
template <int N>
static int exp_var_run(vmx_ctx* c, const uint32_t* a, size_t acap, const uint32_t* E, size_t ecap, bool escalar,
                       int ebits, size_t n, uint32_t* out, size_t ocap) {
  const MontParams<N> M = c->P.params<N>();
  if (n == 0) return VMX_OK;
  if (ebits == 0) {  // everything to the power 0: fill with one
    VMX_LAUNCH(c, k_gather, nblocks(n * (N / 4), 256), 256, 0, reinterpret_cast<const uint4*>(c->P.consts), (size_t)4,
               reinterpret_cast<uint4*>(out), ocap, n, N / 4, (const uint32_t*)nullptr, (const uint32_t*)nullptr,
               (long long)0, (long long)0, 1, (size_t)1);
    VMX_CHECK_LAUNCH();
    return VMX_OK;
  }
#ifndef VMX_HOST_EMUL
  if (n <= c->coop_max) {  // one warp per element: low latency, fills the machine with few elements
    VMX_LAUNCH(c, k_coop_exp<N>, nblocks(n, kCoopWarps), 32 * kCoopWarps, 0, a, acap, E, ecap, escalar ? 1 : 0, ebits,
               n, c->P.consts, M.n0inv, out, ocap);
    VMX_CHECK_LAUNCH();
    c->modmuls += (uint64_t)n * (14 + (uint64_t)ebits + ebits / 4);
    return VMX_OK;
  }
#endif
  const int w = choose_var_window(ebits);
  // bound the per-thread table scratch (2^w entries per element) to ~6 GB, in whole waves
  const size_t wave = wave_threads(c);
  size_t chunk = (size_t)(6e9 / ((double)(1u << w) * N * 4));
  chunk = std::max(wave, chunk / wave * wave);
  if (c->var_chunk) chunk = c->var_chunk;
  chunk = std::min(chunk, n);
  ElemBuf tab;
  VMX_TRY(tab.alloc_elems(c, chunk << w));
  const int nwin = (ebits + w - 1) / w;
  for (size_t i0 = 0; i0 < n; i0 += chunk) {
    const size_t m = std::min(chunk, n - i0);
    VMX_LAUNCH(c, k_exp_var<N>, nblocks(m), kThreads, kThreads * N * 4, a + 4 * i0, acap,
               escalar ? E : E + 4 * i0, ecap, escalar ? 1 : 0, ebits, w, m, tab.d(), tab.cap, c->P.consts,
               out + 4 * i0, ocap, M);
    VMX_CHECK_LAUNCH();
    c->modmuls += (uint64_t)m * ((1u << w) - 2 + (uint64_t)(nwin - 1) * (w + 1));
  }
  return VMX_OK;
}

// out[i] = a[i]^{X[0]} * b[i]^{Y[i]} (k_exp_var2); arrays too small to fill the machine with one thread per element
// take the warp-cooperative path of exp_var_run twice.
template <int N>
static int exp_var2_run(vmx_ctx* c, const uint32_t* a, size_t acap, const uint32_t* X, size_t xcap, int xbits,
                        const uint32_t* b, size_t bcap, const uint32_t* Y, size_t ycap, int ybits, size_t n,
                        uint32_t* out, size_t ocap) {
  const MontParams<N> M = c->P.params<N>();
  if (n == 0) return VMX_OK;
  bool split = xbits == 0 || ybits == 0;
#ifndef VMX_HOST_EMUL
  split = split || n <= c->coop_max;
#endif
  if (split) {
    ElemBuf t;
    VMX_TRY(t.alloc_elems(c, n));
    VMX_TRY(exp_var_run<N>(c, a, acap, X, xcap, true, xbits, n, t.d(), t.cap));
    VMX_TRY(exp_var_run<N>(c, b, bcap, Y, ycap, false, ybits, n, out, ocap));
    VMX_LAUNCH(c, k_mul<N>, nblocks(n), kThreads, 0, t.d(), t.cap, out, ocap, out, ocap, n, M);
    VMX_CHECK_LAUNCH();
    c->modmuls += n;
    return VMX_OK;
  }
  int w = 1;
  {
    double best = 1e300;
    for (int ww = 1; ww <= 6; ww++) {
      const double cost = 2.0 * (1 << ww) + std::max(xbits, ybits) + (double)(xbits + ybits) / ww;
      if (cost < best) { best = cost; w = ww; }
    }
  }
  const size_t wave = wave_threads(c);
  size_t chunk = (size_t)(6e9 / (2.0 * (double)(1u << w) * N * 4));
  chunk = std::max(wave, chunk / wave * wave);
  if (c->var_chunk) chunk = c->var_chunk;
  chunk = std::min(chunk, n);
  ElemBuf tabA, tabB;
  VMX_TRY(tabA.alloc_elems(c, chunk << w));
  VMX_TRY(tabB.alloc_elems(c, chunk << w));
  const int nwin = (std::max(xbits, ybits) + w - 1) / w, nwx = (xbits + w - 1) / w;
  for (size_t i0 = 0; i0 < n; i0 += chunk) {
    const size_t m = std::min(chunk, n - i0);
    VMX_LAUNCH(c, k_exp_var2<N>, nblocks(m), kThreads, kThreads * N * 4, a + 4 * i0, acap, X, xcap, xbits, b + 4 * i0,
               bcap, Y + 4 * i0, ycap, ybits, w, m, tabA.d(), tabB.d(), tabA.cap, c->P.consts, out + 4 * i0, ocap, M);
    VMX_CHECK_LAUNCH();
    c->modmuls += (uint64_t)m * (2 * ((1u << w) - 2) + (uint64_t)(nwin - 1) * (w + 1) + nwx);
  }
  return VMX_OK;
}

// ------------------------------------------------------------------ Pippenger
struct MexpPlan {
  int c = 0, W = 0, J = 0;
  size_t nb = 0;       // W << c bucket segments
  DevBuf seg_off, idx; // bucket accumulation lists
  DevBuf seg2_off, idx2;  // static sub-digit lists
  size_t nseg2 = 0, total2 = 0;
};

static int choose_mexp_window(size_t n, int L) {
  int bc = 4;
  double best = 1e300;
  for (int c = 4; c <= 16; c += 4) {
    const double W = (L + c - 1) / c;
    const double cost = W * ((double)n + (double)(1u << c) * (c / 4) + 60.0 * 30);  // + latency-ish term
    if (cost < best) { best = cost; bc = c; }
  }
  return bc;
}

template <int N>
static int mexp_plan(vmx_ctx* c, const vmx_rarr* e, int L, MexpPlan& P) {
  const size_t n = e->n;
  P.c = c->mexp_window ? c->mexp_window : choose_mexp_window(n, L);
  P.W = (L + P.c - 1) / P.c;
  P.J = P.c / kSubDigit;
  P.nb = (size_t)P.W << P.c;
  const size_t items = n * (size_t)P.W;
  if (items >= 0xffffffffull) { set_error("expProd too large"); return VMX_ESIZE; }
  VMX_TRY(P.seg_off.alloc(c, (P.nb + 1) * 4));
  VMX_TRY(P.idx.alloc(c, items * 4));
  VMX_CU(cudaMemsetAsync(P.seg_off.p, 0, (P.nb + 1) * 4, c->stream));
  VMX_LAUNCH(c, k_digit_hist<N>, nblocks(items, 256), 256, 0, e->d, e->cap, n, P.c, P.W, P.seg_off.as<uint32_t>());
  VMX_CHECK_LAUNCH();
  VMX_TRY(exclusive_scan(c, P.seg_off.as<uint32_t>(), P.nb + 1));
  DevBuf cursor;
  VMX_TRY(cursor.alloc(c, (P.nb + 1) * 4));
  VMX_CU(cudaMemcpyAsync(cursor.p, P.seg_off.p, (P.nb + 1) * 4, cudaMemcpyDeviceToDevice, c->stream));
  VMX_LAUNCH(c, k_digit_scatter<N>, nblocks(items, 256), 256, 0, e->d, e->cap, n, P.c, P.W, cursor.as<uint32_t>(),
             P.idx.as<uint32_t>());
  VMX_CHECK_LAUNCH();
  // static sub-digit lists
  const size_t cnt = (size_t)1 << (P.c - kSubDigit);
  P.nseg2 = (size_t)P.W * P.J * kSubVals;
  P.total2 = P.nseg2 * cnt;
  VMX_TRY(P.seg2_off.alloc(c, (P.nseg2 + 1) * 4));
  VMX_TRY(P.idx2.alloc(c, P.total2 * 4));
  VMX_LAUNCH(c, k_uniform_offsets, nblocks(P.nseg2 + 1, 256), 256, 0, P.seg2_off.as<uint32_t>(), P.nseg2,
             (uint32_t)cnt);
  VMX_CHECK_LAUNCH();
  VMX_LAUNCH(c, k_subdigit_lists, P.nseg2, 128, 0, P.c, P.J, P.idx2.as<uint32_t>());
  VMX_CHECK_LAUNCH();
  return VMX_OK;
}

// Terms per thread in the bucket accumulation.  The threads of a warp finish together with the longest chunk
// among them, so short chunks (a bucket of ~24 terms cut in 3 x 8) keep the lanes balanced at the price of a
// second, much smaller round over the partial products.
static int mexp_chunk() {
  static const int k = [] { const char* e = std::getenv("VMX_MEXP_K"); const int v = e ? std::atoi(e) : 0; return v >= 2 ? v : 8; }();
  return k;
}

// Column `col` of an expProd: buckets, sub-digit products and the weighted sums Y[col * ngroups + g] (Montgomery
// form); mexp_horner then folds the columns' Y into the results.
template <int N>
static int mexp_run(vmx_ctx* c, const MexpPlan& P, const vmx_garr* a, size_t n_terms, uint32_t* Yall, size_t ycap,
                    size_t col) {
  const MontParams<N> M = c->P.params<N>();
  ElemBuf buckets, X, R;
  VMX_TRY(buckets.alloc_elems(c, P.nb));
  VMX_TRY(seg_product<N>(c, c->P, a->d, a->cap, P.idx.as<uint32_t>(), P.seg_off.as<uint32_t>(), P.nb,
                         n_terms * (size_t)P.W, mexp_chunk(), buckets.d(), buckets.cap));
  VMX_TRY(X.alloc_elems(c, P.nseg2));
  VMX_TRY(seg_product<N>(c, c->P, buckets.d(), buckets.cap, P.idx2.as<uint32_t>(), P.seg2_off.as<uint32_t>(),
                         P.nseg2, P.total2, 8, X.d(), X.cap));
  const size_t ngroups = (size_t)P.W * P.J;
  uint32_t* Y = Yall + 4 * col * ngroups;
#ifndef VMX_HOST_EMUL
  VMX_LAUNCH(c, k_coop_weighted_small<N>, nblocks(ngroups, kCoopWarps), 32 * kCoopWarps, 0, X.d(), X.cap, ngroups,
             Y, ycap, c->P.consts, M.n0inv);
  VMX_CHECK_LAUNCH();
#else
  VMX_TRY(R.alloc_elems(c, ngroups));
  VMX_LAUNCH(c, k_weighted_small<N>, nblocks(ngroups, 32), 32, 0, X.d(), X.cap, ngroups, Y, ycap, R.d(), R.cap, M);
  VMX_CHECK_LAUNCH();
#endif
  c->modmuls += ngroups * 28;
  return VMX_OK;
}

// out[oidx + col] = prod_g Y[col * ngroups + g]^(16^g) for col < k: the k Horner chains run side by side
template <int N>
static int mexp_horner(vmx_ctx* c, const MexpPlan& P, const uint32_t* Yall, size_t ycap, size_t k, uint32_t* out,
                       size_t ocap, size_t oidx) {
  const MontParams<N> M = c->P.params<N>();
  const size_t ngroups = (size_t)P.W * P.J;
#ifndef VMX_HOST_EMUL
  VMX_LAUNCH(c, k_coop_horner<N>, k, 32, 0, Yall, ycap, (int)ngroups, out, ocap, oidx, c->P.consts, M.n0inv);
  VMX_CHECK_LAUNCH();
#else
  for (size_t col = 0; col < k; col++) {
    VMX_LAUNCH(c, k_horner<N>, 1, 32, N * 4, Yall + 4 * col * ngroups, ycap, (int)ngroups, out, ocap, oidx + col, M);
    VMX_CHECK_LAUNCH();
  }
#endif
  c->modmuls += k * (ngroups - 1) * 5;
  return VMX_OK;
}

// ------------------------------------------------------------------ ring helpers
// out[i] = a[i] * (element cidx of carr, Montgomery form) [+ addend[i]]   (mod q)
template <int N>
static int ring_mul_const(vmx_ctx* c, const uint32_t* a, size_t acap, const uint32_t* carr, size_t ccap, size_t cidx,
                          const uint32_t* addend, size_t dcap, uint32_t* out, size_t ocap, size_t n) {
  if (!n) return VMX_OK;
  VMX_LAUNCH(c, k_mul_const<N>, nblocks(n), kThreads, 0, a, acap, carr, ccap, cidx, addend, dcap, out, ocap, n,
             c->Q.params<N>());
  VMX_CHECK_LAUNCH();
  c->modmuls += n;
  return VMX_OK;
}

// total = sum_i a[i] (b == null) or sum_i a[i]*b[i] (canonical), written to out (1 element buf)
template <int N>
static int ring_reduce_sum(vmx_ctx* c, const uint32_t* a, size_t acap, const uint32_t* b, size_t bcap, size_t n,
                           ElemBuf& out) {
  const MontParams<N> M = c->Q.params<N>();
  const int K = 16;
  ElemBuf cur, nxt, tmp;
  const uint32_t* src = a;
  size_t scap = acap, m = n;
  const uint32_t* bb = b;
  bool first = true;
  while (first || m > 1) {
    const size_t nch = (m + K - 1) / K;
    VMX_TRY(nxt.alloc_elems(c, nch));
    if (bb) VMX_TRY(tmp.alloc_elems(c, nch));
    VMX_LAUNCH(c, k_chunk_sum<N>, nblocks(nch), kThreads, 0, src, scap, bb, bcap, m, K, nxt.d(), nxt.cap,
               bb ? tmp.d() : (uint32_t*)nullptr, bb ? tmp.cap : (size_t)0, M);
    VMX_CHECK_LAUNCH();
    if (bb) c->modmuls += m;
    std::swap(cur.p, nxt.p); std::swap(cur.c, nxt.c); std::swap(cur.cap, nxt.cap); std::swap(cur.granted, nxt.granted);
    src = cur.d(); scap = cur.cap; m = nch; bb = nullptr; first = false;
  }
  VMX_TRY(out.alloc_elems(c, 1));
  if (b) {  // sum of a*b*R^-1 -> multiply by R^2
    VMX_TRY(ring_mul_const<N>(c, cur.d(), cur.cap, c->Q.consts, 4, 0, nullptr, 0, out.d(), out.cap, 1));
  } else {
    VMX_LAUNCH(c, k_gather, nblocks(N / 4, 32), 32, 0, reinterpret_cast<const uint4*>(cur.d()), cur.cap,
               reinterpret_cast<uint4*>(out.d()), out.cap, (size_t)1, N / 4, (const uint32_t*)nullptr,
               (const uint32_t*)nullptr, (long long)0, (long long)0, 0, (size_t)0);
    VMX_CHECK_LAUNCH();
  }
  return VMX_OK;
}

// Affine scan.  eM = e in Montgomery form (n elements).  want_y = 0: out = recLin(b, e);
// want_y = 1: out = prods(e) (b unused).
template <int N>
static int ring_scan(vmx_ctx* c, const uint32_t* eM, size_t ecap, const uint32_t* b, size_t bcap, size_t n,
                     int want_y, uint32_t* out, size_t ocap) {
  const MontParams<N> M = c->Q.params<N>();
  // elements per thread: a level costs ~2K sequential modmuls of latency (45 us each at 3072 bits, one thread
  // per residue) and there are log_K(n) levels: K = 4 minimises K / log K; short residues are not latency bound
  const int K = N >= 64 ? 4 : 32;
  const size_t nch = (n + K - 1) / K;
  if (nch <= 1) {
    VMX_LAUNCH(c, k_scan_phaseB<N>, 1, kThreads, 0, eM, ecap, b, bcap, n, K, (const uint32_t*)nullptr, (size_t)0,
               (const uint32_t*)nullptr, (size_t)0, want_y, out, ocap, M);
    VMX_CHECK_LAUNCH();
    c->modmuls += n;
    return VMX_OK;
  }
  ElemBuf A, B, AM, IA, IB;
  VMX_TRY(A.alloc_elems(c, nch));
  if (!want_y) VMX_TRY(B.alloc_elems(c, nch));
  VMX_LAUNCH(c, k_scan_phaseA<N>, nblocks(nch), kThreads, 0, eM, ecap, want_y ? (const uint32_t*)nullptr : b, bcap, n,
             K, A.d(), A.cap, want_y ? (uint32_t*)nullptr : B.d(), want_y ? (size_t)0 : B.cap, c->Q.consts, M);
  VMX_CHECK_LAUNCH();
  c->modmuls += (want_y ? 1 : 2) * n;
  // chunk maps compose by the same recurrence: IA = prods(A), IB = recLin(B, A)
  VMX_TRY(AM.alloc_elems(c, nch));
  VMX_TRY(ring_mul_const<N>(c, A.d(), A.cap, c->Q.consts, 4, 0, nullptr, 0, AM.d(), AM.cap, nch));
  if (want_y) {
    VMX_TRY(IA.alloc_elems(c, nch));
    VMX_TRY(ring_scan<N>(c, AM.d(), AM.cap, nullptr, 0, nch, 1, IA.d(), IA.cap));
  } else {
    VMX_TRY(IB.alloc_elems(c, nch));
    VMX_TRY(ring_scan<N>(c, AM.d(), AM.cap, B.d(), B.cap, nch, 0, IB.d(), IB.cap));
  }
  VMX_LAUNCH(c, k_scan_phaseB<N>, nblocks(nch), kThreads, 0, eM, ecap, b, bcap, n, K,
             want_y ? IA.d() : (const uint32_t*)nullptr, want_y ? IA.cap : (size_t)0,
             want_y ? (const uint32_t*)nullptr : IB.d(), want_y ? (size_t)0 : IB.cap, want_y, out, ocap, M);
  VMX_CHECK_LAUNCH();
  c->modmuls += n;
  return VMX_OK;
}

// generic plane-wise copy out[dst(i)] = in[src(i)]
static int gather(vmx_ctx* c, const uint32_t* in, size_t icap, uint32_t* out, size_t ocap, size_t n,
                  const uint32_t* src_idx, const uint32_t* dst_idx, long long src_off, long long dst_off, int limbs = 0) {
  if (!n) return VMX_OK;
  const int planes = (limbs ? limbs : c->nl) / 4;
  VMX_LAUNCH(c, k_gather, nblocks(n * planes, 256), 256, 0, reinterpret_cast<const uint4*>(in), icap,
             reinterpret_cast<uint4*>(out), ocap, n, planes, src_idx, dst_idx, src_off, dst_off, 0, (size_t)0);
  VMX_CHECK_LAUNCH();
  return VMX_OK;
}

static int same_ctx(const void* a, const void* b) {
  if (a != b) { set_error("operands belong to different contexts"); return VMX_EARG; }
  return VMX_OK;
}

// stream bytes [offset, offset + nbytes) of PRGHeuristic(SHA-256); *data points at the first one
static int prg_bytes_dev(vmx_ctx* c, const uint8_t* seed, size_t seedlen, uint64_t offset, size_t nbytes, DevBuf& buf,
                         const uint8_t** data) {
  if (!seed || seedlen < 32 || seedlen > 48) { set_error("PRG seed must be 32..48 bytes"); return VMX_EARG; }
  const size_t first = offset / 32, shift = offset % 32;
  const size_t nblk = (shift + nbytes + 31) / 32;
  if (first + nblk >= 0xffffffffull) { set_error("PRG stream too long"); return VMX_ESIZE; }
  PrgSeed s;
  std::memset(&s, 0, sizeof s);
  std::memcpy(s.bytes, seed, seedlen);
  s.len = (int)seedlen;
  VMX_TRY(buf.alloc(c, nblk * 32));
  if (nblk) {
    VMX_LAUNCH(c, k_prg_bytes, nblocks(nblk, 128), 128, 0, s, first, nblk, buf.as<uint8_t>());
    VMX_CHECK_LAUNCH();
  }
  *data = buf.as<uint8_t>() + shift;
  return VMX_OK;
}

#include "vmx_ec.inl"

}  // namespace vmx

using namespace vmx;

// ====================================================================== C ABI
// out[i] = a[i]^-1 for n residues in Montgomery form (kernels_elem.cuh, "batched inversion"): a product tree of
// arity kInvArity down to ONE residue, which the host inverts by the binary extended Euclid (limbs_inv_mod).
constexpr size_t kInvArity = 4;
template <int N>
static int inv_batch(vmx_ctx* c, const uint32_t* a, size_t acap, size_t n, uint32_t* out, size_t ocap) {
  struct Level { const uint32_t* a; size_t acap, n, T; uint32_t* out; size_t ocap; };
  std::vector<Level> lv;
  std::vector<std::unique_ptr<ElemBuf>> bufs;
  const auto P = c->P.params<N>();
  Level cur{a, acap, n, 0, out, ocap};
  ElemBuf* root = nullptr;
  for (;;) {
    cur.T = (cur.n + kInvArity - 1) / kInvArity;
    bufs.emplace_back(new ElemBuf);  // the partial products = the array of the next level
    ElemBuf* part = bufs.back().get();
    VMX_TRY(part->alloc_elems(c, cur.T));
    VMX_LAUNCH(c, k_inv_up<N>, nblocks(cur.T), kThreads, 0, cur.a, cur.acap, cur.n, cur.T, cur.out, cur.ocap, part->d(),
               part->cap, P);
    VMX_CHECK_LAUNCH();
    c->modmuls += cur.n - cur.T;
    lv.push_back(cur);
    if (cur.T == 1) { root = part; break; }
    bufs.emplace_back(new ElemBuf);  // prefixes, then inverses, of the next level
    ElemBuf* o = bufs.back().get();
    VMX_TRY(o->alloc_elems(c, cur.T));
    cur = Level{part->d(), part->cap, cur.T, 0, o->d(), o->cap};
  }
  // root: (x R)^-1 on the host, back to Montgomery form with two multiplications by R^2
  std::vector<uint32_t> img(root->cap * N), x(N), r(N);
  VMX_CU(cudaMemcpyAsync(img.data(), root->p, img.size() * 4, cudaMemcpyDeviceToHost, c->stream));
  VMX_CU(cudaStreamSynchronize(c->stream));
  for (int j = 0; j < N; j++) x[j] = img[((size_t)(j >> 2) * root->cap) * 4 + (j & 3)];
  if (!limbs_inv_mod(r.data(), x.data(), c->P.n, N)) { set_error("inv: an element of the array is not invertible"); return VMX_EFORMAT; }
  image_put(img, root->cap, 0, r.data(), N);
  VMX_CU(cudaMemcpyAsync(root->p, img.data(), img.size() * 4, cudaMemcpyHostToDevice, c->stream));
  VMX_CU(cudaStreamSynchronize(c->stream));  // img is a stack-lifetime buffer
  for (int k = 0; k < 2; k++) {
    VMX_LAUNCH(c, k_mul_const<N>, 1, kThreads, 0, root->d(), root->cap, c->P.consts, (size_t)4, (size_t)0,
               (const uint32_t*)nullptr, (size_t)0, root->d(), root->cap, (size_t)1, P);
    VMX_CHECK_LAUNCH();
  }
  c->modmuls += 2;
  // down: the inverses of level l + 1 are the chunk inverses of level l
  uint32_t* inv = root->d();
  size_t icap = root->cap;
  for (size_t l = lv.size(); l-- > 0;) {
    const Level& L = lv[l];
    VMX_LAUNCH(c, k_inv_down<N>, nblocks(L.T), kThreads, 0, L.a, L.acap, L.n, L.T, inv, icap, L.out, L.ocap, P);
    VMX_CHECK_LAUNCH();
    c->modmuls += 2 * (L.n - L.T);
    inv = L.out;
    icap = L.ocap;
  }
  return VMX_OK;
}

// Initial values of the tuning knobs (vmx_ctx_set_tuning) from the environment, read when a context is created:
// VMX_COOP_MAX, VMX_VAR_CHUNK, VMX_MEXP_WINDOW, VMX_BIG_BLOCK_MIN, VMX_BIG_CACHE_MAX, VMX_TABLE_BUDGET.  The parity
// tests use them to send the oracle-sized protocol transcripts through the kernels, the block recycling and the
// table eviction that production-sized arrays take.
template <typename Ctx>
static void tuning_from_env(Ctx& c) {
  if (const char* e = std::getenv("VMX_COOP_MAX")) c->coop_max = (size_t)std::strtoull(e, nullptr, 10);
  if (const char* e = std::getenv("VMX_VAR_CHUNK")) c->var_chunk = (size_t)std::strtoull(e, nullptr, 10);
  if (const char* e = std::getenv("VMX_BIG_BLOCK_MIN")) c->big_block_min = (size_t)std::strtoull(e, nullptr, 10);
  if (const char* e = std::getenv("VMX_BIG_CACHE_MAX")) c->big_cache_max = (size_t)std::strtoull(e, nullptr, 10);
  if (const char* e = std::getenv("VMX_TABLE_BUDGET")) c->table_budget = (size_t)std::strtoull(e, nullptr, 10);
  if (const char* e = std::getenv("VMX_MEXP_WINDOW")) {
    const int v = std::atoi(e);
    if (v == 4 || v == 8 || v == 12 || v == 16) c->mexp_window = v;
  }
}

extern "C" {

const char* vmx_last_error(void) { return g_err; }
int vmx_version(void) { return 100; }

int vmx_ctx_create_modp(const uint8_t* p_be, const uint8_t* q_be, const uint8_t* g_be, size_t nbytes, int device,
                        vmx_ctx** out) {
  if (!out) return VMX_EARG;
  *out = nullptr;
  if (!p_be || !q_be || !nbytes) { set_error("null modulus"); return VMX_EARG; }
  (void)g_be;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    (void)cudaGetLastError();
    set_error("no CUDA device: this engine has no CPU path");
    return VMX_ECUDA;
  }
  if (device < 0 || device >= ndev) { set_error("device %d out of range (%d devices)", device, ndev); return VMX_EARG; }
  std::unique_ptr<vmx_ctx> c(new vmx_ctx);
  c->device = device;
  uint32_t tmp[kMaxLimbs];
  if (!be_to_limbs(p_be, nbytes, tmp, kMaxLimbs)) { set_error("modulus larger than 3072 bits"); return VMX_EARG; }
  const int pbits = limbs_bits(tmp, kMaxLimbs);
  c->nl = pbits <= 512 ? 16 : pbits <= 1024 ? 32 : pbits <= 2048 ? 64 : 96;
  c->gl = c->nl;
  if (pbits < 64 || !(tmp[0] & 1)) { set_error("modulus must be odd and >= 64 bits"); return VMX_EARG; }
  std::memcpy(c->P.n, tmp, sizeof tmp);
  c->P.bits = pbits;
  c->P.n0inv = neg_inv32(c->P.n[0]);
  if (!be_to_limbs(q_be, nbytes, tmp, kMaxLimbs)) { set_error("bad group order"); return VMX_EARG; }
  std::memcpy(c->Q.n, tmp, sizeof tmp);
  c->Q.bits = limbs_bits(tmp, kMaxLimbs);
  if (c->Q.bits < 8 || !(tmp[0] & 1) || limbs_cmp(c->Q.n, c->P.n, kMaxLimbs) >= 0) {
    set_error("group order must be odd and smaller than the modulus");
    return VMX_EARG;
  }
  c->Q.n0inv = neg_inv32(c->Q.n[0]);
  c->eb = (size_t)pbits / 8 + 1;
  c->rb = (size_t)c->Q.bits / 8 + 1;
  {  // safe prime?  p == 2q + 1
    uint32_t t2[kMaxLimbs];
    uint32_t carry = 0;
    for (int j = 0; j < kMaxLimbs; j++) { const uint32_t v = c->Q.n[j]; t2[j] = (v << 1) | carry; carry = v >> 31; }
    t2[0] |= 1u;
    c->safe_prime = !carry && limbs_cmp(t2, c->P.n, kMaxLimbs) == 0;
  }
  c->pm2.assign(c->P.n, c->P.n + c->nl);
  { uint32_t two[kMaxLimbs] = {2}; limbs_sub(c->pm2.data(), two, c->nl); }
  tuning_from_env(c);
  VMX_CU(cudaSetDevice(device));
#ifndef VMX_HOST_EMUL
  {
    cudaDeviceProp prop;
    VMX_CU(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    if (prop.major < 10) {
      set_error("device %d is sm_%d%d; this engine is built for sm_100a only", device, prop.major, prop.minor);
      return VMX_ECUDA;
    }
    cudaMemPool_t pool;
    VMX_CU(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t thr = ~0ull;  // keep freed blocks cached in the pool
    VMX_CU(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
  }
#endif
  VMX_CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  void* p = nullptr;
  VMX_CU(cudaMallocAsync(&p, sizeof(int) * 4, c->stream));
  c->d_flag = (int*)p;
  VMX_CU(cudaMallocHost(&p, sizeof(int) * 4));
  c->h_flag = (int*)p;
  VMX_TRY(upload_consts(c.get(), c->P));
  VMX_TRY(upload_consts(c.get(), c->Q));
  *out = c.release();
  return VMX_OK;
}

// ECqPGroup over y^2 = x^3 + a x + b mod p with a base point of prime order n (256-bit p and n):
// replaces arithm.ECqPGroup construction (the reference's default groups are NIST curves,
// demo/mixnet/benchmarks/bench_config:33-50).  All values big-endian unsigned, `nbytes` each.
int vmx_ctx_create_ecq(const uint8_t* p_be, const uint8_t* a_be, const uint8_t* b_be, const uint8_t* gx_be,
                       const uint8_t* gy_be, const uint8_t* n_be, size_t nbytes, int device, vmx_ctx** out) {
  if (!out) return VMX_EARG;
  *out = nullptr;
  if (!p_be || !a_be || !b_be || !gx_be || !gy_be || !n_be || !nbytes) { set_error("null curve parameter"); return VMX_EARG; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    (void)cudaGetLastError();
    set_error("no CUDA device: this engine has no CPU path");
    return VMX_ECUDA;
  }
  if (device < 0 || device >= ndev) { set_error("device %d out of range (%d devices)", device, ndev); return VMX_EARG; }
  std::unique_ptr<vmx_ctx> c(new vmx_ctx);
  c->device = device;
  c->kind = 1;
  c->nl = 8;
  c->gl = 16;
  uint32_t p[8], a[8], b[8], q[8];
  if (!be_to_limbs(p_be, nbytes, p, 8) || !be_to_limbs(a_be, nbytes, a, 8) || !be_to_limbs(b_be, nbytes, b, 8) ||
      !be_to_limbs(n_be, nbytes, q, 8)) {
    set_error("curve parameters must fit 256 bits");
    return VMX_EARG;
  }
  const int pbits = limbs_bits(p, 8), qbits = limbs_bits(q, 8);
  if (pbits <= 224 || !(p[0] & 1) || qbits <= 192 || !(q[0] & 1) || limbs_cmp(a, p, 8) >= 0 || limbs_cmp(b, p, 8) >= 0) {
    set_error("unsupported curve: p and n must be odd, p of 225..256 bits, a, b < p");
    return VMX_EARG;
  }
  std::memset(c->P.n, 0, sizeof c->P.n);
  std::memset(c->Q.n, 0, sizeof c->Q.n);
  std::memcpy(c->P.n, p, sizeof p);
  std::memcpy(c->Q.n, q, sizeof q);
  c->P.bits = pbits;
  c->Q.bits = qbits;
  c->P.n0inv = neg_inv32(p[0]);
  c->Q.n0inv = neg_inv32(q[0]);
  c->cb = (size_t)pbits / 8 + 1;
  c->eb = 2 * c->cb;
  c->rb = (size_t)qbits / 8 + 1;
  // constants of the coordinate field
  vmx::EcCurve& E = c->ecc;
  std::memset(&E, 0, sizeof E);
  std::memcpy(E.F.n, p, sizeof p);
  E.F.n0inv = c->P.n0inv;
  static const uint32_t p256[8] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0, 0, 0, 1, 0xffffffffu};
  E.F.solinas = limbs_cmp(p, p256, 8) == 0 ? 1u : 0u;
  auto to_mont = [&](uint32_t* r, const uint32_t* x) {  // r = x * 2^256 mod p
    std::memcpy(r, x, 32);
    for (int s = 0; s < 256; s++) {
      uint32_t cy = 0;
      for (int j = 0; j < 8; j++) { const uint32_t nc = r[j] >> 31; r[j] = (r[j] << 1) | cy; cy = nc; }
      if (cy || limbs_cmp(r, p, 8) >= 0) limbs_sub(r, p, 8);
    }
  };
  uint32_t one[8] = {1};
  to_mont(E.a, a);
  to_mont(E.b, b);
  to_mont(E.one, one);
  to_mont(E.r2, E.one);
  std::memcpy(E.pm2, p, sizeof p);
  { uint32_t two[8] = {2}; limbs_sub(E.pm2, two, 8); }
  c->ec_sqrt_ok = (p[0] & 3) == 3;
  if (c->ec_sqrt_ok) {  // (p + 1) / 4
    uint32_t t[9];
    std::memcpy(t, p, sizeof p);
    t[8] = 0;
    for (int j = 0; j < 9; j++) { if (++t[j] != 0) break; }
    for (int j = 0; j < 8; j++) E.sqe[j] = (t[j] >> 2) | (t[j + 1] << 30);
  }
  { uint32_t t[8], three[8] = {3}; std::memcpy(t, p, sizeof p); limbs_sub(t, three, 8); E.a_minus3 = limbs_cmp(a, t, 8) == 0 ? 1u : 0u; }
  tuning_from_env(c);
  VMX_CU(cudaSetDevice(device));
#ifndef VMX_HOST_EMUL
  {
    cudaDeviceProp prop;
    VMX_CU(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    if (prop.major < 10) {
      set_error("device %d is sm_%d%d; this engine is built for sm_100a only", device, prop.major, prop.minor);
      return VMX_ECUDA;
    }
    cudaMemPool_t pool;
    VMX_CU(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t thr = ~0ull;
    VMX_CU(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
  }
#endif
  VMX_CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  void* ptr = nullptr;
  VMX_CU(cudaMallocAsync(&ptr, sizeof(int) * 4, c->stream));
  c->d_flag = (int*)ptr;
  VMX_CU(cudaMallocHost(&ptr, sizeof(int) * 4));
  c->h_flag = (int*)ptr;
  VMX_TRY(upload_consts(c.get(), c->P));
  VMX_TRY(upload_consts(c.get(), c->Q));
  {  // the base point must be on the curve
    std::vector<uint8_t> g(2 * c->cb, 0);
    if (nbytes > c->cb) { set_error("coordinate width"); return VMX_EARG; }
    std::memcpy(g.data() + (c->cb - nbytes), gx_be, nbytes);
    std::memcpy(g.data() + c->cb + (c->cb - nbytes), gy_be, nbytes);
    ElemBuf gb;
    const int st = ec_upload_one(c.get(), g.data(), gb);
    if (st != VMX_OK) {
      vmx_ctx_destroy(c.release());
      set_error("the base point is not on the curve");
      return st == VMX_EFORMAT ? VMX_EARG : st;
    }
  }
  *out = c.release();
  return VMX_OK;
}

// ---------------------------------------------------------------- pinned host buffers
// Byte-tree payloads cross PCIe by DMA only when the host side is page-locked: a pageable destination costs a
// staged copy plus a page fault per 4 KB (measured: 39 MB leaves of a 3072-bit array, 10 ms pageable, < 1 ms
// pinned) and the copy sits on the context's stream in front of the next kernel.  Page-locking is itself slow,
// so buffers are pooled process-wide by size; `vmx_host_free` never calls into CUDA (it may run on a finaliser
// thread of the caller's runtime), the pool is trimmed on the next allocation instead.
namespace {
struct HostPool {
  std::mutex mu;
  std::multimap<size_t, void*> idle;
  std::map<void*, size_t> live;
  size_t idle_bytes = 0;
  size_t cap_bytes = (size_t)8 << 30;
};
HostPool& host_pool() {
  static HostPool* p = [] {
    HostPool* q = new HostPool;
    if (const char* e = std::getenv("VMX_HOST_POOL_MB")) q->cap_bytes = (size_t)std::strtoull(e, nullptr, 10) << 20;
    return q;
  }();
  return *p;
}
}  // namespace

int vmx_host_alloc(int device, size_t nbytes, void** out) {
  if (!out) return VMX_EARG;
  *out = nullptr;
  if (!nbytes) nbytes = 1;
  nbytes = (nbytes + 4095) & ~(size_t)4095;
  HostPool& hp = host_pool();
  std::vector<void*> trim;
  {
    std::lock_guard<std::mutex> g(hp.mu);
    auto it = hp.idle.find(nbytes);
    if (it != hp.idle.end()) {
      *out = it->second;
      hp.idle.erase(it);
      hp.idle_bytes -= nbytes;
      hp.live[*out] = nbytes;
      return VMX_OK;
    }
    while (hp.idle_bytes + nbytes > hp.cap_bytes && !hp.idle.empty()) {  // largest first
      auto last = std::prev(hp.idle.end());
      hp.idle_bytes -= last->first;
      trim.push_back(last->second);
      hp.idle.erase(last);
    }
  }
  VMX_CU(cudaSetDevice(device));
  for (void* t : trim) cudaFreeHost(t);
  void* p = nullptr;
  if (cudaMallocHost(&p, nbytes) != cudaSuccess || !p) {
    (void)cudaGetLastError();
    set_error("cannot page-lock %zu bytes of host memory", nbytes);
    return VMX_ENOMEM;
  }
  std::lock_guard<std::mutex> g(hp.mu);
  hp.live[p] = nbytes;
  *out = p;
  return VMX_OK;
}

void vmx_host_free(void* p) {
  if (!p) return;
  HostPool& hp = host_pool();
  std::lock_guard<std::mutex> g(hp.mu);
  auto it = hp.live.find(p);
  if (it == hp.live.end()) return;  // not ours (or freed twice): ignore, never crash a finaliser
  hp.idle.emplace(it->second, p);
  hp.idle_bytes += it->second;
  hp.live.erase(it);
}

size_t vmx_host_pool_bytes(void) {
  HostPool& hp = host_pool();
  std::lock_guard<std::mutex> g(hp.mu);
  size_t t = hp.idle_bytes;
  for (auto& kv : hp.live) t += kv.second;
  return t;
}

void vmx_ctx_destroy(vmx_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (auto& kv : c->tables) cudaFreeAsync(kv.second.d, c->stream);
  for (auto& b : c->big_free) cudaFreeAsync(b.p, c->stream);
  c->big_free.clear();
  cudaFreeAsync(c->P.consts, c->stream);
  cudaFreeAsync(c->Q.consts, c->stream);
  cudaFreeAsync(c->d_flag, c->stream);
  cudaStreamSynchronize(c->stream);
  cudaFreeHost(c->h_flag);
  cudaStreamDestroy(c->stream);
  delete c;
}

size_t vmx_ctx_elem_bytes(const vmx_ctx* c) { return c ? c->eb : 0; }
size_t vmx_ctx_ring_bytes(const vmx_ctx* c) { return c ? c->rb : 0; }
int vmx_ctx_sync(vmx_ctx* c) {
  VMX_ENTER(c);
  VMX_CU(cudaStreamSynchronize(c->stream));
  return VMX_OK;
}
void* vmx_ctx_stream(vmx_ctx* c) { return c ? (void*)c->stream : nullptr; }
int vmx_ctx_set_fixed_window(vmx_ctx* c, int w) {
  if (!c || w < 0 || w > 22) return VMX_EARG;
  c->fixed_window = w;
  return VMX_OK;
}
// Host-side helper of the byte-tree reader: are the n records of 5 + width bytes all leaf headers 0x01 || be32(width)?
// One pass over n cache lines (the reader needs it to skip an array arithmetically before the engine imports it).
int vmx_leaves_uniform(const uint8_t* leaves, size_t n, size_t width) {
  if (!leaves && n) return 0;
  const uint8_t h1 = (uint8_t)(width >> 24), h2 = (uint8_t)(width >> 16), h3 = (uint8_t)(width >> 8), h4 = (uint8_t)width;
  const size_t rec = 5 + width;
  unsigned bad = 0;
  for (size_t i = 0; i < n; i++) {
    const uint8_t* h = leaves + i * rec;
    bad |= (unsigned)(h[0] ^ 1) | (unsigned)(h[1] ^ h1) | (unsigned)(h[2] ^ h2) | (unsigned)(h[3] ^ h3) | (unsigned)(h[4] ^ h4);
  }
  return bad == 0;
}

int vmx_ctx_set_tuning(vmx_ctx* c, const char* key, long long value) {
  if (!c || !key || value < 0) return VMX_EARG;
  const std::string k(key);
  if (k == "coop_max") c->coop_max = (size_t)value;
  else if (k == "var_chunk") c->var_chunk = (size_t)value;
  else if (k == "mexp_window") {
    if (value != 0 && (value % kSubDigit != 0 || value > 16)) { set_error("mexp_window must be 0, 4, 8, 12 or 16"); return VMX_EARG; }
    c->mexp_window = (int)value;
  } else if (k == "big_cache_max") c->big_cache_max = (size_t)value;
  else if (k == "big_block_min") {
    // a block is handed back by the size it was granted with: changing the threshold with blocks outstanding
    // would route them to the wrong allocator, so it may only be set on a context that has no arrays yet
    c->big_block_min = (size_t)value;
  }
  else if (k == "table_max_bytes") c->table_max_bytes = (size_t)value;
  else if (k == "table_budget") c->table_budget = (size_t)value;
  else if (k == "fixed_window") return vmx_ctx_set_fixed_window(c, (int)value);
  else { set_error("unknown tuning key %s", key); return VMX_EARG; }
  return VMX_OK;
}
uint64_t vmx_ctx_launch_count(const vmx_ctx* c) { return c ? c->launches.load() : 0; }
uint64_t vmx_ctx_modmul_count(const vmx_ctx* c) { return c ? c->modmuls.load() : 0; }

// ---------------------------------------------------------------- group arrays: I/O
static int garr_check_members(vmx_ctx* c, const vmx_garr* a, int* ok);

static int garr_import(vmx_ctx* c, size_t n, const uint8_t* be, int hdr, int check_membership, vmx_garr** out) {
  if (!out) return VMX_EARG;
  *out = nullptr;
  VMX_ENTER(c);
  if (n && !be) return VMX_EARG;
  if (c->kind == 1) return ec_import(c, n, be, hdr, out);
  vmx_garr* a = nullptr;
  VMX_TRY(new_garr(c, n, &a));
  std::unique_ptr<vmx_garr, void (*)(vmx_garr*)> guard(a, vmx_garr_free);
  if (n) {
    DevBuf raw;
    const size_t bytes = n * (c->eb + hdr);
    VMX_TRY(raw.alloc(c, bytes));
    VMX_CU(cudaMemcpyAsync(raw.p, be, bytes, cudaMemcpyHostToDevice, c->stream));
    VMX_CU(cudaMemsetAsync(c->d_flag, 0, sizeof(int) * 4, c->stream));
    VMX_DISPATCH(c->nl, VMX_LAUNCH(c, k_from_bytes<N>, nblocks(n, codec_threads(c->eb + hdr)), codec_threads(c->eb + hdr), codec_smem(c->eb + hdr), raw.as<uint8_t>(), n, (int)c->eb, hdr, 0,
                                   a->d, a->cap, c->P.consts, c->d_flag, c->P.params<N>()));
    VMX_CHECK_LAUNCH();
    c->modmuls += n;
    VMX_TRY(read_flags(c, 1));
    if (c->h_flag[0]) { set_error("group element out of range or malformed leaf (flags %d)", c->h_flag[0]); return VMX_EFORMAT; }
    if (check_membership) {
      int ok = 0;
      VMX_TRY(garr_check_members(c, a, &ok));
      if (!ok) { set_error("element not in the order-q subgroup"); return VMX_EFORMAT; }
    }
  }
  *out = guard.release();
  return VMX_OK;
}
int vmx_garr_from_bytes(vmx_ctx* c, size_t n, const uint8_t* be, int check_membership, vmx_garr** out) {
  return garr_import(c, n, be, 0, check_membership, out);
}
int vmx_garr_from_leaves(vmx_ctx* c, size_t n, const uint8_t* leaves, int check_membership, vmx_garr** out) {
  return garr_import(c, n, leaves, 5, check_membership, out);
}

static int exp_scalar_limbs(vmx_ctx* c, const vmx_garr* a, const uint32_t* x, vmx_garr** out);

// element i = (t_i mod p)^((p-1)/q), t_i = i-th `width`-byte big-endian integer of the DEVICE
// buffer d_raw, masked to `bitlen` bits
static int garr_from_raw_dev(vmx_ctx* c, size_t n, const uint8_t* d_raw, size_t width, unsigned bitlen, vmx_garr** out) {
  // cofactor (p-1)/q must be a small integer: find it by repeated addition of q (setup-size host work)
  uint32_t cof = 0;
  {
    std::vector<uint32_t> acc(c->nl + 1, 0), pm1(c->P.n, c->P.n + c->nl);
    pm1.push_back(0);
    pm1[0] -= 1;  // p is odd
    for (cof = 1; cof <= 1u << 20; cof++) {
      uint64_t carry = 0;
      for (int j = 0; j <= c->nl; j++) {
        const uint64_t s = (uint64_t)acc[j] + (j < c->nl ? c->Q.n[j] : 0) + carry;
        acc[j] = (uint32_t)s; carry = s >> 32;
      }
      if (acc == pm1) break;
    }
    if (cof > 1u << 20) { set_error("from_raw: cofactor (p-1)/q is not a small integer"); return VMX_EARG; }
  }
  vmx_garr* t = nullptr;
  VMX_TRY(new_garr(c, n, &t));
  std::unique_ptr<vmx_garr, void (*)(vmx_garr*)> guard(t, vmx_garr_free);
  if (!n) { *out = guard.release(); return VMX_OK; }
  {
    ElemBuf can;
    VMX_TRY(can.alloc_elems(c, n));
    VMX_DISPATCH(c->nl, {
      VMX_LAUNCH(c, k_ring_from_raw<N>, nblocks(n, codec_threads(width)), codec_threads(width), codec_smem(width), d_raw, n, (int)width, (int)bitlen, can.d(),
                 can.cap, c->P.consts, 1, c->P.params<N>());
      VMX_CHECK_LAUNCH();
      // canonical -> Montgomery form
      VMX_LAUNCH(c, k_mul_const<N>, nblocks(n), kThreads, 0, can.d(), can.cap, c->P.consts, (size_t)4, (size_t)0,
                 (const uint32_t*)nullptr, (size_t)0, t->d, t->cap, n, c->P.params<N>());
      VMX_CHECK_LAUNCH();
    });
    c->modmuls += 4 * n;
  }
  uint32_t x[kMaxLimbs] = {cof};
  return exp_scalar_limbs(c, t, x, out);
}

int vmx_garr_from_raw(vmx_ctx* c, size_t n, const uint8_t* be, size_t width, unsigned bitlen, vmx_garr** out) {
  if (!out) return VMX_EARG;
  *out = nullptr;
  VMX_ENTER(c);
  if (c->kind == 1) { set_error("curve groups draw random points with vmx_garr_from_candidates"); return VMX_EARG; }
  if ((n && !be) || !width || width > (size_t)16 * c->nl) { set_error("from_raw: bad width %zu", width); return VMX_EARG; }
  if (bitlen > 8 * width) bitlen = 0;
  DevBuf raw;
  VMX_TRY(raw.alloc(c, n * width));
  if (n) {
    VMX_CU(cudaMemcpyAsync(raw.p, be, n * width, cudaMemcpyHostToDevice, c->stream));
    VMX_CU(cudaStreamSynchronize(c->stream));  // `be` is borrowed for the call only
  }
  return garr_from_raw_dev(c, n, raw.as<uint8_t>(), width, bitlen, out);
}

// PRGHeuristic(SHA-256) output bytes [offset, offset + nbytes) to the host: the expansion of a long
// request (Permutation.random draws size * ~15 bytes, mixnet/ShufflerElGamalSession.java:408-409) runs
// in counter mode on the device instead of one hash call per 32 bytes on the host.
int vmx_prg_bytes_sha256(vmx_ctx* c, const uint8_t* seed, size_t seedlen, uint64_t offset, size_t nbytes, uint8_t* out) {
  VMX_ENTER(c);
  if (nbytes && !out) return VMX_EARG;
  if (!nbytes) return VMX_OK;
  DevBuf buf;
  const uint8_t* data = nullptr;
  VMX_TRY(prg_bytes_dev(c, seed, seedlen, offset, nbytes, buf, &data));
  VMX_CU(cudaMemcpyAsync(out, data, nbytes, cudaMemcpyDeviceToHost, c->stream));
  VMX_CU(cudaStreamSynchronize(c->stream));
  return VMX_OK;
}

int vmx_garr_prg_sha256(vmx_ctx* c, const uint8_t* seed, size_t seedlen, uint64_t offset, size_t n, size_t width,
                        unsigned bitlen, vmx_garr** out) {
  if (!out) return VMX_EARG;
  *out = nullptr;
  VMX_ENTER(c);
  if (!width || width > (size_t)16 * c->nl) { set_error("prg: bad width %zu", width); return VMX_EARG; }
  if (bitlen > 8 * width) bitlen = 0;
  if (c->kind == 1) return ec_random_prg(c, seed, seedlen, offset, n, width, bitlen, out);
  DevBuf raw;
  const uint8_t* data = nullptr;
  VMX_TRY(prg_bytes_dev(c, seed, seedlen, offset, n * width, raw, &data));
  c->prg_consumed = (uint64_t)n * width;
  return garr_from_raw_dev(c, n, data, width, bitlen, out);
}

uint64_t vmx_ctx_prg_consumed(const vmx_ctx* c) { return c ? c->prg_consumed : 0; }

// Ranks of n random keys held in DEVICE memory (see kernels_perm.cuh).  *fallback = 1: not ranked (degenerate
// key distribution or colliding leading words); the caller ranks on the host.
static int permutation_dev(vmx_ctx* c, const uint8_t* d_raw, size_t n, size_t nbytes, unsigned bits, uint32_t* table_out,
                           int* fallback) {
  *fallback = 0;
  if (!n) return VMX_OK;
  if (n >= 0xffffffffull) { set_error("permutation too large"); return VMX_ESIZE; }
  const int lbits = nbytes >= 8 ? (int)(bits - 8 * (nbytes - 8)) : (int)bits;  // significant bits of the lead word
  int B = 1;
  while (((size_t)1 << (B + 1)) * 4 <= n && B + 1 < 28) B++;
  if (B > lbits) B = lbits;
  const size_t nbuckets = (size_t)1 << B;
  DevBuf off, cursor, keys, table;
  VMX_TRY(off.alloc(c, (nbuckets + 1) * 4));
  VMX_TRY(cursor.alloc(c, (nbuckets + 1) * 4));
  VMX_TRY(keys.alloc(c, n * sizeof(PermKey)));
  VMX_TRY(table.alloc(c, n * 4));
  VMX_CU(cudaMemsetAsync(off.p, 0, (nbuckets + 1) * 4, c->stream));
  VMX_LAUNCH(c, k_perm_hist, nblocks(n, 256), 256, 0, d_raw, n, (int)nbytes, (int)bits, lbits, B, off.as<uint32_t>());
  VMX_CHECK_LAUNCH();
  VMX_TRY(exclusive_scan(c, off.as<uint32_t>(), nbuckets + 1));
  VMX_CU(cudaMemcpyAsync(cursor.p, off.p, (nbuckets + 1) * 4, cudaMemcpyDeviceToDevice, c->stream));
  VMX_LAUNCH(c, k_perm_scatter, nblocks(n, 256), 256, 0, d_raw, n, (int)nbytes, (int)bits, lbits, B,
             cursor.as<uint32_t>(), keys.as<PermKey>());
  VMX_CHECK_LAUNCH();
  VMX_CU(cudaMemsetAsync(c->d_flag, 0, sizeof(int) * 4, c->stream));
  VMX_LAUNCH(c, k_perm_rank, nblocks(nbuckets, 128), 128, 0, off.as<uint32_t>(), nbuckets, keys.as<PermKey>(),
             nbytes > 8 ? 1 : 0, table.as<uint32_t>(), c->d_flag);
  VMX_CHECK_LAUNCH();
  VMX_TRY(read_flags(c, 1));
  if (c->h_flag[0]) { *fallback = 1; return VMX_OK; }
  VMX_CU(cudaMemcpyAsync(table_out, table.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
  VMX_CU(cudaStreamSynchronize(c->stream));
  return VMX_OK;
}

int vmx_permutation_prg_sha256(vmx_ctx* c, const uint8_t* seed, size_t seedlen, uint64_t offset, size_t n, size_t nbytes,
                               unsigned bits, uint32_t* table_out, int* fallback) {
  VMX_ENTER(c);
  if (!fallback || (n && !table_out) || !nbytes || bits > 8 * nbytes || bits + 7 < 8 * nbytes || !bits) {
    set_error("permutation: bad key width");
    return VMX_EARG;
  }
  DevBuf raw;
  const uint8_t* data = nullptr;
  VMX_TRY(prg_bytes_dev(c, seed, seedlen, offset, n * nbytes, raw, &data));
  return permutation_dev(c, data, n, nbytes, bits, table_out, fallback);
}

int vmx_permutation_from_raw(vmx_ctx* c, size_t n, const uint8_t* be, size_t nbytes, unsigned bits, uint32_t* table_out,
                             int* fallback) {
  VMX_ENTER(c);
  if (!fallback || (n && (!table_out || !be)) || !nbytes || bits > 8 * nbytes || bits + 7 < 8 * nbytes || !bits) {
    set_error("permutation: bad key width");
    return VMX_EARG;
  }
  DevBuf raw;
  VMX_TRY(raw.alloc(c, n * nbytes));
  if (n) VMX_CU(cudaMemcpyAsync(raw.p, be, n * nbytes, cudaMemcpyHostToDevice, c->stream));
  return permutation_dev(c, raw.as<uint8_t>(), n, nbytes, bits, table_out, fallback);
}

int vmx_garr_from_candidates(vmx_ctx* c, size_t m, const uint8_t* be, size_t width, unsigned bitlen, size_t n_want,
                             vmx_garr** out, size_t* used) {
  if (!out || !used) return VMX_EARG;
  *out = nullptr;
  *used = 0;
  VMX_ENTER(c);
  if (c->kind != 1) { set_error("from_candidates: not a curve group"); return VMX_EARG; }
  if ((m && !be) || !width || width > 64) { set_error("from_candidates: bad width %zu", width); return VMX_EARG; }
  if (bitlen > 8 * width) bitlen = 0;
  vmx_garr* full = nullptr;
  VMX_TRY(new_garr(c, n_want, &full));
  std::unique_ptr<vmx_garr, void (*)(vmx_garr*)> guard(full, vmx_garr_free);
  size_t acc = 0, u = 0;
  if (m && n_want) {
    DevBuf raw;
    VMX_TRY(raw.alloc(c, m * width));
    VMX_CU(cudaMemcpyAsync(raw.p, be, m * width, cudaMemcpyHostToDevice, c->stream));
    VMX_TRY(ec_candidates(c, raw.as<uint8_t>(), m, width, bitlen, full, 0, &acc, &u));
  }
  *used = n_want ? u : 0;
  if (acc == n_want) { *out = guard.release(); return VMX_OK; }
  vmx_garr* part = nullptr;  // fewer accepted than wanted: hand back the accepted prefix
  VMX_TRY(new_garr(c, acc, &part));
  const int st = gather(c, full->d, full->cap, part->d, part->cap, acc, nullptr, nullptr, 0, 0, c->gl);
  if (st != VMX_OK) { vmx_garr_free(part); return st; }
  *out = part;
  return VMX_OK;
}

int vmx_rarr_prg_raw_sha256(vmx_ctx* c, const uint8_t* seed, size_t seedlen, uint64_t offset, size_t n, size_t width,
                            unsigned bitlen, vmx_rarr** out) {
  if (!out) return VMX_EARG;
  *out = nullptr;
  VMX_ENTER(c);
  if (!width || width > (size_t)16 * c->nl) { set_error("prg: bad width %zu", width); return VMX_EARG; }
  if (bitlen > 8 * width) bitlen = 0;
  DevBuf raw;
  const uint8_t* data = nullptr;
  VMX_TRY(prg_bytes_dev(c, seed, seedlen, offset, n * width, raw, &data));
  vmx_rarr* a = nullptr;
  VMX_TRY(new_rarr(c, n, &a));
  std::unique_ptr<vmx_rarr, void (*)(vmx_rarr*)> guard(a, vmx_rarr_free);
  if (n) {
    const int totalbits = bitlen ? (int)bitlen : (int)(8 * width);
    const int need_reduce = totalbits >= c->Q.bits;
    VMX_DISPATCH(c->nl, VMX_LAUNCH(c, k_ring_from_raw<N>, nblocks(n, codec_threads(width)), codec_threads(width), codec_smem(width), data, n, (int)width,
                                   (int)bitlen, a->d, a->cap, c->Q.consts, need_reduce, c->Q.params<N>()));
    VMX_CHECK_LAUNCH();
    if (need_reduce) c->modmuls += 3 * n;
  }
  *out = guard.release();
  return VMX_OK;
}

static int garr_export(const vmx_garr* a, int hdr, uint8_t* be_out) {
  if (!a) return VMX_EARG;
  vmx_ctx* c = a->ctx;
  VMX_ENTER(c);
  if (c->kind == 1) return be_out ? ec_export(a, hdr, be_out) : VMX_EARG;
  if (!a->n) return VMX_OK;
  if (!be_out) return VMX_EARG;
  DevBuf raw;
  const size_t bytes = a->n * (c->eb + hdr);
  VMX_TRY(raw.alloc(c, bytes));
  VMX_DISPATCH(c->nl, VMX_LAUNCH(c, k_to_bytes<N>, nblocks(a->n, codec_threads(c->eb + hdr)), codec_threads(c->eb + hdr), codec_smem(c->eb + hdr), a->d, a->cap, a->n, (int)c->eb, hdr, 0,
                                 raw.as<uint8_t>(), c->P.params<N>()));
  VMX_CHECK_LAUNCH();
  c->modmuls += a->n;
  VMX_CU(cudaMemcpyAsync(be_out, raw.p, bytes, cudaMemcpyDeviceToHost, c->stream));
  VMX_CU(cudaStreamSynchronize(c->stream));
  return VMX_OK;
}
int vmx_garr_to_bytes(const vmx_garr* a, uint8_t* be_out) { return garr_export(a, 0, be_out); }
int vmx_garr_to_leaves(const vmx_garr* a, uint8_t* leaves_out) { return garr_export(a, 5, leaves_out); }

int vmx_garr_fill(vmx_ctx* c, size_t n, const uint8_t* elem_be, vmx_garr** out) {
  if (!out) return VMX_EARG;
  *out = nullptr;
  VMX_ENTER(c);
  ElemBuf one;
  VMX_TRY(upload_one(c, elem_be, true, one));
  vmx_garr* a = nullptr;
  VMX_TRY(new_garr(c, n, &a));
  if (n) {
    const int planes = c->gl / 4;
    VMX_LAUNCH(c, k_gather, nblocks(n * planes, 256), 256, 0, reinterpret_cast<const uint4*>(one.d()), one.cap,
               reinterpret_cast<uint4*>(a->d), a->cap, n, planes, (const uint32_t*)nullptr, (const uint32_t*)nullptr,
               (long long)0, (long long)0, 1, (size_t)0);
    if (cudaGetLastError() != cudaSuccess) { vmx_garr_free(a); set_error("fill launch failed"); return VMX_ECUDA; }
  }
  *out = a;
  return VMX_OK;
}

void vmx_garr_free(vmx_garr* a) {
  if (!a) return;
  cudaSetDevice(a->ctx->device);
  if (a->d) dev_free(a->ctx, a->d, a->granted);
  delete a;
}
size_t vmx_garr_size(const vmx_garr* a) { return a ? a->n : 0; }

// ---------------------------------------------------------------- group arrays: algebra
int vmx_exp_fixed(vmx_ctx* c, const uint8_t* base_be, const vmx_rarr* e, vmx_garr** out) {
  if (!out) return VMX_EARG;
  *out = nullptr;
  VMX_ENTER(c);
  if (!e || !base_be) return VMX_EARG;
  VMX_TRY(same_ctx(e->ctx, c));
  FixedTable T;
  VMX_TRY(get_table(c, base_be, e->n, &T));
  int ebits = 0;
  VMX_TRY(rarr_bitlen(e, &ebits));
  vmx_garr* r = nullptr;
  VMX_TRY(new_garr(c, e->n, &r));
  std::unique_ptr<vmx_garr, void (*)(vmx_garr*)> guard(r, vmx_garr_free);
  if (c->kind == 1) VMX_TRY(ec_exp_fixed_run(c, T, e, ebits, r->d, r->cap));
  else if (e->n) VMX_DISPATCH(c->nl, VMX_TRY(exp_fixed_run<N>(c, T, e, ebits, r->d, r->cap)));
  *out = guard.release();
  return VMX_OK;
}

// PGroupElement.exp(PRingElement) for ONE element: through the cached window table of the base
// when there is one (g, h0, the public key), else on the warp-cooperative multiplier.
int vmx_elem_exp(vmx_ctx* c, const uint8_t* base_be, const uint8_t* e_be, uint8_t* out_be) {
  VMX_ENTER(c);
  if (!base_be || !e_be || !out_be) return VMX_EARG;
  bool have = false;
  FixedTable T;
  {
    std::lock_guard<std::mutex> lk(c->mu);
    auto it = c->tables.find(std::string(reinterpret_cast<const char*>(base_be), c->eb));
    if (it != c->tables.end()) { T = it->second; have = true; }
  }
  vmx_rarr* e = nullptr;
  VMX_TRY(vmx_rarr_from_bytes(c, 1, e_be, &e));
  std::unique_ptr<vmx_rarr, void (*)(vmx_rarr*)> ge(e, vmx_rarr_free);
  ElemBuf res;
  VMX_TRY(res.alloc_gelems(c, 1));
  int ebits = 0;
  VMX_TRY(rarr_bitlen(e, &ebits));
  if (c->kind == 1) {
    if (have) {
      VMX_TRY(ec_exp_fixed_run(c, T, e, ebits, res.d(), res.cap));
    } else {
      ElemBuf base;
      VMX_TRY(upload_one(c, base_be, true, base));
      VMX_TRY(ec_exp_var_run(c, base.d(), base.cap, e->d, e->cap, true, ebits, 1, res.d(), res.cap));
    }
  } else if (have) {
    VMX_DISPATCH(c->nl, VMX_TRY(exp_fixed_run<N>(c, T, e, ebits, res.d(), res.cap)));
  } else {
    ElemBuf base;
    VMX_TRY(upload_one(c, base_be, true, base));
    VMX_DISPATCH(c->nl, VMX_TRY(exp_var_run<N>(c, base.d(), base.cap, e->d, e->cap, true, ebits, 1, res.d(), res.cap)));
  }
  return download_one(c, res.d(), res.cap, 0, true, out_be);
}

// PGroupElement.inv() of ONE element (the divisions C = prod u / prod h and D = B_{N-1} / h0^{prod e},
// hvzk/PoSBasicTW.java:1013-1014): binary extended Euclid on the host, ~1 ms, instead of a
// 3072-step Fermat ladder on one warp (~20 ms).
int vmx_elem_inv(vmx_ctx* c, const uint8_t* in_be, uint8_t* out_be) {
  VMX_ENTER(c);
  if (!in_be || !out_be) return VMX_EARG;
  if (c->kind == 1) return ec_elem_inv(c, in_be, out_be);
  const int N = c->nl;
  std::vector<uint32_t> a(N), r(N);
  if (!be_to_limbs(in_be, c->eb, a.data(), N) || limbs_cmp(a.data(), c->P.n, N) >= 0) {
    set_error("element out of range");
    return VMX_EFORMAT;
  }
  if (!limbs_inv_mod(r.data(), a.data(), c->P.n, N)) { set_error("element not invertible"); return VMX_EFORMAT; }
  std::memset(out_be, 0, c->eb);
  for (size_t b = 0; b < (size_t)4 * N && b < c->eb; b++) out_be[c->eb - 1 - b] = (uint8_t)(r[b / 4] >> (8 * (b % 4)));
  return VMX_OK;
}

int vmx_fixed_precompute(vmx_ctx* c, const uint8_t* base_be, size_t n_hint) {
  VMX_ENTER(c);
  if (!base_be) { set_error("null base"); return VMX_EARG; }
  FixedTable T;
  return get_table(c, base_be, n_hint, &T);
}

int vmx_exp_var(const vmx_garr* a, const vmx_rarr* e, vmx_garr** out) {
  if (!out) return VMX_EARG;
  *out = nullptr;
  if (!a || !e) return VMX_EARG;
  vmx_ctx* c = a->ctx;
  VMX_ENTER(c);
  VMX_TRY(same_ctx(e->ctx, c));
  if (a->n != e->n) { set_error("exp: size mismatch %zu vs %zu", a->n, e->n); return VMX_ESIZE; }
  int ebits = 0;
  VMX_TRY(rarr_bitlen(e, &ebits));
  vmx_garr* r = nullptr;
  VMX_TRY(new_garr(c, a->n, &r));
  std::unique_ptr<vmx_garr, void (*)(vmx_garr*)> guard(r, vmx_garr_free);
  if (c->kind == 1) VMX_TRY(ec_exp_var_run(c, a->d, a->cap, e->d, e->cap, false, ebits, a->n, r->d, r->cap));
  else VMX_DISPATCH(c->nl, VMX_TRY(exp_var_run<N>(c, a->d, a->cap, e->d, e->cap, false, ebits, a->n, r->d, r->cap)));
  *out = guard.release();
  return VMX_OK;
}

int vmx_exp_scalar(const vmx_garr* a, const uint8_t* e_be, vmx_garr** out) {
  if (!out) return VMX_EARG;
  *out = nullptr;
  if (!a || !e_be) return VMX_EARG;
  vmx_ctx* c = a->ctx;
  VMX_ENTER(c);
  ElemBuf e;
  VMX_TRY(upload_one(c, e_be, false, e));
  uint32_t limbs[kMaxLimbs];
  if (!be_to_limbs(e_be, c->rb, limbs, kMaxLimbs)) return VMX_EFORMAT;
  const int ebits = limbs_bits(limbs, kMaxLimbs);
  vmx_garr* r = nullptr;
  VMX_TRY(new_garr(c, a->n, &r));
  std::unique_ptr<vmx_garr, void (*)(vmx_garr*)> guard(r, vmx_garr_free);
  if (c->kind == 1) VMX_TRY(ec_exp_var_run(c, a->d, a->cap, e.d(), e.cap, true, ebits, a->n, r->d, r->cap));
  else VMX_DISPATCH(c->nl, VMX_TRY(exp_var_run<N>(c, a->d, a->cap, e.d(), e.cap, true, ebits, a->n, r->d, r->cap)));
  *out = guard.release();
  return VMX_OK;
}

int vmx_exp_scalar_var(const vmx_garr* a, const uint8_t* x_be, const vmx_garr* b, const vmx_rarr* y, vmx_garr** out) {
  if (!out) return VMX_EARG;
  *out = nullptr;
  if (!a || !x_be || !b || !y) return VMX_EARG;
  vmx_ctx* c = a->ctx;
  VMX_ENTER(c);
  VMX_TRY(same_ctx(b->ctx, c));
  VMX_TRY(same_ctx(y->ctx, c));
  if (a->n != b->n || a->n != y->n) { set_error("exp: size mismatch %zu, %zu, %zu", a->n, b->n, y->n); return VMX_ESIZE; }
  ElemBuf x;
  VMX_TRY(upload_one(c, x_be, false, x));
  uint32_t limbs[kMaxLimbs];
  if (!be_to_limbs(x_be, c->rb, limbs, kMaxLimbs)) return VMX_EFORMAT;
  const int xbits = limbs_bits(limbs, kMaxLimbs);
  int ybits = 0;
  VMX_TRY(rarr_bitlen(y, &ybits));
  vmx_garr* r = nullptr;
  VMX_TRY(new_garr(c, a->n, &r));
  std::unique_ptr<vmx_garr, void (*)(vmx_garr*)> guard(r, vmx_garr_free);
  if (c->kind == 1) {
    VMX_TRY(ec_exp_var2_run(c, a->d, a->cap, x.d(), x.cap, xbits, b->d, b->cap, y->d, y->cap, ybits, a->n, r->d, r->cap));
    *out = guard.release();
    return VMX_OK;
  }
  VMX_DISPATCH(c->nl, VMX_TRY(exp_var2_run<N>(c, a->d, a->cap, x.d(), x.cap, xbits, b->d, b->cap, y->d, y->cap, ybits,
                                              a->n, r->d, r->cap)));
  *out = guard.release();
  return VMX_OK;
}

// a[i]^x for a host-side limb exponent (internal: membership test, inversion)
static int exp_scalar_limbs(vmx_ctx* c, const vmx_garr* a, const uint32_t* x, vmx_garr** out) {
  std::vector<uint32_t> img((size_t)8 * c->nl, 0);
  image_put(img, 8, 0, x, c->nl);
  ElemBuf e;
  VMX_TRY(e.alloc_elems(c, 1));
  VMX_CU(cudaMemcpyAsync(e.p, img.data(), img.size() * 4, cudaMemcpyHostToDevice, c->stream));
  VMX_CU(cudaStreamSynchronize(c->stream));  // img is a stack-lifetime buffer
  const int ebits = limbs_bits(x, c->nl);
  vmx_garr* r = nullptr;
  VMX_TRY(new_garr(c, a->n, &r));
  std::unique_ptr<vmx_garr, void (*)(vmx_garr*)> guard(r, vmx_garr_free);
  VMX_DISPATCH(c->nl, VMX_TRY(exp_var_run<N>(c, a->d, a->cap, e.d(), e.cap, true, ebits, a->n, r->d, r->cap)));
  *out = guard.release();
  return VMX_OK;
}

// Euler criterion x^q == 1 for every element (exact; a Jacobi-symbol kernel is the planned
// replacement).
static int garr_check_members(vmx_ctx* c, const vmx_garr* a, int* ok) {
  if (c->safe_prime) {  // quadratic residues: Legendre symbol by the binary Jacobi algorithm
    *ok = 1;
    if (!a->n) return VMX_OK;
    VMX_CU(cudaMemsetAsync(c->d_flag, 0, sizeof(int) * 4, c->stream));
    VMX_DISPATCH(c->nl, VMX_LAUNCH(c, k_jacobi<N>, nblocks(a->n), kThreads, 0, a->d, a->cap, a->n, c->d_flag,
                                   c->P.params<N>()));
    VMX_CHECK_LAUNCH();
    VMX_TRY(read_flags(c, 1));
    *ok = c->h_flag[0] == 0;
    return VMX_OK;
  }
  vmx_garr* t = nullptr;
  VMX_TRY(exp_scalar_limbs(c, a, c->Q.n, &t));
  std::unique_ptr<vmx_garr, void (*)(vmx_garr*)> guard(t, vmx_garr_free);
  VMX_CU(cudaMemsetAsync(c->d_flag, 0, sizeof(int) * 4, c->stream));
  VMX_DISPATCH(c->nl, VMX_LAUNCH(c, k_differs_from_const<N>, nblocks(a->n, 256), 256, 0, t->d, t->cap, a->n,
                                 c->P.consts, (size_t)4, (size_t)1, c->d_flag));
  VMX_CHECK_LAUNCH();
  VMX_TRY(read_flags(c, 1));
  *ok = c->h_flag[0] == 0;
  return VMX_OK;
}

int vmx_expprod(const vmx_garr* const* a, size_t k, const vmx_rarr* e, uint8_t* out_be) {
  if (!a || !k || !e || !out_be) return VMX_EARG;
  vmx_ctx* c = e->ctx;
  VMX_ENTER(c);
  for (size_t l = 0; l < k; l++) {
    if (!a[l]) return VMX_EARG;
    VMX_TRY(same_ctx(a[l]->ctx, c));
    if (a[l]->n != e->n) { set_error("expProd: size mismatch %zu vs %zu", a[l]->n, e->n); return VMX_ESIZE; }
  }
  if (c->kind == 1) return ec_expprod(c, a, k, e, out_be);
  int L = 0;
  VMX_TRY(rarr_bitlen(e, &L));
  ElemBuf res;
  VMX_TRY(res.alloc_elems(c, k));
  if (e->n == 0 || L == 0) {
    for (size_t l = 0; l < k; l++) VMX_TRY(gather(c, c->P.consts, 4, res.d(), res.cap, 1, nullptr, nullptr, 1, (long long)l));
  } else {
    VMX_DISPATCH(c->nl, {
      MexpPlan P;
      VMX_TRY(mexp_plan<N>(c, e, L, P));
      ElemBuf Y;
      VMX_TRY(Y.alloc_elems(c, k * (size_t)P.W * P.J));
      for (size_t l = 0; l < k; l++) VMX_TRY(mexp_run<N>(c, P, a[l], e->n, Y.d(), Y.cap, l));
      VMX_TRY(mexp_horner<N>(c, P, Y.d(), Y.cap, k, res.d(), res.cap, 0));
    });
  }
  DevBuf raw;
  VMX_TRY(raw.alloc(c, k * c->eb));
  VMX_DISPATCH(c->nl, VMX_LAUNCH(c, k_to_bytes<N>, nblocks(k, codec_threads(c->eb)), codec_threads(c->eb), codec_smem(c->eb), res.d(), res.cap, k, (int)c->eb, 0, 0,
                                 raw.as<uint8_t>(), c->P.params<N>()));
  VMX_CHECK_LAUNCH();
  VMX_CU(cudaMemcpyAsync(out_be, raw.p, k * c->eb, cudaMemcpyDeviceToHost, c->stream));
  VMX_CU(cudaStreamSynchronize(c->stream));
  return VMX_OK;
}

int vmx_mul(const vmx_garr* a, const vmx_garr* b, vmx_garr** out) {
  if (!out) return VMX_EARG;
  *out = nullptr;
  if (!a || !b) return VMX_EARG;
  vmx_ctx* c = a->ctx;
  VMX_ENTER(c);
  VMX_TRY(same_ctx(b->ctx, c));
  if (a->n != b->n) { set_error("mul: size mismatch %zu vs %zu", a->n, b->n); return VMX_ESIZE; }
  vmx_garr* r = nullptr;
  VMX_TRY(new_garr(c, a->n, &r));
  if (c->kind == 1) {
    const int st = ec_mul(c, a, b, r);
    if (st != VMX_OK) { vmx_garr_free(r); return st; }
    *out = r;
    return VMX_OK;
  }
  if (a->n) {
    VMX_DISPATCH(c->nl, VMX_LAUNCH(c, k_mul<N>, nblocks(a->n), kThreads, 0, a->d, a->cap, b->d, b->cap, r->d, r->cap,
                                   a->n, c->P.params<N>()));
    if (cudaGetLastError() != cudaSuccess) { vmx_garr_free(r); set_error("mul launch failed"); return VMX_ECUDA; }
    c->modmuls += a->n;
  }
  *out = r;
  return VMX_OK;
}

int vmx_inv(const vmx_garr* a, vmx_garr** out) {
  if (!out) return VMX_EARG;
  *out = nullptr;
  if (!a) return VMX_EARG;
  vmx_ctx* c = a->ctx;
  VMX_ENTER(c);
  if (c->kind == 1) {
    vmx_garr* r = nullptr;
    VMX_TRY(new_garr(c, a->n, &r));
    const int st = ec_neg(c, a, r);
    if (st != VMX_OK) { vmx_garr_free(r); return st; }
    *out = r;
    return VMX_OK;
  }
  vmx_garr* r = nullptr;
  VMX_TRY(new_garr(c, a->n, &r));
  std::unique_ptr<vmx_garr, void (*)(vmx_garr*)> guard(r, vmx_garr_free);
  if (a->n) VMX_DISPATCH(c->nl, VMX_TRY(inv_batch<N>(c, a->d, a->cap, a->n, r->d, r->cap)));
  *out = guard.release();
  return VMX_OK;
}

int vmx_prod(const vmx_garr* a, uint8_t* out_be) {
  if (!a || !out_be) return VMX_EARG;
  vmx_ctx* c = a->ctx;
  VMX_ENTER(c);
  if (c->kind == 1) return ec_prod(c, a, out_be);
  ElemBuf res;
  VMX_TRY(res.alloc_elems(c, 1));
  DevBuf off;
  VMX_TRY(off.alloc(c, 8));
  const uint32_t h[2] = {0, (uint32_t)a->n};
  VMX_CU(cudaMemcpyAsync(off.p, h, 8, cudaMemcpyHostToDevice, c->stream));
  VMX_CU(cudaStreamSynchronize(c->stream));
  // K scales with n so that the first round still fills the machine
  const int K = (int)std::min<size_t>(64, std::max<size_t>(2, a->n / wave_threads(c)));
  VMX_DISPATCH(c->nl, VMX_TRY(seg_product<N>(c, c->P, a->d, a->cap, nullptr, off.as<uint32_t>(), 1, a->n, K, res.d(),
                                             res.cap)));
  return download_one(c, res.d(), res.cap, 0, true, out_be);
}

static int upload_u32(vmx_ctx* c, const uint32_t* h, size_t n, DevBuf& buf) {
  VMX_TRY(buf.alloc(c, n * 4));
  if (n) {
    VMX_CU(cudaMemcpyAsync(buf.p, h, n * 4, cudaMemcpyHostToDevice, c->stream));
    VMX_CU(cudaStreamSynchronize(c->stream));
  }
  return VMX_OK;
}

// rows_dev[r] = element idx[r] (count rows of nl words, element-major); idx is a HOST list or null
static int pack_rows(vmx_ctx* c, const uint32_t* d, size_t cap, size_t n, const uint32_t* idx, size_t count,
                     void* rows_dev, int limbs) {
  if (!count) return VMX_OK;
  if (!rows_dev) return VMX_EARG;
  DevBuf di;
  if (idx) {
    for (size_t r = 0; r < count; r++) if (idx[r] >= n) { set_error("row index out of range"); return VMX_EARG; }
    VMX_TRY(upload_u32(c, idx, count, di));
  } else if (count > n) { set_error("more rows than elements"); return VMX_ESIZE; }
  const int planes = limbs / 4;
  VMX_LAUNCH(c, k_pack_rows, nblocks(count * planes, 256), 256, 0, reinterpret_cast<const uint4*>(d), cap,
             idx ? di.as<uint32_t>() : (const uint32_t*)nullptr, count, planes, reinterpret_cast<uint4*>(rows_dev));
  VMX_CHECK_LAUNCH();
  return VMX_OK;
}
// element dst_idx[r] of a fresh array of n elements = row r; the rows must cover every element once
static int unpack_rows(vmx_ctx* c, uint32_t* d, size_t cap, size_t n, const void* rows_dev, const uint32_t* dst_idx,
                       size_t count, int limbs) {
  if (count != n) { set_error("unpack_rows: %zu rows for %zu elements", count, n); return VMX_ESIZE; }
  if (!count) return VMX_OK;
  if (!rows_dev) return VMX_EARG;
  DevBuf di;
  if (dst_idx) {
    std::vector<uint8_t> seen(n, 0);
    for (size_t r = 0; r < count; r++) {
      if (dst_idx[r] >= n || seen[dst_idx[r]]) { set_error("unpack_rows: destination list is not a permutation"); return VMX_EARG; }
      seen[dst_idx[r]] = 1;
    }
    VMX_TRY(upload_u32(c, dst_idx, count, di));
  }
  const int planes = limbs / 4;
  VMX_LAUNCH(c, k_unpack_rows, nblocks(count * planes, 256), 256, 0, reinterpret_cast<const uint4*>(rows_dev),
             dst_idx ? di.as<uint32_t>() : (const uint32_t*)nullptr, count, planes, reinterpret_cast<uint4*>(d), cap);
  VMX_CHECK_LAUNCH();
  return VMX_OK;
}


int vmx_permute(const vmx_garr* a, const uint32_t* perm, vmx_garr** out) {
  if (!out) return VMX_EARG;
  *out = nullptr;
  if (!a || (!perm && a->n)) return VMX_EARG;
  vmx_ctx* c = a->ctx;
  VMX_ENTER(c);
  for (size_t i = 0; i < a->n; i++) if (perm[i] >= a->n) { set_error("permutation entry out of range"); return VMX_EARG; }
  DevBuf p;
  VMX_TRY(upload_u32(c, perm, a->n, p));
  vmx_garr* r = nullptr;
  VMX_TRY(new_garr(c, a->n, &r));
  const int s = gather(c, a->d, a->cap, r->d, r->cap, a->n, nullptr, p.as<uint32_t>(), 0, 0, c->gl);
  if (s != VMX_OK) { vmx_garr_free(r); return s; }
  *out = r;
  return VMX_OK;
}

// ---- exchange between the GPUs of one box
size_t vmx_ctx_row_bytes(const vmx_ctx* c) { return c ? (size_t)c->gl * 4 : 0; }
size_t vmx_ctx_ring_row_bytes(const vmx_ctx* c) { return c ? (size_t)c->nl * 4 : 0; }
int vmx_garr_pack_rows(const vmx_garr* a, const uint32_t* idx, size_t count, void* rows_dev) {
  if (!a) return VMX_EARG;
  VMX_ENTER(a->ctx);
  return pack_rows(a->ctx, a->d, a->cap, a->n, idx, count, rows_dev, a->ctx->gl);
}
int vmx_rarr_pack_rows(const vmx_rarr* a, const uint32_t* idx, size_t count, void* rows_dev) {
  if (!a) return VMX_EARG;
  VMX_ENTER(a->ctx);
  return pack_rows(a->ctx, a->d, a->cap, a->n, idx, count, rows_dev, a->ctx->nl);
}
int vmx_garr_unpack_rows(vmx_ctx* c, size_t n, const void* rows_dev, const uint32_t* dst_idx, size_t count, vmx_garr** out) {
  if (!out) return VMX_EARG;
  *out = nullptr;
  VMX_ENTER(c);
  vmx_garr* r = nullptr;
  VMX_TRY(new_garr(c, n, &r));
  const int s = unpack_rows(c, r->d, r->cap, n, rows_dev, dst_idx, count, c->gl);
  if (s != VMX_OK) { vmx_garr_free(r); return s; }
  *out = r;
  return VMX_OK;
}
int vmx_rarr_unpack_rows(vmx_ctx* c, size_t n, const void* rows_dev, const uint32_t* dst_idx, size_t count, vmx_rarr** out) {
  if (!out) return VMX_EARG;
  *out = nullptr;
  VMX_ENTER(c);
  vmx_rarr* r = nullptr;
  VMX_TRY(new_rarr(c, n, &r));
  const int s = unpack_rows(c, r->d, r->cap, n, rows_dev, dst_idx, count, c->nl);
  if (s != VMX_OK) { vmx_rarr_free(r); return s; }
  *out = r;
  return VMX_OK;
}

int vmx_shift_push(const vmx_garr* a, const uint8_t* elem_be, vmx_garr** out) {
  if (!out) return VMX_EARG;
  *out = nullptr;
  if (!a || !elem_be) return VMX_EARG;
  vmx_ctx* c = a->ctx;
  VMX_ENTER(c);
  ElemBuf el;
  VMX_TRY(upload_one(c, elem_be, true, el));
  vmx_garr* r = nullptr;
  VMX_TRY(new_garr(c, a->n, &r));
  std::unique_ptr<vmx_garr, void (*)(vmx_garr*)> guard(r, vmx_garr_free);
  if (a->n) {
    VMX_TRY(gather(c, a->d, a->cap, r->d, r->cap, a->n - 1, nullptr, nullptr, 0, 1, c->gl));
    VMX_TRY(gather(c, el.d(), el.cap, r->d, r->cap, 1, nullptr, nullptr, 0, 0, c->gl));
  }
  *out = guard.release();
  return VMX_OK;
}

int vmx_extract(const vmx_garr* a, const uint8_t* keep, vmx_garr** out) {
  if (!out) return VMX_EARG;
  *out = nullptr;
  if (!a || (!keep && a->n)) return VMX_EARG;
  vmx_ctx* c = a->ctx;
  VMX_ENTER(c);
  std::vector<uint32_t> src;
  for (size_t i = 0; i < a->n; i++) if (keep[i]) src.push_back((uint32_t)i);
  DevBuf p;
  VMX_TRY(upload_u32(c, src.data(), src.size(), p));
  vmx_garr* r = nullptr;
  VMX_TRY(new_garr(c, src.size(), &r));
  const int s = gather(c, a->d, a->cap, r->d, r->cap, src.size(), p.as<uint32_t>(), nullptr, 0, 0, c->gl);
  if (s != VMX_OK) { vmx_garr_free(r); return s; }
  *out = r;
  return VMX_OK;
}

int vmx_slice(const vmx_garr* a, size_t begin, size_t end, vmx_garr** out) {
  if (!out) return VMX_EARG;
  *out = nullptr;
  if (!a) return VMX_EARG;
  if (begin > end || end > a->n) { set_error("slice [%zu,%zu) out of range %zu", begin, end, a->n); return VMX_ESIZE; }
  vmx_ctx* c = a->ctx;
  VMX_ENTER(c);
  vmx_garr* r = nullptr;
  VMX_TRY(new_garr(c, end - begin, &r));
  const int s = gather(c, a->d, a->cap, r->d, r->cap, end - begin, nullptr, nullptr, (long long)begin, 0, c->gl);
  if (s != VMX_OK) { vmx_garr_free(r); return s; }
  *out = r;
  return VMX_OK;
}

static int arrays_equal(vmx_ctx* c, const uint32_t* a, size_t acap, const uint32_t* b, size_t bcap, size_t n, int* eq,
                        int limbs) {
  VMX_CU(cudaMemsetAsync(c->d_flag, 0, sizeof(int) * 4, c->stream));
  if (n) {
    const int planes = limbs / 4;
    VMX_LAUNCH(c, k_equal, nblocks(n * planes, 256), 256, 0, reinterpret_cast<const uint4*>(a), acap,
               reinterpret_cast<const uint4*>(b), bcap, n, planes, c->d_flag);
    VMX_CHECK_LAUNCH();
  }
  VMX_TRY(read_flags(c, 1));
  *eq = c->h_flag[0] == 0;
  return VMX_OK;
}

int vmx_equals(const vmx_garr* a, const vmx_garr* b, int* equal) {
  if (!a || !b || !equal) return VMX_EARG;
  vmx_ctx* c = a->ctx;
  VMX_ENTER(c);
  VMX_TRY(same_ctx(b->ctx, c));
  if (a->n != b->n) { *equal = 0; return VMX_OK; }
  return arrays_equal(c, a->d, a->cap, b->d, b->cap, a->n, equal, c->gl);
}

int vmx_get(const vmx_garr* a, size_t i, uint8_t* out_be) {
  if (!a || !out_be) return VMX_EARG;
  if (i >= a->n) { set_error("index %zu out of range %zu", i, a->n); return VMX_ESIZE; }
  VMX_ENTER(a->ctx);
  return download_one(a->ctx, a->d, a->cap, i, true, out_be);
}

int vmx_expprod_cols(const vmx_garr* const* bases, size_t t, const int64_t* ints, vmx_garr** out) {
  if (!out) return VMX_EARG;
  *out = nullptr;
  if (!bases || !t || !ints) return VMX_EARG;
  vmx_ctx* c = bases[0]->ctx;
  VMX_ENTER(c);
  const size_t n = bases[0]->n;
  if (c->kind == 1) {
    for (size_t j = 0; j < t; j++)
      if (bases[j]->ctx != c || bases[j]->n != n) { set_error("expProd: size mismatch"); return VMX_ESIZE; }
    vmx_garr* r = nullptr;
    VMX_TRY(new_garr(c, n, &r));
    const int st = ec_cols(c, bases, t, ints, r);
    if (st != VMX_OK) { vmx_garr_free(r); return st; }
    *out = r;
    return VMX_OK;
  }
  vmx_garr* acc = nullptr;
  for (size_t j = 0; j < t; j++) {
    if (bases[j]->ctx != c || bases[j]->n != n) { if (acc) vmx_garr_free(acc); set_error("expProd: size mismatch"); return VMX_ESIZE; }
    // term = bases[j]^{|ints[j]|}, inverted if negative
    uint32_t x[kMaxLimbs] = {0};
    const uint64_t mag = ints[j] < 0 ? (uint64_t)(-(ints[j] + 1)) + 1 : (uint64_t)ints[j];
    x[0] = (uint32_t)mag; x[1] = (uint32_t)(mag >> 32);
    vmx_garr* term = nullptr;
    int s = exp_scalar_limbs(c, bases[j], x, &term);
    if (s == VMX_OK && ints[j] < 0) {
      vmx_garr* inv = nullptr;
      s = vmx_inv(term, &inv);
      vmx_garr_free(term);
      term = inv;
    }
    if (s != VMX_OK) { if (acc) vmx_garr_free(acc); return s; }
    if (!acc) { acc = term; continue; }
    vmx_garr* prod = nullptr;
    s = vmx_mul(acc, term, &prod);
    vmx_garr_free(acc);
    vmx_garr_free(term);
    if (s != VMX_OK) return s;
    acc = prod;
  }
  *out = acc;
  return VMX_OK;
}

// ---------------------------------------------------------------- ring arrays
static int rarr_import(vmx_ctx* c, size_t n, const uint8_t* be, int hdr, vmx_rarr** out) {
  if (!out) return VMX_EARG;
  *out = nullptr;
  VMX_ENTER(c);
  if (n && !be) return VMX_EARG;
  vmx_rarr* a = nullptr;
  VMX_TRY(new_rarr(c, n, &a));
  std::unique_ptr<vmx_rarr, void (*)(vmx_rarr*)> guard(a, vmx_rarr_free);
  if (n) {
    DevBuf raw;
    const size_t bytes = n * (c->rb + hdr);
    VMX_TRY(raw.alloc(c, bytes));
    VMX_CU(cudaMemcpyAsync(raw.p, be, bytes, cudaMemcpyHostToDevice, c->stream));
    VMX_CU(cudaMemsetAsync(c->d_flag, 0, sizeof(int) * 4, c->stream));
    VMX_DISPATCH(c->nl, VMX_LAUNCH(c, k_from_bytes<N>, nblocks(n, codec_threads(c->rb + hdr)), codec_threads(c->rb + hdr), codec_smem(c->rb + hdr), raw.as<uint8_t>(), n, (int)c->rb, hdr, 1,
                                   a->d, a->cap, c->Q.consts, c->d_flag, c->Q.params<N>()));
    VMX_CHECK_LAUNCH();
    VMX_TRY(read_flags(c, 1));
    if (c->h_flag[0]) { set_error("ring element out of range or malformed leaf (flags %d)", c->h_flag[0]); return VMX_EFORMAT; }
  }
  *out = guard.release();
  return VMX_OK;
}
int vmx_rarr_from_bytes(vmx_ctx* c, size_t n, const uint8_t* be, vmx_rarr** out) { return rarr_import(c, n, be, 0, out); }
int vmx_rarr_from_leaves(vmx_ctx* c, size_t n, const uint8_t* leaves, vmx_rarr** out) { return rarr_import(c, n, leaves, 5, out); }

int vmx_rarr_from_raw(vmx_ctx* c, size_t n, const uint8_t* be, size_t width, unsigned bitlen, vmx_rarr** out) {
  if (!out) return VMX_EARG;
  *out = nullptr;
  VMX_ENTER(c);
  if ((n && !be) || !width || width > (size_t)16 * c->nl) { set_error("from_raw: bad width %zu", width); return VMX_EARG; }
  if (bitlen > 8 * width) bitlen = 0;
  vmx_rarr* a = nullptr;
  VMX_TRY(new_rarr(c, n, &a));
  std::unique_ptr<vmx_rarr, void (*)(vmx_rarr*)> guard(a, vmx_rarr_free);
  if (n) {
    DevBuf raw;
    VMX_TRY(raw.alloc(c, n * width));
    VMX_CU(cudaMemcpyAsync(raw.p, be, n * width, cudaMemcpyHostToDevice, c->stream));
    const int totalbits = bitlen ? (int)bitlen : (int)(8 * width);
    const int need_reduce = totalbits >= c->Q.bits;
    VMX_DISPATCH(c->nl, VMX_LAUNCH(c, k_ring_from_raw<N>, nblocks(n, codec_threads(width)), codec_threads(width), codec_smem(width), raw.as<uint8_t>(), n, (int)width,
                                   (int)bitlen, a->d, a->cap, c->Q.consts, need_reduce, c->Q.params<N>()));
    VMX_CHECK_LAUNCH();
    if (need_reduce) c->modmuls += 3 * n;
    VMX_CU(cudaStreamSynchronize(c->stream));  // `be` is borrowed for the call only
  }
  *out = guard.release();
  return VMX_OK;
}

int vmx_rarr_prg_sha256(vmx_ctx* c, const uint8_t* seed, size_t seedlen, uint64_t offset, size_t n, unsigned bitlen,
                        vmx_rarr** out) {
  if (!out) return VMX_EARG;
  *out = nullptr;
  VMX_ENTER(c);
  if (!seed || seedlen < 32 || seedlen > 48) { set_error("PRG seed must be 32..48 bytes"); return VMX_EARG; }
  if (!bitlen || (int)bitlen >= c->Q.bits) { set_error("PRG bit length %u must be below |q|", bitlen); return VMX_EARG; }
  const size_t wbytes = (bitlen + 7) / 8;
  if ((offset + n * wbytes) / 32 >= 0xffffffffull) { set_error("PRG stream too long"); return VMX_ESIZE; }
  PrgSeed s;
  std::memset(&s, 0, sizeof s);
  std::memcpy(s.bytes, seed, seedlen);
  s.len = (int)seedlen;
  vmx_rarr* a = nullptr;
  VMX_TRY(new_rarr(c, n, &a));
  if (n) {
    VMX_DISPATCH(c->nl, VMX_LAUNCH(c, k_prg_expand<N>, nblocks(n, 128), 128, 0, s, (size_t)offset, n, (int)bitlen, a->d, a->cap));
    if (cudaGetLastError() != cudaSuccess) { vmx_rarr_free(a); set_error("prg launch failed"); return VMX_ECUDA; }
  }
  *out = a;
  return VMX_OK;
}

static int rarr_export(const vmx_rarr* a, int hdr, uint8_t* be_out) {
  if (!a) return VMX_EARG;
  vmx_ctx* c = a->ctx;
  VMX_ENTER(c);
  if (!a->n) return VMX_OK;
  if (!be_out) return VMX_EARG;
  DevBuf raw;
  const size_t bytes = a->n * (c->rb + hdr);
  VMX_TRY(raw.alloc(c, bytes));
  VMX_DISPATCH(c->nl, VMX_LAUNCH(c, k_to_bytes<N>, nblocks(a->n, codec_threads(c->rb + hdr)), codec_threads(c->rb + hdr), codec_smem(c->rb + hdr), a->d, a->cap, a->n, (int)c->rb, hdr, 1,
                                 raw.as<uint8_t>(), c->Q.params<N>()));
  VMX_CHECK_LAUNCH();
  VMX_CU(cudaMemcpyAsync(be_out, raw.p, bytes, cudaMemcpyDeviceToHost, c->stream));
  VMX_CU(cudaStreamSynchronize(c->stream));
  return VMX_OK;
}
int vmx_rarr_to_bytes(const vmx_rarr* a, uint8_t* be_out) { return rarr_export(a, 0, be_out); }
int vmx_rarr_to_leaves(const vmx_rarr* a, uint8_t* leaves_out) { return rarr_export(a, 5, leaves_out); }

int vmx_rarr_fill(vmx_ctx* c, size_t n, const uint8_t* elem_be, vmx_rarr** out) {
  if (!out) return VMX_EARG;
  *out = nullptr;
  VMX_ENTER(c);
  ElemBuf one;
  VMX_TRY(upload_one(c, elem_be, false, one));
  vmx_rarr* a = nullptr;
  VMX_TRY(new_rarr(c, n, &a));
  if (n) {
    const int planes = c->nl / 4;
    VMX_LAUNCH(c, k_gather, nblocks(n * planes, 256), 256, 0, reinterpret_cast<const uint4*>(one.d()), one.cap,
               reinterpret_cast<uint4*>(a->d), a->cap, n, planes, (const uint32_t*)nullptr, (const uint32_t*)nullptr,
               (long long)0, (long long)0, 1, (size_t)0);
    if (cudaGetLastError() != cudaSuccess) { vmx_rarr_free(a); set_error("fill launch failed"); return VMX_ECUDA; }
  }
  *out = a;
  return VMX_OK;
}

void vmx_rarr_free(vmx_rarr* a) {
  if (!a) return;
  cudaSetDevice(a->ctx->device);
  if (a->d) dev_free(a->ctx, a->d, a->granted);
  delete a;
}
size_t vmx_rarr_size(const vmx_rarr* a) { return a ? a->n : 0; }
int vmx_rarr_bitlen(const vmx_rarr* a, unsigned* bits) {
  if (!a || !bits) return VMX_EARG;
  VMX_ENTER(a->ctx);
  int b = 0;
  VMX_TRY(rarr_bitlen(a, &b));
  *bits = (unsigned)b;
  return VMX_OK;
}

static int ring_binary(const vmx_rarr* a, const vmx_rarr* b, int op, vmx_rarr** out) {
  if (!out) return VMX_EARG;
  *out = nullptr;
  if (!a || (op != 1 && !b)) return VMX_EARG;
  vmx_ctx* c = a->ctx;
  VMX_ENTER(c);
  if (b) {
    VMX_TRY(same_ctx(b->ctx, c));
    if (a->n != b->n) { set_error("ring op: size mismatch %zu vs %zu", a->n, b->n); return VMX_ESIZE; }
  }
  vmx_rarr* r = nullptr;
  VMX_TRY(new_rarr(c, a->n, &r));
  if (a->n) {
    VMX_DISPATCH(c->nl, VMX_LAUNCH(c, k_ring_addsub<N>, nblocks(a->n), kThreads, 0, a->d, a->cap, b ? b->d : a->d,
                                   b ? b->cap : a->cap, r->d, r->cap, a->n, op, c->Q.params<N>()));
    if (cudaGetLastError() != cudaSuccess) { vmx_rarr_free(r); set_error("ring launch failed"); return VMX_ECUDA; }
  }
  *out = r;
  return VMX_OK;
}
int vmx_radd(const vmx_rarr* a, const vmx_rarr* b, vmx_rarr** out) { return ring_binary(a, b, 0, out); }
int vmx_rneg(const vmx_rarr* a, vmx_rarr** out) { return ring_binary(a, nullptr, 1, out); }
int vmx_rsub(const vmx_rarr* a, const vmx_rarr* b, vmx_rarr** out) { return ring_binary(a, b, 2, out); }

int vmx_rmul(const vmx_rarr* a, const vmx_rarr* b, vmx_rarr** out) {
  if (!out) return VMX_EARG;
  *out = nullptr;
  if (!a || !b) return VMX_EARG;
  vmx_ctx* c = a->ctx;
  VMX_ENTER(c);
  VMX_TRY(same_ctx(b->ctx, c));
  if (a->n != b->n) { set_error("ring mul: size mismatch %zu vs %zu", a->n, b->n); return VMX_ESIZE; }
  vmx_rarr* r = nullptr;
  VMX_TRY(new_rarr(c, a->n, &r));
  std::unique_ptr<vmx_rarr, void (*)(vmx_rarr*)> guard(r, vmx_rarr_free);
  if (a->n) {
    // r = a*b*R^-1 ; r = r * R^2 * R^-1 = a*b
    VMX_DISPATCH(c->nl, {
      ElemBuf t;
      VMX_TRY(t.alloc_elems(c, a->n));
      VMX_LAUNCH(c, k_mul<N>, nblocks(a->n), kThreads, 0, a->d, a->cap, b->d, b->cap, t.d(), t.cap, a->n,
                 c->Q.params<N>());
      VMX_CHECK_LAUNCH();
      c->modmuls += a->n;
      VMX_TRY(ring_mul_const<N>(c, t.d(), t.cap, c->Q.consts, 4, 0, nullptr, 0, r->d, r->cap, a->n));
    });
  }
  *out = guard.release();
  return VMX_OK;
}

int vmx_rmuladd(const vmx_rarr* a, const uint8_t* s_be, const vmx_rarr* b, vmx_rarr** out) {
  if (!out) return VMX_EARG;
  *out = nullptr;
  if (!a || !b || !s_be) return VMX_EARG;
  vmx_ctx* c = a->ctx;
  VMX_ENTER(c);
  VMX_TRY(same_ctx(b->ctx, c));
  if (a->n != b->n) { set_error("mulAdd: size mismatch %zu vs %zu", a->n, b->n); return VMX_ESIZE; }
  ElemBuf s, sM;
  VMX_TRY(upload_one(c, s_be, false, s));
  VMX_TRY(sM.alloc_elems(c, 1));
  vmx_rarr* r = nullptr;
  VMX_TRY(new_rarr(c, a->n, &r));
  std::unique_ptr<vmx_rarr, void (*)(vmx_rarr*)> guard(r, vmx_rarr_free);
  VMX_DISPATCH(c->nl, {
    VMX_TRY(ring_mul_const<N>(c, s.d(), s.cap, c->Q.consts, 4, 0, nullptr, 0, sM.d(), sM.cap, 1));
    VMX_TRY(ring_mul_const<N>(c, a->d, a->cap, sM.d(), sM.cap, 0, b->d, b->cap, r->d, r->cap, a->n));
  });
  *out = guard.release();
  return VMX_OK;
}

int vmx_rinner(const vmx_rarr* a, const vmx_rarr* b, uint8_t* out_be) {
  if (!a || !b || !out_be) return VMX_EARG;
  vmx_ctx* c = a->ctx;
  VMX_ENTER(c);
  VMX_TRY(same_ctx(b->ctx, c));
  if (a->n != b->n) { set_error("innerProduct: size mismatch %zu vs %zu", a->n, b->n); return VMX_ESIZE; }
  if (!a->n) { std::memset(out_be, 0, c->rb); return VMX_OK; }
  ElemBuf res;
  VMX_DISPATCH(c->nl, VMX_TRY(ring_reduce_sum<N>(c, a->d, a->cap, b->d, b->cap, a->n, res)));
  return download_one(c, res.d(), res.cap, 0, false, out_be);
}

int vmx_rsum(const vmx_rarr* a, uint8_t* out_be) {
  if (!a || !out_be) return VMX_EARG;
  vmx_ctx* c = a->ctx;
  VMX_ENTER(c);
  if (!a->n) { std::memset(out_be, 0, c->rb); return VMX_OK; }
  ElemBuf res;
  VMX_DISPATCH(c->nl, VMX_TRY(ring_reduce_sum<N>(c, a->d, a->cap, nullptr, 0, a->n, res)));
  return download_one(c, res.d(), res.cap, 0, false, out_be);
}

int vmx_rprod(const vmx_rarr* a, uint8_t* out_be) {
  if (!a || !out_be) return VMX_EARG;
  vmx_ctx* c = a->ctx;
  VMX_ENTER(c);
  ElemBuf aM, res, can;
  VMX_TRY(aM.alloc_elems(c, a->n));
  VMX_TRY(res.alloc_elems(c, 1));
  VMX_TRY(can.alloc_elems(c, 1));
  DevBuf off;
  const uint32_t h[2] = {0, (uint32_t)a->n};
  VMX_TRY(upload_u32(c, h, 2, off));
  const int K = (int)std::min<size_t>(64, std::max<size_t>(2, a->n / wave_threads(c)));
  VMX_DISPATCH(c->nl, {
    VMX_TRY(ring_mul_const<N>(c, a->d, a->cap, c->Q.consts, 4, 0, nullptr, 0, aM.d(), aM.cap, a->n));
    VMX_TRY(seg_product<N>(c, c->Q, aM.d(), aM.cap, nullptr, off.as<uint32_t>(), 1, a->n, K, res.d(), res.cap));
    VMX_LAUNCH(c, k_from_mont<N>, 1, kThreads, 0, res.d(), res.cap, can.d(), can.cap, (size_t)1, c->Q.params<N>());
    VMX_CHECK_LAUNCH();
  });
  return download_one(c, can.d(), can.cap, 0, false, out_be);
}

static int ring_scan_api(const vmx_rarr* b, const vmx_rarr* e, int want_y, vmx_rarr** out, uint8_t* last_be) {
  if (!out) return VMX_EARG;
  *out = nullptr;
  if (!e || (!want_y && !b)) return VMX_EARG;
  vmx_ctx* c = e->ctx;
  VMX_ENTER(c);
  if (b) {
    VMX_TRY(same_ctx(b->ctx, c));
    if (b->n != e->n) { set_error("recLin: size mismatch %zu vs %zu", b->n, e->n); return VMX_ESIZE; }
  }
  const size_t n = e->n;
  vmx_rarr* r = nullptr;
  VMX_TRY(new_rarr(c, n, &r));
  std::unique_ptr<vmx_rarr, void (*)(vmx_rarr*)> guard(r, vmx_rarr_free);
  if (n) {
    ElemBuf eM;
    VMX_TRY(eM.alloc_elems(c, n));
    VMX_DISPATCH(c->nl, {
      VMX_TRY(ring_mul_const<N>(c, e->d, e->cap, c->Q.consts, 4, 0, nullptr, 0, eM.d(), eM.cap, n));
      VMX_TRY(ring_scan<N>(c, eM.d(), eM.cap, b ? b->d : nullptr, b ? b->cap : 0, n, want_y, r->d, r->cap));
    });
    if (last_be) VMX_TRY(download_one(c, r->d, r->cap, n - 1, false, last_be));
  } else if (last_be) {
    std::memset(last_be, 0, c->rb);
  }
  *out = guard.release();
  return VMX_OK;
}
int vmx_rprods(const vmx_rarr* a, vmx_rarr** out) { return ring_scan_api(nullptr, a, 1, out, nullptr); }
int vmx_rreclin(const vmx_rarr* b, const vmx_rarr* e, vmx_rarr** out, uint8_t* last_be) {
  return ring_scan_api(b, e, 0, out, last_be);
}

int vmx_rpermute(const vmx_rarr* a, const uint32_t* perm, vmx_rarr** out) {
  if (!out) return VMX_EARG;
  *out = nullptr;
  if (!a || (!perm && a->n)) return VMX_EARG;
  vmx_ctx* c = a->ctx;
  VMX_ENTER(c);
  for (size_t i = 0; i < a->n; i++) if (perm[i] >= a->n) { set_error("permutation entry out of range"); return VMX_EARG; }
  DevBuf p;
  VMX_TRY(upload_u32(c, perm, a->n, p));
  vmx_rarr* r = nullptr;
  VMX_TRY(new_rarr(c, a->n, &r));
  const int s = gather(c, a->d, a->cap, r->d, r->cap, a->n, nullptr, p.as<uint32_t>(), 0, 0);
  if (s != VMX_OK) { vmx_rarr_free(r); return s; }
  r->bits = a->bits;
  *out = r;
  return VMX_OK;
}

int vmx_rshift_push(const vmx_rarr* a, const uint8_t* elem_be, vmx_rarr** out) {
  if (!out) return VMX_EARG;
  *out = nullptr;
  if (!a || !elem_be) return VMX_EARG;
  vmx_ctx* c = a->ctx;
  VMX_ENTER(c);
  ElemBuf el;
  VMX_TRY(upload_one(c, elem_be, false, el));
  vmx_rarr* r = nullptr;
  VMX_TRY(new_rarr(c, a->n, &r));
  std::unique_ptr<vmx_rarr, void (*)(vmx_rarr*)> guard(r, vmx_rarr_free);
  if (a->n) {
    VMX_TRY(gather(c, a->d, a->cap, r->d, r->cap, a->n - 1, nullptr, nullptr, 0, 1));
    VMX_TRY(gather(c, el.d(), el.cap, r->d, r->cap, 1, nullptr, nullptr, 0, 0));
  }
  *out = guard.release();
  return VMX_OK;
}

int vmx_rslice(const vmx_rarr* a, size_t begin, size_t end, vmx_rarr** out) {
  if (!out) return VMX_EARG;
  *out = nullptr;
  if (!a) return VMX_EARG;
  if (begin > end || end > a->n) { set_error("slice [%zu,%zu) out of range %zu", begin, end, a->n); return VMX_ESIZE; }
  vmx_ctx* c = a->ctx;
  VMX_ENTER(c);
  vmx_rarr* r = nullptr;
  VMX_TRY(new_rarr(c, end - begin, &r));
  const int s = gather(c, a->d, a->cap, r->d, r->cap, end - begin, nullptr, nullptr, (long long)begin, 0);
  if (s != VMX_OK) { vmx_rarr_free(r); return s; }
  *out = r;
  return VMX_OK;
}

int vmx_rget(const vmx_rarr* a, size_t i, uint8_t* out_be) {
  if (!a || !out_be) return VMX_EARG;
  if (i >= a->n) { set_error("index %zu out of range %zu", i, a->n); return VMX_ESIZE; }
  VMX_ENTER(a->ctx);
  return download_one(a->ctx, a->d, a->cap, i, false, out_be);
}

int vmx_requals(const vmx_rarr* a, const vmx_rarr* b, int* equal) {
  if (!a || !b || !equal) return VMX_EARG;
  vmx_ctx* c = a->ctx;
  VMX_ENTER(c);
  VMX_TRY(same_ctx(b->ctx, c));
  if (a->n != b->n) { *equal = 0; return VMX_OK; }
  return arrays_equal(c, a->d, a->cap, b->d, b->cap, a->n, equal, c->nl);
}

// ---------------------------------------------------------------- self test of the cooperative multiplier
int vmx_selftest_coop(const vmx_garr* a, const vmx_garr* b, int* equal) {
  if (!a || !b || !equal) return VMX_EARG;
  vmx_ctx* c = a->ctx;
  VMX_ENTER(c);
  if (c->kind == 1) { set_error("cooperative multiplier: ModPGroup contexts only"); return VMX_EARG; }
  if (a->n != b->n || b->ctx != c) return VMX_ESIZE;
#ifdef VMX_HOST_EMUL
  *equal = 1;
  return VMX_OK;
#else
  ElemBuf x, y;
  VMX_TRY(x.alloc_elems(c, a->n));
  VMX_TRY(y.alloc_elems(c, a->n));
  VMX_DISPATCH(c->nl, {
    VMX_LAUNCH(c, k_mul<N>, nblocks(a->n), kThreads, 0, a->d, a->cap, b->d, b->cap, x.d(), x.cap, a->n, c->P.params<N>());
    VMX_LAUNCH(c, k_coop_mul<N>, nblocks(a->n, kCoopWarps), 32 * kCoopWarps, 0, a->d, a->cap, b->d, b->cap, a->n, y.d(),
               y.cap, c->P.consts, c->P.n0inv);
  });
  VMX_CHECK_LAUNCH();
  return arrays_equal(c, x.d(), x.cap, y.d(), y.cap, a->n, equal, c->nl);
#endif
}

// a[i]^2 by the dedicated squaring (mont_sqr, block-triangular from 32 limbs on) against a[i] * a[i] by the
// multiplication; *ms (optional) = device time of `iters` chained squarings per element
int vmx_selftest_sqr(const vmx_garr* a, int iters, int* equal, float* ms) {
  if (!a || !equal || iters < 1) return VMX_EARG;
  vmx_ctx* c = a->ctx;
  VMX_ENTER(c);
  if (c->kind == 1) { set_error("squaring self test: ModPGroup contexts only"); return VMX_EARG; }
  if (a->n == 0) { *equal = 1; if (ms) *ms = 0.f; return VMX_OK; }
  ElemBuf x, y;
  VMX_TRY(x.alloc_elems(c, a->n));
  VMX_TRY(y.alloc_elems(c, a->n));
#ifndef VMX_HOST_EMUL
  cudaEvent_t e0, e1;
  VMX_CU(cudaEventCreate(&e0));
  VMX_CU(cudaEventCreate(&e1));
#endif
  VMX_DISPATCH(c->nl, {
    const MontParams<N> M = c->P.params<N>();
    // reference: x <- a, then x <- x * x (second operand streamed from x itself before it is overwritten: k_mul
    // reads b[i] word by word while a[i] sits in registers, and writes only at the end)
    VMX_LAUNCH(c, k_mul<N>, nblocks(a->n), kThreads, 0, a->d, a->cap, a->d, a->cap, x.d(), x.cap, a->n, M);
    for (int it = 1; it < iters; it++)
      VMX_LAUNCH(c, k_mul<N>, nblocks(a->n), kThreads, 0, x.d(), x.cap, x.d(), x.cap, x.d(), x.cap, a->n, M);
#ifndef VMX_HOST_EMUL
    VMX_CU(cudaEventRecord(e0, c->stream));
#endif
    VMX_LAUNCH(c, k_sqr_iter<N>, nblocks(a->n), kThreads, kThreads * N * 4, a->d, a->cap, y.d(), y.cap, a->n, iters, M);
#ifndef VMX_HOST_EMUL
    VMX_CU(cudaEventRecord(e1, c->stream));
#endif
  });
  VMX_CHECK_LAUNCH();
  c->modmuls += 2 * (uint64_t)a->n * iters;
#ifndef VMX_HOST_EMUL
  VMX_CU(cudaEventSynchronize(e1));
  float t = 0.f;
  VMX_CU(cudaEventElapsedTime(&t, e0, e1));
  if (ms) *ms = t;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
#else
  if (ms) *ms = 0.f;
#endif
  return arrays_equal(c, x.d(), x.cap, y.d(), y.cap, a->n, equal, c->nl);
}

// a[i]*b[i] on the cooperative multiplier (debug / test hook)
int vmx_debug_coop_mul(const vmx_garr* a, const vmx_garr* b, vmx_garr** out) {
  if (!a || !b || !out) return VMX_EARG;
  *out = nullptr;
  vmx_ctx* c = a->ctx;
  VMX_ENTER(c);
  if (c->kind == 1) { set_error("cooperative multiplier: ModPGroup contexts only"); return VMX_EARG; }
  if (a->n != b->n || b->ctx != c) return VMX_ESIZE;
#ifdef VMX_HOST_EMUL
  return vmx_mul(a, b, out);
#else
  vmx_garr* r = nullptr;
  VMX_TRY(new_garr(c, a->n, &r));
  VMX_DISPATCH(c->nl, VMX_LAUNCH(c, k_coop_mul<N>, nblocks(a->n, kCoopWarps), 32 * kCoopWarps, 0, a->d, a->cap, b->d,
                                 b->cap, a->n, r->d, r->cap, c->P.consts, c->P.n0inv));
  if (cudaGetLastError() != cudaSuccess) { vmx_garr_free(r); set_error("coop launch failed"); return VMX_ECUDA; }
  *out = r;
  return VMX_OK;
#endif
}

// ---------------------------------------------------------------- benchmark hook
int vmx_bench_modmul(vmx_ctx* c, size_t n, int iters, float* ms) {
  VMX_ENTER(c);
  if (!n || iters <= 0 || !ms) return VMX_EARG;
  ElemBuf a, b, o;
  VMX_TRY(a.alloc_elems(c, n));
  VMX_TRY(b.alloc_elems(c, n));
  VMX_TRY(o.alloc_elems(c, n));
  // operands: Montgomery one and R^2 pattern replicated (any residues < p do)
  VMX_LAUNCH(c, k_gather, nblocks(n * (c->nl / 4), 256), 256, 0, reinterpret_cast<const uint4*>(c->P.consts),
             (size_t)4, reinterpret_cast<uint4*>(a.d()), a.cap, n, c->nl / 4, (const uint32_t*)nullptr,
             (const uint32_t*)nullptr, (long long)0, (long long)0, 1, (size_t)1);
  VMX_LAUNCH(c, k_gather, nblocks(n * (c->nl / 4), 256), 256, 0, reinterpret_cast<const uint4*>(c->P.consts),
             (size_t)4, reinterpret_cast<uint4*>(b.d()), b.cap, n, c->nl / 4, (const uint32_t*)nullptr,
             (const uint32_t*)nullptr, (long long)0, (long long)0, 1, (size_t)0);
  VMX_CHECK_LAUNCH();
#ifndef VMX_HOST_EMUL
  cudaEvent_t e0, e1;
  VMX_CU(cudaEventCreate(&e0));
  VMX_CU(cudaEventCreate(&e1));
  VMX_CU(cudaEventRecord(e0, c->stream));
#endif
  VMX_DISPATCH(c->nl, VMX_LAUNCH(c, k_mul_iter<N>, nblocks(n), kThreads, 0, a.d(), b.d(), o.d(), a.cap, n, iters,
                                 c->P.params<N>()));
  VMX_CHECK_LAUNCH();
  c->modmuls += (uint64_t)n * iters;
#ifndef VMX_HOST_EMUL
  VMX_CU(cudaEventRecord(e1, c->stream));
  VMX_CU(cudaEventSynchronize(e1));
  VMX_CU(cudaEventElapsedTime(ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
#else
  *ms = 0.f;
#endif
  return VMX_OK;
}

}  // extern "C"
