// vmnv_native.cpp -- the universal verifier of a Verificatum proof directory as a native pipeline over the C ABI
// of include/vmx.h: what
//     mixnet/MixNetElGamalVerifyFiatShamirSession.java:1318-1668  (verify: keys, shuffles, decryption, plaintexts)
//     hvzk/PoSTW.java:177-260 + hvzk/PoSBasicTW.java:505-514,780-823,970-1066  (proof of a shuffle, verifier)
//     elgamal/DistrElGamalSessionBasic.java:465-727                (combination and batched proof of decryption factors)
//     hvzk/ChallengerRO.java:96-116, distr/IndependentGeneratorsRO.java:110-130, elgamal/ProtocolElGamal.java:659-683
// do with a proof of type "mixing" over a ModPGroup or an ECqPGroup, for ciphertexts of any width, written against nothing but
// the engine's C entry points: byte trees are walked here (headers only -- the engine validates every leaf of an
// array on the device), Fiat-Shamir hashing is OpenSSL's SHA-256 on a worker thread beside the GPU, every group
// and ring operation is one vmx_* call.  The Python mirror (verificatum-vmn_b200/vmnv.py) is the specification it
// is tested against: same verdicts on honest and on corrupted directories (tests/test_vmnv_native.py).
//
// The library binds the engine at run time (vmxv_bind(path of libvmx.so)), so the same object serves the CUDA
// build and the host-emulation build the CPU tests use.  It contains no arithmetic of its own beyond
// subtracting a scalar from q and dividing by an integer below 1010 (Lagrange coefficients).
//
// Build: g++ -std=c++17 -O2 -fPIC -shared -o libvmnv.so vmnv_native.cpp -ldl -lcrypto -lpthread
#include <dlfcn.h>
#include <openssl/evp.h>

#include <atomic>
#include <condition_variable>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/vmx.h"
#include "../../include/vmnv.h"

namespace {

// ------------------------------------------------------------------------------------------------ engine binding
#define VMXV_SYMBOLS(X)                                                                                             \
  X(vmx_last_error) X(vmx_ctx_create_modp) X(vmx_ctx_create_ecq) X(vmx_ctx_destroy) X(vmx_ctx_elem_bytes) X(vmx_ctx_ring_bytes)          \
  X(vmx_garr_from_leaves) X(vmx_garr_to_leaves) X(vmx_garr_fill) X(vmx_garr_free) X(vmx_garr_size)                 \
  X(vmx_garr_prg_sha256) X(vmx_exp_fixed) X(vmx_elem_exp) X(vmx_elem_inv) X(vmx_exp_scalar_var) X(vmx_expprod)     \
  X(vmx_expprod_cols) X(vmx_mul) X(vmx_inv) X(vmx_prod) X(vmx_shift_push) X(vmx_equals) X(vmx_get)                 \
  X(vmx_rarr_from_leaves) X(vmx_rarr_from_bytes) X(vmx_rarr_to_bytes) X(vmx_rarr_prg_sha256) X(vmx_rarr_prg_raw_sha256) X(vmx_rarr_free)      \
  X(vmx_rprod) X(vmx_rmul) X(vmx_radd) X(vmx_leaves_uniform) X(vmx_ctx_launch_count) X(vmx_fixed_precompute)   \
  X(vmx_extract) X(vmx_slice)

struct Api {
#define X(name) decltype(&::name) name = nullptr;
  VMXV_SYMBOLS(X)
#undef X
  void* handle = nullptr;
} api;

struct FailStop : std::runtime_error {  // `v.failStop(...)` of the reference: the directory is unusable
  using std::runtime_error::runtime_error;
};
struct Malformed : std::runtime_error {  // EIOException / ArithmFormatException: the caller decides what it means
  using std::runtime_error::runtime_error;
};

[[noreturn]] void fail_stop(const char* fmt, ...) {
  char buf[400];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw FailStop(buf);
}

using Bytes = std::vector<uint8_t>;
struct Span {
  const uint8_t* p = nullptr;
  size_t n = 0;
  Span() = default;
  Span(const uint8_t* p_, size_t n_) : p(p_), n(n_) {}
  Span sub(size_t off, size_t len) const {
    if (off > n || len > n - off) throw Malformed("truncated");
    return Span(p + off, len);
  }
};

void check(int status, const char* what) {
  if (status == VMX_OK) return;
  if (status == VMX_EFORMAT) throw Malformed(std::string(what) + ": " + api.vmx_last_error());
  throw std::runtime_error(std::string(what) + " failed: " + api.vmx_last_error());
}

// ------------------------------------------------------------------------------------------------ byte trees
constexpr uint8_t NODE = 0, LEAF = 1;
constexpr int kMaxDepth = 64;

uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
void put_be32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }
Bytes header(uint8_t kind, uint32_t count) {
  Bytes h(5);
  h[0] = kind;
  put_be32(h.data() + 1, count);
  return h;
}

struct Hdr { uint8_t kind; uint32_t count; };
Hdr read_hdr(Span s, size_t off) {
  if (off + 5 > s.n) throw Malformed("truncated header");
  const uint8_t k = s.p[off];
  if (k != NODE && k != LEAF) throw Malformed("bad tag");
  return Hdr{k, be32(s.p + off + 1)};
}

// offset one past the subtree at `off` (iterative, depth-capped; equal-width leaf arrays skipped arithmetically)
size_t skip_tree(Span s, size_t off) {
  std::vector<uint64_t> stack{1};
  size_t pos = off;
  bool fresh = false;
  while (!stack.empty()) {
    if (stack.back() == 0) { stack.pop_back(); fresh = false; continue; }
    if (fresh && stack.back() > 1 && pos + 5 <= s.n && s.p[pos] == LEAF) {
      const uint64_t w = be32(s.p + pos + 1), cnt = stack.back();
      if (cnt <= (s.n - pos) / (5 + w) && api.vmx_leaves_uniform(s.p + pos, (size_t)cnt, (size_t)w)) {
        pos += (size_t)(cnt * (5 + w));
        stack.back() = 0;
        continue;
      }
    }
    fresh = false;
    stack.back()--;
    const Hdr h = read_hdr(s, pos);
    pos += 5;
    if (h.kind == LEAF) {
      if (h.count > s.n - pos) throw Malformed("truncated leaf");
      pos += h.count;
    } else if (h.count) {
      if ((int)stack.size() >= kMaxDepth) throw Malformed("byte tree nested too deep");
      stack.push_back(h.count);
      fresh = true;
    }
  }
  return pos;
}

// the children of the node at the start of `s` (exactly `want` of them, or any number if want < 0)
std::vector<Span> children(Span s, int want) {
  const Hdr h = read_hdr(s, 0);
  if (h.kind != NODE) throw Malformed("expected a node");
  if (want >= 0 && h.count != (uint32_t)want) throw Malformed("unexpected number of children");
  if (h.count > s.n) throw Malformed("truncated node");
  std::vector<Span> out;
  size_t pos = 5;
  for (uint32_t i = 0; i < h.count; i++) {
    const size_t end = skip_tree(s, pos);
    out.push_back(Span(s.p + pos, end - pos));
    pos = end;
  }
  return out;
}
// like children(), but the node may have MORE than `n` children: the first n are returned (ByteTreeReader.getNextChild
// n times, as the reference's parsers do)
std::vector<Span> first_children(Span s, int n) {
  const Hdr h = read_hdr(s, 0);
  if (h.kind != NODE || h.count < (uint32_t)n) throw Malformed("too few children");
  std::vector<Span> out;
  size_t pos = 5;
  for (int i = 0; i < n; i++) {
    const size_t end = skip_tree(s, pos);
    out.push_back(Span(s.p + pos, end - pos));
    pos = end;
  }
  return out;
}
Span leaf_payload(Span s, size_t width) {
  const Hdr h = read_hdr(s, 0);
  if (h.kind != LEAF || h.count != width || s.n < 5 + width) throw Malformed("leaf of wrong length");
  return Span(s.p + 5, width);
}

// ------------------------------------------------------------------------------------------------ hashing
struct Sha256 {
  EVP_MD_CTX* c;
  Sha256() : c(EVP_MD_CTX_new()) { EVP_DigestInit_ex(c, EVP_sha256(), nullptr); }
  ~Sha256() { EVP_MD_CTX_free(c); }
  Sha256(const Sha256&) = delete;
  void update(const void* p, size_t n) { if (n) EVP_DigestUpdate(c, p, n); }
  Bytes digest() {
    Bytes out(32);
    unsigned len = 32;
    EVP_DigestFinal_ex(c, out.data(), &len);
    return out;
  }
};
Bytes sha256(const Bytes& a) { Sha256 h; h.update(a.data(), a.size()); return h.digest(); }
// PRGHeuristic(SHA-256): H(seed || be32(0)) || H(seed || be32(1)) || ...
Bytes prg_bytes(const Bytes& seed, size_t n) {
  Bytes out;
  for (uint32_t ctr = 0; out.size() < n; ctr++) {
    Sha256 h;
    uint8_t c4[4];
    put_be32(c4, ctr);
    h.update(seed.data(), seed.size());
    h.update(c4, 4);
    const Bytes d = h.digest();
    out.insert(out.end(), d.begin(), d.end());
  }
  out.resize(n);
  return out;
}

// A random-oracle digest fed from a queue by one worker thread: the caller queues (pointer, length) pieces --
// file contents and serialisations that stay alive until finish() -- and goes on issuing engine calls.
class Oracle {
 public:
  Oracle(const Bytes& prefix, unsigned out_bits) : bits_(out_bits) {
    uint8_t b4[4];
    put_be32(b4, out_bits);
    sha_.update(b4, 4);
    sha_.update(prefix.data(), prefix.size());
    worker_ = std::thread([this] { run(); });
  }
  ~Oracle() { if (worker_.joinable()) { push(nullptr, 0, true); worker_.join(); } }
  void update(Span s) { push(s.p, s.n, false); }
  void update_owned(Bytes b) {  // small pieces (headers, single elements): kept alive here
    owned_.push_back(std::make_unique<Bytes>(std::move(b)));
    push(owned_.back()->data(), owned_.back()->size(), false);
  }
  Bytes finish() {
    push(nullptr, 0, true);
    worker_.join();
    const Bytes seed = sha_.digest();
    Bytes out = prg_bytes(seed, (bits_ + 7) / 8);
    const unsigned extra = (8 - bits_ % 8) % 8;
    if (extra) out[0] &= (uint8_t)(0xFF >> extra);
    return out;
  }
  uint64_t hashed = 0;

 private:
  struct Piece { const uint8_t* p; size_t n; bool stop; };
  void push(const uint8_t* p, size_t n, bool stop) {
    { std::lock_guard<std::mutex> lk(mu_); q_.push_back(Piece{p, n, stop}); }
    cv_.notify_one();
  }
  void run() {
    for (;;) {
      Piece pc;
      { std::unique_lock<std::mutex> lk(mu_); cv_.wait(lk, [this] { return !q_.empty(); }); pc = q_.front(); q_.pop_front(); }
      if (pc.stop) return;
      sha_.update(pc.p, pc.n);
      hashed += pc.n;
    }
  }
  Sha256 sha_;
  unsigned bits_;
  std::thread worker_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<Piece> q_;
  std::vector<std::unique_ptr<Bytes>> owned_;
};

// ------------------------------------------------------------------------------------------------ engine values
struct Ctx {
  vmx_ctx* c = nullptr;
  size_t eb = 0, rb = 0;
  bool curve = false;  // ECqPGroup: an element is a point x || y (two coordinates of cb = eb / 2 bytes, the unit
  size_t cb = 0;       // element all 0xff); on the wire node(leaf x, leaf y), an array node(node(x leaves), node(y leaves))
  Bytes q, g, one;  // big-endian, rb / eb / eb bytes
  ~Ctx() { if (c) api.vmx_ctx_destroy(c); }
};

struct Garr {
  vmx_garr* h = nullptr;
  Garr() = default;
  explicit Garr(vmx_garr* h_) : h(h_) {}
  Garr(Garr&& o) noexcept : h(o.h) { o.h = nullptr; }
  Garr& operator=(Garr&& o) noexcept { if (this != &o) { reset(); h = o.h; o.h = nullptr; } return *this; }
  Garr(const Garr&) = delete;
  ~Garr() { reset(); }
  void reset() { if (h) api.vmx_garr_free(h); h = nullptr; }
};
struct Rarr {
  vmx_rarr* h = nullptr;
  Rarr() = default;
  explicit Rarr(vmx_rarr* h_) : h(h_) {}
  Rarr(Rarr&& o) noexcept : h(o.h) { o.h = nullptr; }
  Rarr& operator=(Rarr&& o) noexcept { if (this != &o) { if (h) api.vmx_rarr_free(h); h = o.h; o.h = nullptr; } return *this; }
  Rarr(const Rarr&) = delete;
  ~Rarr() { if (h) api.vmx_rarr_free(h); }
};
using Elem = Bytes;    // one group element, eb bytes big-endian
using Scalar = Bytes;  // one element of Z_q, rb bytes big-endian

// ---- scalars: the little arithmetic that stays on the host
int cmp_be(const Bytes& a, const Bytes& b) { return a.size() != b.size() ? (a.size() < b.size() ? -1 : 1) : std::memcmp(a.data(), b.data(), a.size()); }
bool is_zero(const Bytes& a) { for (uint8_t v : a) if (v) return false; return true; }
Bytes sub_be(const Bytes& a, const Bytes& b) {  // a - b, a >= b, equal lengths
  Bytes r(a.size());
  int brw = 0;
  for (size_t i = a.size(); i-- > 0;) {
    int d = (int)a[i] - (int)b[i] - brw;
    brw = d < 0;
    r[i] = (uint8_t)(d + (brw ? 256 : 0));
  }
  return r;
}
Scalar scalar_from_u64(const Ctx& C, uint64_t v) {
  Scalar s(C.rb, 0);
  for (size_t i = 0; i < 8 && i < C.rb; i++) s[C.rb - 1 - i] = (uint8_t)(v >> (8 * i));
  return s;
}
Scalar scalar_neg(const Ctx& C, const Scalar& x) { return is_zero(x) ? x : sub_be(C.q, x); }
Scalar scalar_from_bytes(const Ctx& C, const Bytes& be) {  // a non-negative integer below q given in <= rb bytes
  Scalar s(C.rb, 0);
  if (be.size() > C.rb) throw std::runtime_error("scalar too long");
  std::memcpy(s.data() + (C.rb - be.size()), be.data(), be.size());
  return s;
}
uint32_t mod_small(const Bytes& a, uint32_t m) { uint64_t r = 0; for (uint8_t v : a) r = (r * 256 + v) % m; return (uint32_t)r; }
Bytes mul_small_add(const Bytes& a, uint32_t m, uint32_t add) {  // a * m + add, one byte longer
  Bytes r(a.size() + 4, 0);
  uint64_t carry = add;
  for (size_t i = a.size(); i-- > 0;) { carry += (uint64_t)a[i] * m; r[i + 4] = (uint8_t)carry; carry >>= 8; }
  for (size_t i = 4; i-- > 0;) { r[i] = (uint8_t)carry; carry >>= 8; }
  return r;
}
Bytes div_small_exact(const Bytes& a, uint32_t d) {
  Bytes r(a.size());
  uint64_t rem = 0;
  for (size_t i = 0; i < a.size(); i++) { rem = rem * 256 + a[i]; r[i] = (uint8_t)(rem / d); rem %= d; }
  if (rem) throw std::runtime_error("inexact division");
  return r;
}
// c^-1 mod q for 0 < c < 2^31 coprime to q:  (1 + q t) / c with t = (-q)^-1 mod c
Scalar inv_small(const Ctx& C, uint32_t c) {
  if (c == 1) return scalar_from_u64(C, 1);
  const uint32_t qm = mod_small(C.q, c);
  int64_t t = -1;
  for (uint32_t x = 1; x < c; x++) if (((uint64_t)x * qm + 1) % c == 0) { t = x; break; }
  if (t < 0) throw std::runtime_error("not invertible");
  Bytes num = mul_small_add(C.q, (uint32_t)t, 1);
  Bytes quo = div_small_exact(num, c);
  return scalar_from_bytes(C, Bytes(quo.end() - (long)C.rb, quo.end()));
}
// products and sums of scalars mod q go through the engine (one-element ring arrays)
Rarr rarr_of(const Ctx& C, const Scalar& s) {
  vmx_rarr* h = nullptr;
  check(api.vmx_rarr_from_bytes(C.c, 1, s.data(), &h), "vmx_rarr_from_bytes");
  return Rarr(h);
}
Scalar scalar_of(const Ctx& C, const Rarr& r) {
  Scalar s(C.rb);
  check(api.vmx_rarr_to_bytes(r.h, s.data()), "vmx_rarr_to_bytes");
  return s;
}
Scalar scalar_mul(const Ctx& C, const Scalar& a, const Scalar& b) {
  Rarr x = rarr_of(C, a), y = rarr_of(C, b);
  vmx_rarr* h = nullptr;
  check(api.vmx_rmul(x.h, y.h, &h), "vmx_rmul");
  return scalar_of(C, Rarr(h));
}
Scalar scalar_add(const Ctx& C, const Scalar& a, const Scalar& b) {
  Rarr x = rarr_of(C, a), y = rarr_of(C, b);
  vmx_rarr* h = nullptr;
  check(api.vmx_radd(x.h, y.h, &h), "vmx_radd");
  return scalar_of(C, Rarr(h));
}

// ---- single group elements
Elem elem_exp(const Ctx& C, const Elem& b, const Scalar& e) {
  Elem r(C.eb);
  check(api.vmx_elem_exp(C.c, b.data(), e.data(), r.data()), "vmx_elem_exp");
  return r;
}
Elem elem_inv(const Ctx& C, const Elem& b) {
  Elem r(C.eb);
  check(api.vmx_elem_inv(C.c, b.data(), r.data()), "vmx_elem_inv");
  return r;
}
Elem elem_mul(const Ctx& C, const Elem& a, const Elem& b) {
  vmx_garr *x = nullptr, *y = nullptr, *z = nullptr;
  check(api.vmx_garr_fill(C.c, 1, a.data(), &x), "vmx_garr_fill");
  Garr gx(x);
  check(api.vmx_garr_fill(C.c, 1, b.data(), &y), "vmx_garr_fill");
  Garr gy(y);
  check(api.vmx_mul(x, y, &z), "vmx_mul");
  Garr gz(z);
  Elem r(C.eb);
  check(api.vmx_get(z, 0, r.data()), "vmx_get");
  return r;
}
Elem elem_get(const Ctx& C, const Garr& a, size_t i) {
  Elem r(C.eb);
  check(api.vmx_get(a.h, i, r.data()), "vmx_get");
  return r;
}
Elem arr_prod(const Ctx& C, const Garr& a) {
  Elem r(C.eb);
  check(api.vmx_prod(a.h, r.data()), "vmx_prod");
  return r;
}

// a single group element: length, range and membership as PGroup.toElement (the engine's import of a one-element
// array checks all three; a curve point arrives as node(leaf x, leaf y) and is checked against the curve equation)
Elem parse_elem(const Ctx& C, Span t) {
  if (C.curve) {
    const std::vector<Span> xy = children(t, 2);
    const Span x = leaf_payload(xy[0], C.cb), y = leaf_payload(xy[1], C.cb);
    Bytes buf;   // the array form of one point: node(1 leaf x), node(1 leaf y)
    for (Span c : {x, y}) {
      const Bytes n1 = header(NODE, 1), lf = header(LEAF, (uint32_t)C.cb);
      buf.insert(buf.end(), n1.begin(), n1.end());
      buf.insert(buf.end(), lf.begin(), lf.end());
      buf.insert(buf.end(), c.p, c.p + c.n);
    }
    vmx_garr* a = nullptr;
    check(api.vmx_garr_from_leaves(C.c, 1, buf.data(), 1, &a), "point");
    Garr g(a);
    Elem e(x.p, x.p + x.n);
    e.insert(e.end(), y.p, y.p + y.n);
    return e;
  }
  const Hdr h = read_hdr(t, 0);
  if (h.kind != LEAF || h.count != C.eb || t.n < 5 + C.eb) throw Malformed("group element of wrong length");
  vmx_garr* a = nullptr;
  check(api.vmx_garr_from_leaves(C.c, 1, t.p, 1, &a), "element");
  Garr g(a);
  return Elem(t.p + 5, t.p + 5 + C.eb);
}
Scalar parse_scalar(const Ctx& C, Span leaf) {
  const Span v = leaf_payload(leaf, C.rb);
  Scalar s(v.p, v.p + v.n);
  if (cmp_be(s, C.q) >= 0) throw Malformed("ring element out of range");
  return s;
}
// bytes of the serialised leaves of an array of n elements as the engine reads / writes them
size_t garr_leaves_bytes(const Ctx& C, size_t n) { return C.curve ? 2 * (5 + n * (5 + C.cb)) : n * (5 + C.eb); }
// an array of `n` group elements out of node(n leaves) -- over a curve node(node(n x leaves), node(n y leaves)):
// inner headers, leaf headers, range and membership are checked on the device
Garr parse_garr(const Ctx& C, Span node, size_t n) {
  const Hdr h = read_hdr(node, 0);
  if (h.kind != NODE || h.count != (C.curve ? 2 : n)) throw Malformed("array of the wrong size");
  if (C.curve && n > (node.n / 2)) throw Malformed("truncated array");
  const size_t bytes = garr_leaves_bytes(C, n);
  if (node.n < 5 + bytes) throw Malformed("truncated array");
  vmx_garr* a = nullptr;
  check(api.vmx_garr_from_leaves(C.c, n, node.p + 5, 1, &a), "array");
  return Garr(a);
}
Rarr parse_rarr(const Ctx& C, Span node, size_t n) {
  const Hdr h = read_hdr(node, 0);
  if (h.kind != NODE || h.count != n) throw Malformed("array of the wrong size");
  const size_t bytes = n * (5 + C.rb);
  if (node.n < 5 + bytes) throw Malformed("truncated array");
  vmx_rarr* a = nullptr;
  check(api.vmx_rarr_from_leaves(C.c, n, node.p + 5, &a), "ring array");
  return Rarr(a);
}
size_t garr_tree_bytes(const Ctx& C, size_t n) { return 5 + garr_leaves_bytes(C, n); }

// serialisation of an array the way toByteTree() writes it (node header + leaves), kept in `store`
Span garr_tree(const Ctx& C, const Garr& a, size_t n, std::vector<std::unique_ptr<Bytes>>& store) {
  auto buf = std::make_unique<Bytes>(garr_tree_bytes(C, n));
  (*buf)[0] = NODE;
  put_be32(buf->data() + 1, (uint32_t)(C.curve ? 2 : n));
  if (n || C.curve) check(api.vmx_garr_to_leaves(a.h, buf->data() + 5), "vmx_garr_to_leaves");
  store.push_back(std::move(buf));
  return Span(store.back()->data(), store.back()->size());
}
Bytes leaf_tree(const Bytes& e) {
  Bytes t = header(LEAF, (uint32_t)e.size());
  t.insert(t.end(), e.begin(), e.end());
  return t;
}
Bytes elem_tree(const Ctx& C, const Elem& e) {
  if (!C.curve) return leaf_tree(e);
  Bytes t = header(NODE, 2);
  for (int half = 0; half < 2; half++) {
    const Bytes lf = header(LEAF, (uint32_t)C.cb);
    t.insert(t.end(), lf.begin(), lf.end());
    t.insert(t.end(), e.begin() + (long)(half * C.cb), e.begin() + (long)((half + 1) * C.cb));
  }
  return t;
}
size_t elem_tree_bytes(const Ctx& C) { return C.curve ? 5 + 2 * (5 + C.cb) : 5 + C.eb; }
Garr garr_fill(const Ctx& C, size_t n, const Elem& e) {
  vmx_garr* a = nullptr;
  check(api.vmx_garr_fill(C.c, n, e.data(), &a), "vmx_garr_fill");
  return Garr(a);
}
bool garr_equals(const Garr& a, const Garr& b) {
  int eq = 0;
  check(api.vmx_equals(a.h, b.h, &eq), "vmx_equals");
  return eq != 0;
}
Garr garr_mul(const Garr& a, const Garr& b) {
  vmx_garr* o = nullptr;
  check(api.vmx_mul(a.h, b.h, &o), "vmx_mul");
  return Garr(o);
}

// ---- product structure: a "plain" value has `width` components, a ciphertext 2 * width (first all u, then all v)
struct PlainArr { std::vector<Garr> c; };                  // width arrays
struct CiphArr { std::vector<Garr> c; };                   // 2 * width arrays: u_0..u_{w-1}, v_0..v_{w-1}
using PlainElem = std::vector<Elem>;

// node structure of a plain-group value: width 1 -> the value itself, else node(width values)
std::vector<Span> plain_parts(Span s, int width) {
  if (width == 1) return {s};
  return children(s, width);
}
PlainArr parse_plain_arr(const Ctx& C, Span s, int width, size_t n) {
  PlainArr out;
  for (Span part : plain_parts(s, width)) out.c.push_back(parse_garr(C, part, n));
  return out;
}
CiphArr parse_ciph_arr(const Ctx& C, Span s, int width, size_t n) {
  CiphArr out;
  for (Span half : children(s, 2))
    for (Span part : plain_parts(half, width)) out.c.push_back(parse_garr(C, part, n));
  return out;
}
PlainElem parse_plain_elem(const Ctx& C, Span s, int width) {
  PlainElem out;
  for (Span part : plain_parts(s, width)) out.push_back(parse_elem(C, part));
  return out;
}
size_t plain_arr_tree_bytes(const Ctx& C, int width, size_t n) { return width == 1 ? garr_tree_bytes(C, n) : 5 + width * garr_tree_bytes(C, n); }
size_t ciph_arr_tree_bytes(const Ctx& C, int width, size_t n) { return 5 + 2 * plain_arr_tree_bytes(C, width, n); }

void hash_plain_elem(const Ctx& C, Oracle& o, const PlainElem& e) {
  if (e.size() > 1) o.update_owned(header(NODE, (uint32_t)e.size()));
  for (const Elem& x : e) o.update_owned(elem_tree(C, x));
}
void hash_plain_arr(const Ctx& C, Oracle& o, const std::vector<Garr>& comps, size_t first, int width, size_t n,
                    std::vector<std::unique_ptr<Bytes>>& store) {
  if (width > 1) o.update_owned(header(NODE, (uint32_t)width));
  for (int i = 0; i < width; i++) o.update(garr_tree(C, comps[first + i], n, store));
}
void hash_ciph_arr(const Ctx& C, Oracle& o, const CiphArr& a, int width, size_t n, std::vector<std::unique_ptr<Bytes>>& store) {
  o.update_owned(header(NODE, 2));
  hash_plain_arr(C, o, a.c, 0, width, n, store);
  hash_plain_arr(C, o, a.c, (size_t)width, width, n, store);
}

// the batching vector of a proof: n integers of `bits` bits from PRG(seed) as elements of Z_q
// (hvzk/PoSBasicTW.java:533-538); integers as wide as q (a 256-bit curve order) are reduced as they are drawn
Rarr batch_vector(const Ctx& C, const Bytes& seed, size_t n, unsigned bits) {
  size_t qbits = 8 * C.q.size();
  for (size_t i = 0; i < C.q.size(); i++) if (C.q[i]) { qbits = 8 * (C.q.size() - i); for (uint8_t v = C.q[i]; !(v & 0x80); v <<= 1) qbits--; break; }
  vmx_rarr* h = nullptr;
  if (bits < qbits) check(api.vmx_rarr_prg_sha256(C.c, seed.data(), seed.size(), 0, n, bits, &h), "vmx_rarr_prg_sha256");
  else check(api.vmx_rarr_prg_raw_sha256(C.c, seed.data(), seed.size(), 0, n, (bits + 7) / 8, bits, &h), "vmx_rarr_prg_raw_sha256");
  return Rarr(h);
}

// prod_i arrays[j][i]^e[i] for every j in one engine call (the exponent digits are sorted once)
std::vector<Elem> expprod_many(const Ctx& C, const std::vector<const Garr*>& arrays, const Rarr& e) {
  std::vector<const vmx_garr*> hs;
  for (const Garr* a : arrays) hs.push_back(a->h);
  Bytes out(arrays.size() * C.eb);
  check(api.vmx_expprod(hs.data(), hs.size(), e.h, out.data()), "vmx_expprod");
  std::vector<Elem> r;
  for (size_t j = 0; j < arrays.size(); j++) r.emplace_back(out.begin() + (long)(j * C.eb), out.begin() + (long)((j + 1) * C.eb));
  return r;
}

// ------------------------------------------------------------------------------------------------ the session
struct Session {
  Ctx& C;
  explicit Session(Ctx& c) : C(c) {}
  const vmxv_params* P;
  std::map<std::string, Span> files;
  Bytes prefix;  // rho
  int width = 1;
  int party = 0;   // the party whose proof is being verified (test vectors)
  uint64_t hashed = 0;

  Span file(const std::string& name) const {
    auto it = files.find(name);
    if (it == files.end()) fail_stop("Can not find %s in proof directory!", name.c_str());
    return it->second;
  }
  bool has(const std::string& name) const { return files.count(name) != 0; }
  std::string text(const std::string& name) const { const Span s = file(name); return std::string((const char*)s.p, s.n); }

  Bytes challenge_finish(Oracle& o) { Bytes r = o.finish(); hashed += o.hashed; return r; }

  // the scalar test vectors of `vmnv -t` (mixnet/MixNetElGamalVerifyFiatShamirTool.java:82-224) in the order the
  // reference prints them, one per line: name '@' party (0: none) '=' value
  std::string vectors;
  void record(const char* name, int party, const std::string& value) {
    vectors += std::string(name) + "@" + std::to_string(party) + "=" + value + "\n";
  }
  static std::string hex_of(const Bytes& b) {
    static const char* d = "0123456789abcdef";
    std::string o;
    for (uint8_t v : b) { o += d[v >> 4]; o += d[v & 15]; }
    return o;
  }
  static std::string decimal_of(Bytes b) {   // a non-negative big-endian integer (LargeInteger.toString)
    std::string o;
    size_t first = 0;
    while (first < b.size()) {
      unsigned rem = 0;
      for (size_t i = first; i < b.size(); i++) { const unsigned cur = rem * 256 + b[i]; b[i] = (uint8_t)(cur / 10); rem = cur % 10; }
      o += (char)('0' + rem);
      while (first < b.size() && b[first] == 0) first++;
    }
    if (o.empty()) o = "0";
    return std::string(o.rbegin(), o.rend());
  }

  // ---- hvzk/PoSTW.java:177-260 over hvzk/PoSBasicTW.java: one proof of a shuffle
  // (w == nullptr: hvzk/PoSCTW.java:137-210 over hvzk/PoSCBasicTW.java -- the same proof without the ciphertexts)
  bool verify_shuffle(const Garr& h, Span hTree, const Elem& h0, size_t n, const Garr& u, Span uTree, const CiphArr* w,
                      const CiphArr* wp, Span wFile, Span wpFile, Span commitFile, Span replyFile, const Elem& y);
  // ---- hvzk/CCPoSW.java:160-260 over hvzk/CCPoSBasicW.java: one commitment-consistent proof of a shuffle
  bool verify_ccpos(const Garr& h, Span hTree, size_t n, const Garr& u, Span uTree, const CiphArr& w, const CiphArr& wp,
                    Span wFile, Span wpFile, Span commitFile, Span replyFile, const Elem& y);

  void run(vmxv_report* rep);
};

// Integer.parseInt: an optional sign and decimal digits, nothing else (no blanks, no trailing bytes)
bool parse_int_strict(const std::string& s, long* out) {
  size_t i = (!s.empty() && (s[0] == '+' || s[0] == '-')) ? 1 : 0;
  if (i >= s.size() || s.size() > 11) return false;
  for (size_t j = i; j < s.size(); j++) if (s[j] < '0' || s[j] > '9') return false;
  *out = std::strtol(s.c_str(), nullptr, 10);
  return true;
}

std::string two(int l) { char b[16]; snprintf(b, sizeof b, "%02d", l); return b; }

bool Session::verify_shuffle(const Garr& h, Span hTree, const Elem& h0, size_t n, const Garr& u, Span uTree,
                             const CiphArr* wq, const CiphArr* wpq, Span wFile, Span wpFile, Span commitFile,
                             Span replyFile, const Elem& y) {
  const bool ciph = wq != nullptr;   // PoS (ciphertexts) or PoSC (commitments only)
  const int W = width, K = ciph ? 2 * W : 0;
  std::vector<std::unique_ptr<Bytes>> store;  // serialisations that must outlive the hashing
  // seed = RO(rho || node(g, h, u, pk, w, w'))  (PoSTW.java:118-124), hashed beside the imports below;
  //        RO(rho || node(g, h, u)) for a proof of a shuffle of commitments (PoSCTW.java:90-92)
  Oracle seedO(prefix, 256);
  seedO.update_owned(header(NODE, ciph ? 6 : 3));
  seedO.update_owned(elem_tree(C, C.g));
  seedO.update(hTree);
  seedO.update(uTree);
  if (ciph) {
    seedO.update_owned(header(NODE, 2));
    PlainElem gs((size_t)W, C.g), ys((size_t)W, y);
    hash_plain_elem(C, seedO, gs);
    hash_plain_elem(C, seedO, ys);
    // w and w' were parsed from these files (fail-stop otherwise); their trees are the files when canonical
    if (wFile.n == ciph_arr_tree_bytes(C, W, n)) seedO.update(wFile); else hash_ciph_arr(C, seedO, *wq, W, n, store);
    if (wpFile.n == ciph_arr_tree_bytes(C, W, n)) seedO.update(wpFile); else hash_ciph_arr(C, seedO, *wpq, W, n, store);
  }

  // commitment (:780-823): node(B, A', B', C', D', F'); anything malformed -> all trivial
  Garr B, Bp;
  Elem Ap, Cp, Dp;
  std::vector<Elem> Fp;  // 2 * width components
  bool malformed = false;
  try {
    const std::vector<Span> ch = first_children(commitFile, ciph ? 6 : 5);
    B = parse_garr(C, ch[0], n);
    Ap = parse_elem(C, ch[1]);
    Bp = parse_garr(C, ch[2], n);
    Cp = parse_elem(C, ch[3]);
    Dp = parse_elem(C, ch[4]);
    if (ciph)
      for (Span half : children(ch[5], 2))
        for (Span part : plain_parts(half, W)) Fp.push_back(parse_elem(C, part));
  } catch (const Malformed&) {
    malformed = true;
  }
  if (malformed) {
    B = garr_fill(C, n, C.one);
    Bp = garr_fill(C, n, C.one);
    Ap = Cp = Dp = C.one;
    Fp.assign((size_t)K, C.one);
  }
  // replies (:970-990): node(k_A, k_B, k_C, k_D, k_E, k_F); malformed -> reject once the challenge is derived
  Scalar kA, kC, kD;
  std::vector<Scalar> kF;  // width components
  Rarr kB, kE;
  bool parsed = true;
  try {
    const std::vector<Span> ch = first_children(replyFile, ciph ? 6 : 5);
    kA = parse_scalar(C, ch[0]);
    kB = parse_rarr(C, ch[1], n);
    kC = parse_scalar(C, ch[2]);
    kD = parse_scalar(C, ch[3]);
    kE = parse_rarr(C, ch[4], n);
    if (ciph)
      for (Span part : plain_parts(ch[5], W)) kF.push_back(parse_scalar(C, part));
  } catch (const Malformed&) {
    parsed = false;
  }

  // everything in the five checks that is a function of the proof alone, queued while the seed is hashed
  Elem Cc, rightA, rightC, rightD;
  std::vector<Elem> rightF;
  Garr rightB, BshiftInv;
  if (parsed) {
    Cc = elem_mul(C, arr_prod(C, u), elem_inv(C, arr_prod(C, h)));                      // :1013
    std::vector<const Garr*> arrs{&h};
    if (ciph) for (const Garr& a : wpq->c) arrs.push_back(&a);
    const std::vector<Elem> pe = expprod_many(C, arrs, kE);                              // :1021, :1063
    rightA = elem_mul(C, elem_exp(C, C.g, kA), pe[0]);
    vmx_garr* t = nullptr;
    check(api.vmx_exp_fixed(C.c, C.g.data(), kB.h, &t), "vmx_exp_fixed");                // :1030
    rightB = Garr(t);
    check(api.vmx_shift_push(B.h, h0.data(), &t), "vmx_shift_push");                     // :1031
    Garr Bshift(t);
    check(api.vmx_inv(Bshift.h, &t), "vmx_inv");
    BshiftInv = Garr(t);
    rightC = elem_exp(C, C.g, kC);                                                       // :1048
    rightD = elem_exp(C, C.g, kD);                                                       // :1055
    for (int i = 0; i < K; i++) {                                                        // :1063 pk^-k_F * prod w'^k_E
      const Scalar nk = scalar_neg(C, kF[(size_t)(i % W)]);
      rightF.push_back(elem_mul(C, elem_exp(C, i < W ? C.g : y, nk), pe[(size_t)(1 + i)]));
    }
  }
  const Bytes prgSeed = challenge_finish(seedO);

  // batching vector (:533-538) and challenge v = RO(rho || node(leaf(seed), commitment))  (PoSTW.java:146-147)
  Rarr e = batch_vector(C, prgSeed, n, (unsigned)P->ebitlenro);
  Oracle chalO(prefix, (unsigned)P->vbitlenro);
  chalO.update_owned(header(NODE, 2));
  chalO.update_owned(leaf_tree(prgSeed));
  const size_t plainElemBytes = W == 1 ? elem_tree_bytes(C) : 5 + (size_t)W * elem_tree_bytes(C);
  const bool commit_canonical = !malformed && read_hdr(commitFile, 0).count == (ciph ? 6u : 5u) &&
      commitFile.n == 5 + 2 * garr_tree_bytes(C, n) + 3 * elem_tree_bytes(C) + (ciph ? 5 + 2 * plainElemBytes : 0);
  if (commit_canonical) {
    chalO.update(commitFile);
  } else {
    chalO.update_owned(header(NODE, ciph ? 6 : 5));
    chalO.update(garr_tree(C, B, n, store));
    chalO.update_owned(elem_tree(C, Ap));
    chalO.update(garr_tree(C, Bp, n, store));
    chalO.update_owned(elem_tree(C, Cp));
    chalO.update_owned(elem_tree(C, Dp));
    if (ciph) {
      chalO.update_owned(header(NODE, 2));
      for (int half = 0; half < 2; half++) {
        PlainElem pe_(Fp.begin() + half * W, Fp.begin() + (half + 1) * W);
        hash_plain_elem(C, chalO, pe_);
      }
    }
  }
  // A = prod u^e, F = prod w^e  (:407-410), while the challenge is hashed
  std::vector<const Garr*> arrs{&u};
  if (ciph) for (const Garr& a : wq->c) arrs.push_back(&a);
  const std::vector<Elem> AF = expprod_many(C, arrs, e);
  const Bytes vBytes = challenge_finish(chalO);
  record(ciph ? "PoS.s" : "PoSC.s", party, hex_of(prgSeed));
  record(ciph ? "PoS.v" : "PoSC.v", party, decimal_of(vBytes));
  if (!parsed) return false;
  const Scalar v = scalar_from_bytes(C, vBytes);

  // the five checks (:1008-1066)
  Scalar eprod(C.rb);
  check(api.vmx_rprod(e.h, eprod.data()), "vmx_rprod");
  const Elem D = elem_mul(C, elem_get(C, B, n - 1), elem_inv(C, elem_exp(C, h0, eprod)));          // :1014
  const bool okA = elem_mul(C, elem_exp(C, AF[0], v), Ap) == rightA;                               // :1020-1021
  vmx_garr* t = nullptr;
  check(api.vmx_exp_scalar_var(B.h, v.data(), BshiftInv.h, kE.h, &t), "vmx_exp_scalar_var");      // :1028,1032
  Garr both(t);
  Garr left = garr_mul(both, Bp);
  const bool okB = garr_equals(left, rightB);                                                      // :1035
  const bool okC = elem_mul(C, elem_exp(C, Cc, v), Cp) == rightC;                                  // :1048
  const bool okD = elem_mul(C, elem_exp(C, D, v), Dp) == rightD;                                   // :1055
  bool okF = true;
  for (int i = 0; i < K; i++) okF = okF && elem_mul(C, elem_exp(C, AF[(size_t)(1 + i)], v), Fp[(size_t)i]) == rightF[(size_t)i];   // :1062-1063
  return okA && okB && okC && okD && okF;
}

bool Session::verify_ccpos(const Garr& h, Span hTree, size_t n, const Garr& u, Span uTree, const CiphArr& w, const CiphArr& wp,
                           Span wFile, Span wpFile, Span commitFile, Span replyFile, const Elem& y) {
  const int W = width, K = 2 * W;
  std::vector<std::unique_ptr<Bytes>> store;
  // seed = RO(rho || node(g, h, u, pk, w, w'))  (CCPoSW.java:92-98)
  Oracle seedO(prefix, 256);
  seedO.update_owned(header(NODE, 6));
  seedO.update_owned(elem_tree(C, C.g));
  seedO.update(hTree);
  seedO.update(uTree);
  seedO.update_owned(header(NODE, 2));
  {
    PlainElem gs((size_t)W, C.g), ys((size_t)W, y);
    hash_plain_elem(C, seedO, gs);
    hash_plain_elem(C, seedO, ys);
  }
  if (wFile.n == ciph_arr_tree_bytes(C, W, n)) seedO.update(wFile); else hash_ciph_arr(C, seedO, w, W, n, store);
  if (wpFile.n == ciph_arr_tree_bytes(C, W, n)) seedO.update(wpFile); else hash_ciph_arr(C, seedO, wp, W, n, store);
  // commitment (CCPoSBasicW.java:408-431): node(A', B'), B' a ciphertext; malformed -> both trivial
  Elem Ap;
  std::vector<Elem> Bp;  // 2 * width components
  try {
    const std::vector<Span> ch = first_children(commitFile, 2);
    Ap = parse_elem(C, ch[0]);
    for (Span half : children(ch[1], 2))
      for (Span part : plain_parts(half, W)) Bp.push_back(parse_elem(C, part));
  } catch (const Malformed&) {
    Ap = C.one;
    Bp.assign((size_t)K, C.one);
  }
  // reply (:519-552): node(k_A, k_B, k_E); malformed -> reject once the challenge is derived
  Scalar kA;
  std::vector<Scalar> kB;  // width components
  Rarr kE;
  bool parsed = true;
  try {
    const std::vector<Span> ch = first_children(replyFile, 3);
    kA = parse_scalar(C, ch[0]);
    for (Span part : plain_parts(ch[1], W)) kB.push_back(parse_scalar(C, part));
    kE = parse_rarr(C, ch[2], n);
  } catch (const Malformed&) {
    parsed = false;
  }
  // the right-hand sides g^k_A * prod h^k_E and pk^-k_B * prod w'^k_E (:554-579), queued while the seed is hashed
  Elem rightA;
  std::vector<Elem> rightB;
  if (parsed) {
    std::vector<const Garr*> arrs{&h};
    for (const Garr& a : wp.c) arrs.push_back(&a);
    const std::vector<Elem> pe = expprod_many(C, arrs, kE);
    rightA = elem_mul(C, elem_exp(C, C.g, kA), pe[0]);
    for (int i = 0; i < K; i++) {
      const Scalar nk = scalar_neg(C, kB[(size_t)(i % W)]);
      rightB.push_back(elem_mul(C, elem_exp(C, i < W ? C.g : y, nk), pe[(size_t)(1 + i)]));
    }
  }
  const Bytes prgSeed = challenge_finish(seedO);
  Rarr e = batch_vector(C, prgSeed, n, (unsigned)P->ebitlenro);
  Oracle chalO(prefix, (unsigned)P->vbitlenro);
  chalO.update_owned(header(NODE, 2));
  chalO.update_owned(leaf_tree(prgSeed));
  chalO.update_owned(header(NODE, 2));
  chalO.update_owned(elem_tree(C, Ap));
  chalO.update_owned(header(NODE, 2));
  for (int half = 0; half < 2; half++) {
    PlainElem pe_(Bp.begin() + half * W, Bp.begin() + (half + 1) * W);
    hash_plain_elem(C, chalO, pe_);
  }
  // A = prod u^e, B = prod w^e  (:493-506), while the challenge is hashed
  std::vector<const Garr*> arrs{&u};
  for (const Garr& a : w.c) arrs.push_back(&a);
  const std::vector<Elem> AB = expprod_many(C, arrs, e);
  const Bytes vBytes = challenge_finish(chalO);
  record("CCPoS.s", party, hex_of(prgSeed));
  record("CCPoS.v", party, decimal_of(vBytes));
  if (!parsed) return false;
  const Scalar v = scalar_from_bytes(C, vBytes);
  bool ok = elem_mul(C, elem_exp(C, AB[0], v), Ap) == rightA;
  for (int i = 0; i < K; i++) ok = ok && elem_mul(C, elem_exp(C, AB[(size_t)(1 + i)], v), Bp[(size_t)i]) == rightB[(size_t)i];
  return ok;
}

// modified Lagrange coefficients (elgamal/DistrElGamalSessionBasic.java:290-452): small signed integers
static const int kOddPrimeMax = 1009;
uint64_t prime_log(uint64_t number, uint64_t prime) { uint64_t a = 1, b = 1; while (b <= number) { a = b; b *= prime; } return a; }
bool is_odd_prime(int n) { if (n < 3 || n % 2 == 0) return false; for (int d = 3; d * d <= n; d += 2) if (n % d == 0) return false; return true; }
Scalar prod_factor(const Ctx& C, int k) {
  if (k > kOddPrimeMax) fail_stop("Too many parties!");
  Scalar res = scalar_from_u64(C, 1);
  int prime = 2, next = 3;
  while (prime <= k) {
    res = scalar_mul(C, res, scalar_from_u64(C, prime_log((uint64_t)k, (uint64_t)prime)));
    prime = next;
    do { next += 2; } while (!is_odd_prime(next));
  }
  return scalar_mul(C, res, res);
}
Scalar scalar_small_signed(const Ctx& C, int v) {  // v mod q for a small integer
  return v >= 0 ? scalar_from_u64(C, (uint64_t)v) : scalar_neg(C, scalar_from_u64(C, (uint64_t)(-v)));
}
Scalar inv_small_signed(const Ctx& C, int v) {
  const Scalar i = inv_small(C, (uint32_t)(v < 0 ? -v : v));
  return v < 0 ? scalar_neg(C, i) : i;
}
// the coefficient as the integer of smallest absolute value representing it, and as a scalar
bool small_signed_of(const Ctx& C, const Scalar& s, int64_t* out) {
  auto fits = [&](const Bytes& b, int64_t* v) {
    for (size_t i = 0; i + 8 < b.size(); i++) if (b[i]) return false;
    uint64_t x = 0;
    for (size_t i = b.size() >= 8 ? b.size() - 8 : 0; i < b.size(); i++) x = (x << 8) | b[i];
    if (x >> 62) return false;
    *v = (int64_t)x;
    return true;
  };
  int64_t v;
  if (fits(s, &v)) { *out = v; return true; }
  if (fits(sub_be(C.q, s), &v)) { *out = -v; return true; }
  return false;
}
std::vector<Scalar> lagrange(const Ctx& C, const std::vector<bool>& correct, int k, int threshold) {
  const Scalar pf = prod_factor(C, k);
  std::vector<Scalar> out;
  for (int i = 1; (int)out.size() < threshold && i <= k; i++) {
    if (!correct[(size_t)i]) continue;
    Scalar res = pf;
    int t = 0;
    for (int l = 1; t < threshold && l <= k; l++) {
      if (!correct[(size_t)l]) continue;
      if (l != i) {
        res = scalar_mul(C, res, scalar_small_signed(C, l));
        res = scalar_mul(C, res, inv_small_signed(C, l - i));
      }
      t++;
    }
    out.push_back(res);
  }
  if ((int)out.size() < threshold) fail_stop("Attempting to combine too few decryption factors!");
  return out;
}

void Session::run(vmxv_report* rep) {
  const int k = P->k, threshold = P->threshold;
  const size_t kVerdicts = sizeof rep->shuffles / sizeof rep->shuffles[0];
  // ---- header files (MixNetElGamalVerifyFiatShamirSession.java:1318-1360)
  if (text("version") != P->version) fail_stop("Mismatching versions!");
  // determineType :329-358, determineSessionParams :984-1005
  const std::string type = text("type");
  if (type != "mixing" && type != "shuffling" && type != "decryption") fail_stop("Unknown type of proof!");
  if (P->expected_type && P->expected_type[0] && type != P->expected_type)
    fail_stop("Attempting to verify proof of %s, but proof is a proof of %s!", P->expected_type, type.c_str());
  rep->type = type == "mixing" ? 0 : type == "shuffling" ? 1 : 2;
  const std::string auxsid = text("auxsid");
  bool sid_ok = !auxsid.empty() && auxsid.size() <= 1024;
  for (char ch : auxsid) sid_ok = sid_ok && (std::isalnum((unsigned char)ch) || ch == '_' || ch == ' ') && (unsigned char)ch < 128;
  if (!sid_ok) fail_stop("Can not read auxsid from file!");
  if (P->expected_auxsid && P->expected_auxsid[0] && auxsid != P->expected_auxsid)
    fail_stop("The given auxiliary session identifier does not match the one in the proof!");
  record("par.k", 0, std::to_string(k));
  record("par.lambda", 0, std::to_string(threshold));
  record("par.n_e", 0, std::to_string(P->ebitlenro));
  record("par.n_r", 0, std::to_string(P->rbitlen));
  record("par.n_v", 0, std::to_string(P->vbitlenro));
  record("par.s_Gq", 0, P->pgroup_string);
  record("par.version", 0, P->version);
  bool dec = !P->nodec, posc = !P->noposc, ccpos = !P->noccpos;
  if (type == "shuffling") dec = false;
  else if (type == "decryption") posc = ccpos = false;
  width = 1;
  if (ccpos || dec) {
    long wv = 0;
    if (!parse_int_strict(text("width"), &wv)) fail_stop("Can not parse width given in file!");
    if (wv < 1 || wv > 1024 || (P->expected_width > 0 && wv != P->expected_width)) fail_stop("Mismatching or invalid width!");
    width = (int)wv;
    record("par.omega", 0, std::to_string(width));
  }
  const int W = width;
  // ---- global prefix (:158-189)
  {
    auto sleaf = [](const std::string& s) { Bytes t = header(LEAF, (uint32_t)s.size()); t.insert(t.end(), s.begin(), s.end()); return t; };
    auto ileaf = [](int v) { Bytes t = header(LEAF, 4); t.resize(9); put_be32(t.data() + 5, (uint32_t)v); return t; };
    Bytes t = header(NODE, 8);
    for (const Bytes& part : {sleaf(P->version), sleaf(std::string(P->sid) + "." + auxsid), ileaf(P->rbitlen), ileaf(P->vbitlenro),
                              ileaf(P->ebitlenro), sleaf("PRGHeuristic(SHA-256)"), sleaf(P->pgroup_string),
                              sleaf("HashfunctionHeuristic(SHA-256)")})
      t.insert(t.end(), part.begin(), part.end());
    prefix = sha256(t);
  }
  // ---- keys (:195-266)
  Elem y;
  std::vector<Elem> coeffs;
  try {
    const std::vector<Span> pk = children(file("FullPublicKey.bt"), 2);
    const Elem g0 = parse_elem(C, pk[0]);
    y = parse_elem(C, pk[1]);
    if (g0 != C.g) fail_stop("Basic public key is not the standard generator!");
  } catch (const Malformed&) {
    fail_stop("Could not read full El Gamal public key from file!");
  }
  // window tables of the bases every proof raises to full-length exponents: g (arrays and single elements), the
  // public key y and h0 (single elements: a small table turns a 3071-step ladder on one warp, ~20 ms, into a few
  // multiplications per thread); they stay with the cached context for the next verification
  check(api.vmx_fixed_precompute(C.c, y.data(), 16), "vmx_fixed_precompute");
  if (dec) {   // readMixServerPKeys :228-266
    try {
      for (Span c : children(file("proofs/PolynomialInExponent.bt"), threshold)) coeffs.push_back(parse_elem(C, c));
    } catch (const Malformed&) {
      fail_stop("Unable to read polynomial in exponent from file!");
    }
    if (coeffs[0] != y) fail_stop("Mismatching public keys!");
    std::vector<Elem> pkeys((size_t)k + 1);
    for (int l = 1; l <= k; l++) {  // PolynomialInExponent.evaluate(l) = prod_i coeffs[i]^(l^i), by Horner's rule in the
      Elem acc = coeffs.back();     // exponent (l^i itself leaves 64 bits for a hundred parties and a threshold of 11)
      const Scalar ls = scalar_from_u64(C, (uint64_t)l);
      for (size_t i = coeffs.size() - 1; i-- > 0;) acc = elem_mul(C, elem_exp(C, acc, ls), coeffs[i]);
      pkeys[(size_t)l] = acc;
    }
  }
  record("par.sid", 0, P->sid);
  record("der.rho", 0, hex_of(prefix));
  const bool precomp = has("proofs/maxciph");                                                     // :946-948
  int active = 0;
  {
    long a = 0;
    if (!parse_int_strict(text("proofs/activethreshold"), &a)) fail_stop("Can not parse active threshold given in file!");
    if (a > k || a < threshold) fail_stop("Active threshold out of range!");
    active = (int)a;
    record("par.lambda", 0, std::to_string(active));
  }
  // ---- input ciphertexts (readCiphertexts :1017-1046; readArray with size 0: the size is that of the first array)
  auto read_ciph = [&](Span s, const std::string& name, size_t n) {
    try {
      return parse_ciph_arr(C, s, W, n);
    } catch (const Malformed& e) {
      fail_stop("Unable to read array %s! (%s)", name.c_str(), e.what());
    }
  };
  std::unique_ptr<CiphArr> ciphertexts;
  std::string ctName;
  Span ctFile;
  size_t n = 0;
  if (ccpos || dec) {
    if (ccpos || type == "decryption") {
      ctName = "Ciphertexts.bt";
      file(ctName);
    } else if (has("proofs/Ciphertexts" + two(active) + ".bt")) {
      ctName = "proofs/Ciphertexts" + two(active) + ".bt";
    }
    if (!ctName.empty()) {
      ctFile = file(ctName);
      try {
        Span s = children(ctFile, 2)[0];
        if (W > 1) s = children(s, W)[0];
        if (C.curve) s = children(s, 2)[0];   // an array over a curve is node(x leaves, y leaves)
        const Hdr h = read_hdr(s, 0);
        if (h.kind != NODE) throw Malformed("array expected");
        n = h.count;
      } catch (const Malformed&) {
        fail_stop("Unable to read ciphertexts!");
      }
      if (n == 0) fail_stop("No ciphertexts!");
      ciphertexts = std::make_unique<CiphArr>(read_ciph(ctFile, ctName, n));
    }
  }
  // The seed of the decryption proof is RO(node(node(g, L), node(node(coeffs), node(f_1..f_k)))) (:1586-1600), L the list
  // that is decrypted: it depends on files only, so it is hashed on its worker thread WHILE the shuffles are verified,
  // on the premise that L is the last output file (the last shuffle is valid) and the files canonical; checked
  // below, hashed again if it does not hold.
  std::unique_ptr<Oracle> spec;
  std::string lastName;
  if (dec) {
    lastName = ctName;
    if (ccpos) {
      lastName = "proofs/Ciphertexts" + two(active) + ".bt";
      if (!has(lastName)) lastName = "ShuffledCiphertexts.bt";
    }
    bool all = !lastName.empty() && has(lastName) && n > 0 && files[lastName].n == ciph_arr_tree_bytes(C, W, n);
    for (int l = 1; l <= k && all; l++) {
      const std::string nm = "proofs/DecryptionFactors" + two(l) + ".bt";
      all = has(nm) && files[nm].n == plain_arr_tree_bytes(C, W, n);
    }
    if (all) {
      spec = std::make_unique<Oracle>(prefix, 256);
      spec->update_owned(header(NODE, 2));
      spec->update_owned(header(NODE, 2));
      spec->update_owned(elem_tree(C, C.g));
      spec->update(files[lastName]);
      spec->update_owned(header(NODE, 2));
      spec->update_owned(header(NODE, (uint32_t)coeffs.size()));
      for (const Elem& c : coeffs) spec->update_owned(elem_tree(C, c));
      spec->update_owned(header(NODE, (uint32_t)k));
      for (int l = 1; l <= k; l++) spec->update(files["proofs/DecryptionFactors" + two(l) + ".bt"]);
    }
  }
  // ---- shuffles (:1378-1530)
  const CiphArr* inp = ciphertexts.get();
  Span inpFile = ctFile;
  std::vector<std::unique_ptr<CiphArr>> outputs;
  int valid = 0;
  if (posc || ccpos) {
    size_t maxciph = n;
    if (precomp) {                                                                                // getMaxciph :541-548
      long mc = 0;
      if (!parse_int_strict(text("proofs/maxciph"), &mc)) fail_stop("Can not parse maxciph file!");
      record("par.N_0", 0, std::to_string(mc));
      if (mc < 1) fail_stop("Invalid maxciph!");
      maxciph = (size_t)mc;
    } else if (!ciphertexts) {
      fail_stop("No ciphertexts!");
    }
    // independent generators (distr/IndependentGeneratorsRO.java:110-130)
    Garr h;
    {
      // the serialisation of the generators is hashed into the seed of every proof: a maxciph no file of the
      // directory can answer to is refused before anything of that size is allocated
      size_t largest = 0, pbits0 = 0;
      for (const auto& f : files) largest = f.second.n > largest ? f.second.n : largest;
      for (size_t i = 0; i < P->nbytes; i++) if (P->p_be[i]) { pbits0 = 8 * (P->nbytes - i); for (uint8_t v = P->p_be[i]; !(v & 0x80); v <<= 1) pbits0--; break; }
      if (maxciph > largest / (5 + (pbits0 + 7) / 8)) fail_stop("maxciph exceeds what the proof directory can hold!");
      Sha256 d;
      uint8_t b4[4];
      put_be32(b4, 256);
      d.update(b4, 4);
      d.update(prefix.data(), prefix.size());
      const std::string sid = "generators";
      Bytes leaf = header(LEAF, (uint32_t)sid.size());
      leaf.insert(leaf.end(), sid.begin(), sid.end());
      d.update(leaf.data(), leaf.size());
      const Bytes seed = prg_bytes(d.digest(), 32);
      size_t pbits = 0;
      for (size_t i = 0; i < P->nbytes; i++) if (P->p_be[i]) { pbits = 8 * (P->nbytes - i); for (uint8_t v = P->p_be[i]; !(v & 0x80); v <<= 1) pbits--; break; }
      const size_t bits = pbits + (size_t)P->rbitlen;
      vmx_garr* a = nullptr;
      check(api.vmx_garr_prg_sha256(C.c, seed.data(), seed.size(), 0, maxciph, (bits + 7) / 8, (unsigned)bits, &a), "vmx_garr_prg_sha256");
      h = Garr(a);
    }
    const Elem h0 = elem_get(C, h, 0);
    check(api.vmx_fixed_precompute(C.c, C.g.data(), maxciph), "vmx_fixed_precompute");
    check(api.vmx_fixed_precompute(C.c, h0.data(), 16), "vmx_fixed_precompute");
    std::vector<std::unique_ptr<Bytes>> hStore;
    const Span hTree = garr_tree(C, h, maxciph, hStore);   // the generators are hashed into the seed of every proof
    Garr shrunkH;                                          // getShrunkGenerators :1059-1068
    Span shrunkHTree;
    if (ccpos && precomp) {
      if (n > maxciph) fail_stop("Too few generators have been derived!");
      vmx_garr* a = nullptr;
      check(api.vmx_slice(h.h, 0, n, &a), "vmx_slice");
      shrunkH = Garr(a);
      shrunkHTree = garr_tree(C, shrunkH, n, hStore);
    }
    for (int l = 1; l <= active; l++) {
      bool verdict = true;
      const std::string pcName = "proofs/PermutationCommitment" + two(l) + ".bt";
      if (!((posc && precomp && !ccpos && has(pcName))                                            // getPoSCActive :958
            || has("proofs/CCPoSCommitment" + two(l) + ".bt") || has("proofs/PoSCommitment" + two(l) + ".bt")))   // :972-976
        continue;
      rep->n_shuffles = l;
      party = l;
      // readPermutationCommitment :626-641: fail-stop when missing or malformed
      const Span pcFile = file(pcName);
      Garr u;
      try {
        u = parse_garr(C, pcFile, maxciph);
      } catch (const Malformed& e) {
        fail_stop("Unable to read array %s! (%s)", pcName.c_str(), e.what());
      }
      std::vector<std::unique_ptr<Bytes>> uStore;
      Span uTree = pcFile.n == garr_tree_bytes(C, maxciph) ? pcFile : garr_tree(C, u, maxciph, uStore);
      if (posc && precomp) {                                                                      // verifyPoSC :652-705
        const bool ok = verify_shuffle(h, hTree, h0, maxciph, u, uTree, nullptr, nullptr, Span(), Span(),
                                       file("proofs/PoSCCommitment" + two(l) + ".bt"), file("proofs/PoSCReply" + two(l) + ".bt"), y);
        if ((size_t)l <= kVerdicts) rep->poscs[l - 1] = ok ? 1 : -1;
        if (!ok) {   // "Setting permutation commitment to list of generators."
          verdict = false;
          vmx_garr* cp = nullptr;
          check(api.vmx_slice(h.h, 0, maxciph, &cp), "vmx_slice");
          u = Garr(cp);
          uTree = hTree;
        }
      }
      if (ccpos) {
        std::string name = "proofs/Ciphertexts" + two(l) + ".bt";
        if (l == active && !has(name)) name = "ShuffledCiphertexts.bt";
        const Span outFile = file(name);
        // fail-stop if malformed; an invalid PROOF keeps the input
        outputs.push_back(std::make_unique<CiphArr>(read_ciph(outFile, name, n)));
        bool ok;
        if (precomp) {
          // shrinkPermComm :714-745: a keep list that cannot be read or keeps the wrong number is fail-stop
          Span flags;
          try {
            flags = leaf_payload(file("proofs/KeepList" + two(l) + ".bt"), maxciph);
            for (size_t i = 0; i < maxciph; i++) if (flags.p[i] > 1) throw Malformed("boolean");
          } catch (const Malformed&) {
            fail_stop("Unable to open keeplist of Party %d!", l);
          }
          size_t total = 0;
          for (size_t i = 0; i < maxciph; i++) total += flags.p[i];
          if (total != n) fail_stop("Wrong number of true elements in keep list of Party %d!", l);
          vmx_garr* a = nullptr;
          check(api.vmx_extract(u.h, flags.p, &a), "vmx_extract");
          Garr shrunkU(a);
          u.reset();
          std::vector<std::unique_ptr<Bytes>> suStore;
          const Span suTree = garr_tree(C, shrunkU, n, suStore);
          ok = verify_ccpos(shrunkH, shrunkHTree, n, shrunkU, suTree, *inp, *outputs.back(), inpFile, outFile,
                            file("proofs/CCPoSCommitment" + two(l) + ".bt"), file("proofs/CCPoSReply" + two(l) + ".bt"), y);
          verdict = verdict && ok;
        } else {
          verdict = verify_shuffle(h, hTree, h0, n, u, uTree, inp, outputs.back().get(), inpFile, outFile,
                                   file("proofs/PoSCommitment" + two(l) + ".bt"), file("proofs/PoSReply" + two(l) + ".bt"), y);
        }
        if (verdict) { inp = outputs.back().get(); inpFile = outFile; }
        else outputs.pop_back();
      }
      if ((size_t)l <= kVerdicts) rep->shuffles[l - 1] = verdict ? 1 : -1;   // (all are counted; 0: the party took no part)
      valid += verdict ? 1 : 0;
    }
    rep->valid_proofs = valid;
    rep->enough_valid_proofs = valid >= threshold ? 1 : 0;
  } else {
    rep->enough_valid_proofs = 1;
  }
  if (!dec) {
    rep->decryption = -1;
    rep->plaintexts = -1;
    rep->accepted = rep->enough_valid_proofs;
    return;
  }
  if (!inp) fail_stop("No ciphertexts to decrypt!");
  const CiphArr& mixed = *inp;
  // ---- decryption (:1535-1665)
  std::vector<bool> correct((size_t)k + 1, false);
  try {
    const Span flags = leaf_payload(file("proofs/CorrectIndices.bt"), (size_t)k + 1);
    for (int i = 0; i <= k; i++) {
      if (flags.p[i] > 1) throw Malformed("boolean");
      correct[(size_t)i] = flags.p[i] == 1;
    }
  } catch (const Malformed&) {
    fail_stop("Failed to read indices of correct decryption factors!");
  }
  int ncorrect = 0;
  for (int l = 1; l <= k; l++) ncorrect += correct[(size_t)l] ? 1 : 0;
  if (ncorrect < threshold) fail_stop("Too few correct decryption factors!");
  std::vector<PlainArr> f((size_t)k + 1);
  std::vector<Span> fFile((size_t)k + 1);
  for (int l = 1; l <= k; l++) {
    const std::string nm = "proofs/DecryptionFactors" + two(l) + ".bt";
    fFile[(size_t)l] = file(nm);
    try {
      f[(size_t)l] = parse_plain_arr(C, fFile[(size_t)l], W, n);
    } catch (const Malformed& e) {
      fail_stop("Unable to read array %s! (%s)", nm.c_str(), e.what());
    }
  }
  // combineDecryptionFactors (DistrElGamalSessionBasic.java:465-503)
  const std::vector<Scalar> lam = lagrange(C, correct, k, threshold);
  std::vector<int64_t> lamInt;
  for (const Scalar& s : lam) {
    int64_t v;
    if (!small_signed_of(C, s, &v)) fail_stop("Lagrange coefficient out of range");
    lamInt.push_back(v);
  }
  std::vector<int> used;
  for (int l = 1; (int)used.size() < threshold && l <= k; l++) if (correct[(size_t)l]) used.push_back(l);
  PlainArr combined;
  for (int c = 0; c < W; c++) {
    std::vector<const vmx_garr*> bases;
    for (int l : used) bases.push_back(f[(size_t)l].c[(size_t)c].h);
    vmx_garr* o = nullptr;
    check(api.vmx_expprod_cols(bases.data(), bases.size(), lamInt.data(), &o), "vmx_expprod_cols");
    combined.c.emplace_back(o);
  }
  // seed = RO(rho || node(node(g, L), node(node(coeffs), node(f_1 .. f_k))))  (:1586-1600)
  std::vector<std::unique_ptr<Bytes>> store;
  Bytes prgSeed;
  if (spec && inpFile.p == files[lastName].p) {   // the premise held: the mixed list IS the last output file
    prgSeed = challenge_finish(*spec);
  } else {
    spec.reset();   // (joins its worker)
    Oracle seedO(prefix, 256);
    seedO.update_owned(header(NODE, 2));
    seedO.update_owned(header(NODE, 2));
    seedO.update_owned(elem_tree(C, C.g));
    if (inpFile.n == ciph_arr_tree_bytes(C, W, n)) seedO.update(inpFile); else hash_ciph_arr(C, seedO, mixed, W, n, store);
    seedO.update_owned(header(NODE, 2));
    seedO.update_owned(header(NODE, (uint32_t)coeffs.size()));
    for (const Elem& c : coeffs) seedO.update_owned(elem_tree(C, c));
    seedO.update_owned(header(NODE, (uint32_t)k));
    for (int l = 1; l <= k; l++) {
      if (fFile[(size_t)l].n == plain_arr_tree_bytes(C, W, n)) seedO.update(fFile[(size_t)l]);
      else hash_plain_arr(C, seedO, f[(size_t)l].c, 0, W, n, store);
    }
    prgSeed = challenge_finish(seedO);
  }
  record("Dec.s", 0, hex_of(prgSeed));
  Rarr e = batch_vector(C, prgSeed, n, (unsigned)P->ebitlenro);
  // A = prod u^e, combinedB = prod combined^e  (:524-526, :683-685)
  std::vector<const Garr*> arrs;
  for (int c = 0; c < W; c++) arrs.push_back(&mixed.c[(size_t)c]);
  for (int c = 0; c < W; c++) arrs.push_back(&combined.c[(size_t)c]);
  const std::vector<Elem> AB = expprod_many(C, arrs, e);
  // commitments (:549-565): node(y', B'); malformed -> verdict false, ONE substituted
  std::vector<Elem> yp((size_t)k + 1);
  std::vector<PlainElem> Bp((size_t)k + 1);
  std::vector<Scalar> kx((size_t)k + 1);
  for (int l = 1; l <= k; l++) {
    try {
      const std::vector<Span> ch = first_children(file("proofs/DecrFactCommitment" + two(l) + ".bt"), 2);
      yp[(size_t)l] = parse_elem(C, ch[0]);
      Bp[(size_t)l] = parse_plain_elem(C, ch[1], W);
    } catch (const Malformed&) {
      yp[(size_t)l] = C.one;
      Bp[(size_t)l].assign((size_t)W, C.one);
    }
  }
  Oracle chalO(prefix, (unsigned)P->vbitlenro);
  chalO.update_owned(header(NODE, 2));
  chalO.update_owned(leaf_tree(prgSeed));
  chalO.update_owned(header(NODE, (uint32_t)k));
  for (int l = 1; l <= k; l++) {
    chalO.update_owned(header(NODE, 2));
    chalO.update_owned(elem_tree(C, yp[(size_t)l]));
    hash_plain_elem(C, chalO, Bp[(size_t)l]);
  }
  const Bytes decV = challenge_finish(chalO);
  record("Dec.v", 0, decimal_of(decV));
  const Scalar v = scalar_from_bytes(C, decV);
  for (int l = 1; l <= k; l++) {
    try {
      kx[(size_t)l] = parse_scalar(C, file("proofs/DecrFactReply" + two(l) + ".bt"));
    } catch (const Malformed&) {
      kx[(size_t)l] = scalar_from_u64(C, 0);
    }
  }
  // combine (:642-678) and the combined check (:693-700)
  Elem cyp = C.one;
  PlainElem cBp((size_t)W, C.one);
  Scalar ckx = scalar_from_u64(C, 0);
  for (size_t t = 0; t < used.size(); t++) {
    const int l = used[t];
    cyp = elem_mul(C, cyp, elem_exp(C, yp[(size_t)l], lam[t]));
    for (int c = 0; c < W; c++) cBp[(size_t)c] = elem_mul(C, cBp[(size_t)c], elem_exp(C, Bp[(size_t)l][(size_t)c], lam[t]));
    ckx = scalar_add(C, ckx, scalar_mul(C, kx[(size_t)l], lam[t]));
  }
  bool decOk = elem_mul(C, elem_exp(C, elem_inv(C, y), v), cyp) == elem_exp(C, C.g, ckx);
  for (int c = 0; c < W; c++)
    decOk = decOk && elem_mul(C, elem_exp(C, AB[(size_t)(W + c)], v), cBp[(size_t)c]) == elem_exp(C, AB[(size_t)c], ckx);
  rep->decryption = decOk ? 1 : 0;
  if (!decOk) fail_stop("Verify combined proof of decryption... failed!");
  // ---- plaintexts (:1267-1275)
  PlainArr plain;
  try {
    plain = parse_plain_arr(C, file("Plaintexts.bt"), W, n);
  } catch (const Malformed& e) {
    fail_stop("Unable to read array Plaintexts.bt! (%s)", e.what());
  }
  bool match = true;
  for (int c = 0; c < W; c++) {
    Garr computed = garr_mul(mixed.c[(size_t)(W + c)], combined.c[(size_t)c]);
    match = match && garr_equals(plain.c[(size_t)c], computed);
  }
  rep->plaintexts = match ? 1 : 0;
  if (!match) fail_stop("Plaintexts are incorrect!");
  rep->accepted = rep->enough_valid_proofs;
}

}  // namespace

extern "C" {

int vmxv_bind(const char* libvmx_path) {
  if (api.handle) return 0;
  void* h = dlopen(libvmx_path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) { fprintf(stderr, "vmxv_bind: %s\n", dlerror()); return -1; }
#define X(name)                                                                            \
  api.name = reinterpret_cast<decltype(api.name)>(dlsym(h, #name));                        \
  if (!api.name) { fprintf(stderr, "vmxv_bind: %s misses %s\n", libvmx_path, #name); dlclose(h); return -2; }
  VMXV_SYMBOLS(X)
#undef X
  api.handle = h;
  return 0;
}

int vmxv_verify(const vmxv_params* P, const vmxv_file* files, size_t nfiles, vmxv_report* rep) {
  if (!P || !rep || (nfiles && !files)) return -1;
  std::memset(rep, 0, sizeof *rep);
  rep->decryption = rep->plaintexts = -1;
  if (!api.handle) { snprintf(rep->error, sizeof rep->error, "vmxv_bind was not called"); return -1; }
  {  // usage errors: nothing of the directory is looked at
    const char* bad = nullptr;
    if (P->kind != 0 && P->kind != 1) bad = "kind must be 0 (ModPGroup) or 1 (ECqPGroup)";
    else if (!P->p_be || !P->q_be || !P->g_be || P->nbytes == 0 || P->nbytes > (1u << 16)) bad = "group parameters missing";
    else if (P->kind == 1 && (!P->a_be || !P->b_be || !P->gy_be)) bad = "curve parameters missing";
    else if (P->k < 1 || P->k > kOddPrimeMax || P->threshold < 1 || P->threshold > P->k) bad = "parties / threshold out of range";
    else if (P->vbitlenro < 1 || P->ebitlenro < 1 || P->rbitlen < 0 || P->vbitlenro > 4096 || P->ebitlenro > 4096 ||
             P->rbitlen > 4096) bad = "bit lengths out of range";
    else if (!P->version || !P->sid || !P->pgroup_string) bad = "version / sid / pgroup string missing";
    for (size_t i = 0; i < nfiles && !bad; i++)
      if (!files[i].name || (files[i].size && !files[i].data)) bad = "a file without name or data";
    if (bad) { snprintf(rep->error, sizeof rep->error, "%s", bad); return -1; }
  }
  try {
    // engine contexts are kept between calls (keyed by modulus and device): their fixed-base tables and recycled
    // blocks serve the next verification over the same group, as one vmnv process verifying several proofs would
    static std::mutex cache_mu;
    // never destroyed: at process exit the CUDA runtime may already be gone when static destructors run
    static auto& cache = *new std::map<std::string, std::unique_ptr<Ctx>>();
    std::unique_lock<std::mutex> cache_lock(cache_mu);   // one verification at a time per process
    std::string key = std::string((const char*)P->p_be, P->nbytes) + std::string((const char*)P->g_be, P->nbytes) +
                      std::string((const char*)P->q_be, P->nbytes) +
                      "#" + std::to_string(P->device) + "#" + std::to_string(P->kind);
    if (P->kind == 1) key += std::string((const char*)P->a_be, P->nbytes) + std::string((const char*)P->b_be, P->nbytes);
    auto it = cache.find(key);
    if (it == cache.end()) {
      vmx_ctx* c = nullptr;
      const int st = P->kind == 1
          ? api.vmx_ctx_create_ecq(P->p_be, P->a_be, P->b_be, P->g_be, P->gy_be, P->q_be, P->nbytes, P->device, &c)
          : api.vmx_ctx_create_modp(P->p_be, P->q_be, P->g_be, P->nbytes, P->device, &c);
      if (st != VMX_OK) {
        snprintf(rep->error, sizeof rep->error, "context: %s", api.vmx_last_error());
        return -1;
      }
      auto ctx = std::make_unique<Ctx>();
      ctx->c = c;
      ctx->eb = api.vmx_ctx_elem_bytes(c);
      ctx->rb = api.vmx_ctx_ring_bytes(c);
      auto fit = [](const uint8_t* be, size_t nbytes, size_t w) {
        Bytes out(w, 0);
        for (size_t i = 0; i < nbytes && i < w; i++) out[w - 1 - i] = be[nbytes - 1 - i];
        return out;
      };
      ctx->q = fit(P->q_be, P->nbytes, ctx->rb);
      if (P->kind == 1) {
        ctx->curve = true;
        ctx->cb = ctx->eb / 2;
        ctx->g = fit(P->g_be, P->nbytes, ctx->cb);
        const Bytes gy = fit(P->gy_be, P->nbytes, ctx->cb);
        ctx->g.insert(ctx->g.end(), gy.begin(), gy.end());
        ctx->one = Bytes(ctx->eb, 0xff);
      } else {
        ctx->g = fit(P->g_be, P->nbytes, ctx->eb);
        ctx->one = Bytes(ctx->eb, 0);
        ctx->one.back() = 1;
      }
      if (cache.size() >= 4) cache.clear();   // a verifier sees one group, a test suite a few
      it = cache.emplace(key, std::move(ctx)).first;
    }
    Session S(*it->second);
    S.P = P;
    vmx_ctx* c = S.C.c;
    for (size_t i = 0; i < nfiles; i++) S.files[files[i].name] = Span(files[i].data, files[i].size);
    const uint64_t l0 = api.vmx_ctx_launch_count(c);
    try {
      S.run(rep);
    } catch (const FailStop& e) {
      rep->fail_stop = 1;
      rep->accepted = 0;
      snprintf(rep->error, sizeof rep->error, "%s", e.what());
    } catch (const Malformed& e) {  // a malformed file outside the places where trivial values are substituted
      rep->fail_stop = 1;
      rep->accepted = 0;
      snprintf(rep->error, sizeof rep->error, "Malformed proof directory: %s", e.what());
    }
    snprintf(rep->test_vectors, sizeof rep->test_vectors, "%s", S.vectors.c_str());
    rep->hashed_bytes = S.hashed;
    rep->launches = api.vmx_ctx_launch_count(c) - l0;
    return 0;
  } catch (const std::exception& e) {
    snprintf(rep->error, sizeof rep->error, "%s", e.what());
    return -2;
  }
}

}  // extern "C"
