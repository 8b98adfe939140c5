// vmx_ec.inl -- host orchestration of the ECqPGroup engine (included by vmx.cu inside namespace vmx,
// after the shared helpers).  An EC context has kind = 1, nl = 8 (coordinate field and exponent
// ring residues), gl = 16 (an affine point).  The C ABI entry points of include/vmx.h branch here.

// ------------------------------------------------------------------ buffers
struct LimbBuf : DevBuf {  // limb-major temporary with an explicit limb count
  size_t cap = 0;
  int alloc_limbs(vmx_ctx* ctx, size_t n, int limbs) { cap = cap_for(n); return alloc(ctx, cap * (size_t)limbs * 4); }
  uint32_t* d() const { return as<uint32_t>(); }
};
constexpr int kAffLimbs = 16, kJacLimbs = 24;
static inline unsigned ec_blocks(size_t n) { return nblocks(n, kEcThreads); }
// kernels are instantiated per reduction (P-256 shift-add / generic Montgomery): pick the context's
#define VMX_EC_LAUNCH(c, kernel, grid, block, smem, ...)                                   \
  do {                                                                                     \
    if ((c)->ecc.F.solinas) VMX_LAUNCH(c, kernel<true>, grid, block, smem, __VA_ARGS__);   \
    else VMX_LAUNCH(c, kernel<false>, grid, block, smem, __VA_ARGS__);                     \
  } while (0)

// field multiplications per point operation (accounting only)
constexpr uint64_t kMulMadd = 11, kMulAdd = 16, kMulDbl = 8, kMulNorm = 11;

// ------------------------------------------------------------------ batched inversion / normalisation
static int ec_batch_inv(vmx_ctx* c, const uint32_t* in, size_t icap, size_t n, uint32_t* out, size_t ocap) {
  if (!n) return VMX_OK;
#ifndef VMX_HOST_EMUL
  const size_t per = (size_t)kEcThreads * kInvK;
  if (n > 32) {
    if (n <= per) {
      VMX_EC_LAUNCH(c, k_fp_inv_block, 1, kEcThreads, 0, in, icap, n, out, ocap, c->ecc);
      VMX_CHECK_LAUNCH();
      c->modmuls += 7 * n + 400;
      return VMX_OK;
    }
    const size_t nb = (n + per - 1) / per;
    LimbBuf excl, totals, totals_inv;
    VMX_TRY(excl.alloc_limbs(c, nb * kEcThreads, 8));
    VMX_TRY(totals.alloc_limbs(c, nb, 8));
    VMX_TRY(totals_inv.alloc_limbs(c, nb, 8));
    VMX_EC_LAUNCH(c, k_fp_inv_up, nb, kEcThreads, 0, in, icap, n, excl.d(), excl.cap, totals.d(), totals.cap, c->ecc);
    VMX_CHECK_LAUNCH();
    VMX_TRY(ec_batch_inv(c, totals.d(), totals.cap, nb, totals_inv.d(), totals_inv.cap));
    VMX_EC_LAUNCH(c, k_fp_inv_down, nb, kEcThreads, 0, in, icap, n, excl.d(), excl.cap, totals_inv.d(), totals_inv.cap, out,
               ocap, c->ecc);
    VMX_CHECK_LAUNCH();
    c->modmuls += 7 * n;
    return VMX_OK;
  }
#endif
  VMX_EC_LAUNCH(c, k_fp_inv_each, ec_blocks(n), kEcThreads, 0, in, icap, n, out, ocap, c->ecc);
  VMX_CHECK_LAUNCH();
  c->modmuls += 384 * n;
  return VMX_OK;
}

// jac (24-limb scratch, n points) -> affine points at out[dst(i)] (see k_ec_finish for the mapping)
static int ec_normalize(vmx_ctx* c, const uint32_t* jac, size_t jcap, size_t n, uint32_t* out, size_t ocap,
                        size_t dst_off = 0, int tw = 0, int tj = 0) {
  if (!n) return VMX_OK;
  LimbBuf zinv;
  VMX_TRY(zinv.alloc_limbs(c, n, 8));
  VMX_TRY(ec_batch_inv(c, jac + 16 * jcap, jcap, n, zinv.d(), zinv.cap));
  VMX_EC_LAUNCH(c, k_ec_finish, ec_blocks(n), kEcThreads, 0, jac, jcap, zinv.d(), zinv.cap, n, out, ocap, dst_off, tw, tj,
             c->ecc);
  VMX_CHECK_LAUNCH();
  c->modmuls += 4 * n;
  return VMX_OK;
}

// ------------------------------------------------------------------ codec
// layout 0: n * (x || y), cb bytes each.  layout 5 (byte-tree form of an array, [VCR-mem] SURVEY.md §8c):
// node(n)[leaf(x_0) .. leaf(x_{n-1})] node(n)[leaf(y_0) ..], i.e. the children of the 2-node an ECqPGroup
// array serialises to.
static size_t ec_wire_bytes(const vmx_ctx* c, size_t n, int hdr) {
  return hdr ? 2 * (5 + n * (5 + c->cb)) : n * 2 * c->cb;
}
static bool ec_node_hdr_ok(const uint8_t* h, size_t n) {
  return h[0] == 0 && h[1] == (uint8_t)(n >> 24) && h[2] == (uint8_t)(n >> 16) && h[3] == (uint8_t)(n >> 8) && h[4] == (uint8_t)n;
}
static void ec_node_hdr_put(uint8_t* h, size_t n) {
  h[0] = 0; h[1] = (uint8_t)(n >> 24); h[2] = (uint8_t)(n >> 16); h[3] = (uint8_t)(n >> 8); h[4] = (uint8_t)n;
}

static int ec_import_dev(vmx_ctx* c, size_t n, const uint8_t* d_raw, int hdr, uint32_t* out, size_t ocap) {
  const size_t cb = c->cb;
  const uint8_t* rx = hdr ? d_raw + 10 : d_raw;
  const uint8_t* ry = hdr ? d_raw + 5 + n * (5 + cb) + 10 : d_raw + cb;
  const size_t stride = hdr ? 5 + cb : 2 * cb;
  VMX_CU(cudaMemsetAsync(c->d_flag, 0, sizeof(int) * 4, c->stream));
  VMX_EC_LAUNCH(c, k_ec_from_bytes, ec_blocks(n), kEcThreads, 0, rx, ry, stride, n, (int)cb, hdr, out, ocap, c->d_flag, c->ecc);
  VMX_CHECK_LAUNCH();
  c->modmuls += 6 * n;
  VMX_TRY(read_flags(c, 1));
  if (c->h_flag[0]) {
    set_error("curve point out of range, not on the curve or malformed leaf (flags %d)", c->h_flag[0]);
    return VMX_EFORMAT;
  }
  return VMX_OK;
}

static int ec_import(vmx_ctx* c, size_t n, const uint8_t* be, int hdr, vmx_garr** out) {
  vmx_garr* a = nullptr;
  VMX_TRY(new_garr(c, n, &a));
  std::unique_ptr<vmx_garr, void (*)(vmx_garr*)> guard(a, vmx_garr_free);
  if (hdr && be) {
    if (!ec_node_hdr_ok(be, n) || !ec_node_hdr_ok(be + 5 + n * (5 + c->cb), n)) {
      set_error("point array: malformed coordinate nodes");
      return VMX_EFORMAT;
    }
  }
  if (n) {
    DevBuf raw;
    const size_t bytes = ec_wire_bytes(c, n, hdr);
    VMX_TRY(raw.alloc(c, bytes));
    VMX_CU(cudaMemcpyAsync(raw.p, be, bytes, cudaMemcpyHostToDevice, c->stream));
    VMX_TRY(ec_import_dev(c, n, raw.as<uint8_t>(), hdr, a->d, a->cap));
  }
  *out = guard.release();
  return VMX_OK;
}

static int ec_export(const vmx_garr* a, int hdr, uint8_t* be_out) {
  vmx_ctx* c = a->ctx;
  const size_t n = a->n, cb = c->cb;
  if (hdr) { ec_node_hdr_put(be_out, n); ec_node_hdr_put(be_out + 5 + n * (5 + cb), n); }
  if (!n) return VMX_OK;
  DevBuf raw;
  const size_t bytes = ec_wire_bytes(c, n, hdr);
  VMX_TRY(raw.alloc(c, bytes));
  uint8_t* rx = hdr ? raw.as<uint8_t>() + 10 : raw.as<uint8_t>();
  uint8_t* ry = hdr ? raw.as<uint8_t>() + 5 + n * (5 + cb) + 10 : raw.as<uint8_t>() + cb;
  const size_t stride = hdr ? 5 + cb : 2 * cb;
  VMX_EC_LAUNCH(c, k_ec_to_bytes, ec_blocks(n), kEcThreads, 0, a->d, a->cap, n, (int)cb, hdr, rx, ry, stride, c->ecc);
  VMX_CHECK_LAUNCH();
  c->modmuls += 2 * n;
  if (hdr) {  // the two node headers are host-written: copy the leaf runs only
    const size_t run = n * (5 + cb);
    VMX_CU(cudaMemcpyAsync(be_out + 5, raw.as<uint8_t>() + 5, run, cudaMemcpyDeviceToHost, c->stream));
    VMX_CU(cudaMemcpyAsync(be_out + 10 + run, raw.as<uint8_t>() + 10 + run, run, cudaMemcpyDeviceToHost, c->stream));
  } else {
    VMX_CU(cudaMemcpyAsync(be_out, raw.p, bytes, cudaMemcpyDeviceToHost, c->stream));
  }
  VMX_CU(cudaStreamSynchronize(c->stream));
  return VMX_OK;
}

// one point given as x || y bytes -> 1-element affine temporary (range and on-curve checked)
static int ec_upload_one(vmx_ctx* c, const uint8_t* be, ElemBuf& buf) {
  DevBuf raw;
  VMX_TRY(raw.alloc(c, 2 * c->cb));
  VMX_CU(cudaMemcpyAsync(raw.p, be, 2 * c->cb, cudaMemcpyHostToDevice, c->stream));
  VMX_TRY(buf.alloc_limbs(c, 1, kAffLimbs));
  return ec_import_dev(c, 1, raw.as<uint8_t>(), 0, buf.d(), buf.cap);
}

static int ec_download_one(vmx_ctx* c, const uint32_t* d, size_t cap, size_t idx, uint8_t* out_be) {
  DevBuf raw;
  VMX_TRY(raw.alloc(c, 2 * c->cb));
  VMX_EC_LAUNCH(c, k_ec_to_bytes, 1, kEcThreads, 0, d + 4 * idx, cap, (size_t)1, (int)c->cb, 0, raw.as<uint8_t>(),
             raw.as<uint8_t>() + c->cb, 2 * c->cb, c->ecc);
  VMX_CHECK_LAUNCH();
  VMX_CU(cudaMemcpyAsync(out_be, raw.p, 2 * c->cb, cudaMemcpyDeviceToHost, c->stream));
  VMX_CU(cudaStreamSynchronize(c->stream));
  return VMX_OK;
}

// ------------------------------------------------------------------ fixed base
// an entry is 64 B and costs one mixed addition + its share of the batched inversion to build: wide windows
// are cheap (w = 20 at n = 10^6: 13 windows x 2^20 x 64 B = 872 MB, built in a few ms, 13 additions per point)
static double ec_fixed_cost(int w, size_t n, int ebits) {
  const double nwin = (ebits + w - 1) / w;
  return nwin * (kFixedReuse * (double)kMulMadd * (double)n + 2.0 * kMulMadd * (double)(1u << w));
}
static int ec_choose_fixed_window(const vmx_ctx* c, size_t n) {
  if (c->fixed_window) return c->fixed_window;
  const int ebits = c->Q.bits;
  double best = 1e300;
  int bw = 4;
  for (int w = 4; w <= 22; w++) {
    const double nwin = (ebits + w - 1) / w;
    if (nwin * (double)(1u << w) * kAffLimbs * 4 > 4e9) break;
    const double cost = ec_fixed_cost(w, n, ebits);
    if (cost < best) { best = cost; bw = w; }
  }
  return bw;
}

static int ec_build_table(vmx_ctx* c, const uint32_t* base, size_t bcap, int w, FixedTable& T) {
  const int ebits = c->Q.bits;
  T.w = w;
  T.nwin = (ebits + w - 1) / w;
  const size_t entries = (size_t)T.nwin << w;
  T.cap = cap_for(entries);
  void* p = nullptr;
  if (cudaMallocAsync(&p, T.cap * kAffLimbs * 4, c->stream) != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("fixed-base table allocation failed (%zu bytes)", T.cap * kAffLimbs * 4);
    return VMX_ENOMEM;
  }
  T.d = (uint32_t*)p;
  VMX_CU(cudaMemsetAsync(T.d, 0xff, T.cap * kAffLimbs * 4, c->stream));  // every entry = unit until written
  const int qlen = T.nwin * w;
  LimbBuf Qj, Q;
  VMX_TRY(Qj.alloc_limbs(c, qlen, kJacLimbs));
  VMX_TRY(Q.alloc_limbs(c, qlen, kAffLimbs));
  VMX_EC_LAUNCH(c, k_ec_dbl_chain, 1, 32, 0, base, bcap, Qj.d(), Qj.cap, qlen, c->ecc);
  VMX_CHECK_LAUNCH();
  c->modmuls += kMulDbl * qlen;
  VMX_TRY(ec_normalize(c, Qj.d(), Qj.cap, qlen, Q.d(), Q.cap));
  LimbBuf lvl;
  VMX_TRY(lvl.alloc_limbs(c, (size_t)T.nwin << (w - 1), kJacLimbs));
  for (int j = 0; j < w; j++) {
    const size_t items = (size_t)T.nwin << j;
    VMX_EC_LAUNCH(c, k_ec_table_level, ec_blocks(items), kEcThreads, 0, T.d, T.cap, w, T.nwin, j, qlen, Q.d(), Q.cap,
               lvl.d(), lvl.cap, c->ecc);
    VMX_CHECK_LAUNCH();
    c->modmuls += kMulMadd * items;
    VMX_TRY(ec_normalize(c, lvl.d(), lvl.cap, items, T.d, T.cap, 0, w, j));
  }
  return VMX_OK;
}

static int ec_exp_fixed_run(vmx_ctx* c, const FixedTable& T, const vmx_rarr* e, int ebits, uint32_t* out, size_t ocap) {
  const size_t n = e->n;
  if (!n) return VMX_OK;
  const int nwin = std::min(T.nwin, std::max(1, (ebits + T.w - 1) / T.w));
  LimbBuf jac;
  VMX_TRY(jac.alloc_limbs(c, n, kJacLimbs));
  VMX_EC_LAUNCH(c, k_ec_exp_fixed, ec_blocks(n), kEcThreads, 0, T.d, T.cap, T.w, nwin, e->d, e->cap, n, jac.d(), jac.cap,
             c->ecc);
  VMX_CHECK_LAUNCH();
  c->modmuls += kMulMadd * n * nwin;
  return ec_normalize(c, jac.d(), jac.cap, n, out, ocap);
}

// ------------------------------------------------------------------ variable base
static int ec_exp_var_run(vmx_ctx* c, const uint32_t* a, size_t acap, const uint32_t* E, size_t ecap, bool escalar,
                          int ebits, size_t n, uint32_t* out, size_t ocap) {
  if (!n) return VMX_OK;
  LimbBuf jac, tab;
  VMX_TRY(jac.alloc_limbs(c, n, kJacLimbs));
  size_t chunk = (size_t)(6e9 / (15.0 * kJacLimbs * 4));
  chunk = std::min(std::max<size_t>(chunk & ~(size_t)1023, 1024), n);
  VMX_TRY(tab.alloc_limbs(c, chunk * 15, kJacLimbs));
  const int nwin = (ebits + 3) / 4;
  for (size_t i0 = 0; i0 < n; i0 += chunk) {
    const size_t m = std::min(chunk, n - i0);
    VMX_EC_LAUNCH(c, k_ec_exp_var, ec_blocks(m), kEcThreads, 0, a + 4 * i0, acap, escalar ? E : E + 4 * i0, ecap,
               escalar ? 1 : 0, ebits, m, tab.d(), tab.cap, jac.d() + 4 * i0, jac.cap, c->ecc);
    VMX_CHECK_LAUNCH();
    c->modmuls += (uint64_t)m * (14 * kMulMadd + (uint64_t)nwin * (4 * kMulDbl + kMulAdd));
  }
  return ec_normalize(c, jac.d(), jac.cap, n, out, ocap);
}

// out[i] = X[0] * a[i] + Y[i] * b[i] (k_ec_exp_var2), normalised to affine
static int ec_exp_var2_run(vmx_ctx* c, const uint32_t* a, size_t acap, const uint32_t* X, size_t xcap, int xbits,
                           const uint32_t* b, size_t bcap, const uint32_t* Y, size_t ycap, int ybits, size_t n,
                           uint32_t* out, size_t ocap) {
  if (!n) return VMX_OK;
  LimbBuf jac, tabA, tabB;
  VMX_TRY(jac.alloc_limbs(c, n, kJacLimbs));
  size_t chunk = (size_t)(6e9 / (30.0 * kJacLimbs * 4));
  chunk = std::min(std::max<size_t>(chunk & ~(size_t)1023, 1024), n);
  VMX_TRY(tabA.alloc_limbs(c, chunk * 15, kJacLimbs));
  VMX_TRY(tabB.alloc_limbs(c, chunk * 15, kJacLimbs));
  const int nwx = (xbits + 3) / 4, nwy = (ybits + 3) / 4, nwin = std::max(nwx, nwy);
  for (size_t i0 = 0; i0 < n; i0 += chunk) {
    const size_t m = std::min(chunk, n - i0);
    VMX_EC_LAUNCH(c, k_ec_exp_var2, ec_blocks(m), kEcThreads, 0, a + 4 * i0, acap, X, xcap, xbits, b + 4 * i0, bcap,
               Y + 4 * i0, ycap, ybits, m, tabA.d(), tabB.d(), tabA.cap, jac.d() + 4 * i0, jac.cap, c->ecc);
    VMX_CHECK_LAUNCH();
    c->modmuls += (uint64_t)m * (28 * kMulMadd + (uint64_t)nwin * 4 * kMulDbl + (uint64_t)(nwx + nwy) * kMulAdd);
  }
  return ec_normalize(c, jac.d(), jac.cap, n, out, ocap);
}

// ------------------------------------------------------------------ segmented sums
// out[s] = sum_{k in [seg_off[s], seg_off[s+1])} V[idx ? idx[k] : k] as Jacobian points (empty -> unit);
// V affine (vjac = 0) or Jacobian.  Same chunked rounds as seg_product.
static int ec_seg_sum(vmx_ctx* c, const uint32_t* V, size_t vcap, int vjac, const uint32_t* idx, const uint32_t* seg_off,
                      size_t nseg, size_t total_bound, int K, uint32_t* out, size_t ocap) {
  DevBuf off_keep;
  LimbBuf val_keep;
  const uint32_t* cur_V = V;
  size_t cur_vcap = vcap;
  int cur_jac = vjac;
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
    const bool last = (c->h_flag[0] == 0);
    const uint32_t* nch_dev = chunk_off.as<uint32_t>() + nseg;
    c->modmuls += (cur_jac ? kMulAdd : kMulMadd) * cur_total;
    if (last) {
      VMX_EC_LAUNCH(c, k_ec_seg_sum, ec_blocks(nseg), kEcThreads, 0, cur_V, cur_vcap, cur_jac, cur_idx, chunks.as<Chunk>(),
                 nch_dev, out, ocap, c->ecc);
      VMX_CHECK_LAUNCH();
      return VMX_OK;
    }
    LimbBuf part;
    VMX_TRY(part.alloc_limbs(c, nch_bound, kJacLimbs));
    VMX_EC_LAUNCH(c, k_ec_seg_sum, ec_blocks(nch_bound), kEcThreads, 0, cur_V, cur_vcap, cur_jac, cur_idx,
               chunks.as<Chunk>(), nch_dev, part.d(), part.cap, c->ecc);
    VMX_CHECK_LAUNCH();
    std::swap(val_keep.p, part.p); std::swap(val_keep.c, part.c); std::swap(val_keep.cap, part.cap);
    std::swap(val_keep.granted, part.granted);
    std::swap(off_keep.p, chunk_off.p); std::swap(off_keep.c, chunk_off.c); std::swap(off_keep.granted, chunk_off.granted);
    cur_V = val_keep.d();
    cur_vcap = val_keep.cap;
    cur_jac = 1;
    cur_idx = nullptr;
    cur_off = off_keep.as<uint32_t>();
    cur_total = nch_bound;
  }
  set_error("segmented sum did not converge");
  return VMX_ECUDA;
}

// ------------------------------------------------------------------ Pippenger
// Y[col * ngroups + g] for one column (array) of an expProd
static int ec_mexp_run(vmx_ctx* c, const MexpPlan& P, const vmx_garr* a, size_t n_terms, uint32_t* Y, size_t ycap,
                       size_t col) {
  LimbBuf buckets, X;
  VMX_TRY(buckets.alloc_limbs(c, P.nb, kJacLimbs));
  VMX_TRY(ec_seg_sum(c, a->d, a->cap, 0, P.idx.as<uint32_t>(), P.seg_off.as<uint32_t>(), P.nb, n_terms * (size_t)P.W,
                     mexp_chunk(), buckets.d(), buckets.cap));
  VMX_TRY(X.alloc_limbs(c, P.nseg2, kJacLimbs));
  VMX_TRY(ec_seg_sum(c, buckets.d(), buckets.cap, 1, P.idx2.as<uint32_t>(), P.seg2_off.as<uint32_t>(), P.nseg2, P.total2,
                     8, X.d(), X.cap));
  const size_t ngroups = (size_t)P.W * P.J;
  VMX_EC_LAUNCH(c, k_ec_weighted_small, ec_blocks(ngroups), kEcThreads, 0, X.d(), X.cap, ngroups, Y + 4 * col * ngroups,
             ycap, c->ecc);
  VMX_CHECK_LAUNCH();
  c->modmuls += ngroups * 30 * kMulAdd;
  return VMX_OK;
}

static int ec_expprod(vmx_ctx* c, const vmx_garr* const* a, size_t k, const vmx_rarr* e, uint8_t* out_be) {
  int L = 0;
  VMX_TRY(rarr_bitlen(e, &L));
  LimbBuf res;
  VMX_TRY(res.alloc_limbs(c, k, kAffLimbs));
  if (e->n == 0 || L == 0) {
    VMX_CU(cudaMemsetAsync(res.p, 0xff, res.cap * kAffLimbs * 4, c->stream));
  } else {
    MexpPlan P;
    VMX_TRY(mexp_plan<8>(c, e, L, P));
    const size_t ngroups = (size_t)P.W * P.J;
    LimbBuf Y, R;
    VMX_TRY(Y.alloc_limbs(c, k * ngroups, kJacLimbs));
    for (size_t l = 0; l < k; l++) VMX_TRY(ec_mexp_run(c, P, a[l], e->n, Y.d(), Y.cap, l));
    VMX_TRY(R.alloc_limbs(c, k, kJacLimbs));
    VMX_EC_LAUNCH(c, k_ec_horner, nblocks(k, 32), 32, 0, Y.d(), Y.cap, (int)ngroups, (int)k, R.d(), R.cap, c->ecc);
    VMX_CHECK_LAUNCH();
    c->modmuls += k * ngroups * (4 * kMulDbl + kMulAdd);
    VMX_TRY(ec_normalize(c, R.d(), R.cap, k, res.d(), res.cap));
  }
  for (size_t l = 0; l < k; l++) VMX_TRY(ec_download_one(c, res.d(), res.cap, l, out_be + l * c->eb));
  return VMX_OK;
}

static int ec_prod(vmx_ctx* c, const vmx_garr* a, uint8_t* out_be) {
  LimbBuf rj, res;
  VMX_TRY(rj.alloc_limbs(c, 1, kJacLimbs));
  VMX_TRY(res.alloc_limbs(c, 1, kAffLimbs));
  DevBuf off;
  VMX_TRY(off.alloc(c, 8));
  const uint32_t h[2] = {0, (uint32_t)a->n};
  VMX_CU(cudaMemcpyAsync(off.p, h, 8, cudaMemcpyHostToDevice, c->stream));
  VMX_CU(cudaStreamSynchronize(c->stream));
  const int K = (int)std::min<size_t>(64, std::max<size_t>(2, a->n / ((size_t)c->sm_count * 1024)));
  VMX_TRY(ec_seg_sum(c, a->d, a->cap, 0, nullptr, off.as<uint32_t>(), 1, a->n, K, rj.d(), rj.cap));
  VMX_TRY(ec_normalize(c, rj.d(), rj.cap, 1, res.d(), res.cap));
  return ec_download_one(c, res.d(), res.cap, 0, out_be);
}

// ------------------------------------------------------------------ element-wise
static int ec_mul(vmx_ctx* c, const vmx_garr* a, const vmx_garr* b, vmx_garr* r) {
  if (!a->n) return VMX_OK;
  LimbBuf jac;
  VMX_TRY(jac.alloc_limbs(c, a->n, kJacLimbs));
  VMX_EC_LAUNCH(c, k_ec_add, ec_blocks(a->n), kEcThreads, 0, a->d, a->cap, b->d, b->cap, a->n, jac.d(), jac.cap, c->ecc);
  VMX_CHECK_LAUNCH();
  c->modmuls += kMulMadd * a->n;
  return ec_normalize(c, jac.d(), jac.cap, a->n, r->d, r->cap);
}

static int ec_neg(vmx_ctx* c, const vmx_garr* a, vmx_garr* r) {
  if (!a->n) return VMX_OK;
  VMX_EC_LAUNCH(c, k_ec_neg, ec_blocks(a->n), kEcThreads, 0, a->d, a->cap, a->n, r->d, r->cap, c->ecc);
  VMX_CHECK_LAUNCH();
  return VMX_OK;
}

static int ec_cols(vmx_ctx* c, const vmx_garr* const* bases, size_t t, const int64_t* ints, vmx_garr* r) {
  if (t > 8) { set_error("expProd: at most 8 columns on a curve group"); return VMX_EARG; }
  const size_t n = r->n;
  if (!n) return VMX_OK;
  EcCols A;
  std::memset(&A, 0, sizeof A);
  A.t = (int)t;
  uint64_t work = 0;
  for (size_t j = 0; j < t; j++) {
    A.d[j] = bases[j]->d;
    A.cap[j] = bases[j]->cap;
    A.k[j] = (long long)ints[j];
    uint64_t mag = ints[j] < 0 ? (uint64_t)(-(ints[j] + 1)) + 1 : (uint64_t)ints[j];
    int bits = 0;
    while (mag) { bits++; mag >>= 1; }
    work += (uint64_t)bits * (kMulDbl + kMulMadd / 2) + kMulAdd;
  }
  LimbBuf jac;
  VMX_TRY(jac.alloc_limbs(c, n, kJacLimbs));
  VMX_EC_LAUNCH(c, k_ec_cols, ec_blocks(n), kEcThreads, 0, A, n, jac.d(), jac.cap, c->ecc);
  VMX_CHECK_LAUNCH();
  c->modmuls += work * n;
  return ec_normalize(c, jac.d(), jac.cap, n, r->d, r->cap);
}

// -(x, y) = (x, p - y) on the wire form of ONE point (PGroupElement.inv, hvzk/PoSBasicTW.java:1013-1014)
static int ec_elem_inv(vmx_ctx* c, const uint8_t* in_be, uint8_t* out_be) {
  const size_t cb = c->cb;
  std::memcpy(out_be, in_be, 2 * cb);
  bool allff = true;
  for (size_t k = 0; k < 2 * cb; k++) allff = allff && in_be[k] == 0xff;
  if (allff) return VMX_OK;
  uint32_t y[8], r[8];
  if (!be_to_limbs(in_be + cb, cb, y, 8) || limbs_cmp(y, c->P.n, 8) >= 0) { set_error("point out of range"); return VMX_EFORMAT; }
  bool zero = true;
  for (int j = 0; j < 8; j++) zero = zero && y[j] == 0;
  if (zero) return VMX_OK;
  std::memcpy(r, c->P.n, sizeof r);
  limbs_sub(r, y, 8);
  std::memset(out_be + cb, 0, cb);
  for (size_t b = 0; b < 32 && b < cb; b++) out_be[2 * cb - 1 - b] = (uint8_t)(r[b / 4] >> (8 * (b % 4)));
  return VMX_OK;
}

// ------------------------------------------------------------------ random elements
// Candidates j = 0..m-1 from d_raw (m * width bytes): accepted ones are appended to out[have..want) in
// stream order.  *accepted = how many were appended, *used = candidates consumed (all m unless the
// array became full, then up to and including the one that filled it).
static int ec_candidates(vmx_ctx* c, const uint8_t* d_raw, size_t m, size_t width, unsigned bitlen, vmx_garr* out,
                         size_t have, size_t* accepted, size_t* used) {
  *accepted = 0;
  *used = m;
  if (!m) return VMX_OK;
  if (!c->ec_sqrt_ok) { set_error("random curve points need p = 3 mod 4"); return VMX_EARG; }
  LimbBuf xs, cand;
  DevBuf ok, pos;
  VMX_TRY(xs.alloc_limbs(c, m, 8));
  VMX_TRY(cand.alloc_limbs(c, m, kAffLimbs));
  VMX_TRY(ok.alloc(c, (m + 1) * 4));
  VMX_TRY(pos.alloc(c, (m + 1) * 4));
  VMX_LAUNCH(c, k_ring_from_raw<8>, nblocks(m, codec_threads(width)), codec_threads(width), codec_smem(width), d_raw, m, (int)width, (int)bitlen, xs.d(), xs.cap,
             c->P.consts, 1, c->P.params<8>());
  VMX_CHECK_LAUNCH();
  VMX_CU(cudaMemsetAsync(ok.p, 0, (m + 1) * 4, c->stream));
  VMX_EC_LAUNCH(c, k_ec_candidates, ec_blocks(m), kEcThreads, 0, xs.d(), xs.cap, m, cand.d(), cand.cap, ok.as<uint32_t>(),
             c->ecc);
  VMX_CHECK_LAUNCH();
  c->modmuls += 400 * m;
  VMX_CU(cudaMemcpyAsync(pos.p, ok.p, (m + 1) * 4, cudaMemcpyDeviceToDevice, c->stream));
  VMX_TRY(exclusive_scan(c, pos.as<uint32_t>(), m + 1));
  VMX_CU(cudaMemsetAsync(c->d_flag, 0, sizeof(int) * 4, c->stream));
  VMX_LAUNCH(c, k_ec_compact, nblocks(m, 256), 256, 0, reinterpret_cast<const uint4*>(cand.d()), cand.cap,
             ok.as<uint32_t>(), pos.as<uint32_t>(), m, have, out->n, reinterpret_cast<uint4*>(out->d), out->cap,
             reinterpret_cast<uint32_t*>(c->d_flag));
  VMX_CHECK_LAUNCH();
  VMX_CU(cudaMemcpyAsync(c->h_flag + 1, pos.as<uint32_t>() + m, 4, cudaMemcpyDeviceToHost, c->stream));
  VMX_TRY(read_flags(c, 1));
  const size_t total = (size_t)(uint32_t)c->h_flag[1];
  const size_t room = out->n - have;
  *accepted = std::min(total, room);
  if (total >= room) *used = (size_t)(uint32_t)c->h_flag[0];
  return VMX_OK;
}

static int ec_random_prg(vmx_ctx* c, const uint8_t* seed, size_t seedlen, uint64_t offset, size_t n, size_t width,
                         unsigned bitlen, vmx_garr** out) {
  vmx_garr* a = nullptr;
  VMX_TRY(new_garr(c, n, &a));
  std::unique_ptr<vmx_garr, void (*)(vmx_garr*)> guard(a, vmx_garr_free);
  size_t have = 0;
  uint64_t cand_used = 0;
  for (int round = 0; have < n; round++) {
    if (round > 200) { set_error("random points: too many rejected candidates"); return VMX_ECUDA; }
    const size_t m = 2 * (n - have) + 64;
    DevBuf raw;
    const uint8_t* data = nullptr;
    VMX_TRY(prg_bytes_dev(c, seed, seedlen, offset + cand_used * width, m * width, raw, &data));
    size_t acc = 0, used = 0;
    VMX_TRY(ec_candidates(c, data, m, width, bitlen, a, have, &acc, &used));
    have += acc;
    cand_used += used;
  }
  c->prg_consumed = cand_used * width;
  *out = guard.release();
  return VMX_OK;
}
