/*
 * ORACLE / CPU BASELINE (test infrastructure; never linked into or loaded by the product).
 *
 * GMP-backed array operations with the algorithms the reference's native stack uses
 * (verificatum-gmpmee behind verificatum-vmgj behind verificatum-vcr 3.1.0; none of them is
 * vendored in /root/reference, see SURVEY.md §0/§2.1 -- the algorithms are restated from their
 * published descriptions):
 *   ref_fixed_exp   fixed-base exponentiation with a 2^w-ary window table      (gmpmee fpowm; called
 *                   for PGroupElement.exp(PRingElementArray), e.g. ShufflerElGamalSession.java:407)
 *   ref_powm_array  one modular exponentiation per element, GMP mpz_powm         (VMG.powm; PoSBasicTW.java:1028,1032)
 *   ref_expprod     simultaneous exponentiation in blocks of k bases             (gmpmee spowm; PoSBasicTW.java:408,1021)
 *   ref_mul_array   element-wise product                                        (PoSBasicTW.java:448,610)
 * All of them split the array over `threads` pthreads (VCR's ArrayWorker does the same over
 * the host cores).  Operands are fixed-width big-endian byte strings, element after element.
 *
 * The image has libgmp.so.10 but no gmp.h: the handful of prototypes used are declared by hand
 * (GMP 6 ABI: mpz_t is {int alloc; int size; mp_limb_t* d}).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { int _mp_alloc; int _mp_size; unsigned long* _mp_d; } __mpz_struct;
typedef __mpz_struct mpz_t[1];
extern void __gmpz_init(mpz_t);
extern void __gmpz_clear(mpz_t);
extern void __gmpz_set(mpz_t, const mpz_t);
extern void __gmpz_set_ui(mpz_t, unsigned long);
extern void __gmpz_import(mpz_t, size_t, int, size_t, int, size_t, const void*);
extern void* __gmpz_export(void*, size_t*, int, size_t, int, size_t, const mpz_t);
extern void __gmpz_powm(mpz_t, const mpz_t, const mpz_t, const mpz_t);
extern void __gmpz_mul(mpz_t, const mpz_t, const mpz_t);
extern void __gmpz_mod(mpz_t, const mpz_t, const mpz_t);
extern int __gmpz_tstbit(const mpz_t, unsigned long);
extern size_t __gmpz_sizeinbase(const mpz_t, int);
extern int __gmpz_jacobi(const mpz_t, const mpz_t);
extern void __gmpz_neg(mpz_t, const mpz_t);
#define mpz_init __gmpz_init
#define mpz_clear __gmpz_clear
#define mpz_set __gmpz_set
#define mpz_set_ui __gmpz_set_ui
#define mpz_powm __gmpz_powm
#define mpz_mul __gmpz_mul
#define mpz_mod __gmpz_mod
#define mpz_tstbit __gmpz_tstbit
#define mpz_sizeinbase __gmpz_sizeinbase

static void get(mpz_t x, const uint8_t* p, size_t w) { __gmpz_import(x, w, 1, 1, 1, 0, p); }
static void put(uint8_t* p, size_t w, const mpz_t x) {
  size_t cnt = 0;
  uint8_t tmp[4096];
  memset(p, 0, w);
  __gmpz_export(tmp, &cnt, 1, 1, 1, 0, x);
  if (cnt > w) cnt = w;
  memcpy(p + (w - cnt), tmp, cnt);
}
static unsigned window(const mpz_t e, unsigned long pos, int w) {
  unsigned v = 0;
  for (int b = 0; b < w; b++) v |= (unsigned)__gmpz_tstbit(e, pos + b) << b;
  return v;
}

typedef struct {
  int kind, tid, threads;
  size_t n, ew, xw;
  const uint8_t *a, *b, *e, *mod;
  uint8_t* out;
  int a_scalar, e_scalar, e_neg;
  /* fixed base */
  __mpz_struct* table; int w, nwin;
  /* expprod */
  int k; __mpz_struct* partial;
} job_t;

static void range(const job_t* j, size_t* lo, size_t* hi) {
  *lo = j->n * (size_t)j->tid / j->threads;
  *hi = j->n * (size_t)(j->tid + 1) / j->threads;
}

static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  size_t lo, hi;
  range(j, &lo, &hi);
  mpz_t m, x, y, e, t;
  mpz_init(m); mpz_init(x); mpz_init(y); mpz_init(e); mpz_init(t);
  get(m, j->mod, j->ew);
  if (j->kind == 0) { /* powm per element */
    for (size_t i = lo; i < hi; i++) {
      get(x, j->a + (j->a_scalar ? 0 : i * j->ew), j->ew);
      get(e, j->e + (j->e_scalar ? 0 : i * j->xw), j->xw);
      if (j->e_neg) __gmpz_neg(e, e);   /* GMP inverts the base first */
      mpz_powm(y, x, e, m);
      put(j->out + i * j->ew, j->ew, y);
    }
  } else if (j->kind == 1) { /* mul */
    for (size_t i = lo; i < hi; i++) {
      get(x, j->a + i * j->ew, j->ew);
      get(y, j->b + i * j->ew, j->ew);
      mpz_mul(t, x, y); mpz_mod(t, t, m);
      put(j->out + i * j->ew, j->ew, t);
    }
  } else if (j->kind == 2) { /* fixed base with table */
    for (size_t i = lo; i < hi; i++) {
      get(e, j->e + i * j->xw, j->xw);
      mpz_set_ui(y, 1);
      for (int k = 0; k < j->nwin; k++) {
        unsigned d = window(e, (unsigned long)k * j->w, j->w);
        if (!d) continue;
        mpz_mul(t, y, &j->table[((size_t)k << j->w) + d]); mpz_mod(y, t, m);
      }
      put(j->out + i * j->ew, j->ew, y);
    }
  } else if (j->kind == 3) { /* simultaneous exponentiation of terms [lo, hi) in blocks of k */
    const int k = j->k;
    __mpz_struct* tab = (__mpz_struct*)malloc(sizeof(__mpz_struct) << k);
    __mpz_struct* es = (__mpz_struct*)malloc(sizeof(__mpz_struct) * k);
    for (int s = 0; s < (1 << k); s++) mpz_init(&tab[s]);
    for (int s = 0; s < k; s++) mpz_init(&es[s]);
    mpz_set_ui(&j->partial[j->tid], 1);
    for (size_t b0 = lo; b0 < hi; b0 += k) {
      const int kk = (int)((hi - b0 < (size_t)k) ? hi - b0 : k);
      size_t bits = 0;
      mpz_set_ui(&tab[0], 1);
      for (int s = 0; s < kk; s++) {
        get(x, j->a + (b0 + s) * j->ew, j->ew);
        get(&es[s], j->e + (b0 + s) * j->xw, j->xw);
        size_t bl = es[s]._mp_size ? mpz_sizeinbase(&es[s], 2) : 0;
        if (bl > bits) bits = bl;
        for (int r = 0; r < (1 << s); r++) { mpz_mul(t, &tab[r], x); mpz_mod(&tab[(1 << s) + r], t, m); }
      }
      mpz_set_ui(y, 1);
      for (long bit = (long)bits - 1; bit >= 0; bit--) {
        mpz_mul(t, y, y); mpz_mod(y, t, m);
        unsigned idx = 0;
        for (int s = 0; s < kk; s++) idx |= (unsigned)mpz_tstbit(&es[s], (unsigned long)bit) << s;
        if (idx) { mpz_mul(t, y, &tab[idx]); mpz_mod(y, t, m); }
      }
      mpz_mul(t, &j->partial[j->tid], y); mpz_mod(&j->partial[j->tid], t, m);
    }
    for (int s = 0; s < (1 << k); s++) mpz_clear(&tab[s]);
    for (int s = 0; s < k; s++) mpz_clear(&es[s]);
    free(tab); free(es);
  }
  mpz_clear(m); mpz_clear(x); mpz_clear(y); mpz_clear(e); mpz_clear(t);
  return NULL;
}

static void run(job_t* proto, int threads) {
  if (threads < 1) threads = 1;
  if ((size_t)threads > proto->n && proto->n) threads = (int)proto->n;
  pthread_t th[256];
  job_t jobs[256];
  if (threads > 256) threads = 256;
  for (int t = 0; t < threads; t++) { jobs[t] = *proto; jobs[t].tid = t; jobs[t].threads = threads; }
  for (int t = 1; t < threads; t++) pthread_create(&th[t], NULL, worker, &jobs[t]);
  worker(&jobs[0]);
  for (int t = 1; t < threads; t++) pthread_join(th[t], NULL);
}

/* out[i] = a[i or 0]^{e[i or 0]} mod m */
void ref_powm_array(uint8_t* out, const uint8_t* a, int a_scalar, const uint8_t* e, int e_scalar, size_t n,
                    const uint8_t* mod, size_t ew, size_t xw, int threads) {
  job_t j; memset(&j, 0, sizeof j);
  j.kind = 0; j.n = n; j.ew = ew; j.xw = xw; j.a = a; j.e = e; j.mod = mod; j.out = out; j.a_scalar = a_scalar; j.e_scalar = e_scalar;
  run(&j, threads);
}

/* out[i] = a[i]^{-e} mod m for one small magnitude e: how a negative integer exponent (the modified Lagrange
 * coefficients of DistrElGamalSessionBasic.java:465-503) is applied -- an inversion and a short power, not a
 * |q|-bit exponentiation by q - e */
void ref_powm_array_neg(uint8_t* out, const uint8_t* a, const uint8_t* e_abs, size_t n, const uint8_t* mod, size_t ew,
                        size_t xw, int threads) {
  job_t j; memset(&j, 0, sizeof j);
  j.kind = 0; j.n = n; j.ew = ew; j.xw = xw; j.a = a; j.e = e_abs; j.mod = mod; j.out = out; j.e_scalar = 1; j.e_neg = 1;
  run(&j, threads);
}

void ref_mul_array(uint8_t* out, const uint8_t* a, const uint8_t* b, size_t n, const uint8_t* mod, size_t ew, int threads) {
  job_t j; memset(&j, 0, sizeof j);
  j.kind = 1; j.n = n; j.ew = ew; j.a = a; j.b = b; j.mod = mod; j.out = out;
  run(&j, threads);
}

/* Window table of `base` (gmpmee fpowm_precomp): entry (k, d) = base^(d * 2^(w k)). */
typedef struct { __mpz_struct* table; int w, nwin; } fixed_table_t;

void* ref_fixed_table_create(const uint8_t* base, const uint8_t* mod, size_t ew, int ebits, int w) {
  fixed_table_t* T = (fixed_table_t*)malloc(sizeof(fixed_table_t));
  mpz_t m, q, t;
  mpz_init(m); mpz_init(q); mpz_init(t);
  get(m, mod, ew); get(q, base, ew);
  T->w = w; T->nwin = (ebits + w - 1) / w;
  const size_t entries = (size_t)T->nwin << w;
  T->table = (__mpz_struct*)malloc(sizeof(__mpz_struct) * entries);
  for (size_t s = 0; s < entries; s++) mpz_init(&T->table[s]);
  for (int k = 0; k < T->nwin; k++) {  /* q = base^(2^(w k)) */
    __mpz_struct* R = &T->table[(size_t)k << w];
    mpz_set_ui(&R[0], 1);
    for (unsigned d = 1; d < (1u << w); d++) { mpz_mul(t, &R[d - 1], q); mpz_mod(&R[d], t, m); }
    for (int s = 0; s < w; s++) { mpz_mul(t, q, q); mpz_mod(q, t, m); }
  }
  mpz_clear(m); mpz_clear(q); mpz_clear(t);
  return T;
}

void ref_fixed_table_free(void* h) {
  fixed_table_t* T = (fixed_table_t*)h;
  const size_t entries = (size_t)T->nwin << T->w;
  for (size_t s = 0; s < entries; s++) mpz_clear(&T->table[s]);
  free(T->table);
  free(T);
}

/* out[i] = base^{e[i]} through the table (gmpmee fpowm) */
void ref_fixed_exp(uint8_t* out, void* h, const uint8_t* e, size_t n, const uint8_t* mod, size_t ew, size_t xw,
                   int threads) {
  fixed_table_t* T = (fixed_table_t*)h;
  job_t j; memset(&j, 0, sizeof j);
  j.table = T->table; j.w = T->w; j.nwin = T->nwin;
  j.kind = 2; j.n = n; j.ew = ew; j.xw = xw; j.e = e; j.mod = mod; j.out = out;
  run(&j, threads);
}

/* out[i] = 1 if the Legendre/Jacobi symbol (a[i] | m) is 1 and 0 < a[i] < m, else 0
 * (membership in the order-q subgroup of a safe-prime group; mpz_jacobi, as VCR's natives use) */
void ref_jacobi_array(uint8_t* out, const uint8_t* a, size_t n, const uint8_t* mod, size_t ew) {
  mpz_t m, x;
  mpz_init(m); mpz_init(x);
  get(m, mod, ew);
  for (size_t i = 0; i < n; i++) {
    get(x, a + i * ew, ew);
    out[i] = (uint8_t)(x->_mp_size > 0 && __gmpz_jacobi(x, m) == 1);
  }
  mpz_clear(m); mpz_clear(x);
}

/* out = prod_i a[i]^{e[i]} */
void ref_expprod(uint8_t* out, const uint8_t* a, const uint8_t* e, size_t n, const uint8_t* mod, size_t ew, size_t xw,
                 int k, int threads) {
  job_t j; memset(&j, 0, sizeof j);
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  if ((size_t)threads > n && n) threads = (int)n;
  j.partial = (__mpz_struct*)malloc(sizeof(__mpz_struct) * 256);
  for (int t = 0; t < 256; t++) { mpz_init(&j.partial[t]); mpz_set_ui(&j.partial[t], 1); }
  j.kind = 3; j.n = n; j.ew = ew; j.xw = xw; j.a = a; j.e = e; j.mod = mod; j.k = k;
  if (n) run(&j, threads);
  mpz_t m, y, t;
  mpz_init(m); mpz_init(y); mpz_init(t);
  get(m, mod, ew);
  mpz_set_ui(y, 1);
  for (int s = 0; s < threads; s++) { mpz_mul(t, y, &j.partial[s]); mpz_mod(y, t, m); }
  put(out, ew, y);
  for (int s = 0; s < 256; s++) mpz_clear(&j.partial[s]);
  free(j.partial);
  mpz_clear(m); mpz_clear(y); mpz_clear(t);
}
