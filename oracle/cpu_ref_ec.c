/*
 * ORACLE / CPU BASELINE for curve groups (test infrastructure; never linked into or loaded by the product).
 *
 * GMP-backed array operations over a prime curve y^2 = x^3 + a x + b with the algorithms the reference's native
 * stack uses for ECqPGroup (verificatum-vec behind verificatum-vecj behind verificatum-vcr 3.1.0; not vendored in
 * /root/reference, SURVEY.md §0/§2.1 -- restated from their published descriptions: Jacobian coordinates on mpz,
 * `fmul` = fixed-base multiplication with a window table, `smul` = simultaneous multiplication in blocks of k):
 *   ref_ec_exp_array   one scalar multiplication per point (4-bit windows)          (VEC.mul; PoSBasicTW.java:1028,1032)
 *   ref_ec_fixed_*     fixed-base table of affine points + mixed additions          (VEC.fmul; ShufflerElGamalSession.java:407)
 *   ref_ec_expprod     simultaneous multiplication of terms in blocks of k          (VEC.smul; PoSBasicTW.java:408,1021)
 *   ref_ec_mul_array   element-wise point addition                                  (VEC.add; PoSBasicTW.java:448,610)
 * Arrays are split over `threads` pthreads.  A point is x || y, 32 big-endian bytes each; the unit element is 64
 * bytes 0xff.  Scalars are `xw` big-endian bytes.  curve = p || a || b (32 bytes each).
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
extern void __gmpz_add(mpz_t, const mpz_t, const mpz_t);
extern void __gmpz_sub(mpz_t, const mpz_t, const mpz_t);
extern void __gmpz_mul(mpz_t, const mpz_t, const mpz_t);
extern void __gmpz_mul_2exp(mpz_t, const mpz_t, unsigned long);
extern void __gmpz_mod(mpz_t, const mpz_t, const mpz_t);
extern int __gmpz_invert(mpz_t, const mpz_t, const mpz_t);
extern int __gmpz_tstbit(const mpz_t, unsigned long);
extern size_t __gmpz_sizeinbase(const mpz_t, int);
#define SGN(x) ((x)->_mp_size)

typedef struct { mpz_t X, Y, Z; } jac_t;               /* Z = 0: the unit element */
typedef struct { mpz_t p, a, t1, t2, t3, t4, t5, t6; } ctx_t;

static void get(mpz_t x, const uint8_t* p, size_t w) { __gmpz_import(x, w, 1, 1, 1, 0, p); }
static void put32(uint8_t* p, const mpz_t x) {
  size_t cnt = 0;
  uint8_t tmp[64];
  memset(p, 0, 32);
  __gmpz_export(tmp, &cnt, 1, 1, 1, 0, x);
  if (cnt > 32) cnt = 32;
  memcpy(p + (32 - cnt), tmp, cnt);
}
static void mulm(ctx_t* c, mpz_t r, const mpz_t a, const mpz_t b) { __gmpz_mul(r, a, b); __gmpz_mod(r, r, c->p); }
static void subm(ctx_t* c, mpz_t r, const mpz_t a, const mpz_t b) { __gmpz_sub(r, a, b); __gmpz_mod(r, r, c->p); }
static void jac_init(jac_t* P) { __gmpz_init(P->X); __gmpz_init(P->Y); __gmpz_init(P->Z); }
static void jac_clear(jac_t* P) { __gmpz_clear(P->X); __gmpz_clear(P->Y); __gmpz_clear(P->Z); }
static void jac_set(jac_t* R, const jac_t* P) { __gmpz_set(R->X, P->X); __gmpz_set(R->Y, P->Y); __gmpz_set(R->Z, P->Z); }
static void jac_inf(jac_t* P) { __gmpz_set_ui(P->X, 1); __gmpz_set_ui(P->Y, 1); __gmpz_set_ui(P->Z, 0); }
static void ctx_init(ctx_t* c, const uint8_t* curve) {
  __gmpz_init(c->p); __gmpz_init(c->a); __gmpz_init(c->t1); __gmpz_init(c->t2); __gmpz_init(c->t3);
  __gmpz_init(c->t4); __gmpz_init(c->t5); __gmpz_init(c->t6);
  get(c->p, curve, 32); get(c->a, curve + 32, 32);
}
static void ctx_clear(ctx_t* c) {
  __gmpz_clear(c->p); __gmpz_clear(c->a); __gmpz_clear(c->t1); __gmpz_clear(c->t2); __gmpz_clear(c->t3);
  __gmpz_clear(c->t4); __gmpz_clear(c->t5); __gmpz_clear(c->t6);
}
static int is_unit_bytes(const uint8_t* p) { for (int i = 0; i < 64; i++) if (p[i] != 0xff) return 0; return 1; }
static void jac_load(jac_t* P, const uint8_t* p) {
  if (is_unit_bytes(p)) { jac_inf(P); return; }
  get(P->X, p, 32); get(P->Y, p + 32, 32); __gmpz_set_ui(P->Z, 1);
}
static void jac_store(ctx_t* c, uint8_t* out, const jac_t* P) {   /* to affine: one inversion */
  if (SGN(P->Z) == 0) { memset(out, 0xff, 64); return; }
  __gmpz_invert(c->t1, P->Z, c->p);
  mulm(c, c->t2, c->t1, c->t1);             /* Z^-2 */
  mulm(c, c->t3, P->X, c->t2);
  put32(out, c->t3);
  mulm(c, c->t2, c->t2, c->t1);             /* Z^-3 */
  mulm(c, c->t3, P->Y, c->t2);
  put32(out + 32, c->t3);
}

static void jac_dbl(ctx_t* c, jac_t* P) {
  if (SGN(P->Z) == 0) return;
  if (SGN(P->Y) == 0) { jac_inf(P); return; }
  mulm(c, c->t1, P->Y, P->Y);               /* Y^2 */
  mulm(c, c->t2, P->X, c->t1);              /* X Y^2 */
  __gmpz_mul_2exp(c->t2, c->t2, 2); __gmpz_mod(c->t2, c->t2, c->p);   /* S = 4 X Y^2 */
  mulm(c, c->t3, P->X, P->X);               /* X^2 */
  __gmpz_mul_2exp(c->t4, c->t3, 1); __gmpz_add(c->t3, c->t3, c->t4);  /* 3 X^2 */
  mulm(c, c->t4, P->Z, P->Z);
  mulm(c, c->t4, c->t4, c->t4);             /* Z^4 */
  mulm(c, c->t4, c->t4, c->a);
  __gmpz_add(c->t3, c->t3, c->t4); __gmpz_mod(c->t3, c->t3, c->p);    /* M */
  mulm(c, P->Z, P->Y, P->Z);
  __gmpz_mul_2exp(P->Z, P->Z, 1); __gmpz_mod(P->Z, P->Z, c->p);       /* Z' = 2 Y Z */
  mulm(c, c->t4, c->t3, c->t3);
  __gmpz_mul_2exp(c->t5, c->t2, 1);
  subm(c, P->X, c->t4, c->t5);              /* X' = M^2 - 2 S */
  mulm(c, c->t1, c->t1, c->t1);             /* Y^4 */
  __gmpz_mul_2exp(c->t1, c->t1, 3);
  subm(c, c->t2, c->t2, P->X);
  mulm(c, c->t2, c->t2, c->t3);
  subm(c, P->Y, c->t2, c->t1);              /* Y' = M (S - X') - 8 Y^4 */
}

/* P += Q (Q Jacobian, possibly with Z = 1) */
static void jac_add(ctx_t* c, jac_t* P, const jac_t* Q) {
  if (SGN(Q->Z) == 0) return;
  if (SGN(P->Z) == 0) { jac_set(P, Q); return; }
  mulm(c, c->t1, Q->Z, Q->Z);               /* Z2^2 */
  mulm(c, c->t2, P->Z, P->Z);               /* Z1^2 */
  mulm(c, c->t3, P->X, c->t1);              /* U1 */
  mulm(c, c->t4, Q->X, c->t2);              /* U2 */
  mulm(c, c->t1, c->t1, Q->Z);
  mulm(c, c->t1, c->t1, P->Y);              /* S1 */
  mulm(c, c->t2, c->t2, P->Z);
  mulm(c, c->t2, c->t2, Q->Y);              /* S2 */
  subm(c, c->t4, c->t4, c->t3);             /* H */
  subm(c, c->t2, c->t2, c->t1);             /* R */
  if (SGN(c->t4) == 0) {
    if (SGN(c->t2) == 0) jac_dbl(c, P); else jac_inf(P);
    return;
  }
  mulm(c, P->Z, P->Z, Q->Z);
  mulm(c, P->Z, P->Z, c->t4);               /* Z3 = H Z1 Z2 */
  mulm(c, c->t5, c->t4, c->t4);             /* H^2 */
  mulm(c, c->t6, c->t5, c->t4);             /* H^3 */
  mulm(c, c->t3, c->t3, c->t5);             /* U1 H^2 */
  mulm(c, P->X, c->t2, c->t2);
  __gmpz_sub(P->X, P->X, c->t6);
  __gmpz_mul_2exp(c->t5, c->t3, 1);
  subm(c, P->X, P->X, c->t5);               /* X3 = R^2 - H^3 - 2 U1 H^2 */
  subm(c, c->t3, c->t3, P->X);
  mulm(c, c->t3, c->t3, c->t2);
  mulm(c, c->t1, c->t1, c->t6);
  subm(c, P->Y, c->t3, c->t1);              /* Y3 = R (U1 H^2 - X3) - S1 H^3 */
}

static unsigned window(const mpz_t e, unsigned long pos, int w) {
  unsigned v = 0;
  for (int b = 0; b < w; b++) v |= (unsigned)__gmpz_tstbit(e, pos + b) << b;
  return v;
}

typedef struct {
  int kind, tid, threads;
  size_t n, xw;
  const uint8_t *a, *b, *e, *curve;
  uint8_t* out;
  int a_scalar, e_scalar;
  jac_t* table; int w, nwin;     /* fixed base */
  int k; jac_t* partial;         /* expprod */
} job_t;

static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  const size_t lo = j->n * (size_t)j->tid / j->threads, hi = j->n * (size_t)(j->tid + 1) / j->threads;
  ctx_t c;
  ctx_init(&c, j->curve);
  mpz_t e;
  __gmpz_init(e);
  jac_t P, Q;
  jac_init(&P); jac_init(&Q);
  if (j->kind == 0) {            /* scalar multiplication per point, 4-bit windows */
    jac_t tab[16];
    for (int d = 0; d < 16; d++) jac_init(&tab[d]);
    for (size_t i = lo; i < hi; i++) {
      jac_load(&Q, j->a + (j->a_scalar ? 0 : i * 64));
      get(e, j->e + (j->e_scalar ? 0 : i * j->xw), j->xw);
      jac_inf(&tab[0]);
      for (int d = 1; d < 16; d++) { jac_set(&tab[d], &tab[d - 1]); jac_add(&c, &tab[d], &Q); }
      const size_t bits = SGN(e) ? __gmpz_sizeinbase(e, 2) : 0;
      jac_inf(&P);
      for (long k = (long)((bits + 3) / 4) - 1; k >= 0; k--) {
        for (int s = 0; s < 4; s++) jac_dbl(&c, &P);
        const unsigned d = window(e, (unsigned long)k * 4, 4);
        if (d) jac_add(&c, &P, &tab[d]);
      }
      jac_store(&c, j->out + i * 64, &P);
    }
    for (int d = 0; d < 16; d++) jac_clear(&tab[d]);
  } else if (j->kind == 1) {     /* point addition */
    for (size_t i = lo; i < hi; i++) {
      jac_load(&P, j->a + i * 64);
      jac_load(&Q, j->b + i * 64);
      jac_add(&c, &P, &Q);
      jac_store(&c, j->out + i * 64, &P);
    }
  } else if (j->kind == 2) {     /* fixed base through the table */
    for (size_t i = lo; i < hi; i++) {
      get(e, j->e + i * j->xw, j->xw);
      jac_inf(&P);
      for (int k = 0; k < j->nwin; k++) {
        const unsigned d = window(e, (unsigned long)k * j->w, j->w);
        if (d) jac_add(&c, &P, &j->table[((size_t)k << j->w) + d]);
      }
      jac_store(&c, j->out + i * 64, &P);
    }
  } else if (j->kind == 3) {     /* simultaneous multiplication of terms [lo, hi) in blocks of k */
    const int k = j->k;
    jac_t* tab = (jac_t*)malloc(sizeof(jac_t) << k);
    __mpz_struct* es = (__mpz_struct*)malloc(sizeof(__mpz_struct) * k);
    for (int s = 0; s < (1 << k); s++) jac_init(&tab[s]);
    for (int s = 0; s < k; s++) __gmpz_init(&es[s]);
    jac_inf(&j->partial[j->tid]);
    for (size_t b0 = lo; b0 < hi; b0 += k) {
      const int kk = (int)((hi - b0 < (size_t)k) ? hi - b0 : k);
      size_t bits = 0;
      jac_inf(&tab[0]);
      for (int s = 0; s < kk; s++) {
        jac_load(&Q, j->a + (b0 + s) * 64);
        get(&es[s], j->e + (b0 + s) * j->xw, j->xw);
        const size_t bl = es[s]._mp_size ? __gmpz_sizeinbase(&es[s], 2) : 0;
        if (bl > bits) bits = bl;
        for (int r = 0; r < (1 << s); r++) { jac_set(&tab[(1 << s) + r], &tab[r]); jac_add(&c, &tab[(1 << s) + r], &Q); }
      }
      jac_inf(&P);
      for (long bit = (long)bits - 1; bit >= 0; bit--) {
        jac_dbl(&c, &P);
        unsigned idx = 0;
        for (int s = 0; s < kk; s++) idx |= (unsigned)__gmpz_tstbit(&es[s], (unsigned long)bit) << s;
        if (idx) jac_add(&c, &P, &tab[idx]);
      }
      jac_add(&c, &j->partial[j->tid], &P);
    }
    for (int s = 0; s < (1 << k); s++) jac_clear(&tab[s]);
    for (int s = 0; s < k; s++) __gmpz_clear(&es[s]);
    free(tab); free(es);
  }
  jac_clear(&P); jac_clear(&Q);
  __gmpz_clear(e);
  ctx_clear(&c);
  return NULL;
}

static int clamp_threads(int threads, size_t n) {
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  if ((size_t)threads > n && n) threads = (int)n;
  return threads;
}

static void run(job_t* proto, int threads) {
  threads = clamp_threads(threads, proto->n);
  pthread_t th[256];
  job_t jobs[256];
  for (int t = 0; t < threads; t++) { jobs[t] = *proto; jobs[t].tid = t; jobs[t].threads = threads; }
  for (int t = 1; t < threads; t++) pthread_create(&th[t], NULL, worker, &jobs[t]);
  worker(&jobs[0]);
  for (int t = 1; t < threads; t++) pthread_join(th[t], NULL);
}

/* out[i] = e[i or 0] * a[i or 0] */
void ref_ec_exp_array(uint8_t* out, const uint8_t* a, int a_scalar, const uint8_t* e, int e_scalar, size_t n,
                      const uint8_t* curve, size_t xw, int threads) {
  job_t j; memset(&j, 0, sizeof j);
  j.kind = 0; j.n = n; j.xw = xw; j.a = a; j.e = e; j.curve = curve; j.out = out; j.a_scalar = a_scalar; j.e_scalar = e_scalar;
  if (n) run(&j, threads);
}

void ref_ec_mul_array(uint8_t* out, const uint8_t* a, const uint8_t* b, size_t n, const uint8_t* curve, int threads) {
  job_t j; memset(&j, 0, sizeof j);
  j.kind = 1; j.n = n; j.a = a; j.b = b; j.curve = curve; j.out = out;
  if (n) run(&j, threads);
}

/* Window table of `base`: entry (k, d) = d * 2^(w k) * base, kept affine (Z = 1) so that additions are mixed. */
typedef struct { jac_t* table; int w, nwin; } ec_table_t;

void* ref_ec_fixed_table_create(const uint8_t* base, const uint8_t* curve, int ebits, int w) {
  ec_table_t* T = (ec_table_t*)malloc(sizeof(ec_table_t));
  ctx_t c;
  ctx_init(&c, curve);
  T->w = w; T->nwin = (ebits + w - 1) / w;
  const size_t entries = (size_t)T->nwin << w;
  T->table = (jac_t*)malloc(sizeof(jac_t) * entries);
  for (size_t s = 0; s < entries; s++) jac_init(&T->table[s]);
  jac_t Q;
  jac_init(&Q);
  jac_load(&Q, base);
  uint8_t buf[64];
  for (int k = 0; k < T->nwin; k++) {    /* Q = 2^(w k) * base */
    jac_t* R = &T->table[(size_t)k << w];
    jac_inf(&R[0]);
    for (unsigned d = 1; d < (1u << w); d++) {
      jac_set(&R[d], &R[d - 1]);
      jac_add(&c, &R[d], &Q);
    }
    for (unsigned d = 1; d < (1u << w); d++) { jac_store(&c, buf, &R[d]); jac_load(&R[d], buf); }   /* normalise */
    for (int s = 0; s < w; s++) jac_dbl(&c, &Q);
    jac_store(&c, buf, &Q); jac_load(&Q, buf);
  }
  jac_clear(&Q);
  ctx_clear(&c);
  return T;
}

void ref_ec_fixed_table_free(void* h) {
  ec_table_t* T = (ec_table_t*)h;
  const size_t entries = (size_t)T->nwin << T->w;
  for (size_t s = 0; s < entries; s++) jac_clear(&T->table[s]);
  free(T->table);
  free(T);
}

void ref_ec_fixed_exp(uint8_t* out, void* h, const uint8_t* e, size_t n, const uint8_t* curve, size_t xw, int threads) {
  ec_table_t* T = (ec_table_t*)h;
  job_t j; memset(&j, 0, sizeof j);
  j.table = T->table; j.w = T->w; j.nwin = T->nwin;
  j.kind = 2; j.n = n; j.xw = xw; j.e = e; j.curve = curve; j.out = out;
  if (n) run(&j, threads);
}

/* out = sum_i e[i] * a[i] */
void ref_ec_expprod(uint8_t* out, const uint8_t* a, const uint8_t* e, size_t n, const uint8_t* curve, size_t xw, int k,
                    int threads) {
  job_t j; memset(&j, 0, sizeof j);
  threads = clamp_threads(threads, n);
  j.partial = (jac_t*)malloc(sizeof(jac_t) * 256);
  for (int t = 0; t < 256; t++) { jac_init(&j.partial[t]); jac_inf(&j.partial[t]); }
  j.kind = 3; j.n = n; j.xw = xw; j.a = a; j.e = e; j.curve = curve; j.k = k;
  if (n) run(&j, threads);
  ctx_t c;
  ctx_init(&c, curve);
  jac_t S;
  jac_init(&S); jac_inf(&S);
  for (int s = 0; s < threads; s++) jac_add(&c, &S, &j.partial[s]);
  jac_store(&c, out, &S);
  jac_clear(&S);
  for (int s = 0; s < 256; s++) jac_clear(&j.partial[s]);
  free(j.partial);
  ctx_clear(&c);
}
