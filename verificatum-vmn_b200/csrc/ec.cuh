// ec.cuh -- point arithmetic on y^2 = x^3 + a x + b over a 256-bit prime field (fp256.cuh), the
// group law behind ECqPGroup (reference default: NIST curves, demo/mixnet/benchmarks/bench_config:33-50;
// BASELINE.json config 5 is P-256).
//
// Arrays hold AFFINE points (x, y in Montgomery form, 16 limbs = 4 uint4 planes, the unit element
// as the all-ones word pattern, which is not a residue): the canonical form, so `equals` is a
// bit compare, serialisation is a byte swap, and every accumulation loop adds an affine operand
// to a Jacobian accumulator (8M + 3S instead of 12M + 4S).  Kernels produce Jacobian results
// into a 24-limb scratch array; one batched inversion (kernels_ec.cuh) brings a whole array
// back to affine form at ~11 field multiplications per point.
//
// Exactness: the special cases of the addition law (equal points, opposite points, the unit on
// either side) are branched on explicitly -- inputs are adversarial in a verifier -- so every
// result is the true group element; bit-exact parity with the oracle (oracle/ec.py) follows from
// uniqueness of the affine form.
#pragma once
#include "fp256.cuh"
#include "layout.cuh"

namespace vmx {

struct EcCurve {
  Fp256 F;
  uint32_t a[8], b[8];  // curve coefficients, Montgomery form
  uint32_t one[8];      // R mod p
  uint32_t r2[8];       // R^2 mod p
  uint32_t pm2[8];      // p - 2 (inversion exponent), plain
  uint32_t sqe[8];      // (p + 1) / 4 (square-root exponent, p = 3 mod 4), plain
  uint32_t a_minus3;    // 1: a = p - 3
};

struct Jac { uint32_t X[8], Y[8], Z[8]; };  // Z = 0: the unit element

VMX_DEV bool jac_is_inf(const Jac& P) { return fp_is_zero(P.Z); }
VMX_DEV void jac_set_inf(Jac& P, const EcCurve& C) {
  fp_copy(P.X, C.one); fp_copy(P.Y, C.one);
#pragma unroll
  for (int j = 0; j < 8; j++) P.Z[j] = 0;
}
VMX_DEV void jac_from_affine(Jac& P, const uint32_t (&x)[8], const uint32_t (&y)[8], const EcCurve& C) {
  fp_copy(P.X, x); fp_copy(P.Y, y); fp_copy(P.Z, C.one);
}

// the unit element in an affine array: every word 0xffffffff
VMX_DEV bool aff_is_inf(const uint32_t (&x)[8]) {
  uint32_t m = 0xffffffffu;
#pragma unroll
  for (int j = 0; j < 8; j++) m &= x[j];
  return m == 0xffffffffu;
}

// P <- 2P
template <bool SOL>
VMX_DEV void jac_dbl(Jac& P, const EcCurve& C) {
  const Fp256& F = C.F;
  uint32_t t0[8], t1[8], t2[8], t3[8];
  if (SOL || C.a_minus3) {  // the P-256 instantiation carries the a = -3 formula only
    // dbl-2001-b: 3M + 5S
    fp_sqr<SOL>(t0, P.Z, F);            // delta
    fp_sqr<SOL>(t1, P.Y, F);            // gamma
    fp_mul<SOL>(t2, P.X, t1, F);        // beta
    fp_sub(t3, P.X, t0, F);
    fp_add(t0, P.X, t0, F);        // (t0 keeps delta no longer; Z3 needs it -> recompute below)
    fp_mul<SOL>(t3, t3, t0, F);
    fp_dbl(t0, t3, F);
    fp_add(t3, t0, t3, F);         // alpha = 3 (X - delta)(X + delta)
    // Z3 = (Y + Z)^2 - gamma - delta = 2 Y Z
    fp_mul<SOL>(t0, P.Y, P.Z, F);
    fp_dbl(P.Z, t0, F);
    // X3 = alpha^2 - 8 beta
    fp_dbl(t2, t2, F); fp_dbl(t2, t2, F);  // 4 beta
    fp_sqr<SOL>(t0, t3, F);
    fp_sub(t0, t0, t2, F);
    fp_sub(P.X, t0, t2, F);
    // Y3 = alpha (4 beta - X3) - 8 gamma^2
    fp_sub(t2, t2, P.X, F);
    fp_mul<SOL>(t2, t3, t2, F);
    fp_sqr<SOL>(t1, t1, F);
    fp_dbl(t1, t1, F); fp_dbl(t1, t1, F); fp_dbl(t1, t1, F);
    fp_sub(P.Y, t2, t1, F);
  } else {
    // general a: M = 3 X^2 + a Z^4, S = 4 X Y^2
    fp_sqr<SOL>(t0, P.Z, F);
    fp_sqr<SOL>(t0, t0, F);
    fp_mul<SOL>(t0, t0, C.a, F);        // a Z^4
    fp_sqr<SOL>(t1, P.X, F);
    fp_dbl(t2, t1, F);
    fp_add(t1, t2, t1, F);
    fp_add(t3, t1, t0, F);         // M
    fp_sqr<SOL>(t1, P.Y, F);            // Y^2
    fp_mul<SOL>(t2, P.X, t1, F);
    fp_dbl(t2, t2, F); fp_dbl(t2, t2, F);  // S
    fp_mul<SOL>(t0, P.Y, P.Z, F);
    fp_dbl(P.Z, t0, F);            // Z3 = 2 Y Z
    fp_sqr<SOL>(t0, t3, F);
    fp_sub(t0, t0, t2, F);
    fp_sub(P.X, t0, t2, F);        // X3 = M^2 - 2 S
    fp_sub(t2, t2, P.X, F);
    fp_mul<SOL>(t2, t3, t2, F);
    fp_sqr<SOL>(t1, t1, F);
    fp_dbl(t1, t1, F); fp_dbl(t1, t1, F); fp_dbl(t1, t1, F);
    fp_sub(P.Y, t2, t1, F);        // Y3 = M (S - X3) - 8 Y^4
  }
}

// The doubling branch of an addition (both operands equal) is never taken on honest inputs: it lives in
// ONE out-of-line copy per instantiation so that the hot loops stay small enough for the instruction cache
// (an inlined point operation is 30-50 KB of straight-line SASS).
#ifndef VMX_HOST_EMUL
#define VMX_NOINLINE __device__ __noinline__
#else
#define VMX_NOINLINE inline
#endif
template <bool SOL>
VMX_NOINLINE void jac_dbl_rare(Jac* P, const EcCurve* C) {
  Jac T = *P;
  jac_dbl<SOL>(T, *C);
  *P = T;
}

// P <- P + (x2, y2), (x2, y2) an affine point that is not the unit.  8M + 3S.
template <bool SOL>
VMX_DEV void jac_madd(Jac& P, const uint32_t (&x2)[8], const uint32_t (&y2)[8], const EcCurve& C) {
  const Fp256& F = C.F;
  if (jac_is_inf(P)) { jac_from_affine(P, x2, y2, C); return; }
  uint32_t zz[8], u2[8], s2[8], h[8], r[8];
  fp_sqr<SOL>(zz, P.Z, F);
  fp_mul<SOL>(u2, x2, zz, F);
  fp_mul<SOL>(s2, P.Z, zz, F);
  fp_mul<SOL>(s2, y2, s2, F);
  fp_sub(h, u2, P.X, F);
  fp_sub(r, s2, P.Y, F);
  if (fp_is_zero(h)) {
    if (fp_is_zero(r)) { Jac T = P; jac_dbl_rare<SOL>(&T, &C); P = T; } else jac_set_inf(P, C);
    return;
  }
  fp_mul<SOL>(P.Z, P.Z, h, F);          // Z3 = Z1 H
  fp_sqr<SOL>(zz, h, F);                // HH
  fp_mul<SOL>(h, h, zz, F);             // HHH
  fp_mul<SOL>(u2, P.X, zz, F);          // V = X1 HH
  fp_sqr<SOL>(s2, r, F);
  fp_sub(s2, s2, h, F);
  fp_sub(s2, s2, u2, F);
  fp_sub(P.X, s2, u2, F);          // X3 = r^2 - HHH - 2V
  fp_sub(u2, u2, P.X, F);
  fp_mul<SOL>(u2, r, u2, F);
  fp_mul<SOL>(h, P.Y, h, F);
  fp_sub(P.Y, u2, h, F);           // Y3 = r (V - X3) - Y1 HHH
}

// P <- P + Q (both Jacobian).  12M + 4S.
template <bool SOL>
VMX_DEV void jac_add(Jac& P, const Jac& Q, const EcCurve& C) {
  const Fp256& F = C.F;
  if (jac_is_inf(Q)) return;
  if (jac_is_inf(P)) { P = Q; return; }
  uint32_t z1z1[8], z2z2[8], u1[8], u2[8], s1[8], s2[8];
  fp_sqr<SOL>(z1z1, P.Z, F);
  fp_sqr<SOL>(z2z2, Q.Z, F);
  fp_mul<SOL>(u1, P.X, z2z2, F);
  fp_mul<SOL>(u2, Q.X, z1z1, F);
  fp_mul<SOL>(s1, Q.Z, z2z2, F);
  fp_mul<SOL>(s1, P.Y, s1, F);
  fp_mul<SOL>(s2, P.Z, z1z1, F);
  fp_mul<SOL>(s2, Q.Y, s2, F);
  fp_sub(u2, u2, u1, F);           // H
  fp_sub(s2, s2, s1, F);           // r
  if (fp_is_zero(u2)) {
    if (fp_is_zero(s2)) { Jac T = P; jac_dbl_rare<SOL>(&T, &C); P = T; } else jac_set_inf(P, C);
    return;
  }
  fp_mul<SOL>(P.Z, P.Z, Q.Z, F);
  fp_mul<SOL>(P.Z, P.Z, u2, F);         // Z3 = Z1 Z2 H
  fp_sqr<SOL>(z1z1, u2, F);             // HH
  fp_mul<SOL>(u2, u2, z1z1, F);         // HHH
  fp_mul<SOL>(u1, u1, z1z1, F);         // V = U1 HH
  fp_sqr<SOL>(z2z2, s2, F);
  fp_sub(z2z2, z2z2, u2, F);
  fp_sub(z2z2, z2z2, u1, F);
  fp_sub(P.X, z2z2, u1, F);        // X3
  fp_sub(u1, u1, P.X, F);
  fp_mul<SOL>(u1, s2, u1, F);
  fp_mul<SOL>(s1, s1, u2, F);
  fp_sub(P.Y, u1, s1, F);          // Y3
}

// y^2 == x^3 + a x + b ?   (x, y in Montgomery form)
template <bool SOL>
VMX_DEV bool ec_on_curve(const uint32_t (&x)[8], const uint32_t (&y)[8], const EcCurve& C) {
  const Fp256& F = C.F;
  uint32_t l[8], r[8];
  fp_sqr<SOL>(l, y, F);
  fp_sqr<SOL>(r, x, F);
  fp_add(r, r, C.a, F);
  fp_mul<SOL>(r, r, x, F);
  fp_add(r, r, C.b, F);
  return fp_eq(l, r);
}

// ---- element access.  Affine arrays: 16 limbs (x = limbs 0..7, y = limbs 8..15).
VMX_DEV void ec_load_affine(uint32_t (&x)[8], uint32_t (&y)[8], const uint32_t* d, size_t cap, size_t i) {
  const uint4* p = reinterpret_cast<const uint4*>(d) + i;
  const uint4 v0 = p[0], v1 = p[cap], v2 = p[2 * cap], v3 = p[3 * cap];
  x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w; x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
  y[0] = v2.x; y[1] = v2.y; y[2] = v2.z; y[3] = v2.w; y[4] = v3.x; y[5] = v3.y; y[6] = v3.z; y[7] = v3.w;
}
VMX_DEV void ec_store_affine(const uint32_t (&x)[8], const uint32_t (&y)[8], uint32_t* d, size_t cap, size_t i) {
  uint4* p = reinterpret_cast<uint4*>(d) + i;
  p[0] = make_uint4(x[0], x[1], x[2], x[3]);
  p[cap] = make_uint4(x[4], x[5], x[6], x[7]);
  p[2 * cap] = make_uint4(y[0], y[1], y[2], y[3]);
  p[3 * cap] = make_uint4(y[4], y[5], y[6], y[7]);
}
VMX_DEV void ec_store_affine_inf(uint32_t* d, size_t cap, size_t i) {
  uint4* p = reinterpret_cast<uint4*>(d) + i;
  const uint4 f = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
  p[0] = f; p[cap] = f; p[2 * cap] = f; p[3 * cap] = f;
}
// 8-limb residue arrays (2 planes)
VMX_DEV void fp_load(uint32_t (&x)[8], const uint32_t* d, size_t cap, size_t i) {
  const uint4* p = reinterpret_cast<const uint4*>(d) + i;
  const uint4 v0 = p[0], v1 = p[cap];
  x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w; x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
}
VMX_DEV void fp_store(const uint32_t (&x)[8], uint32_t* d, size_t cap, size_t i) {
  uint4* p = reinterpret_cast<uint4*>(d) + i;
  p[0] = make_uint4(x[0], x[1], x[2], x[3]);
  p[cap] = make_uint4(x[4], x[5], x[6], x[7]);
}
// Jacobian scratch arrays: 24 limbs (X, Y, Z), 6 planes
VMX_DEV void ec_load_jac(Jac& P, const uint32_t* d, size_t cap, size_t i) {
  fp_load(P.X, d, cap, i);
  fp_load(P.Y, d + 8 * cap, cap, i);
  fp_load(P.Z, d + 16 * cap, cap, i);
}
VMX_DEV void ec_store_jac(const Jac& P, uint32_t* d, size_t cap, size_t i) {
  fp_store(P.X, d, cap, i);
  fp_store(P.Y, d + 8 * cap, cap, i);
  fp_store(P.Z, d + 16 * cap, cap, i);
}

}  // namespace vmx
