// ptx_arith.cuh -- carry-chain primitives for multi-word integer arithmetic on sm_100a.
//
// Every 32x32+64 multiply-accumulate is written as the PTX pair
//     mad{c}.lo.cc.u32 lo, a, b, lo ; madc.hi.cc.u32 hi, a, b, hi
// which ptxas (12.9, sm_100a) fuses into ONE `IMAD.WIDE.U32.X Rd, Pout, Ra, Rb, Rc, Pin`
// provided (lo,hi) can live in an even-aligned register pair.  That instruction issues at
// 1 warp-instr / 4 clk / SMSP on B200 (measured, profiles/r01_ubench_imad.txt), i.e.
// 32 MAC/clk/SM -- the roofline unit of this engine.
//
// The asm statements are `volatile` so that the (invisible to the compiler) carry flag
// dependency between consecutive statements is preserved.
//
// When compiled with -DVMX_HOST_EMUL (tests only, g++), the same primitives are emulated with
// an explicit carry variable so the limb algorithms can be unit-tested on the CPU against
// Python bigints.  The shipped library never defines VMX_HOST_EMUL.
#pragma once
#include <cstdint>

#ifdef VMX_HOST_EMUL
#define VMX_DEV inline
namespace vmx_emul { static thread_local uint32_t CC = 0; }
// (lo,hi) += a*b                     sets carry
VMX_DEV void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  unsigned __int128 s = (unsigned __int128)a * b + (((uint64_t)hi << 32) | lo);
  lo = (uint32_t)s; hi = (uint32_t)(s >> 32); vmx_emul::CC = (uint32_t)(s >> 64);
}
// (lo,hi) += a*b + carry             sets carry
VMX_DEV void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  unsigned __int128 s = (unsigned __int128)a * b + (((uint64_t)hi << 32) | lo) + vmx_emul::CC;
  lo = (uint32_t)s; hi = (uint32_t)(s >> 32); vmx_emul::CC = (uint32_t)(s >> 64);
}
// (dlo,dhi) = a*b + (clo,chi) [+ carry]   sets carry
VMX_DEV void mad_wide_cc3(uint32_t& dlo, uint32_t& dhi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
  unsigned __int128 s = (unsigned __int128)a * b + (((uint64_t)chi << 32) | clo);
  dlo = (uint32_t)s; dhi = (uint32_t)(s >> 32); vmx_emul::CC = (uint32_t)(s >> 64);
}
VMX_DEV void madc_wide_cc3(uint32_t& dlo, uint32_t& dhi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
  unsigned __int128 s = (unsigned __int128)a * b + (((uint64_t)chi << 32) | clo) + vmx_emul::CC;
  dlo = (uint32_t)s; dhi = (uint32_t)(s >> 32); vmx_emul::CC = (uint32_t)(s >> 64);
}
VMX_DEV void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  uint64_t s = (uint64_t)a * b; lo = (uint32_t)s; hi = (uint32_t)(s >> 32);
}
VMX_DEV void add_cc(uint32_t& d, uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b; d = (uint32_t)s; vmx_emul::CC = (uint32_t)(s >> 32); }
VMX_DEV void addc_cc(uint32_t& d, uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b + vmx_emul::CC; d = (uint32_t)s; vmx_emul::CC = (uint32_t)(s >> 32); }
VMX_DEV void addc(uint32_t& d, uint32_t a, uint32_t b) { d = a + b + vmx_emul::CC; }
VMX_DEV void sub_cc(uint32_t& d, uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a - b; d = (uint32_t)s; vmx_emul::CC = (uint32_t)((s >> 32) & 1); }
VMX_DEV void subc_cc(uint32_t& d, uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a - b - vmx_emul::CC; d = (uint32_t)s; vmx_emul::CC = (uint32_t)((s >> 32) & 1); }
VMX_DEV void subc(uint32_t& d, uint32_t a, uint32_t b) { d = a - b - vmx_emul::CC; }
#else
#define VMX_DEV __device__ __forceinline__
VMX_DEV void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
VMX_DEV void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
VMX_DEV void mad_wide_cc3(uint32_t& dlo, uint32_t& dhi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;" : "=r"(dlo), "=r"(dhi) : "r"(a), "r"(b), "r"(clo), "r"(chi));
}
VMX_DEV void madc_wide_cc3(uint32_t& dlo, uint32_t& dhi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;" : "=r"(dlo), "=r"(dhi) : "r"(a), "r"(b), "r"(clo), "r"(chi));
}
VMX_DEV void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
}
VMX_DEV void add_cc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
VMX_DEV void addc_cc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
VMX_DEV void addc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("addc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
VMX_DEV void sub_cc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
VMX_DEV void subc_cc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
VMX_DEV void subc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("subc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
#endif
