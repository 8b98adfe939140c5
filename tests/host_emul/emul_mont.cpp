// Host emulation harness (TESTS ONLY): runs the limb-level Montgomery algorithm of
// csrc/mont.cuh on the CPU through the VMX_HOST_EMUL primitives and prints results for
// comparison with Python bigints.  Usage: emul_mont <N> < input  (hex words: n[N] n0inv, then
// pairs of a[N] b[N] until EOF) -> prints r[N] per pair.
#define VMX_HOST_EMUL 1
#include <cstdio>
#include <cstdlib>
#include "../../verificatum-vmn_b200/csrc/mont.cuh"

template <int N> int run() {
  vmx::MontParams<N> M;
  for (int i = 0; i < N; i++) if (scanf("%x", &M.n[i]) != 1) return 1;
  if (scanf("%x", &M.n0inv) != 1) return 1;
  uint32_t a[N], b[N];
  for (;;) {
    for (int i = 0; i < N; i++) if (scanf("%x", &a[i]) != 1) return 0;
    for (int i = 0; i < N; i++) if (scanf("%x", &b[i]) != 1) return 1;
    vmx::mont_mul<N>(a, [&](int i) { return vmx::Word2{b[i], b[i + 1]}; }, M);
    for (int i = 0; i < N; i++) printf("%08x ", a[i]);
    printf("\n");
  }
}
int main(int argc, char** argv) {
  int n = atoi(argv[1]);
  if (n == 96) return run<96>();
  if (n == 64) return run<64>();
  if (n == 32) return run<32>();
  if (n == 16) return run<16>();
  return 2;
}
