#define VMX_HOST_EMUL 1
#include <cstdio>
#include <cstdlib>
#include "../verificatum-vmn_b200/csrc/fp256.cuh"
using namespace vmx;
int main() {
  Fp256 F;
  for (int i = 0; i < 8; i++) if (scanf("%x", &F.n[i]) != 1) return 1;
  if (scanf("%x %x", &F.n0inv, &F.solinas) != 2) return 1;
  uint32_t a[8], b[8], r[8];
  for (;;) {
    for (int i = 0; i < 8; i++) if (scanf("%x", &a[i]) != 1) return 0;
    for (int i = 0; i < 8; i++) if (scanf("%x", &b[i]) != 1) return 1;
    fp_mul(r, a, b, F); for (int i = 0; i < 8; i++) printf("%08x ", r[i]); printf("\n");
    fp_add(r, a, b, F); for (int i = 0; i < 8; i++) printf("%08x ", r[i]); printf("\n");
    fp_sub(r, a, b, F); for (int i = 0; i < 8; i++) printf("%08x ", r[i]); printf("\n");
    fp_neg(r, a, F); for (int i = 0; i < 8; i++) printf("%08x ", r[i]); printf("\n");
  }
}
