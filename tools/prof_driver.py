"""Targeted driver for ncu captures and micro-timings of the big kernels (N = 100k, 3072 bit)."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
vmx = importlib.import_module("verificatum-vmn_b200")
A = vmx.arithm
groups = importlib.import_module("verificatum-vmn_b200.groups")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 3072
p, q, g = groups.rfc3526(bits)
G = A.ModPGroup(p, q, g)
R = G.getPRing()
rs = vmx.crypto.PRGHeuristic(); rs.setSeed(bytes(range(32)))
stream = torch.cuda.ExternalStream(G._lib.vmx_ctx_stream(G.ctx))
def timed(label, fn, reps=2):
    fn(); G.sync()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    mm0 = G.modmul_count()
    e0.record(stream)
    for _ in range(reps): fn()
    e1.record(stream); e1.synchronize()
    ms = e0.elapsed_time(e1) / reps
    mm = (G.modmul_count() - mm0) / reps
    print("%-28s %9.3f ms  %12.0f modmuls  %6.1f%% of IMAD peak" % (label, ms, mm, 100 * mm * (2*(bits//32)**2 + bits//32) / (ms * 1e-3) / 9.26e12), flush=True)
e = R.randomElementArray(n, rs, 100)
X = G.randomElementArray(n, rs, 100)
kE = R.toElementArray(A.LargeIntegerArray.random(n, 613, rs, R))
e256 = R.toElementArray(A.LargeIntegerArray.random(n, 256, rs, R))
v = R.toElement(int.from_bytes(bytes(range(32)), "big"))
quick = len(sys.argv) > 3 and sys.argv[3] == "quick"
for w in ([0] if quick else [0, 16, 17, 18]):
    if w == 18 and n < 500000:
        continue
    G.set_tuning(fixed_window=w)
    timed("exp_fixed (3071 bit) w=%s" % (w or "auto"), lambda: G.getg().exp(e).free())
G.set_tuning(fixed_window=0)
Y = G.randomElementArray(n, rs, 100)
timed("expMulExp (256 / 613 bit)", lambda: X.expMulExp(v, Y, kE).free())
if n <= 200000:
    full = R.toElement(G.q - 12345678901234567890123)
    timed("exp_scalar (3071 bit)", lambda: X.exp(full).free(), reps=1)
timed("exp_var (613 bit)", lambda: X.exp(kE).free())
timed("exp_scalar (256 bit)", lambda: X.exp(v).free())
timed("expProd (613 bit)", lambda: X.expProd(kE))
timed("expProd (256 bit)", lambda: X.expProd(e256))
timed("mul", lambda: X.mul(X).free())
timed("prod", lambda: X.prod())
if quick:
    sys.exit(0)
m = X.to_matrix()
t0 = time.time(); G.toElementArray(n, m, check_membership=False).free(); G.sync(); t1 = time.time()
G.toElementArray(n, m, check_membership=True).free(); G.sync(); t2 = time.time()
timed("import no check", lambda: G.toElementArray(n, m, check_membership=False).free())
timed("import + Jacobi membership", lambda: G.toElementArray(n, m, check_membership=True).free())
print("host wall: import %.1f ms, import+membership %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
# dedicated squaring against the multiplication: `iters` chained operations per element
import ctypes as C
eq, ms, ms2 = C.c_int(), C.c_float(), C.c_float()
iters = 64
vmx._native.check(G._lib.vmx_selftest_sqr(X.h, iters, C.byref(eq), C.byref(ms)))
vmx._native.check(G._lib.vmx_selftest_sqr(X.h, iters, C.byref(eq), C.byref(ms)))
vmx._native.check(G._lib.vmx_bench_modmul(G.ctx, n, iters, C.byref(ms2)))
vmx._native.check(G._lib.vmx_bench_modmul(G.ctx, n, iters, C.byref(ms2)))
macs = 2 * (bits // 32) ** 2 + bits // 32
print("mont_sqr  x%d: %8.3f ms (equal=%d)  %5.1f%% of IMAD peak counted as modmuls" % (iters, ms.value, eq.value, 100 * n * iters * macs / (ms.value * 1e-3) / 9.26e12))
print("mont_mul  x%d: %8.3f ms             %5.1f%% of IMAD peak" % (iters, ms2.value, 100 * n * iters * macs / (ms2.value * 1e-3) / 9.26e12))
print("sqr / mul time = %.3f" % (ms.value / ms2.value))
