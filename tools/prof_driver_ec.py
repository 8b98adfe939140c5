"""Micro-timings of the curve-group kernels (P-256).  Usage: prof_driver_ec.py [n]"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
vmx = importlib.import_module("verificatum-vmn_b200")
A = vmx.arithm
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
G = A.ECqPGroup("P-256")
R = G.getPRing()
rs = vmx.crypto.PRGHeuristic(); rs.setSeed(bytes(range(32)))
stream = torch.cuda.ExternalStream(G._lib.vmx_ctx_stream(G.ctx))
PEAK = 9.26e12
def timed(label, fn, reps=2):
    fn(); G.sync()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    mm0 = G.modmul_count()
    e0.record(stream)
    for _ in range(reps): fn()
    e1.record(stream); e1.synchronize()
    ms = e0.elapsed_time(e1) / reps
    mm = (G.modmul_count() - mm0) / reps
    print("%-28s %9.3f ms  %14.0f fieldmuls  %7.2f Gmul/s  %5.1f%% of IMAD peak @136 MAC" % (label, ms, mm, mm / ms / 1e6, 100 * mm * 136 / (ms * 1e-3) / PEAK), flush=True)
t0 = time.time()
e = R.randomElementArray(n, rs, 100)
X = G.randomElementArray(n, rs, 100)
G.sync()
print("setup (random exponents + %d random points): %.1f ms" % (n, (time.time() - t0) * 1e3))
e256 = R.randomElementArray(n, rs, 100)
v = R.toElement(int.from_bytes(bytes(range(32)), "big"))
t0 = time.time(); G.getg().exp(e).free(); G.sync(); print("first fixed exp incl. table build: %.1f ms" % ((time.time() - t0) * 1e3))
timed("exp_fixed (256 bit)", lambda: G.getg().exp(e).free())
timed("exp_var (256 bit)", lambda: X.exp(e256).free())
timed("exp_scalar (256 bit)", lambda: X.exp(v).free())
timed("expProd (256 bit)", lambda: X.expProd(e256))
timed("mul", lambda: X.mul(X).free())
timed("inv", lambda: X.inv().free())
timed("prod", lambda: X.prod())
timed("random points", lambda: G.randomElementArray(n, rs, 100).free(), reps=1)
m = X.toByteTree().to_bytes()
timed("import (on-curve check)", lambda: G.toElementArray(n, vmx.eio.ByteTreeReader(m)).free())
X._leaves = None
timed("export", lambda: (setattr(X, "_leaves", None), X.leaves()))
