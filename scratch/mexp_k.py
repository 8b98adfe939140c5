"""A/B of the Pippenger chunk size (VMX_MEXP_K) and timings of the fixed-base kernel with the wider tables."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
vmx = importlib.import_module("verificatum-vmn_b200")
A = vmx.arithm
groups = importlib.import_module("verificatum-vmn_b200.groups")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
G = A.ModPGroup(*groups.rfc3526(3072))
R = G.getPRing()
rs = vmx.crypto.PRGHeuristic(); rs.setSeed(bytes(range(32)))
stream = torch.cuda.ExternalStream(G._lib.vmx_ctx_stream(G.ctx))
def timed(label, fn, reps=3):
    fn(); G.sync()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    mm0 = G.modmul_count()
    e0.record(stream)
    for _ in range(reps): fn()
    e1.record(stream); e1.synchronize()
    ms = e0.elapsed_time(e1) / reps
    mm = (G.modmul_count() - mm0) / reps
    print("K=%s %-22s %9.3f ms %12.0f modmuls %6.1f%% of IMAD peak" % (os.environ.get("VMX_MEXP_K", "32"), label, ms, mm, 100 * mm * 18528 / (ms * 1e-3) / 9.26e12), flush=True)
X = G.randomElementArray(n, rs, 100)
kE = R.toElementArray(A.LargeIntegerArray.random(n, 613, rs, R))
e256 = R.toElementArray(A.LargeIntegerArray.random(n, 256, rs, R))
timed("expProd (613 bit)", lambda: X.expProd(kE))
timed("expProd (256 bit)", lambda: X.expProd(e256))
if os.environ.get("VMX_MEXP_K", "32") == "32":
    e = R.randomElementArray(n, rs, 100)
    timed("exp_fixed (3071 bit)", lambda: G.getg().exp(e).free())
