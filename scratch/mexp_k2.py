import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
vmx = importlib.import_module("verificatum-vmn_b200")
A = vmx.arithm
groups = importlib.import_module("verificatum-vmn_b200.groups")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
G = A.ModPGroup(*groups.rfc3526(3072))
R = G.getPRing()
rs = vmx.crypto.PRGHeuristic(); rs.setSeed(bytes(range(32)))
X = G.randomElementArray(n, rs, 100)
kE = R.toElementArray(A.LargeIntegerArray.random(n, 613, rs, R))
e256 = R.toElementArray(A.LargeIntegerArray.random(n, 256, rs, R))
G.sync()
for label, e in (("613", kE), ("256", e256), ("613", kE), ("256", e256)):
    for rep in range(4):
        l0 = G.launch_count(); t0 = time.time(); X.expProd(e); G.sync(); dt = time.time() - t0
        print("K=%s expProd %s rep %d: %.2f ms host wall, %d launches, bitlen %d" % (os.environ.get("VMX_MEXP_K", "32"), label, rep, dt * 1e3, G.launch_count() - l0, e.bitLength()), flush=True)
