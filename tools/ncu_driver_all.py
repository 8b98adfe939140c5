"""ncu target: the three dominant kernels of a step at n elements (3072 bit), tables built before:
k_exp_fixed<96> (fixed base, full-length exponents), k_exp_var2<96> (x^v * y^k, 256 / 613 bit), k_seg_prod<96>
(Pippenger bucket accumulation of a 613-bit multi-exponentiation)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
vmx = importlib.import_module("verificatum-vmn_b200")
A = vmx.arithm
groups = importlib.import_module("verificatum-vmn_b200.groups")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
G = A.ModPGroup(*groups.rfc3526(3072))
R = G.getPRing()
rs = vmx.crypto.PRGHeuristic(); rs.setSeed(bytes(range(32)))
e = R.randomElementArray(n, rs, 100)
X = G.randomElementArray(n, rs, 100)
Y = G.randomElementArray(n, rs, 100)
kE = R.toElementArray(A.LargeIntegerArray.random(n, 613, rs, R))
v = R.toElement(int.from_bytes(bytes(range(32)), "big"))
G.precomputeFixedBase(G.getg(), n)
G.sync()
for _ in range(2):
    G.getg().exp(e).free()
    X.expMulExp(v, Y, kE).free()
    X.expProd(kE)
G.sync()
print("done")
