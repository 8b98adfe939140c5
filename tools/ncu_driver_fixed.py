"""ncu target: one fixed-base exponentiation of n 3071-bit exponents (k_exp_fixed<96>), table built before."""
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
G.precomputeFixedBase(G.getg(), n)
G.sync()
G.getg().exp(e).free()
G.getg().exp(e).free()
G.sync()
print("done")
