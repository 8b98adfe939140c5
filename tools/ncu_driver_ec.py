"""ncu target for the curve kernels: exactly one fixed-base exp, one variable-base exp, one expProd and one mul
over n P-256 points, in that order (table construction first).  Usage under ncu:
  ncu --set full --clock-control none -k regex:"^k_ec_exp_fixed|^k_ec_exp_var|^k_ec_seg_sum|^k_ec_add" -c 4 ...
(the first k_ec_seg_sum launch of the expProd is the bucket accumulation over all n * W terms)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
vmx = importlib.import_module("verificatum-vmn_b200")
A = vmx.arithm
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
G = A.ECqPGroup("P-256")
R = G.getPRing()
rs = vmx.crypto.PRGHeuristic(); rs.setSeed(bytes(range(32)))
e = R.randomElementArray(n, rs, 100)
X = G.randomElementArray(n, rs, 100)
G.precomputeFixedBase(G.getg(), n)
G.sync()
F = G.getg().exp(e)
V = X.exp(e)
P = X.expProd(e)
M = X.mul(F)
G.sync()
print("done", n)
