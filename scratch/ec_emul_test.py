import os, sys, random, importlib
sys.path.insert(0, "/root/repo")
import __graft_entry__ as ge
os.environ["VMX_LIBRARY_PATH"] = ge.build_host_emul()
vmx = importlib.import_module("verificatum-vmn_b200")
A = vmx.arithm
from oracle import ec as oec
from oracle.crypto import SeededRandomSource
OG = oec.ECqPGroup("P-256")
G = A.ECqPGroup("P-256")
print("ctx ok", G.elem_bytes, G.ring_bytes)
def to_oracle(el):
    x, y = G._unpack(el.value)
    return oec.UNIT if (x, y) == (-1, -1) else oec.ECPoint(x, y)
def from_oracle(P):
    return G.getONE() if P.is_unit() else A.PGroupElement(G, G._pack(P.x, P.y))
rnd = random.Random(7)
n = 37
exps = [rnd.randrange(OG.q) for _ in range(n)]
exps[3] = 0; exps[5] = 1; exps[6] = OG.q - 1
R = G.getPRing()
e = R.toElementArray([R.toElement(x) for x in exps])
# fixed base
arr = G.getg().exp(e)
got = [to_oracle(x) for x in arr.elements()]
want = [OG.op_exp(OG.g, x) for x in exps]
assert got == want, "exp_fixed"
print("exp_fixed ok")
# mul, inv
b = G.getg().exp(R.toElementArray([R.toElement(rnd.randrange(OG.q)) for _ in range(n)]))
bo = [to_oracle(x) for x in b.elements()]
m = arr.mul(b)
assert [to_oracle(x) for x in m.elements()] == [OG.op_mul(u, v) for u, v in zip(want, bo)], "mul"
mi = arr.mul(arr.inv())
assert all(to_oracle(x).is_unit() for x in mi.elements()), "inv"
dbl = arr.mul(arr)
assert [to_oracle(x) for x in dbl.elements()] == [OG.op_mul(u, u) for u in want], "dbl"
print("mul/inv ok")
# exp_var, exp_scalar
f = [rnd.randrange(OG.q) for _ in range(n)]; f[0] = 0; f[1] = 1
fa = R.toElementArray([R.toElement(x) for x in f])
v = arr.exp(fa)
assert [to_oracle(x) for x in v.elements()] == [OG.op_exp(u, k) for u, k in zip(want, f)], "exp_var"
sc = R.toElement(rnd.randrange(2**200))
s = arr.exp(sc)
assert [to_oracle(x) for x in s.elements()] == [OG.op_exp(u, sc.value) for u in want], "exp_scalar"
print("exp_var/scalar ok")
# expProd, prod
ep = arr.expProd(fa)
acc = oec.UNIT
for u, k in zip(want, f): acc = OG.op_mul(acc, OG.op_exp(u, k))
assert to_oracle(ep) == acc, "expProd"
pr = arr.prod()
acc = oec.UNIT
for u in want: acc = OG.op_mul(acc, u)
assert to_oracle(pr) == acc, "prod"
print("expProd/prod ok")
# serialisation
from oracle import bytetree as obt
tb = arr.toByteTree().to_bytes()
assert tb == OG.leaf_array_tree(want).to_bytes(), "array tree"
back = G.toElementArray(n, vmx.eio.ByteTreeReader(tb))
assert back.equals(arr)
assert G.getg().toByteTree().to_bytes() == OG.leaf_tree(OG.g).to_bytes()
print("byte trees ok")
# random elements
seed = bytes(range(32))
prg = vmx.crypto.PRGHeuristic(); prg.setSeed(seed)
ra = G.randomElementArray(20, prg, 100)
ors = SeededRandomSource(seed)
assert [to_oracle(x) for x in ra.elements()] == OG.random_array(20, ors, 100), "random"
nxt = prg.getBytes(8); assert nxt == ors.get_bytes(8), "prg position"
print("random ok")
# single element ops
g = G.getg()
x = R.toElement(rnd.randrange(OG.q))
assert to_oracle(g.exp(x)) == OG.op_exp(OG.g, x.value)
h = from_oracle(want[8])
assert to_oracle(h.exp(x)) == OG.op_exp(want[8], x.value)
assert to_oracle(h.inv()) == OG.op_inv(want[8])
assert to_oracle(h.mul(g)) == OG.op_mul(want[8], OG.g)
assert to_oracle(h.mul(h.inv())).is_unit()
# cols
cols = G.expProd([arr, b, m], [3, -2, 1], 3)
assert [to_oracle(x) for x in cols.elements()] == [OG.op_mul(OG.op_mul(OG.op_exp(u, 3), OG.op_inv(OG.op_exp(v, 2))), w) for u, v, w in zip(want, bo, [OG.op_mul(u, v) for u, v in zip(want, bo)])], "cols"
# movement
sp = arr.shiftPush(g)
assert [to_oracle(x) for x in sp.elements()] == [OG.g] + want[:-1]
assert to_oracle(arr.get(4)) == want[4]
assert [to_oracle(x) for x in arr.copyOfRange(2, 9).elements()] == want[2:9]
print("ALL OK")
