import os, sys, importlib, random, time
pass
sys.path.insert(0, ".")
vmx = importlib.import_module("verificatum-vmn_b200")
A = vmx.arithm
import numpy as np
# 512-bit safe prime group for speed: find one deterministically
def is_pp(n):
    return pow(2, n-1, n) == 1 and pow(3, n-1, n) == 1
random.seed(5)
def gen_safe(bits):
    while True:
        q = random.getrandbits(bits-1) | (1 << (bits-2)) | 1
        if q % 3 != 2: continue
        if is_pp(q) and is_pp(2*q+1): return 2*q+1, q
p, q = gen_safe(512)
g = 4
G = A.ModPGroup(p, q, g)
R = G.getPRing()
n = 50
rnd = random.Random(1)
xs = [pow(g, rnd.randrange(q), p) for _ in range(n)]
es = [rnd.randrange(q) for _ in range(n)]
t0=time.time()
X = G.toElementArray([A.PGroupElement(G, x) for x in xs])
E = R.toElementArray([A.PFieldElement(R, e) for e in es])
assert [e.value for e in X.elements()] == xs
assert [e.value for e in E.elements()] == es
print("codec ok", time.time()-t0)
Y = X.mul(X); assert [e.value for e in Y.elements()] == [x*x % p for x in xs]; print("mul ok")
F = G.getg().exp(E); assert [e.value for e in F.elements()] == [pow(g, e, p) for e in es]; print("exp_fixed ok", time.time()-t0)
V = X.exp(E); assert [e.value for e in V.elements()] == [pow(x, e, p) for x, e in zip(xs, es)]; print("exp_var ok", time.time()-t0)
s = A.PFieldElement(R, rnd.randrange(1<<200))
S = X.exp(s); assert [e.value for e in S.elements()] == [pow(x, s.value, p) for x in xs]; print("exp_scalar ok")
ep = X.expProd(E)
ref = 1
for x, e in zip(xs, es): ref = ref * pow(x, e, p) % p
assert ep.value == ref, (hex(ep.value), hex(ref)); print("expprod ok", time.time()-t0)
pr = X.prod(); ref = 1
for x in xs: ref = ref*x % p
assert pr.value == ref; print("prod ok")
# ---- ring ops
n2 = 300
as_ = [rnd.randrange(q) for _ in range(n2)]
bs_ = [rnd.randrange(q) for _ in range(n2)]
Ar = R.toElementArray([A.PFieldElement(R, e) for e in as_])
Br = R.toElementArray([A.PFieldElement(R, e) for e in bs_])
vals = lambda arr: [e.value for e in arr.elements()]
assert vals(Ar.add(Br)) == [(a+b) % q for a, b in zip(as_, bs_)]
assert vals(Ar.neg()) == [(-a) % q for a in as_]
assert vals(Ar.mul(Br)) == [(a*b) % q for a, b in zip(as_, bs_)]
sc = A.PFieldElement(R, rnd.randrange(q))
assert vals(Ar.mulAdd(sc, Br)) == [(a*sc.value+b) % q for a, b in zip(as_, bs_)]
assert Ar.innerProduct(Br).value == sum(a*b for a, b in zip(as_, bs_)) % q
assert Ar.sum().value == sum(as_) % q
ref = 1
for a in as_: ref = ref*a % q
assert Ar.prod().value == ref
pr = []; acc = 1
for a in as_: acc = acc*a % q; pr.append(acc)
assert vals(Ar.prods()) == pr
x, d = Br.recLin(Ar)
xs_ = [bs_[0]]
for i in range(1, n2): xs_.append((xs_[-1]*as_[i] + bs_[i]) % q)
assert vals(x) == xs_ and d.value == xs_[-1]
print("ring ok")
perm = list(range(n2)); rnd.shuffle(perm)
P = A.Permutation(perm)
out = vals(Ar.permute(P))
exp_ = [0]*n2
for i in range(n2): exp_[perm[i]] = as_[i]
assert out == exp_
assert vals(Ar.shiftPush(sc)) == [sc.value] + as_[:-1]
assert vals(Ar.copyOfRange(3, 17)) == as_[3:17]
assert Ar.get(7).value == as_[7]
assert Ar.equals(Ar) and not Ar.equals(Br)
assert X.shiftPush(G.getg()).elements()[0].value == g
keep = [i % 3 == 0 for i in range(n)]
assert [e.value for e in X.extract(keep).elements()] == [x for x, k in zip(xs, keep) if k]
print("movement ok")
# PRG
from importlib import import_module
cr = vmx.crypto
prg = cr.PRGHeuristic(); prg.setSeed(bytes(range(32)))
lia = A.LargeIntegerArray.random(77, 100, prg, R)
ev = vals(R.toElementArray(lia))
prg2 = cr.PRGHeuristic(); prg2.setSeed(bytes(range(32)))
exp_ = []
for i in range(77):
    b = prg2.getBytes(13); exp_.append(int.from_bytes(b, "big") & ((1<<100)-1))
assert ev == exp_
assert prg.getBytes(40) == prg2.getBytes(40)
print("prg ok")
rs = cr.PRGHeuristic(); rs.setSeed(bytes(range(32)))
ra = R.randomElementArray(20, rs, 100)
rs2 = cr.PRGHeuristic(); rs2.setSeed(bytes(range(32)))
bits = q.bit_length()+100; w=(bits+7)//8
assert vals(ra) == [(int.from_bytes(rs2.getBytes(w), "big") & ((1<<bits)-1)) % q for _ in range(20)]
rs = cr.PRGHeuristic(); rs.setSeed(bytes(range(32)))
ga = G.randomElementArray(10, rs, 100)
rs2 = cr.PRGHeuristic(); rs2.setSeed(bytes(range(32)))
bits = p.bit_length()+100; w=(bits+7)//8
assert [e.value for e in ga.elements()] == [pow((int.from_bytes(rs2.getBytes(w), "big") & ((1<<bits)-1)) % p, 2, p) for _ in range(10)]
print("random arrays ok")
# membership
try:
    bad = G.toElementArray(2, np.frombuffer(b"".join(v.to_bytes(G.elem_bytes,"big") for v in [xs[0], p-1]), dtype=np.uint8))
    assert False
except A.ArithmFormatException: print("membership reject ok")
inv = X.inv(); assert [e.value for e in inv.elements()] == [pow(x, -1, p) for x in xs]
cols = G.expProd([X, Y], [3, -2], 3)
assert [e.value for e in cols.elements()] == [pow(x,3,p)*pow(pow(x*x%p,2,p),-1,p)%p for x in xs]
print("inv/cols ok")
# ---- 3072-bit (RFC 3526) at a size that fills the machine
sys.path.insert(0, "oracle")
import gen_groups
for bits in (2048, 3072):
    p = gen_groups.modp(bits, gen_groups.GROUPS[bits]); q = (p-1)//2; g = 2
    G = A.ModPGroup(p, q, g); R = G.getPRing()
    n = 2000
    es = [rnd.randrange(q) for _ in range(n)]
    E = R.toElementArray([A.PFieldElement(R, e) for e in es])
    t0 = time.time(); F = G.getg().exp(E); G.sync(); t1 = time.time()
    fv = [e.value for e in F.elements()]
    assert fv == [pow(g, e, p) for e in es]; print(bits, "exp_fixed ok", t1-t0)
    ks = [rnd.randrange(1 << 613) for _ in range(n)]
    K = R.toElementArray([A.PFieldElement(R, e) for e in ks])
    t0 = time.time(); V = F.exp(K); G.sync(); t1 = time.time()
    assert [e.value for e in V.elements()] == [pow(x, e, p) for x, e in zip(fv, ks)]; print(bits, "exp_var ok", t1-t0)
    t0 = time.time(); ep = F.expProd(K); t1 = time.time()
    ref = 1
    for x, e in zip(fv, ks): ref = ref * pow(x, e, p) % p
    assert ep.value == ref; print(bits, "expprod ok", t1-t0)
    # raw modmul throughput
    import ctypes as C
    ms = C.c_float()
    lib = vmx._native.load()
    for it in (16, 64):
        vmx._native.check(lib.vmx_bench_modmul(G.ctx, 148*256*4, it, C.byref(ms)))
        nl = bits // 32
        print(bits, "bench_modmul iters", it, ms.value, "ms", 148*256*4*it/(ms.value*1e-3), "modmul/s", 148*256*4*it/(ms.value*1e-3)*(2*nl*nl+nl)/9.26e12, "of IMAD peak")
print("ALL OK")
# cooperative multiplier self test on every size
import ctypes as C
for bits, (p_, q_, g_) in (("512", vmx.arithm and importlib.import_module("verificatum-vmn_b200.groups").test512()),
                           ("2048", importlib.import_module("verificatum-vmn_b200.groups").rfc3526(2048)),
                           ("3072", importlib.import_module("verificatum-vmn_b200.groups").rfc3526(3072))):
    G = A.ModPGroup(p_, q_, g_); R = G.getPRing()
    rs = cr.PRGHeuristic(); rs.setSeed(bytes(range(32)))
    X1 = G.randomElementArray(5000, rs, 100); X2 = G.randomElementArray(5000, rs, 100)
    eq = C.c_int()
    vmx._native.check(vmx._native.load().vmx_selftest_coop(X1.h, X2.h, C.byref(eq)))
    print(bits, "coop selftest equal =", eq.value); assert eq.value == 1
    # edge: p-1 squared
    E1 = G.toElementArray([A.PGroupElement(G, p_-1), A.PGroupElement(G, 1), A.PGroupElement(G, p_-2)])
    vmx._native.check(vmx._native.load().vmx_selftest_coop(E1.h, E1.h, C.byref(eq))); assert eq.value == 1
    t0=time.time(); y = G.getg().exp(R.toElement(q_-5)); t1=time.time()
    assert y.value == pow(g_, q_-5, p_); print(bits, "single exp ok", t1-t0)
    t0=time.time(); yi = y.inv(); t1=time.time(); assert yi.value == pow(y.value, -1, p_); print(bits, "single inv ok", t1-t0)
print("COOP OK")
