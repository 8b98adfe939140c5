import sys, importlib, ctypes as C
sys.path.insert(0, ".")
vmx = importlib.import_module("verificatum-vmn_b200"); A = vmx.arithm; cr = vmx.crypto
groups = importlib.import_module("verificatum-vmn_b200.groups")
lib = vmx._native.load()
for name, (p, q, g) in (("512", groups.test512()), ("2048", groups.rfc3526(2048)), ("3072", groups.rfc3526(3072))):
    G = A.ModPGroup(p, q, g)
    rs = cr.PRGHeuristic(); rs.setSeed(bytes(range(32)))
    X1 = G.randomElementArray(9, rs, 100); X2 = G.randomElementArray(9, rs, 100)
    h = C.c_void_p(); vmx._native.check(lib.vmx_debug_coop_mul(X1.h, X2.h, C.byref(h)))
    Y = A.PGroupElementArray(G, h)
    a = [e.value for e in X1.elements()]; b = [e.value for e in X2.elements()]; y = [e.value for e in Y.elements()]
    bad = 0
    for i in range(9):
        want = a[i]*b[i] % p
        if y[i] != want:
            bad += 1
            if bad <= 2:
                # to_bytes multiplies by R^-1: compare Montgomery residues to see the error pattern
                R = 1 << (32 * ((p.bit_length()+31)//32 if p.bit_length() > 512 else 16))
                d = (y[i] - want) % p
                print(name, "elem", i, "mismatch; diff*R mod p =", hex(d * R % p)[:80], " (p-diff)*R:", hex((p-d)*R % p)[:80])
    print(name, "bad", bad, "of 9")
