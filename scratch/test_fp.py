import random, subprocess
P=0xFFFFFFFF00000001000000000000000000000000FFFFFFFFFFFFFFFFFFFFFFFF
Q=0xFFFFFFFF00000000FFFFFFFFFFFFFFFFBCE6FAADA7179E84F3B9CAC2FC632551
K=2**256-2**32-977
def limbs(x): return " ".join("%08x"%((x>>(32*i))&0xffffffff) for i in range(8))
def run(n, sol):
    n0inv=(-pow(n,-1,2**32))%2**32
    rnd=random.Random(1)
    cases=[(rnd.randrange(n),rnd.randrange(n)) for _ in range(300)]
    cases+=[(n-1,n-1),(0,0),(1,n-1),(n-1,1),(0,5),(2**255%n,2**255%n)]
    inp=limbs(n)+" %08x %x\n"%(n0inv,sol)+"\n".join(limbs(a)+" "+limbs(b) for a,b in cases)+"\n"
    out=subprocess.run(["scratch/emul_fp"],input=inp,capture_output=True,text=True).stdout.strip().split("\n")
    Rinv=pow(2**256,-1,n)
    bad=0
    for k,(a,b) in enumerate(cases):
        got=[sum(int(w,16)<<(32*i) for i,w in enumerate(l.split())) for l in out[4*k:4*k+4]]
        exp=[a*b*Rinv%n,(a+b)%n,(a-b)%n,(-a)%n]
        if got!=exp: bad+=1; print("MISMATCH",k,[g==e for g,e in zip(got,exp)])
    print(hex(n)[:12],"solinas",sol,"bad",bad,"of",len(cases))
run(P,1); run(P,0); run(Q,0); run(K,0)
