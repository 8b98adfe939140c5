"""Host timeline (every C-ABI call with its start / end) of device-resident re-encrypt + prove + verify steps:
where does the host block, and how long is the GPU left without work?  usage: trace_device_step.py N out.json"""
import importlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["VMX_TRACE"] = "1"
import torch
vmx = importlib.import_module("verificatum-vmn_b200")
A = vmx.arithm
hvzk = importlib.import_module("verificatum-vmn_b200.hvzk")
mixnet = importlib.import_module("verificatum-vmn_b200.mixnet")
groups = importlib.import_module("verificatum-vmn_b200.groups")
tr = importlib.import_module("verificatum-vmn_b200._trace")
crypto = vmx.crypto
n = int(sys.argv[1])
G = A.ModPGroup(*groups.rfc3526(3072))
stream = torch.cuda.ExternalStream(G._lib.vmx_ctx_stream(G.ctx))
def prg(label):
    r = crypto.PRGHeuristic(); r.setSeed(crypto.HashfunctionHeuristic("SHA-256").hash(label.encode())); return r
rs0 = prg("setup")
x = G.getPRing().randomElement(rs0, 100)
pk = A.PPGroup(G, 2).product(G.getg(), G.getg().exp(x))
w = mixnet.demoCiphertexts(pk, n, rs0)
params = mixnet.SessionParams(pGroupString="trace")
h = mixnet.ShufflerSession(G, pk, params, prg("p")).deriveGenerators(n)
seed = bytes(range(32)); challenge = int.from_bytes(bytes(range(32)), "big")
def step(i):
    rs = prg("step%d" % i)
    P = hvzk.PoSBasicTW(256, 256, 100, crypto.PRGHeuristic(), rs)
    V = hvzk.PoSBasicTW(256, 256, 100, crypto.PRGHeuristic(), rs)
    with tr.span("phase.reencrypt"):
        s = G.getPRing().randomElementArray(n, rs, 100)
        f = pk.exp(s)
        pi = A.Permutation.random(n, rs, 100, G)
        re = w.mul(f); f.free()
        out = re.permute(pi.inv()); re.free()
    with tr.span("phase.precompute"):
        P.precompute(G.getg(), h, pi)
    with tr.span("phase.commit"):
        P.setInstance(pk, w, out, s); P.commit(seed)
    with tr.span("phase.reply"):
        P.reply(challenge)
    with tr.span("phase.computeAF"):
        V.precompute(G.getg(), h); V.setInstance(pk, w, out); V.u = P.u; V.setBatchVector(seed); V.computeAF()
    with tr.span("phase.checks"):
        V.B, V.Ap, V.Bp, V.Cp, V.Dp, V.Fp = P.B, P.Ap, P.Bp, P.Cp, P.Dp, P.Fp
        V.setChallenge(challenge)
        V.k_A, V.k_B, V.k_C, V.k_D, V.k_E, V.k_F = P.k_A, P.k_B, P.k_C, P.k_D, P.k_E, P.k_F
        assert V.verifyParsed()
    V.e.free(); V.u = V.e = V.B = V.Bp = V.k_B = V.k_E = None
    P.free(); s.free(); out.free()
for i in range(2):
    step(i)
G.sync()
tr.start()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
t0 = time.time(); e0.record(stream)
for i in range(2, 5):
    with tr.span("step"):
        step(i)
e1.record(stream); e1.synchronize()
print("3 steps: device %.1f ms, host %.1f ms" % (e0.elapsed_time(e1), (time.time() - t0) * 1e3))
json.dump(tr.stop(), open(sys.argv[2], "w"))
