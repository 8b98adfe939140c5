"""Shared test cases: the same seeded instance built for the oracle (Python ints) and for the
engine (device arrays), so that transcripts can be compared byte for byte."""
import importlib

from oracle import arithm as oar
from oracle import ec as oec
from oracle import protocols as opr
from oracle.crypto import SeededRandomSource


def groups_mod():
    return importlib.import_module("verificatum-vmn_b200.groups")


def group_params(bits: int):
    g = groups_mod()
    return g.test512() if bits == 512 else g.rfc3526(bits)


def oracle_group(spec):
    """spec: bit size of a safe-prime ModPGroup, or the name of a curve ("P-256")."""
    if isinstance(spec, str):
        return oec.ECqPGroup(spec)
    return oar.ModPGroup(*group_params(spec))


def engine_group(vmx, spec):
    A = vmx.arithm
    if isinstance(spec, str):
        return A.ECqPGroup(spec)
    return A.ModPGroup(*group_params(spec))


def elem_value(el):
    """Engine PGroupElement -> the oracle's value of it (int, or ECPoint on a curve)."""
    G = getattr(el, "group", None)
    if G is None:  # ring element
        return el.value
    if getattr(G, "is_curve", False):
        x, y = G._unpack(el.value)
        return oec.UNIT if (x, y) == (-1, -1) else oec.ECPoint(x, y)
    return el.value


def engine_elem(vmx, G, v):
    """Oracle value -> engine PGroupElement."""
    A = vmx.arithm
    if getattr(G, "is_curve", False):
        return G.getONE() if v.is_unit() else A.PGroupElement(G, G._pack(v.x, v.y))
    return A.PGroupElement(G, v)


def seed(label: str) -> bytes:
    import hashlib
    return hashlib.sha256(("vmx-test/" + label).encode()).digest()


class OracleCase:
    def __init__(self, bits: int, n: int, label: str = "case"):
        self.G = oracle_group(bits)
        self.n = n
        self.params = opr.Params(pgroup_string="test-%s" % bits)
        rs = SeededRandomSource(seed(label + "/setup"))
        self.x = oar.ring_random_element(self.G, rs, 100)
        self.pk = (self.G.g, self.G.op_exp(self.G.g, self.x))
        self.w = opr.demo_ciphertexts(self.G, self.pk, n, rs)
        self.h = opr.independent_generators(self.G, "sha256", self.params.prefix(), "generators", n, self.params.rbitlen)


class EngineCase:
    def __init__(self, vmx, bits: int, n: int, label: str = "case"):
        A = vmx.arithm
        mix = importlib.import_module("verificatum-vmn_b200.mixnet")
        self.G = engine_group(vmx, bits)
        self.n = n
        self.params = mix.SessionParams(pGroupString="test-%s" % bits)
        rs = vmx.crypto.PRGHeuristic()
        rs.setSeed(seed(label + "/setup"))
        self.x = self.G.getPRing().randomElement(rs, 100)
        y = self.G.getg().exp(self.x)
        self.pk = A.PPGroup(self.G, 2).product(self.G.getg(), y)
        self.w = mix.demoCiphertexts(self.pk, n, rs)
        self.mix = mix

    def session(self, label: str):
        rs = None
        if label is not None:
            import importlib as il
            cr = il.import_module("verificatum-vmn_b200.crypto")
            rs = cr.PRGHeuristic()
            rs.setSeed(seed(label))
        return self.mix.ShufflerSession(self.G, self.pk, self.params, rs)


def col_values(arr):
    """Engine array (possibly product) -> nested tuple of lists of ints, like the oracle's."""
    if hasattr(arr, "comps"):
        return tuple(col_values(c) for c in arr.comps)
    return [elem_value(e) for e in arr.elements()]
