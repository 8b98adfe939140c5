"""The GMP back end of the oracle (oracle/cpu_ref.c through oracle/accel.py) against the Python-integer
oracle (oracle/arithm.py) and Python's own pow(): fixed-base tables (gmpmee fpowm), per-element powm,
simultaneous exponentiation (gmpmee spowm), element-wise product, Legendre/Jacobi symbol -- at the three
ModPGroup sizes of the parity tests, with the edge exponents (0, 1, q - 1, short, top bit set) and the edge
bases (1, p - 1, g).  This pins the checker the GPU production-shape tests (tests/test_gpu_parity.py,
production_kernels) and the CPU baseline of bench.py rely on."""
import random

import pytest

from oracle import accel
from oracle import arithm as oar
from tests.cases import group_params


def _instance(bits, n, seed):
    p, q, g = group_params(bits)
    rnd = random.Random(seed)
    bases = [pow(g, rnd.randrange(1, q), p) for _ in range(n)]
    bases[:3] = [1, p - 1, g]
    exps = [rnd.randrange(q) for _ in range(n)]
    exps[:6] = [0, 1, q - 1, 2, (1 << (q.bit_length() - 1)) % q, rnd.randrange(1 << 64)]
    return p, q, g, rnd, bases, exps


@pytest.mark.parametrize("bits,n", [(512, 200), (2048, 24), (3072, 10)])
def test_powm_and_mul_arrays(bits, n):
    p, q, g, rnd, bases, exps = _instance(bits, n, bits)
    acc = accel.Accel(p, threads=3, q=q)
    assert acc.exp_var(bases, exps) == [pow(b, e, p) for b, e in zip(bases, exps)]
    for s in (0, 1, q >> 1, rnd.randrange(1 << 256), rnd.randrange(q)):
        assert acc.exp_var(bases, s) == [pow(b, s, p) for b in bases], s
    # small negative integers mod q (Lagrange coefficients) take an inversion and a short power: the same value
    # as the plain power on members of the order-q subgroup, which is all the protocol oracle applies it to
    # (arrays are validated on import); p - 1 is left out here for that reason
    members = [b for b in bases if pow(b, q, p) == 1]
    assert len(members) >= n - 1
    for s in (q - 1, q - 5, q - (1 << 40)):
        assert acc.exp_var(members, s) == [pow(b, s, p) for b in members], s
    other = bases[::-1]
    assert acc.mul(bases, other) == [a * b % p for a, b in zip(bases, other)]


@pytest.mark.parametrize("bits,n", [(512, 200), (2048, 20), (3072, 8)])
@pytest.mark.parametrize("window", [4, 8])
def test_fixed_base_tables(bits, n, window):
    p, q, g, rnd, bases, exps = _instance(bits, n, bits + window)
    acc = accel.Accel(p, threads=2, fixed_window=window, q=q)
    for base in (g, bases[7], 1, p - 1):
        assert acc.exp_fixed(base, exps, q.bit_length()) == [pow(base, e, p) for e in exps]
    short = [rnd.randrange(1 << 100) for _ in range(n)]
    assert acc.exp_fixed(g, short, q.bit_length()) == [pow(g, e, p) for e in short]


@pytest.mark.parametrize("bits,n", [(512, 203), (2048, 21), (3072, 9)])
@pytest.mark.parametrize("k", [1, 5, 7])
def test_simultaneous_exponentiation(bits, n, k):
    p, q, g, rnd, bases, exps = _instance(bits, n, bits + k)
    G = oar.ModPGroup(p, q, g)
    e256 = [rnd.randrange(1 << 256) for _ in range(n)]
    want, want256 = oar.g_exp_prod(G, bases, exps), oar.g_exp_prod(G, bases, e256)
    for threads in (1, 4):
        acc = accel.Accel(p, threads=threads, spowm_width=k, q=q)
        assert acc.expprod(bases, exps) == want
        assert acc.expprod(bases, e256) == want256
        assert acc.expprod(bases[:1], exps[:1]) == pow(bases[0], exps[0], p)
        assert acc.expprod(bases, [0] * n) == 1


@pytest.mark.parametrize("bits", [512, 2048, 3072])
def test_jacobi_is_eulers_criterion(bits):
    p, q, g, rnd, bases, exps = _instance(bits, 10, bits)
    cand = [rnd.randrange(1, p) for _ in range(60)] + [1, 2, 3, 4, p - 1, p - 2, (p - 1) // 2, (p + 1) // 2]
    acc = accel.Accel(p, q=q)
    assert acc.members(cand) == [pow(v, q, p) == 1 for v in cand]


def test_installed_back_end_is_the_python_oracle():
    """accel.install() reroutes oracle.arithm's array functions; same values as before the rerouting, also on
    product (tuple) arrays and on the empty array."""
    p, q, g, rnd, bases, exps = _instance(512, 37, 99)
    G = oar.ModPGroup(p, q, g)
    prod_arr = (bases, bases[::-1])
    want = (oar.g_exp(G, bases, exps), oar.g_exp(G, g, exps), oar.g_exp(G, bases, exps[9]),
            oar.g_exp_prod(G, prod_arr, exps), oar.g_mul(G, prod_arr, prod_arr), oar.g_exp(G, prod_arr, exps))
    undo = accel.install(G, threads=2)
    try:
        got = (oar.g_exp(G, bases, exps), oar.g_exp(G, g, exps), oar.g_exp(G, bases, exps[9]),
               oar.g_exp_prod(G, prod_arr, exps), oar.g_mul(G, prod_arr, prod_arr), oar.g_exp(G, prod_arr, exps))
        assert oar.g_exp(G, [], []) == [] and oar.g_exp_prod(G, [], []) == 1
    finally:
        undo()
    assert got == want
