"""Generate the RFC 3526 MODP safe primes from their defining formula and verify them.

TEST INFRASTRUCTURE (oracle side).  The reference would obtain its fixed groups from
`vog -gen ModPGroup -fixed <bits>` (demo/mixnet/.conf:189-195), whose tables live in the
un-vendored VCR jar; SURVEY.md §8d therefore fixes the RFC 3526 groups as the synthetic
inputs.  RFC 3526: p = 2^n - 2^(n-64) - 1 + 2^64 * ( floor(2^(n-130) * pi) + c ).
"""
import sys


def pi_scaled(bits):
    """floor(pi * 2^bits) via Machin's formula with integer arithmetic."""
    guard = 64
    one = 1 << (bits + guard)

    def arctan_inv(x):
        total = term = one // x
        x2 = x * x
        n = 1
        sign = -1
        while term:
            term //= x2
            n += 2
            total += sign * (term // n)
            sign = -sign
        return total

    pi = 4 * (4 * arctan_inv(5) - arctan_inv(239))
    return pi >> guard


def modp(n, c):
    return (1 << n) - (1 << (n - 64)) - 1 + (1 << 64) * ((pi_scaled(n - 130)) + c)


def is_probable_prime(n, rounds=8):
    if n < 2:
        return False
    for sp in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if n % sp == 0:
            return n == sp
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in (2, 3, 5, 7, 11, 13, 17, 19)[:rounds]:
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


GROUPS = {2048: 124476, 3072: 1690314}

if __name__ == "__main__":
    for bits, c in GROUPS.items():
        p = modp(bits, c)
        q = (p - 1) // 2
        assert p.bit_length() == bits
        assert is_probable_prime(p) and is_probable_prime(q), bits
        assert pow(2, q, p) == 1  # g = 2 generates the order-q subgroup (p = 7 mod 8)
        h = "%x" % p
        print(bits, h[:32], "...", h[-32:])
