"""Fixed safe-prime groups used as synthetic inputs.

The reference takes its fixed groups from `vog -gen ModPGroup -fixed <bits>`
(demo/mixnet/.conf:189-195), whose tables live inside the un-vendored VCR jar; the RFC 3526
MODP groups are used instead: p = 2^n - 2^(n-64) - 1 + 2^64 * (floor(2^(n-130) * pi) + c), p = 2q + 1,
p = 7 mod 8, so g = 2 generates the subgroup of order q.  oracle/gen_groups.py re-derives the primes
from the formula and tests/test_oracle_formats.py checks these constants against it."""

RFC3526_HEX = {
    2048: (
        "ffffffffffffffffc90fdaa22168c234c4c6628b80dc1cd129024e088a67cc74020bbea63b139b22514a08798e3404dd"
        "ef9519b3cd3a431b302b0a6df25f14374fe1356d6d51c245e485b576625e7ec6f44c42e9a637ed6b0bff5cb6f406b7ed"
        "ee386bfb5a899fa5ae9f24117c4b1fe649286651ece45b3dc2007cb8a163bf0598da48361c55d39a69163fa8fd24cf5f"
        "83655d23dca3ad961c62f356208552bb9ed529077096966d670c354e4abc9804f1746c08ca18217c32905e462e36ce3b"
        "e39e772c180e86039b2783a2ec07a28fb5c55df06f4c52c9de2bcbf6955817183995497cea956ae515d2261898fa0510"
        "15728e5a8aacaa68ffffffffffffffff"
    ),
    3072: (
        "ffffffffffffffffc90fdaa22168c234c4c6628b80dc1cd129024e088a67cc74020bbea63b139b22514a08798e3404dd"
        "ef9519b3cd3a431b302b0a6df25f14374fe1356d6d51c245e485b576625e7ec6f44c42e9a637ed6b0bff5cb6f406b7ed"
        "ee386bfb5a899fa5ae9f24117c4b1fe649286651ece45b3dc2007cb8a163bf0598da48361c55d39a69163fa8fd24cf5f"
        "83655d23dca3ad961c62f356208552bb9ed529077096966d670c354e4abc9804f1746c08ca18217c32905e462e36ce3b"
        "e39e772c180e86039b2783a2ec07a28fb5c55df06f4c52c9de2bcbf6955817183995497cea956ae515d2261898fa0510"
        "15728e5a8aaac42dad33170d04507a33a85521abdf1cba64ecfb850458dbef0a8aea71575d060c7db3970f85a6e1e4c7"
        "abf5ae8cdb0933d71e8c94e04a25619dcee3d2261ad2ee6bf12ffa06d98a0864d87602733ec86a64521f2b18177b200c"
        "bbe117577a615d6c770988c0bad946e208e24fa074e5ab3143db5bfce0fd108e4b82d120a93ad2caffffffffffffffff"
    ),
}


def rfc3526(bits: int):
    """(p, q, g) of the RFC 3526 MODP group of the given size (2048 or 3072)."""
    p = int(RFC3526_HEX[bits], 16)
    return p, (p - 1) // 2, 2


# 512-bit safe prime for fast tests (the reference's own unit test runs on ModPGroup(512):
# hvzk/TestPoSCBasicTW.java:74-85); p = 2q + 1, g = 4.
TEST512_P = int(
    "b510b7368262b646857d9bd8ea2dcaa73205ef6d7197620363efbabe6cc520a8476fc1995e4ec1e22dae8e62cc0a5aed"
    "6056701e606ee05fb0139347abba0a17"
    , 16)


def test512():
    return TEST512_P, (TEST512_P - 1) // 2, 4
