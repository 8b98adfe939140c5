"""vmx -- B200-native engine for the batched group-exponentiation hot path of the Verificatum
mix-net (re-encryption, proofs of shuffle, decryption-factor proofs).

The directory name carries a hyphen (it is the repository's product name), so import it with

    import importlib; vmx = importlib.import_module("verificatum-vmn_b200")

Sub-modules: `arithm` (GPU-backed mirror of com.verificatum.arithm arrays), `eio` (byte trees),
`crypto` (PRG / random oracle), `hvzk` (PoSBasicTW, PoSCBasicTW, CCPoSBasicW, ChallengerRO),
`mixnet` (re-encryption shuffle), `elgamal` (decryption-factor proofs), `_native` (C ABI).
"""
from . import _native  # noqa: F401
from . import eio, crypto, arithm  # noqa: F401

__all__ = ["_native", "eio", "crypto", "arithm"]
