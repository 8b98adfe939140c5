"""ORACLE (test infrastructure only -- never imported by the product path).

Byte-tree codec of the Verificatum wire/disk format.  The codec itself lives in the
un-vendored dependency verificatum-vcr 3.1.0 (`com.verificatum.eio.ByteTree*`,
pinned by /root/reference configure.ac:35); the rules below restate the public verifier
specification (SURVEY.md §8c [VCR-mem]) and are validated against the one binary fixture the
reference tree ships: the hex-marshalled ModPGroup in
demo/mixnet/benchmarks/bench_config:43 (see tests/test_oracle_formats.py).

    node  = 0x00 || be32(#children) || children
    leaf  = 0x01 || be32(#bytes)    || bytes

Call sites in the reference that fix how trees are composed:
hvzk/PoSBasicTW.java:694-699 (commitment), :880-886 (reply), hvzk/PoSTW.java:118-124
(challenge data), elgamal/ProtocolElGamalInterfaceRaw.java:53-56 (column-wise product arrays).
"""
from __future__ import annotations

import struct
from typing import List, Sequence, Tuple, Union

NODE = 0
LEAF = 1


class ByteTree:
    """Immutable byte tree: either a leaf (bytes) or a node (list of ByteTree)."""

    __slots__ = ("value", "children")

    def __init__(self, value: Union[bytes, Sequence["ByteTree"]]):
        if isinstance(value, (bytes, bytearray, memoryview)):
            self.value = bytes(value)
            self.children = None
        else:
            self.value = None
            self.children = list(value)

    def is_leaf(self) -> bool:
        return self.children is None

    def to_bytes(self) -> bytes:
        out: List[bytes] = []
        self._emit(out)
        return b"".join(out)

    def _emit(self, out: List[bytes]) -> None:
        if self.children is None:
            out.append(struct.pack(">BI", LEAF, len(self.value)))
            out.append(self.value)
        else:
            out.append(struct.pack(">BI", NODE, len(self.children)))
            for c in self.children:
                c._emit(out)

    def update(self, digest) -> None:
        """Feed the serialisation into a hashlib object (ByteTreeBasic.update)."""
        if self.children is None:
            digest.update(struct.pack(">BI", LEAF, len(self.value)))
            digest.update(self.value)
        else:
            digest.update(struct.pack(">BI", NODE, len(self.children)))
            for c in self.children:
                c.update(digest)

    def __eq__(self, other) -> bool:
        return isinstance(other, ByteTree) and self.to_bytes() == other.to_bytes()

    def __repr__(self) -> str:
        if self.children is None:
            return "Leaf(%d)" % len(self.value)
        return "Node(%s)" % ", ".join(repr(c) for c in self.children)


def leaf(data: bytes) -> ByteTree:
    return ByteTree(bytes(data))


def node(*children: ByteTree) -> ByteTree:
    if len(children) == 1 and not isinstance(children[0], ByteTree):
        return ByteTree(list(children[0]))
    return ByteTree(list(children))


class EIOError(ValueError):
    """Malformed byte tree (EIOException in the reference)."""


MAX_DEPTH = 64   # the trees of a proof directory are at most ~6 deep; deeper input is malformed


def parse(data: bytes, offset: int = 0, depth: int = 0) -> Tuple[ByteTree, int]:
    if depth > MAX_DEPTH:
        raise EIOError("byte tree nested too deep")
    if offset + 5 > len(data):
        raise EIOError("truncated header")
    kind, n = struct.unpack_from(">BI", data, offset)
    offset += 5
    if kind == LEAF:
        if offset + n > len(data):
            raise EIOError("truncated leaf")
        return ByteTree(data[offset:offset + n]), offset + n
    if kind == NODE:
        kids = []
        for _ in range(n):
            c, offset = parse(data, offset, depth + 1)
            kids.append(c)
        return ByteTree(kids), offset
    raise EIOError("bad tag %d" % kind)


def from_bytes(data: bytes) -> ByteTree:
    t, end = parse(data, 0)
    if end != len(data):
        raise EIOError("trailing bytes")
    return t


def read(data: bytes) -> ByteTree:
    """The byte tree at the start of `data`, as a ByteTreeReaderF over a file sees it: what follows the tree is never
    read (no caller in the reference checks for the end of the file)."""
    return parse(data, 0)[0]


# ---------------------------------------------------------------- integers
def int_byte_length(x: int) -> int:
    """Length of BigInteger.toByteArray() for x >= 0 (two's complement, minimal)."""
    return x.bit_length() // 8 + 1


def int_to_bytes(x: int, length: int | None = None) -> bytes:
    """Big-endian two's complement (LargeInteger.toByteArray); fixed width if given."""
    if length is None:
        length = (x.bit_length() if x >= 0 else (~x).bit_length()) // 8 + 1
    return x.to_bytes(length, "big", signed=True)


def bytes_to_int(b: bytes) -> int:
    return int.from_bytes(b, "big", signed=True)


def int_leaf(x: int, length: int | None = None) -> ByteTree:
    return leaf(int_to_bytes(x, length))


def int32_leaf(x: int) -> ByteTree:
    """ByteTree.intToByteTree: a 4-byte leaf."""
    return leaf(struct.pack(">i", x))


def string_leaf(s: str) -> ByteTree:
    return leaf(s.encode("ascii"))


def bool_array_leaf(flags: Sequence[bool]) -> ByteTree:
    return leaf(bytes(1 if f else 0 for f in flags))
