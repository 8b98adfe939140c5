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

    __slots__ = ("value", "children", "declared")

    def __init__(self, value: Union[bytes, Sequence["ByteTree"]], declared: int | None = None):
        if isinstance(value, (bytes, bytearray, memoryview)):
            self.value = bytes(value)
            self.children = None
            self.declared = None
        else:
            self.value = None
            self.children = list(value)
            # the number of children the node's header announces: more than len(children) only for the root of a
            # tree obtained with read() whose later children are missing or malformed
            self.declared = len(self.children) if declared is None else declared

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
    """The byte tree at the start of `data`, as the reference's code sees a file or a message through a ByteTreeReader:
    what follows the tree is never read (no caller checks for the end of the file), and the children of the ROOT are
    reached one getNextChild() at a time -- a parser that takes the first k children of a node never learns that a
    later one is missing or malformed.  The root therefore keeps the longest prefix of well-formed children, with the
    count its header announces in `declared`; `first(t, k)` / `exact(t, k)` are the two ways the reference's parsers
    consume a node.  Everything below the root is parsed strictly (getNextChild skips over the whole child)."""
    if len(data) < 5:
        raise EIOError("truncated header")
    kind, n = struct.unpack_from(">BI", data, 0)
    if kind != NODE:
        return parse(data, 0)[0]
    kids, offset = [], 5
    for _ in range(n):
        try:
            c, offset = parse(data, offset, 1)
        except EIOError:
            break
        kids.append(c)
    return ByteTree(kids, declared=n)


def first(t: ByteTree, k: int) -> List[ByteTree]:
    """getNextChild() k times: the node announces at least k children and the first k are well formed."""
    if t.is_leaf() or t.declared < k or len(t.children) < k:
        raise EIOError("expected a node of at least %d children" % k)
    return t.children[:k]


def exact(t: ByteTree, k: int) -> List[ByteTree]:
    """A node that announces exactly k children (getRemaining() == k), all of them well formed."""
    if t.is_leaf() or t.declared != k or len(t.children) != k:
        raise EIOError("expected a node of %d children" % k)
    return t.children


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
