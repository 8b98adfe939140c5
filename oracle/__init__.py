"""ORACLE -- test infrastructure only.

A CPU restatement of the reference's algorithms for the hot path (byte trees, PRG / random
oracle, group and ring array semantics, PoSBasicTW / PoSCBasicTW / CCPoSBasicW, the
re-encryption shuffle, decryption-factor proofs), each function citing the file:line of
/root/reference it follows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg may import it -- always as the checker or as the reported CPU baseline, never as the thing
shipped.  Nothing under verificatum-vmn_b200/ imports this package.

Parity status: UNPINNED against an actual Java/GMP run (no JVM in the build image, and the
reference ships no golden vectors).  Pinned by the fixtures in tests/golden/
(in-tree marshalled ModPGroup, PRG / RO known answers) and by exact integer arithmetic.
"""
