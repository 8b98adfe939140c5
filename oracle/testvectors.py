"""ORACLE (test infrastructure).  The text format of the reference's `vmnv -t` test vectors
(mixnet/MixNetElGamalVerifyFiatShamir.java:334-388, mixnet/MixNetElGamalVerifyFiatShamirTool.java:292-309):

    <blank line>
    TEST VECTOR
    <name> - <description>
    <value>

with the vectors of one party's proof between "###################### BEGIN PARTY l ######################" and
"####################### END PARTY l #######################" lines.  `parse` reads such output (interleaved with
whatever else `vmnv -v` prints) into {(name, party or None): [values in order of appearance]}; `render` writes the
(name, party, value) lists the verifiers of this repository record in the same format, so that an operator can diff
the two, and so that the loader of tests/test_reference_pin.py is tested before a real dump arrives."""
from __future__ import annotations

import re
from typing import Dict, List, Optional, Sequence, Tuple

# names whose values are single lines of digits / hexadecimal digits / short strings
SCALAR_NAMES = ("par.version", "par.sid", "par.k", "par.lambda", "par.n_e", "par.n_r", "par.n_v", "par.omega", "par.N_0",
                "der.rho", "PoS.s", "PoS.v", "PoSC.s", "PoSC.v", "CCPoS.s", "CCPoS.v", "Dec.s", "Dec.v")
_BEGIN = re.compile(r"^#+ BEGIN PARTY (\d+) #+$")
_END = re.compile(r"^#+ END PARTY (\d+) #+$")
_HEAD = re.compile(r"^\s*([A-Za-z]+(?:\.[A-Za-z_0-9]+)?) - ")


def parse(text: str) -> Dict[Tuple[str, Optional[int]], List[str]]:
    out: Dict[Tuple[str, Optional[int]], List[str]] = {}
    party: Optional[int] = None
    lines = text.splitlines()
    i = 0
    while i < len(lines):
        ln = lines[i].strip()
        m = _BEGIN.match(ln)
        if m:
            party = int(m.group(1))
        elif _END.match(ln):
            party = None
        elif ln == "TEST VECTOR" and i + 1 < len(lines):
            h = _HEAD.match(lines[i + 1])
            if h:
                j = i + 2
                value = []
                while j < len(lines) and lines[j].strip() and lines[j].strip() != "TEST VECTOR" and not lines[j].startswith("#"):
                    value.append(lines[j].strip())
                    j += 1
                if h.group(1) in SCALAR_NAMES:   # (one line; whatever follows without a blank line is other output)
                    value = value[:1]
                out.setdefault((h.group(1), party), []).append("\n".join(value))
                i = j
                continue
        i += 1
    return out


def render(vectors: Sequence[Tuple[str, Optional[int], str]]) -> str:
    parts: List[str] = []
    party: Optional[int] = None
    for name, l, value in vectors:
        if l != party:
            if party is not None:
                parts.append("\n####################### END PARTY %d #######################" % party)
            if l is not None:
                parts.append("\n###################### BEGIN PARTY %d ######################" % l)
            party = l
        parts.append("\nTEST VECTOR\n%s - %s\n%s" % (name, "(description)", value))
    if party is not None:
        parts.append("\n####################### END PARTY %d #######################" % party)
    return "\n".join(parts) + "\n"


def same_value(name: str, ours: str, theirs: str) -> bool:
    """Seeds and the prefix are hexadecimal strings; challenges are integers whose radix in LargeInteger.toString is
    [VCR-mem] (the tool's own description says hexadecimal, the call site prints toString()): both are accepted."""
    if name.endswith(".v"):
        want = int(ours)
        for radix in (10, 16):
            try:
                if int(theirs, radix) == want:
                    return True
            except ValueError:
                pass
        return False
    if name.endswith(".s") or name == "der.rho":
        return ours.lower() == theirs.lower()
    return ours == theirs
