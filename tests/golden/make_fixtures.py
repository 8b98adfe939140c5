"""Regenerates tests/golden/*.json from the reference tree (run in the build container, where
/root/reference is mounted; the GPU box only sees the committed outputs).

  modpgroup_bench_config.json   the hex-marshalled ModPGroup of
                                /root/reference/demo/mixnet/benchmarks/bench_config:43 -- the only binary
                                byte-tree fixture the reference ships (real VCR output) -- and the explicit
                                (p, g) of the same group from demo/mixnet/group_descriptions:32.
  prg_ro_kat.json               SHA-256 known-answer values of the verifier specification for PRGHeuristic
                                and RandomOracle (seed 00 01 .. 1f); recalled values,
                                written down here once so that later edits of oracle/crypto.py cannot drift.
"""
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def main():
    line = open(os.path.join(REF, "demo/mixnet/benchmarks/bench_config")).read().splitlines()[42]
    hexstr = re.search(r"BENCH_PGROUP,([0-9a-f]+)", line).group(1)
    gd = open(os.path.join(REF, "demo/mixnet/group_descriptions")).read()
    m = re.search(r'ModPGroup_safeprime_15492=\$\(vog -gen ModPGroup -explic "([0-9a-f]+)" "([0-9a-f]+)"', gd)
    out = {"source": "demo/mixnet/benchmarks/bench_config:43, demo/mixnet/group_descriptions:32",
           "marshalled_hex": hexstr}
    if m:
        out["explicit_p_hex"], out["explicit_g_hex"] = m.group(1), m.group(2)
    json.dump(out, open(os.path.join(HERE, "modpgroup_bench_config.json"), "w"), indent=1)
    kat = {"seed_hex": bytes(range(32)).hex(),
           "prg_sha256_first_64_bytes_hex": "70f4003d52b6eb03da852e93256b5986b5d4883098bb7973bc5318cc66637a84"
                                            "04a6950a06d3e3308ad7d3606ef810eb124e3943404ca746a12c51c7bf776839",
           "ro_sha256": {"65": "001a8d6b6f65899ba5", "261": "1c04f57d5f5856824bca3af0ca466e283593bfc556ae2e9f4829c7ba8eb76db878"}}
    json.dump(kat, open(os.path.join(HERE, "prg_ro_kat.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
