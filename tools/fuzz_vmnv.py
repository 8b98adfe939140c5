"""Differential fuzzing of the native universal verifier (libvmnv.so) against the Python mirror of the reference's
vmnv, on the host-emulation build of the engine (tests/parity_bodies.py: native_vmnv_fuzz).

    python tools/fuzz_vmnv.py [--rounds 1000] [--spec 512|P-256] [--width 1] [--seed label] [--asan]

--asan re-executes under the AddressSanitizer / UBSan builds of both libraries (built like tools/asan_emul.sh).
Test infrastructure: the emulation build runs every kernel body on the CPU; nothing here is a product path."""
import argparse
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=1000)
    ap.add_argument("--spec", default="512")
    ap.add_argument("--n", type=int, default=3)
    ap.add_argument("--k", type=int, default=3)
    ap.add_argument("--threshold", type=int, default=2)
    ap.add_argument("--width", type=int, default=1)
    ap.add_argument("--seed", default="fuzz")
    ap.add_argument("--mode", default="mixing", choices=["mixing", "shuffling", "decryption"])
    ap.add_argument("--maxciph", type=int, default=0, help="pre-compute for this many ciphertexts first (0: no pre-computation)")
    ap.add_argument("--asan", action="store_true")
    ap.add_argument("--oracle", action="store_true", help="the oracle's vmnv (oracle/protocols.py) as a third party")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()

    if args.asan and not os.environ.get("VMNV_FUZZ_CHILD"):
        out = os.path.join(ROOT, "build", "asan")
        os.makedirs(out, exist_ok=True)
        csrc = os.path.join(ROOT, "verificatum-vmn_b200", "csrc")
        flags = ["g++", "-std=c++17", "-O1", "-g", "-fno-omit-frame-pointer", "-fsanitize=address,undefined",
                 "-fno-sanitize-recover=undefined", "-fPIC", "-shared"]
        subprocess.run(flags + ["-DVMX_HOST_EMUL", "-x", "c++", "-o", os.path.join(out, "libvmx_emul_asan.so"), "vmx.cu"],
                       check=True, cwd=csrc)
        subprocess.run(flags + ["-o", os.path.join(out, "libvmnv_asan.so"), "vmnv_native.cpp", "-ldl", "-lcrypto",
                                "-lpthread"], check=True, cwd=csrc)
        env = dict(os.environ, VMNV_FUZZ_CHILD="1", VMX_EMUL_LIBRARY=os.path.join(out, "libvmx_emul_asan.so"),
                   VMNV_LIBRARY_PATH=os.path.join(out, "libvmnv_asan.so"),
                   ASAN_OPTIONS="detect_leaks=0:abort_on_error=0:halt_on_error=1",
                   UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1")
        pre = [subprocess.check_output(["gcc", "-print-file-name=" + l]).decode().strip() for l in ("libasan.so", "libubsan.so")]
        env["LD_PRELOAD"] = " ".join(pre)
        sys.exit(subprocess.call([sys.executable] + sys.argv, env=env))

    os.environ.setdefault("VMX_BUFFER_MIN", "256")
    os.environ.setdefault("VMX_MEXP_SORT_MIN", "1")
    import importlib
    import __graft_entry__ as ge
    os.environ["VMX_LIBRARY_PATH"] = os.environ.get("VMX_EMUL_LIBRARY") or ge.build_host_emul()
    if not os.environ.get("VMNV_LIBRARY_PATH"):
        ge.build_vmnv()
    vmx = importlib.import_module("verificatum-vmn_b200")
    from tests import parity_bodies as pb
    spec = int(args.spec) if args.spec.isdigit() else args.spec
    t0 = time.time()
    tally = pb.native_vmnv_fuzz(vmx, spec, args.n, args.rounds, seed_label=args.seed, k=args.k, threshold=args.threshold,
                                width=args.width, log=(lambda m: print(m, flush=True)) if args.verbose else None,
                                mode=args.mode, maxciph=args.maxciph or None, oracle=args.oracle)
    print("spec=%s n=%d k=%d threshold=%d width=%d mode=%s maxciph=%s seed=%r rounds=%d: native == mirror%s on every round; "
          "outcomes %s; %.0f s%s"
          % (args.spec, args.n, args.k, args.threshold, args.width, args.mode, args.maxciph or None, args.seed, args.rounds,
             " == oracle" if args.oracle else "", dict(sorted(tally.items())),
             time.time() - t0, " (ASan + UBSan builds)" if os.environ.get("VMNV_FUZZ_CHILD") else ""))


if __name__ == "__main__":
    main()
