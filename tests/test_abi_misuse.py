"""The C ABI never aborts on bad arguments ("never abort on bad input"; errors are status codes,
include/vmx.h): null handles and buffers, mismatched sizes, indices out of range, element counts whose byte size leaves
64 bits.  Every probe runs in ONE child process on the host-emulation build: a crash is a failed test, not a lost
test session."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROBES = [
    # (expression, expected status: None = any non-zero status code, or the exact value)
    ("lib.vmx_mul(a3, a5, C.byref(out))", 2),
    ("lib.vmx_mul(null, a5, C.byref(out))", 5),
    ("lib.vmx_mul(a3, a3, None)", 5),
    ("lib.vmx_exp_var(a3, r5, C.byref(out))", 2),
    ("lib.vmx_exp_var(a3, null, C.byref(out))", 5),
    ("lib.vmx_get(a3, 3, buf)", 2),
    ("lib.vmx_get(a3, 2**63, buf)", 2),
    ("lib.vmx_get(null, 0, buf)", 5),
    ("lib.vmx_slice(a5, 4, 2, C.byref(out))", 2),
    ("lib.vmx_slice(a5, 0, 6, C.byref(out))", 2),
    ("lib.vmx_extract(a5, None, C.byref(out))", 5),
    ("lib.vmx_extract(null, b'\\x01' * 5, C.byref(out))", 5),
    ("lib.vmx_equals(a3, null, C.byref(eq))", 5),
    ("lib.vmx_prod(null, buf)", 5),
    ("lib.vmx_prod(a3, None)", 5),
    ("lib.vmx_expprod(None, 1, r3, buf)", 5),
    ("lib.vmx_expprod((C.c_void_p * 1)(a3), 1, r5, buf)", 2),
    ("lib.vmx_expprod((C.c_void_p * 1)(null), 1, r3, buf)", 5),
    ("lib.vmx_expprod((C.c_void_p * 1)(a3), 0, r3, buf)", 5),
    ("lib.vmx_expprod_cols((C.c_void_p * 2)(a3, a5), 2, (C.c_int64 * 2)(1, 2), C.byref(out))", 2),
    ("lib.vmx_expprod_cols(None, 2, (C.c_int64 * 2)(1, 2), C.byref(out))", 5),
    ("lib.vmx_exp_scalar_var(a3, one, a5, r3, C.byref(out))", 2),
    ("lib.vmx_exp_scalar_var(a3, None, a3, r3, C.byref(out))", 5),
    ("lib.vmx_shift_push(a3, None, C.byref(out))", 5),
    ("lib.vmx_permute(a3, None, C.byref(out))", 5),
    ("lib.vmx_permute(a3, (C.c_uint32 * 3)(0, 7, 1), C.byref(out))", 5),
    ("lib.vmx_garr_from_bytes(ctx, 3, None, 1, C.byref(out))", 5),
    ("lib.vmx_garr_from_bytes(None, 3, one * 3, 1, C.byref(out))", 5),
    ("lib.vmx_garr_from_leaves(ctx, 2**40, one, 1, C.byref(out))", None),
    ("lib.vmx_garr_fill(ctx, 2**62, one, C.byref(out))", 2),
    ("lib.vmx_rarr_fill(ctx, 2**62, one, C.byref(out))", 2),
    ("lib.vmx_garr_prg_sha256(ctx, b'x' * 32, 32, 0, 2**61, 80, 600, C.byref(out))", None),
    ("lib.vmx_exp_fixed(ctx, None, r3, C.byref(out))", 5),
    ("lib.vmx_exp_fixed(ctx, one, null, C.byref(out))", 5),
    ("lib.vmx_elem_exp(ctx, None, one, buf)", 5),
    ("lib.vmx_elem_inv(ctx, (0).to_bytes(eb, 'big'), buf)", 1),
    ("lib.vmx_rmuladd(r3, None, r3, C.byref(out))", 5),
    ("lib.vmx_rmuladd(r3, one, r5, C.byref(out))", 2),
    ("lib.vmx_rreclin(r3, r5, C.byref(out), buf)", 2),
    ("lib.vmx_rinner(r3, r5, buf)", 2),
    ("lib.vmx_rget(r3, 9, buf)", 2),
    ("lib.vmx_rslice(r5, 3, 1, C.byref(out))", 2),
    ("lib.vmx_ctx_set_tuning(ctx, None, 1)", 5),
    ("lib.vmx_ctx_set_tuning(ctx, b'nonsense', 1)", 5),
    ("lib.vmx_ctx_set_fixed_window(ctx, 99)", 5),
    ("lib.vmx_fixed_precompute(ctx, None, 10)", 5),
    ("lib.vmx_garr_pack_rows(a3, None, 2, None)", 5),
    ("lib.vmx_rarr_prg_sha256(ctx, None, 32, 0, 4, 100, C.byref(out))", 5),
    ("lib.vmx_rarr_prg_sha256(ctx, b'x' * 32, 32, 0, 4, 0, C.byref(out))", 5),
    ("lib.vmx_garr_prg_sha256(ctx, b'x' * 32, 32, 0, 4, 0, 0, C.byref(out))", 5),
    ("lib.vmx_garr_free(null)", "void"),
    ("lib.vmx_rarr_free(null)", "void"),
    ("lib.vmx_ctx_destroy(None)", "void"),
    ("lib.vmx_host_free(None)", "void"),
]

CHILD = r'''
import ctypes as C, importlib, os, sys
sys.path.insert(0, %(root)r)
os.environ["VMX_LIBRARY_PATH"] = %(lib)r
vmx = importlib.import_module("verificatum-vmn_b200")
lib = vmx._native.load()
g = importlib.import_module("verificatum-vmn_b200.groups")
G = vmx.arithm.ModPGroup(*g.test512())
ctx = G.ctx
eb, rb = lib.vmx_ctx_elem_bytes(ctx), lib.vmx_ctx_ring_bytes(ctx)
one = (1).to_bytes(eb, "big")
def garr(n):
    h = C.c_void_p(); assert lib.vmx_garr_fill(ctx, n, one, C.byref(h)) == 0; return h
def rarr(n):
    h = C.c_void_p(); assert lib.vmx_rarr_fill(ctx, n, (3).to_bytes(rb, "big"), C.byref(h)) == 0; return h
out, buf, eq, null = C.c_void_p(), C.create_string_buffer(4 * eb), C.c_int(), C.c_void_p()
a3, a5, r3, r5 = garr(3), garr(5), rarr(3), rarr(5)
for i, (expr, want) in enumerate(%(probes)r):
    print("PROBE", i, flush=True)
    rc = eval(expr)
    print("RC", i, rc, flush=True)
# and the context is still usable afterwards
h = C.c_void_p()
assert lib.vmx_mul(a3, a3, C.byref(h)) == 0 and lib.vmx_garr_size(h) == 3
print("ALIVE")
'''


def test_bad_arguments_are_status_codes(emul_lib):
    code = CHILD % {"root": ROOT, "lib": emul_lib, "probes": PROBES}
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = r.stdout.splitlines()
    last = [ln for ln in lines if ln.startswith("PROBE")]
    assert r.returncode == 0 and lines and lines[-1] == "ALIVE", \
        "the library died in %s:\n%s" % (PROBES[int(last[-1].split()[1])][0] if last else "set-up", r.stderr[-1500:])
    rcs = {int(ln.split()[1]): ln.split()[2] for ln in lines if ln.startswith("RC")}
    for i, (expr, want) in enumerate(PROBES):
        got = rcs[i]
        if want == "void":
            assert got == "None", (expr, got)
        elif want is None:
            assert got not in ("0", "None"), (expr, got)
        else:
            assert got == str(want), (expr, got)
