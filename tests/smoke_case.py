"""__graft_entry__.smoke(): one small invocation of the hot path on cuda:0 (re-encrypt + prove +
verify of 64 ciphertexts over the 3072-bit group) checked byte for byte against the oracle."""
import importlib


def run():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("smoke() needs cuda:0; the engine has no CPU path")
    vmx = importlib.import_module("verificatum-vmn_b200")
    from tests import parity_bodies as pb
    pb.transcript_parity(vmx, 3072, 64)
    pb.accept_reject_properties(vmx, 3072, 3000)
    print("smoke ok: 3072-bit shuffle of 64 ciphertexts is byte-identical to the oracle; N=3000 proof verifies")
