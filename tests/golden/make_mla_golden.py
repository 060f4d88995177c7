"""Golden vectors for absorbed-MLA decode from the reference's own torch restatement `attention_ref_torch`
(archive/ktransformers/operators/flashinfer_wrapper.py:30-76 — the function the reference's MLAWrapper self-test
compares flashinfer against, :343-380 there).  The function source is executed as is (it is pure torch); the absorbed
form is fed to it the way the reference does: q = [q_nope' | q_pe] (576 wide), k = [ckv | k_pe] broadcast over heads,
v = ckv broadcast over heads.

    python tests/golden/make_mla_golden.py      (build container only: needs /root/reference)

Output: tests/golden/mla_ref.npz — per case: q_nope, q_pe, kv rows (bf16 bit patterns, uint16), sm_scale, the reference
output (bf16 bits) and its base-2 LSE (fp32).
"""
import math
import os

import numpy as np
import torch

OUT = os.path.dirname(os.path.abspath(__file__))
src = open("/root/reference/archive/ktransformers/operators/flashinfer_wrapper.py").read()
fn_src = src[src.index("def attention_ref_torch("):src.index("class MLAWrapper")]
ns = {"torch": torch, "math": math}
exec(fn_src, ns)
attention_ref_torch = ns["attention_ref_torch"]


def bits(t):
    return t.view(torch.int16).numpy().view(np.uint16)


out = {}
torch.manual_seed(20260923)
for name, B, H, L, scale in (("a", 1, 16, 100, (128 + 64) ** -0.5), ("b", 2, 128, 65, 0.1352337788608801), ("c", 1, 128, 33, 1.0)):
    q_nope = (torch.randn(B, H, 512) * 0.5).to(torch.bfloat16)
    q_pe = (torch.randn(B, H, 64) * 0.5).to(torch.bfloat16)
    kv = torch.randn(B, L, 576).to(torch.bfloat16)
    if name == "c":   # scores that grow along the sequence: exercises the running-maximum update
        kv = (kv.float() * torch.linspace(0.05, 1.0, L)[None, :, None]).to(torch.bfloat16)
    q = torch.cat([q_nope, q_pe], -1)                                   # [B, H, 576]  (qo_len = 1)
    k = kv[:, :, None, :].expand(B, L, H, 576).reshape(B * L, H, 576)
    v = kv[:, :, None, :512].expand(B, L, H, 512).reshape(B * L, H, 512)
    o_ref, lse2 = attention_ref_torch(B, q, k, v, False, scale)
    out.update({f"{name}_q_nope": bits(q_nope), f"{name}_q_pe": bits(q_pe), f"{name}_kv": bits(kv), f"{name}_scale": np.float32(scale),
                f"{name}_out": bits(o_ref.contiguous()), f"{name}_lse2": lse2.reshape(B, H).float().numpy()})
    print(name, o_ref.shape, lse2.shape, float(o_ref.float().abs().max()))
np.savez_compressed(os.path.join(OUT, "mla_ref.npz"), **out)
