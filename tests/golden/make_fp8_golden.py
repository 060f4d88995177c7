"""Mints tests/golden/fp8_ref.npz from the UNMODIFIED reference's Triton kernels (fp8gemm.py), executed on the CPU by Triton's
interpreter.  Run in the build container only (needs /root/reference):
    TRITON_INTERPRET=1 python tests/golden/make_fp8_golden.py
Two artifacts of the interpreter's software casts are part of these vectors (tests/test_oracle_pinned.py accounts for both):
fp32 -> e4m3 drops the carry when rounding crosses a binade, and fp32 -> bf16 truncates."""
import importlib.util
import os
import sys

import numpy as np

os.environ["TRITON_INTERPRET"] = "1"
import torch  # noqa: E402
import triton  # noqa: E402

REF = "/root/reference/archive/ktransformers/ktransformers_ext/triton/fp8gemm.py"
spec = importlib.util.spec_from_file_location("fp8gemm_ref", REF)
m = importlib.util.module_from_spec(spec)
spec.loader.exec_module(m)

out = {}
for name, (T, K, N, seed) in {"a": (3, 256, 256, 0), "b": (1, 512, 384, 1), "c": (8, 1024, 200, 2)}.items():
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(T, K, generator=g) / 10).to(torch.bfloat16)
    if name == "c":
        x[2, 128:256] *= 50          # a block with a large scale next to small ones
    w = (torch.randn(N, K, generator=g) * 0.7).to(torch.float8_e4m3fn)
    ws = (torch.rand((N + 127) // 128, K // 128, generator=g) * 0.02 + 0.001).float()
    q, s = m.act_quant(x, 128)                                    # the reference's wrapper, untouched
    c = torch.empty(T, N, dtype=torch.bfloat16)
    kern = m.fp8_gemm_kernel.fn                                   # the @triton.jit body under the autotuner
    cfg = m.fp8_gemm_configs[0].kwargs                            # the reference's own first config (16 x 32 x 128)
    grid = (triton.cdiv(T, cfg["BLOCK_SIZE_M"]), triton.cdiv(N, cfg["BLOCK_SIZE_N"]))
    kern[grid](q, w, c, s, ws, T, N, K, **cfg)
    out[f"{name}_x"] = x.view(torch.int16).numpy().view(np.uint16)
    out[f"{name}_w"] = w.view(torch.uint8).numpy()
    out[f"{name}_ws"] = ws.numpy()
    out[f"{name}_q"] = q.view(torch.uint8).numpy()
    out[f"{name}_s"] = s.numpy()
    out[f"{name}_c"] = c.view(torch.int16).numpy().view(np.uint16)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fp8_ref.npz"), **out)
print("wrote fp8_ref.npz:", {k: v.shape for k, v in out.items()})
