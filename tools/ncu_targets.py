"""Launches the kernels tools/gpu_profile_r02.sh captures with ncu: MLA decode (ctx 32768) and the o_proj-shaped dense linear."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ktransformers_b200 import native  # noqa: E402
from ktransformers_b200.util.synth import synth_blocks  # noqa: E402

lib = native.lib()
S = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
bf = torch.bfloat16
ctx, page = 32768, 64
npg = ctx // page
kv = torch.randn((npg, page, 576), device="cuda", dtype=bf)
pt = torch.arange(npg, dtype=torch.int32, device="cuda")[None].contiguous()
kl = torch.tensor([ctx], dtype=torch.int32, device="cuda")
qn = (torch.randn(1, 128, 512, device="cuda") * 0.5).to(bf); qp = (torch.randn(1, 128, 64, device="cuda") * 0.5).to(bf)
out = torch.zeros(1, 128, 512, dtype=bf, device="cuda")
wsb = lib.ktb200_mla_workspace_bytes(1, 128, 0); ws = torch.zeros(wsb, dtype=torch.uint8, device="cuda")
p = native.MlaParams(1, 128, page, npg, 0, 0.072, qn.data_ptr(), qp.data_ptr(), kv.data_ptr(), pt.data_ptr(), kl.data_ptr(), out.data_ptr(), None, ws.data_ptr(), wsb, npg * page)
w = synth_blocks(12, 7168 * 16384, "cuda", 1)
h = C.c_void_p()
native.check(lib.ktb200_linear_create(16384, 7168, w.data_ptr(), 12, 30, 8, 0, C.byref(h)))
native.check(lib.ktb200_linear_load_weights(h, S()))
x = torch.randn(1, 16384, device="cuda").to(bf); y = torch.zeros(1, 7168, dtype=bf, device="cuda")
for _ in range(4):
    native.check(lib.ktb200_mla_decode(C.byref(p), S()))
    native.check(lib.ktb200_linear_forward(h, 1, x.data_ptr(), y.data_ptr(), None, None, S()))
torch.cuda.synchronize()
print("done")
