"""Per-op timing of the non-MoE parts of a DeepSeek-V3 decode layer (CUDA events, L2 flushed between repetitions)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ktransformers_b200 import native  # noqa: E402
from ktransformers_b200.util.synth import synth_blocks  # noqa: E402

lib = native.lib()
Q4_K, Q6_K, BF16 = 12, 14, 30
S = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
bf = torch.bfloat16


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)) * 1e3


def linear(inf, outf, t):
    w = synth_blocks(t, outf * inf, "cuda", inf + outf)
    h = C.c_void_p()
    native.check(lib.ktb200_linear_create(inf, outf, w.data_ptr(), t, BF16, 8, 0, C.byref(h)))
    native.check(lib.ktb200_linear_load_weights(h, S()))
    x = torch.randn(1, inf, device="cuda").to(bf); y = torch.zeros(1, outf, dtype=bf, device="cuda")
    us = timeit(lambda: native.check(lib.ktb200_linear_forward(h, 1, x.data_ptr(), y.data_ptr(), None, None, S())))
    nbytes = w.numel()
    print(f"linear {inf:6d} -> {outf:6d} type {t}: {us:7.1f} us  {nbytes / us / 1e3:7.0f} GB/s", flush=True)
    return w


keep = [linear(7168, 1536, Q4_K), linear(7168, 576, Q4_K), linear(1536, 24576, Q4_K), linear(16384, 7168, Q4_K), linear(7168, 129280, Q6_K)]
# dense MLP
H, DI = 7168, 18432
gw, uw, dw = synth_blocks(Q4_K, DI * H, "cuda", 1), synth_blocks(Q4_K, DI * H, "cuda", 2), synth_blocks(Q6_K, H * DI, "cuda", 3)
mh = C.c_void_p()
native.check(lib.ktb200_mlp_create(H, DI, gw.data_ptr(), uw.data_ptr(), dw.data_ptr(), Q4_K, Q4_K, Q6_K, BF16, 8, 0, C.byref(mh)))
native.check(lib.ktb200_mlp_load_weights(mh, S()))
x = torch.randn(1, H, device="cuda").to(bf); y = torch.zeros(1, H, dtype=bf, device="cuda")
us = timeit(lambda: native.check(lib.ktb200_mlp_forward(mh, 1, x.data_ptr(), y.data_ptr(), 0, None, S())))
print(f"dense MLP 7168 -> 18432 -> 7168: {us:7.1f} us  {(gw.numel() + uw.numel() + dw.numel()) / us / 1e3:7.0f} GB/s", flush=True)
# bmm
q = torch.randn(128, 1, 192, device="cuda").to(bf); wuk = torch.randn(128, 128, 512, device="cuda").to(bf); o = torch.zeros(128, 1, 512, dtype=bf, device="cuda")
print(f"bmm W_UK absorb: {timeit(lambda: torch.bmm(q[:, :, :128], wuk, out=o)):7.1f} us  ({wuk.numel() * 2 / 1e6:.1f} MB)")
lat = torch.randn(128, 1, 512, device="cuda").to(bf); o2 = torch.zeros(128, 1, 128, dtype=bf, device="cuda")
print(f"bmm W_UV        : {timeit(lambda: torch.bmm(lat, wuk.transpose(1, 2), out=o2)):7.1f} us")
# norms / prep
w = torch.ones(H, dtype=bf, device="cuda"); r = torch.randn(1, H, device="cuda").to(bf); d = torch.randn(1, H, device="cuda").to(bf); out = torch.zeros_like(r)
print(f"add_rmsnorm 7168: {timeit(lambda: native.check(lib.ktb200_add_rmsnorm(r.data_ptr(), d.data_ptr(), w.data_ptr(), 1e-6, out.data_ptr(), 1, H, S()))):7.1f} us")
# MLA
for ctx in (1024, 4096, 32768):
    page = 64
    npg = ctx // page + 1
    kv = torch.randn((npg, page, 576), device="cuda", dtype=bf)
    pt = torch.arange(npg, dtype=torch.int32, device="cuda")[None].contiguous()
    kl = torch.tensor([ctx], dtype=torch.int32, device="cuda")
    qn = torch.randn(1, 128, 512, device="cuda").to(bf); qp = torch.randn(1, 128, 64, device="cuda").to(bf)
    oo = torch.zeros(1, 128, 512, dtype=bf, device="cuda")
    wsb = lib.ktb200_mla_workspace_bytes(1, 128, 0); ws = torch.zeros(wsb, dtype=torch.uint8, device="cuda")
    p = native.MlaParams(1, 128, page, npg, 0, 0.072, qn.data_ptr(), qp.data_ptr(), kv.data_ptr(), pt.data_ptr(), kl.data_ptr(), oo.data_ptr(), None, ws.data_ptr(), wsb, npg * page)
    print(f"mla decode ctx {ctx}: {timeit(lambda: native.check(lib.ktb200_mla_decode(C.byref(p), S()))):7.1f} us")
