"""In-graph per-op cost of the attention layer's kernels: each op repeated 61x with distinct weights in one CUDA graph."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ktransformers_b200 import native  # noqa: E402
from ktransformers_b200.util.synth import synth_blocks  # noqa: E402

lib = native.lib()
Q4_K, BF16 = 12, 30
S = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
bf = torch.bfloat16
N = 61


def graph_time(fn):
    fn(); torch.cuda.synchronize()
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5 / N * 1e3


def linear_set(inf, outf):
    hs = []
    for i in range(N):
        w = synth_blocks(Q4_K, outf * inf, "cuda", i)
        h = C.c_void_p()
        native.check(lib.ktb200_linear_create(inf, outf, w.data_ptr(), Q4_K, BF16, 8, 0, C.byref(h)))
        native.check(lib.ktb200_linear_load_weights(h, S()))
        hs.append((h, w))
    x = torch.randn(1, inf, device="cuda").to(bf); y = torch.zeros(1, outf, dtype=bf, device="cuda")
    us = graph_time(lambda: [native.check(lib.ktb200_linear_forward(h, 1, x.data_ptr(), y.data_ptr(), None, None, S())) for h, _ in hs])
    mb = outf * inf * 144 / 256 / 1e6
    print(f"linear {inf:6d} -> {outf:6d}: {us:6.1f} us in-graph  ({mb:5.1f} MB, {mb / us * 1e3:6.0f} GB/s)", flush=True)
    for h, _ in hs:
        lib.ktb200_linear_destroy(h)


for shp in ((7168, 2112), (1536, 24576), (16384, 7168)):
    linear_set(*shp)
H = 7168
w = torch.ones(H, dtype=bf, device="cuda"); r = torch.randn(1, H, device="cuda").to(bf); d = torch.randn(1, H, device="cuda").to(bf); out = torch.zeros_like(r)
print(f"add_rmsnorm 7168 : {graph_time(lambda: [native.check(lib.ktb200_add_rmsnorm(r.data_ptr(), d.data_ptr(), w.data_ptr(), 1e-6, out.data_ptr(), 1, H, S())) for _ in range(N)]):6.1f} us")
wuk = [(torch.randn(128, 128, 512, device="cuda") * 0.05).to(bf) for _ in range(N)]
q = torch.randn(1, 128 * 192, device="cuda").to(bf); qa = torch.zeros(128, 512, dtype=bf, device="cuda"); o = torch.zeros(128, 128, dtype=bf, device="cuda")
print(f"absorb_q (16.8 MB): {graph_time(lambda: [native.check(lib.ktb200_mla_absorb_q(q.data_ptr(), 192, 128 * 192, wk.data_ptr(), 128, 128, 512, qa.data_ptr(), 1, S())) for wk in wuk]):6.1f} us")
print(f"absorb_o (16.8 MB): {graph_time(lambda: [native.check(lib.ktb200_mla_absorb_o(qa.data_ptr(), wk.data_ptr(), 128, 128, 512, o.data_ptr(), 1, S())) for wk in wuk]):6.1f} us")
for ctx in (1024, 4096):
    page = 64
    npg = ctx // page + 1
    kvs = [torch.randn((npg, page, 576), device="cuda", dtype=bf) for _ in range(N)]
    pt = torch.arange(npg, dtype=torch.int32, device="cuda")[None].contiguous()
    kl = torch.tensor([ctx], dtype=torch.int32, device="cuda")
    qn = torch.randn(1, 128, 512, device="cuda").to(bf); qp = torch.randn(1, 128, 64, device="cuda").to(bf)
    oo = torch.zeros(1, 128, 512, dtype=bf, device="cuda")
    wsb = lib.ktb200_mla_workspace_bytes(1, 128, 0); ws = torch.zeros(wsb, dtype=torch.uint8, device="cuda")
    ps = [native.MlaParams(1, 128, page, npg, 0, 0.072, qn.data_ptr(), qp.data_ptr(), kv.data_ptr(), pt.data_ptr(), kl.data_ptr(), oo.data_ptr(), None, ws.data_ptr(), wsb, npg * page) for kv in kvs]
    print(f"mla decode ctx {ctx}: {graph_time(lambda: [native.check(lib.ktb200_mla_decode(C.byref(p), S())) for p in ps]):6.1f} us (2 launches)")
