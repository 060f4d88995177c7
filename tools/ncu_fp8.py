"""ncu target: the FP8 linear at the lm_head and o_proj shapes, bs 1 (weights cycled so that every call streams from HBM)."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ktransformers_b200 import native
lib = native.lib(); s = torch.cuda.current_stream().cuda_stream
for K, N in ((7168, 129280), (16384, 7168)):
    hs, keep = [], []
    for c in range(2 if N > 100000 else 4):
        w = torch.randint(0, 120, (N, K), dtype=torch.uint8, device="cuda"); ws = torch.rand(((N + 127) // 128, K // 128), device="cuda") * 0.01 + 0.001
        h = C.c_void_p(); native.check(lib.ktb200_fp8_linear_create(K, N, w.data_ptr(), ws.data_ptr(), 30, 0, C.byref(h))); hs.append(h); keep.append((w, ws))
    x = (torch.randn(1, K, device="cuda") / 10).bfloat16(); y = torch.zeros(1, N, dtype=torch.bfloat16, device="cuda")
    for i in range(3 * len(hs)):
        native.check(lib.ktb200_fp8_linear_forward(hs[i % len(hs)], 1, x.data_ptr(), y.data_ptr(), None, s))
    torch.cuda.synchronize()
