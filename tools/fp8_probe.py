"""Times ktb200_fp8_linear_forward at the DeepSeek-V3 projection shapes (bs = 1 and 8): GB/s of e4m3 weight bytes against the HBM roofline.
Several weight copies are cycled so that no call finds its weights in L2 (126 MB)."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ktransformers_b200 import native
lib = native.lib()
s = torch.cuda.current_stream().cuda_stream
for name, K, N in (("q_a + kv_a", 7168, 2112), ("q_b", 1536, 24576), ("o_proj", 16384, 7168), ("shared gate/up", 7168, 4096), ("lm_head", 7168, 129280)):
    copies = max(2, int(400e6 // (K * N)) + 1)
    hs, keep = [], []
    for c in range(copies):
        w = torch.randint(0, 120, (N, K), dtype=torch.uint8, device="cuda")          # positive finite e4m3 bit patterns
        ws = torch.rand(((N + 127) // 128, K // 128), device="cuda") * 0.01 + 0.001
        h = C.c_void_p(); native.check(lib.ktb200_fp8_linear_create(K, N, w.data_ptr(), ws.data_ptr(), 30, 0, C.byref(h)))
        hs.append(h); keep.append((w, ws))
    for T in (1, 8):
        x = (torch.randn(T, K, device="cuda") / 10).bfloat16(); y = torch.zeros(T, N, dtype=torch.bfloat16, device="cuda")
        def run(i): native.check(lib.ktb200_fp8_linear_forward(hs[i % copies], T, x.data_ptr(), y.data_ptr(), None, s))
        for i in range(copies): run(i)
        torch.cuda.synchronize()
        n = 6 * copies
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for i in range(n): run(i)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / n * 1e3
        print(f"{name:15s} {K:6d} -> {N:6d}  T={T}: {us:7.1f} us  {K * N / us / 1e3:7.0f} GB/s  ({K * N / 1e6:.1f} MB, {copies} copies, back-to-back launches)", flush=True)
    for h in hs: lib.ktb200_fp8_linear_destroy(h)
