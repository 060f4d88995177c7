"""Helpers for the GPU parity tests: drive the C-ABI (include/ktb200.h) with torch-owned device memory."""
import ctypes as C

import numpy as np
import torch

from ktransformers_b200 import native

NP_HID = {0: np.float32, 1: np.float16, 30: np.uint16}
TORCH_HID = {0: torch.float32, 1: torch.float16, 30: torch.bfloat16}


def dev(a, dtype=None):
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint16:
        a = a.view(np.int16)
    t = torch.from_numpy(a)
    if dtype is not None:
        t = t.view(dtype)
    return t.cuda()


def stream():
    return torch.cuda.current_stream().cuda_stream


class Moe:
    """ktb200_moe handle owning device copies of raw ggml blocks (uploaded from numpy or given as cuda tensors)."""

    def __init__(self, E, k, H, I, gate, up, down, gt, ut, dt, hidden_type, max_tokens=64, use_silu=1, offset=0):
        self.lib = native.lib()
        self.E, self.k, self.H, self.I, self.hidden_type = E, k, H, I, hidden_type
        self.gate = gate if isinstance(gate, torch.Tensor) else dev(gate)
        self.up = up if isinstance(up, torch.Tensor) else dev(up)
        self.down = down if isinstance(down, torch.Tensor) else dev(down)
        cfg = native.MoeConfig(E, k, H, I, 64, 10, max_tokens, use_silu, self.gate.data_ptr(), self.up.data_ptr(),
                               self.down.data_ptr(), gt, ut, dt, hidden_type, offset)
        self.h = C.c_void_p()
        native.check(self.lib.ktb200_moe_create(C.byref(cfg), torch.cuda.current_device(), C.byref(self.h)))
        native.check(self.lib.ktb200_moe_load_weights(self.h, stream()))

    def forward(self, ids, w, x, bsz=None, out=None):
        """numpy in, numpy out (x / out in the numpy carrier of hidden_type: bf16 as uint16 bits)."""
        qlen, k = ids.shape
        ids_d, w_d = dev(ids.astype(np.int64)), dev(w.astype(np.float32))
        x_d = dev(x, TORCH_HID[self.hidden_type] if self.hidden_type == 30 else None)
        out_d = torch.zeros((qlen, self.H), dtype=TORCH_HID[self.hidden_type], device="cuda") if out is None else out
        bsz_d = torch.tensor([bsz], dtype=torch.int32, device="cuda") if bsz is not None else None
        native.check(self.lib.ktb200_moe_forward(self.h, qlen, k, ids_d.data_ptr(), w_d.data_ptr(), x_d.data_ptr(),
                                                 out_d.data_ptr(), bsz_d.data_ptr() if bsz_d is not None else None, stream()))
        torch.cuda.synchronize()
        o = out_d.cpu()
        return o.view(torch.int16).numpy().view(np.uint16) if self.hidden_type == 30 else o.numpy()

    def forward_host(self, ids, w, x):
        qlen, k = ids.shape
        out = np.zeros((qlen, self.H), NP_HID[self.hidden_type])
        ids, w, x = np.ascontiguousarray(ids, np.int64), np.ascontiguousarray(w, np.float32), np.ascontiguousarray(x)
        native.check(self.lib.ktb200_moe_forward_host(self.h, qlen, k, ids.ctypes.data, w.ctypes.data, x.ctypes.data,
                                                      out.ctypes.data, stream()))
        return out

    def close(self):
        if self.h:
            self.lib.ktb200_moe_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Mlp:
    def __init__(self, H, I, g, u, d, gt, ut, dt, hidden_type):
        self.lib = native.lib()
        self.keep = tuple(t if isinstance(t, torch.Tensor) else dev(t) for t in (g, u, d))
        self.h = C.c_void_p()
        native.check(self.lib.ktb200_mlp_create(H, I, self.keep[0].data_ptr(), self.keep[1].data_ptr(), self.keep[2].data_ptr(),
                                                gt, ut, dt, hidden_type, 64, torch.cuda.current_device(), C.byref(self.h)))
        native.check(self.lib.ktb200_mlp_load_weights(self.h, stream()))

    def close(self):
        if self.h:
            self.lib.ktb200_mlp_destroy(self.h)
            self.h = None


def moe_forward_shared(moe: "Moe", mlp: "Mlp", ids, w, x):
    qlen, k = ids.shape
    ids_d, w_d = dev(ids.astype(np.int64)), dev(w.astype(np.float32))
    x_d = dev(x, TORCH_HID[moe.hidden_type] if moe.hidden_type == 30 else None)
    out_d = torch.zeros((qlen, moe.H), dtype=TORCH_HID[moe.hidden_type], device="cuda")
    native.check(moe.lib.ktb200_moe_forward_shared(moe.h, mlp.h if mlp is not None else None, qlen, k, ids_d.data_ptr(), w_d.data_ptr(),
                                                   x_d.data_ptr(), out_d.data_ptr(), None, stream()))
    torch.cuda.synchronize()
    o = out_d.cpu()
    return o.view(torch.int16).numpy().view(np.uint16) if moe.hidden_type == 30 else o.numpy()


def linear_forward(in_size, out_size, proj, proj_type, hidden_type, x, bias=None):
    lib = native.lib()
    p = proj if isinstance(proj, torch.Tensor) else dev(proj)
    h = C.c_void_p()
    native.check(lib.ktb200_linear_create(in_size, out_size, p.data_ptr(), proj_type, hidden_type, 64, torch.cuda.current_device(), C.byref(h)))
    native.check(lib.ktb200_linear_load_weights(h, stream()))
    x_d = dev(x, TORCH_HID[hidden_type] if hidden_type == 30 else None)
    out = torch.zeros((x.shape[0], out_size), dtype=TORCH_HID[hidden_type], device="cuda")
    b = dev(bias.astype(np.float32)) if bias is not None else None
    native.check(lib.ktb200_linear_forward(h, x.shape[0], x_d.data_ptr(), out.data_ptr(), b.data_ptr() if b is not None else None, None, stream()))
    torch.cuda.synchronize()
    lib.ktb200_linear_destroy(h)
    o = out.cpu()
    return o.view(torch.int16).numpy().view(np.uint16) if hidden_type == 30 else o.numpy()


def mlp_forward(H, I, g, u, d, gt, ut, dt, hidden_type, x, accumulate_into=None):
    lib = native.lib()
    gd, ud, dd = (t if isinstance(t, torch.Tensor) else dev(t) for t in (g, u, d))
    h = C.c_void_p()
    native.check(lib.ktb200_mlp_create(H, I, gd.data_ptr(), ud.data_ptr(), dd.data_ptr(), gt, ut, dt, hidden_type, 64, torch.cuda.current_device(), C.byref(h)))
    native.check(lib.ktb200_mlp_load_weights(h, stream()))
    x_d = dev(x, TORCH_HID[hidden_type] if hidden_type == 30 else None)
    if accumulate_into is not None:
        out = dev(accumulate_into, TORCH_HID[hidden_type] if hidden_type == 30 else None).clone()
    else:
        out = torch.zeros((x.shape[0], H), dtype=TORCH_HID[hidden_type], device="cuda")
    native.check(lib.ktb200_mlp_forward(h, x.shape[0], x_d.data_ptr(), out.data_ptr(), int(accumulate_into is not None), None, stream()))
    torch.cuda.synchronize()
    lib.ktb200_mlp_destroy(h)
    o = out.cpu()
    return o.view(torch.int16).numpy().view(np.uint16) if hidden_type == 30 else o.numpy()


def quantize(x, hidden_type, act_type):
    lib = native.lib()
    rows, cols = x.shape
    x_d = dev(x, TORCH_HID[hidden_type] if hidden_type == 30 else None)
    per = {15: (256, 292), 8: (32, 34)}[act_type]
    out = torch.zeros(rows * cols // per[0] * per[1], dtype=torch.uint8, device="cuda")
    native.check(lib.ktb200_quantize_activations(x_d.data_ptr(), hidden_type, rows, cols, act_type, out.data_ptr(), stream()))
    torch.cuda.synchronize()
    return out.cpu().numpy().reshape(rows, -1)


def dequantize(raw, ggml_type, n, out_type=0):
    lib = native.lib()
    r = raw if isinstance(raw, torch.Tensor) else dev(raw)
    out = torch.zeros(n, dtype=TORCH_HID[out_type], device="cuda")
    native.check(lib.ktb200_dequantize(r.data_ptr(), ggml_type, n, out.data_ptr(), out_type, stream()))
    torch.cuda.synchronize()
    return out.cpu()


def gate_forward(x, W, bias, top_k, n_group, topk_group, scoring=0, topk_method=0, norm=1, scale=2.5, hidden_type=0, want_logits=False):
    lib = native.lib()
    T, H = x.shape
    E = W.shape[0]
    Wd, bd = dev(W.astype(np.float32)), (dev(bias.astype(np.float32)) if bias is not None else None)
    x_d = dev(x, TORCH_HID[hidden_type] if hidden_type == 30 else None)
    idx = torch.zeros((T, top_k), dtype=torch.int64, device="cuda")
    w = torch.zeros((T, top_k), dtype=torch.float32, device="cuda")
    logits = torch.zeros((T, E), dtype=torch.float32, device="cuda") if want_logits else None
    cfg = native.GateConfig(E, H, top_k, n_group, topk_group, scoring, topk_method, norm, scale, Wd.data_ptr(),
                            bd.data_ptr() if bd is not None else None, hidden_type)
    native.check(lib.ktb200_moe_gate_forward(C.byref(cfg), T, x_d.data_ptr(), idx.data_ptr(), w.data_ptr(),
                                             logits.data_ptr() if logits is not None else None, None, stream()))
    torch.cuda.synchronize()
    return idx.cpu().numpy(), w.cpu().numpy(), (logits.cpu().numpy() if logits is not None else None)


class Gate:
    """ktb200_gate_config with device-resident router weights."""

    def __init__(self, W, bias, top_k, n_group, topk_group, scoring=0, topk_method=0, norm=1, scale=2.5, hidden_type=0):
        self.Wd = dev(W.astype(np.float32))
        self.bd = dev(bias.astype(np.float32)) if bias is not None else None
        self.top_k, self.E = top_k, W.shape[0]
        self.cfg = native.GateConfig(W.shape[0], W.shape[1], top_k, n_group, topk_group, scoring, topk_method, norm, scale,
                                     self.Wd.data_ptr(), self.bd.data_ptr() if self.bd is not None else None, hidden_type)


def moe_block_forward(gate: "Gate", moe: "Moe", mlp, x, repeats=1, graph=False):
    """ktb200_moe_block_forward -> (out, idx, w) as numpy; `repeats` back-to-back calls (the barrier words must reset),
    optionally captured in a CUDA graph and replayed."""
    lib = native.lib()
    qlen = x.shape[0]
    x_d = dev(x, TORCH_HID[moe.hidden_type] if moe.hidden_type == 30 else None)
    out_d = torch.zeros((qlen, moe.H), dtype=TORCH_HID[moe.hidden_type], device="cuda")
    idx = torch.zeros((qlen, gate.top_k), dtype=torch.int64, device="cuda")
    w = torch.zeros((qlen, gate.top_k), dtype=torch.float32, device="cuda")

    def call(st):
        native.check(lib.ktb200_moe_block_forward(C.byref(gate.cfg), moe.h, mlp.h if mlp is not None else None, qlen, x_d.data_ptr(),
                                                  out_d.data_ptr(), idx.data_ptr(), w.data_ptr(), None, st))
    call(stream())
    torch.cuda.synchronize()
    if graph:
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        with torch.cuda.graph(g, stream=side):
            call(torch.cuda.current_stream().cuda_stream)
        for _ in range(repeats):
            g.replay()
    else:
        for _ in range(repeats - 1):
            call(stream())
    torch.cuda.synchronize()
    o = out_d.cpu()
    o = o.view(torch.int16).numpy().view(np.uint16) if moe.hidden_type == 30 else o.numpy()
    return o, idx.cpu().numpy(), w.cpu().numpy()


def mla_decode(q_nope, q_pe, kv_cache, page_table, kv_len, sm_scale, num_kv_splits=0):
    """numpy float32 arrays holding bf16 values -> (out [B,H,512] float32, lse [B,H])"""
    lib = native.lib()
    B, Hq, _ = q_nope.shape
    to_bf = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(torch.bfloat16).cuda()
    qn, qp, kv = to_bf(q_nope), to_bf(q_pe), to_bf(kv_cache)
    pt = torch.from_numpy(np.ascontiguousarray(page_table.astype(np.int32))).cuda()
    kl = torch.from_numpy(np.ascontiguousarray(kv_len.astype(np.int32))).cuda()
    out = torch.zeros((B, Hq, 512), dtype=torch.bfloat16, device="cuda")
    lse = torch.zeros((B, Hq), dtype=torch.float32, device="cuda")
    ws_bytes = lib.ktb200_mla_workspace_bytes(B, Hq, 0)
    ws = torch.zeros(ws_bytes, dtype=torch.uint8, device="cuda")
    p = native.MlaParams(B, Hq, kv_cache.shape[1], page_table.shape[1], num_kv_splits, float(sm_scale), qn.data_ptr(), qp.data_ptr(),
                         kv.data_ptr(), pt.data_ptr(), kl.data_ptr(), out.data_ptr(), lse.data_ptr(), ws.data_ptr(), ws_bytes)
    native.check(lib.ktb200_mla_decode(C.byref(p), stream()))
    torch.cuda.synchronize()
    return out.float().cpu().numpy(), lse.cpu().numpy()


def mla_kv_write(kv_cache_t, page_size, ckv, k_pe, page_idx, page_off):
    lib = native.lib()
    native.check(lib.ktb200_mla_kv_write(kv_cache_t.data_ptr(), page_size, ckv.data_ptr(), k_pe.data_ptr(), page_idx.data_ptr(),
                                         page_off.data_ptr(), ckv.shape[0], stream()))
    torch.cuda.synchronize()
