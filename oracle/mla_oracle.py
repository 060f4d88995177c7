"""TEST INFRASTRUCTURE ONLY — numpy restatement of absorbed-MLA paged decode attention.

Reference math: archive/ktransformers/operators/attention.py:395-478 (absorbed decode),
flashinfer_wrapper.attention_ref_torch (archive/ktransformers/operators/flashinfer_wrapper.py:30-76) and
the Triton split-KV kernel (triton_attention.py:16-163: fp32 scores/softmax, P cast to the KV dtype before P.V).
    s[h,t] = (q_nope[h] . ckv[t] + q_pe[h] . k_pe[t]) * sm_scale ; p = softmax_t(s) ; out[h] = sum_t p[h,t] ckv[t]
The reference's own tolerances for this op are loose (rel-mean < 2e-1, kt-kernel/examples/test_mla.py:724); the GPU
tests use much tighter ones.  Parity status: PINNED — tests/golden/mla_ref.npz holds outputs of the reference's own
`attention_ref_torch` (executed from /root/reference by tests/golden/make_mla_golden.py); tests/test_oracle_pinned.py
checks this restatement against them (bf16 output rounding, base-2 LSE to 1e-4).
"""
from __future__ import annotations

import numpy as np


def bf16_round(x: np.ndarray) -> np.ndarray:
    i = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    r = ((i + (0x7FFF + ((i >> 16) & 1))) >> 16).astype(np.uint32) << 16
    return r.view(np.float32)


def gather_kv(kv_cache: np.ndarray, page_table_row: np.ndarray, length: int, page_size: int) -> np.ndarray:
    pages = page_table_row[: (length + page_size - 1) // page_size]
    return kv_cache[pages].reshape(-1, kv_cache.shape[-1])[:length]


def mla_decode(q_nope, q_pe, kv_cache, page_table, kv_len, sm_scale, p_bf16: bool = True):
    """q_nope [B,H,512], q_pe [B,H,64], kv_cache [pages,page,576] (float32 holding bf16 values),
    page_table [B,max_pages] int, kv_len [B].  Returns out [B,H,512] float32 and lse [B,H] (natural log)."""
    B, H, _ = q_nope.shape
    page_size = kv_cache.shape[1]
    out = np.zeros((B, H, 512), np.float32)
    lse = np.zeros((B, H), np.float32)
    for b in range(B):
        kv = gather_kv(kv_cache, page_table[b], int(kv_len[b]), page_size).astype(np.float64)
        s = (q_nope[b].astype(np.float64) @ kv[:, :512].T + q_pe[b].astype(np.float64) @ kv[:, 512:].T) * sm_scale
        m = s.max(axis=1, keepdims=True)
        e = np.exp(s - m)
        den = e.sum(axis=1, keepdims=True)
        p = e / den
        if p_bf16:
            # the kernels cast exp(s-m) (not p) to bf16 before P.V and divide by the fp32 sum afterwards
            pv = bf16_round(e.astype(np.float32)).astype(np.float64) @ kv[:, :512] / den
        else:
            pv = p @ kv[:, :512]
        out[b] = pv.astype(np.float32)
        lse[b] = (m[:, 0] + np.log(den[:, 0])).astype(np.float32)
    return out, lse
