"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement of the reference's FP8 linear — DeepSeek's 128-block e4m3 scheme as the reference ships it:
    act_quant      archive/ktransformers/ktransformers_ext/triton/fp8gemm.py:10-47   s = max|x| / 448 per 128 values, y = e4m3(x / s)
    weight_dequant fp8gemm.py:50-95                                                  w * scale_inv[row // 128][col // 128]
    fp8_gemm       fp8gemm.py:104-192                                                acc += dot(a_blk, b_blk) * a_s * b_s per 128 of K, fp32
    KLinearFP8     archive/ktransformers/operators/linear.py:388-435                 forward = act_quant -> fp8_gemm -> cast back

Pinned: tests/golden/fp8_ref.npz holds outputs of the reference's OWN Triton kernels, run on the CPU by Triton's interpreter
(TRITON_INTERPRET=1; tests/golden/make_fp8_golden.py — the autotuner needs a device to time configs, so the script launches the
un-tuned `fp8_gemm_kernel.fn` with the first of the reference's own configs).  tests/test_oracle_pinned.py checks this file
against those vectors: scales exactly; quantised bytes exactly except for two artifacts of the interpreter's software cast — it
drops the carry when round-to-nearest crosses a binade (124.16 -> 64; the GPU's cvt.rn and this file give 128) and rounds exact
ties away from zero instead of to even — and the test checks that these are the ONLY differences; the GEMM's fp32 accumulator (on the golden's own bytes) bit for bit — the interpreter narrows it to
bf16 by truncation where the GPU rounds to nearest even, so the golden equals the accumulator's upper 16 bits."""
from __future__ import annotations

import numpy as np
import torch

BLOCK = 128
E4M3_MAX = 448.0


def to_e4m3_bytes(x_f32: np.ndarray) -> np.ndarray:
    """fp32 -> e4m3 (round to nearest even), as raw bytes."""
    return torch.from_numpy(np.ascontiguousarray(x_f32, dtype=np.float32)).to(torch.float8_e4m3fn).view(torch.uint8).numpy()


def e4m3_bytes_to_f32(b: np.ndarray) -> np.ndarray:
    return torch.from_numpy(np.ascontiguousarray(b, dtype=np.uint8)).view(torch.float8_e4m3fn).to(torch.float32).numpy()


def act_quant(x_f32: np.ndarray, block: int = BLOCK):
    """fp8gemm.py:19-27 — x [..., K] (already widened to fp32: `.to(tl.float32)`), returns (e4m3 bytes [..., K], s fp32 [..., K/block])."""
    x = np.ascontiguousarray(x_f32, dtype=np.float32)
    xb = x.reshape(*x.shape[:-1], x.shape[-1] // block, block)
    s = (np.abs(xb).max(axis=-1) / np.float32(E4M3_MAX)).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        y = (xb / s[..., None]).astype(np.float32)          # an all-zero block divides 0 / 0 exactly like the reference
    return to_e4m3_bytes(y).reshape(x.shape), s


def weight_dequant(w_bytes: np.ndarray, scale_inv: np.ndarray, block: int = BLOCK) -> np.ndarray:
    """fp8gemm.py:63-73 — w [N][K] e4m3 bytes, scale_inv [ceil(N/block)][ceil(K/block)] fp32."""
    w = e4m3_bytes_to_f32(w_bytes)
    N, K = w.shape
    s = np.repeat(np.repeat(scale_inv.astype(np.float32), block, axis=0)[:N], block, axis=1)[:, :K]
    return (w * s).astype(np.float32)


def fp8_gemm(a_bytes: np.ndarray, a_s: np.ndarray, b_bytes: np.ndarray, b_s: np.ndarray, block: int = BLOCK) -> np.ndarray:
    """fp8gemm.py:150-172 — a [M][K], a_s [M][K/block], b [N][K], b_s [ceil(N/block)][K/block] -> fp32 accumulator [M][N].
    Products of two e4m3 values are exact in fp32; the block dot is summed in float64 and rounded once (tl.dot's order is
    unspecified), then `* a_s[:, None] * b_s[None, :]` and the running sum are fp32 in the reference's order."""
    a, b = e4m3_bytes_to_f32(a_bytes).astype(np.float64), e4m3_bytes_to_f32(b_bytes).astype(np.float64)
    M, K = a.shape
    N = b.shape[0]
    acc = np.zeros((M, N), np.float32)
    rows = np.arange(N) // block
    for i in range(K // block):
        d = (a[:, i * block:(i + 1) * block] @ b[:, i * block:(i + 1) * block].T).astype(np.float32)
        acc = (acc + (d * a_s[:, i].astype(np.float32)[:, None]) * b_s[rows, i].astype(np.float32)[None, :]).astype(np.float32)
    return acc


def linear_forward(x_f32: np.ndarray, w_bytes: np.ndarray, scale_inv: np.ndarray) -> np.ndarray:
    """KLinearFP8.forward (linear.py:409-414) up to the final cast: fp32 accumulator [T][N]; the caller rounds to bf16."""
    q, s = act_quant(x_f32)
    return fp8_gemm(q, s, w_bytes, scale_inv)
