"""Synthetic GGUF weight blocks of the real shapes (there are no checkpoints in this environment).

`synth_blocks` fills a tensor with random but WELL-FORMED ggml blocks of a given type: random quants and
sub-scales, fp16 super-block scales chosen so that dequantised weights are O(1) — the same distribution
class as `randn` weights pushed through ggml's from_float (kt-kernel/bench/bench_moe.py:166-170), without
needing a CPU quantiser for 7 GB of weights per layer.  Works on CPU and CUDA tensors.
"""
from __future__ import annotations

import torch

from .custom_gguf import GGML_QUANT_SIZES, GGMLQuantizationType

# byte offset of the fp16 scale fields inside a block, and the magnitude that makes values O(1)
_SCALE_FIELDS = {
    GGMLQuantizationType.Q4_K: ([0], [2], 1.0 / (40 * 8)),       # d at 0, dmin at 2 ; value ~ d*sc*q
    GGMLQuantizationType.Q5_K: ([0], [2], 1.0 / (40 * 16)),
    GGMLQuantizationType.Q6_K: ([208], [], 1.0 / (64 * 20)),
    GGMLQuantizationType.Q2_K: ([80], [82], 1.0 / (8 * 2)),
    GGMLQuantizationType.Q3_K: ([108], [], 1.0 / (20 * 3)),
    GGMLQuantizationType.IQ4_XS: ([0], [], 1.0 / (20 * 60)),
    GGMLQuantizationType.Q8_0: ([0], [], 1.0 / 70),
}


def synth_blocks(ggml_type: int, n_elements: int, device="cpu", seed: int = 0) -> torch.Tensor:
    t = GGMLQuantizationType(int(ggml_type))
    epb, bpb = GGML_QUANT_SIZES[t]
    assert n_elements % epb == 0
    nb = n_elements // epb
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    out = torch.empty((nb, bpb), dtype=torch.uint8, device=device)
    # fill in chunks to bound temporary memory
    step = max(1, (1 << 28) // bpb)
    for b0 in range(0, nb, step):
        b1 = min(nb, b0 + step)
        out[b0:b1] = torch.randint(0, 256, (b1 - b0, bpb), dtype=torch.uint8, device=device, generator=gen)
    d_offs, dmin_offs, mag = _SCALE_FIELDS[t]
    for offs, scale in ((d_offs, mag), (dmin_offs, mag * 0.5)):
        for off in offs:
            vals = (torch.rand(nb, device=device, generator=gen) * 0.5 + 0.75) * scale
            h = vals.to(torch.float16).view(torch.int16)
            out[:, off] = (h & 0xFF).to(torch.uint8)
            out[:, off + 1] = ((h >> 8) & 0xFF).to(torch.uint8)
    return out.reshape(-1)
