"""Golden routing for the reference's own gate test recipe (kt-kernel/examples/test_gate.py:17-35, 197-215):
torch.manual_seed(42); weights = randn(256, 7168); bias = randn(256); input = randn(64, 7168), DeepSeek-V3 routing.
The expected ids / weights come from the reference's pure-torch MoEGate.forward source
(archive/ktransformers/models/modeling_deepseek_v3.py:430-481) executed here on CPU; the inputs are NOT stored (9 MB):
the test regenerates them from the seed and checks `probe` (a few sampled values) to detect a torch RNG change.

    python tests/golden/make_gate_seed42.py      (build container only: needs /root/reference)
"""
import os
import types

import numpy as np
import torch

OUT = os.path.dirname(os.path.abspath(__file__))
src = open("/root/reference/archive/ktransformers/models/modeling_deepseek_v3.py").read()
ns = {"torch": torch, "nn": torch.nn, "F": torch.nn.functional, "math": __import__("math")}
exec(src[src.index("class MoEGate(nn.Module):"):src.index("class DeepseekV3MoE(nn.Module):")], ns)

E, H, k, T = 256, 7168, 8, 64
torch.manual_seed(42)
W = torch.randn((E, H), dtype=torch.float32)
bias = torch.randn((E,), dtype=torch.float32)
x = torch.randn(T, H, dtype=torch.float32)
cfg = types.SimpleNamespace(num_experts_per_tok=k, n_routed_experts=E, routed_scaling_factor=2.5, scoring_func="sigmoid",
                            topk_method="noaux_tc", n_group=8, topk_group=4, norm_topk_prob=True, hidden_size=H)
gate = ns["MoEGate"](cfg)
with torch.no_grad():
    gate.weight.copy_(W)
    gate.e_score_correction_bias.copy_(bias)
    idx, w = gate(x[None])
probe = np.array([W[0, 0], W[255, 7167], bias[7], x[0, 0], x[63, 7167]], np.float32)
np.savez_compressed(os.path.join(OUT, "gate_seed42.npz"), idx=idx.numpy(), w=w.numpy(), probe=probe)
print("ids[0]", sorted(idx[0].tolist()), "probe", probe)
