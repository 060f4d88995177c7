"""Debug: Kimi-K2-shaped MoE block (E=384) qlen=8 — fused launch vs separate launches vs oracle, per token."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_util as G
from ktransformers_b200 import native
from ktransformers_b200.util.synth import synth_blocks
from oracle.bindings import BF16, Q4_K, Q6_K, Oracle, bf16_to_f32, f32_to_bf16_bits
oracle = Oracle()
E, H, I, k, ng, tg, scale = 384, 7168, 2048, 8, 1, 1, 2.827
sy = lambda t, n, s: synth_blocks(t, n, device="cuda", seed=s)
gate_w, up_w, down_w = sy(Q4_K, E * I * H, 301), sy(Q4_K, E * I * H, 302), sy(Q6_K, E * H * I, 303)
sg, su, sd = sy(Q4_K, I * H, 304), sy(Q4_K, I * H, 305), sy(Q6_K, H * I, 306)
down_raw = down_w.clone()
sg_np, su_np, sd_np = sg.cpu().numpy(), su.cpu().numpy(), sd.cpu().numpy()
gb, db = gate_w.numel() // E, down_w.numel() // E
m = G.Moe(E, k, H, I, gate_w, up_w, down_w, Q4_K, Q4_K, Q6_K, BF16, max_tokens=8)
mlp = G.Mlp(H, I, sg, su, sd, Q4_K, Q4_K, Q6_K, BF16)
rng = np.random.default_rng(E + H)
W = rng.standard_normal((E, H)).astype(np.float32); bias = rng.standard_normal(E).astype(np.float32)
gate = G.Gate(W, bias, k, ng, tg, scale=scale, hidden_type=BF16)
for qlen in (1, 8):
    xb = f32_to_bf16_bits((rng.standard_normal((qlen, H)) / 100).astype(np.float32))
    out, idx, w = G.moe_block_forward(gate, m, mlp, xb)
    sep = G.moe_forward_shared(m, mlp, idx, w, xb)
    routed_only = m.forward(idx, w, xb)
    sel = sorted(set(idx.reshape(-1).tolist())); remap = {e: i for i, e in enumerate(sel)}
    g_np = torch.cat([gate_w[e * gb:(e + 1) * gb] for e in sel]).cpu().numpy()
    u_np = torch.cat([up_w[e * gb:(e + 1) * gb] for e in sel]).cpu().numpy()
    d_np = torch.cat([down_raw[e * db:(e + 1) * db] for e in sel]).cpu().numpy()
    ids_l = np.vectorize(remap.get)(idx).astype(np.int64)
    routed = oracle.moe_forward(len(sel), H, I, g_np, u_np, d_np, Q4_K, Q4_K, Q6_K, BF16, ids_l, w, xb)
    shared = oracle.mlp_forward(H, I, sg_np, su_np, sd_np, Q4_K, Q4_K, Q6_K, BF16, xb)
    want = bf16_to_f32((torch.from_numpy(routed.view(np.int16)).view(torch.bfloat16) + torch.from_numpy(shared.view(np.int16)).view(torch.bfloat16)).view(torch.int16).numpy().view(np.uint16))
    o, s_, r_, ro = bf16_to_f32(out), bf16_to_f32(sep), bf16_to_f32(routed), bf16_to_f32(routed_only)
    print(f"qlen={qlen} fused==separate: {np.array_equal(out, sep)}")
    for t in range(qlen):
        print(f"  t={t} ids={idx[t].tolist()} |want|max={np.abs(want[t]).max():.2f} fused err={np.abs(o[t]-want[t]).max():.3f} sep err={np.abs(s_[t]-want[t]).max():.3f} routed-only err={np.abs(ro[t]-r_[t]).max():.3f}")
