"""Minimal DeepSeek-V3 MoE block definitions — the injection targets of the B200 rule files.

Only the classes the hot path's YAML rules match against are defined here (the reference carries the
whole HF model, archive/ktransformers/models/modeling_deepseek_v3.py; everything outside the MoE block
is out of scope).  Semantics restate the reference:
    DeepseekV3MLP   :385-397     down(silu(gate x) * up x)
    MoEGate         :400-481     sigmoid scoring + noaux_tc grouped top-k
    DeepseekV3MoE   :483-616     gate -> routed experts (+ shared experts)
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn
from transformers.configuration_utils import PretrainedConfig


class DeepseekV3Config(PretrainedConfig):
    model_type = "deepseek_v3"

    def __init__(self, hidden_size=7168, intermediate_size=18432, moe_intermediate_size=2048, n_routed_experts=256,
                 n_shared_experts=1, num_experts_per_tok=8, n_group=8, topk_group=4, routed_scaling_factor=2.5,
                 norm_topk_prob=True, scoring_func="sigmoid", topk_method="noaux_tc", hidden_act="silu",
                 num_hidden_layers=61, first_k_dense_replace=3, moe_layer_freq=1, num_attention_heads=128,
                 q_lora_rank=1536, kv_lora_rank=512, qk_rope_head_dim=64, qk_nope_head_dim=128, v_head_dim=128,
                 vocab_size=129280, rms_norm_eps=1e-6, ep_size=1, **kwargs):
        self.hidden_size = hidden_size
        self.intermediate_size = intermediate_size
        self.moe_intermediate_size = moe_intermediate_size
        self.n_routed_experts = n_routed_experts
        self.n_shared_experts = n_shared_experts
        self.num_experts_per_tok = num_experts_per_tok
        self.n_group = n_group
        self.topk_group = topk_group
        self.routed_scaling_factor = routed_scaling_factor
        self.norm_topk_prob = norm_topk_prob
        self.scoring_func = scoring_func
        self.topk_method = topk_method
        self.hidden_act = hidden_act
        self.num_hidden_layers = num_hidden_layers
        self.first_k_dense_replace = first_k_dense_replace
        self.moe_layer_freq = moe_layer_freq
        self.num_attention_heads = num_attention_heads
        self.q_lora_rank = q_lora_rank
        self.kv_lora_rank = kv_lora_rank
        self.qk_rope_head_dim = qk_rope_head_dim
        self.qk_nope_head_dim = qk_nope_head_dim
        self.v_head_dim = v_head_dim
        self.vocab_size = vocab_size
        self.rms_norm_eps = rms_norm_eps
        self.ep_size = ep_size
        super().__init__(**kwargs)


class DeepseekV3MLP(nn.Module):
    def __init__(self, config, hidden_size=None, intermediate_size=None):
        super().__init__()
        self.config = config
        self.hidden_size = hidden_size or config.hidden_size
        self.intermediate_size = intermediate_size or config.intermediate_size
        self.gate_proj = nn.Linear(self.hidden_size, self.intermediate_size, bias=False)
        self.up_proj = nn.Linear(self.hidden_size, self.intermediate_size, bias=False)
        self.down_proj = nn.Linear(self.intermediate_size, self.hidden_size, bias=False)
        self.act_fn = F.silu

    def forward(self, x):
        return self.down_proj(self.act_fn(self.gate_proj(x)) * self.up_proj(x))


class MoEGate(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.top_k = config.num_experts_per_tok
        self.n_routed_experts = config.n_routed_experts
        self.routed_scaling_factor = config.routed_scaling_factor
        self.scoring_func = config.scoring_func
        self.topk_method = config.topk_method
        self.n_group = config.n_group
        self.topk_group = config.topk_group
        self.norm_topk_prob = config.norm_topk_prob
        self.gating_dim = config.hidden_size
        self.weight = nn.Parameter(torch.empty((self.n_routed_experts, self.gating_dim)))
        if self.topk_method == "noaux_tc":
            self.e_score_correction_bias = nn.Parameter(torch.empty((self.n_routed_experts)))
        if self.weight.device.type != "meta":
            nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
            if self.topk_method == "noaux_tc":
                nn.init.zeros_(self.e_score_correction_bias)

    def forward(self, hidden_states):
        bsz, seq_len, h = hidden_states.shape
        x = hidden_states.view(-1, h)
        logits = F.linear(x.type(torch.float32), self.weight.type(torch.float32), None)
        if self.scoring_func != "sigmoid":
            raise NotImplementedError(f"insupportable scoring function for MoE gating: {self.scoring_func}")
        scores = logits.sigmoid()
        if self.topk_method != "noaux_tc":
            raise NotImplementedError(f"insupportable TopK function for MoE gating: {self.topk_method}")
        n = bsz * seq_len
        choice = scores.view(n, -1) + self.e_score_correction_bias.unsqueeze(0)
        group_scores = choice.view(n, self.n_group, -1).topk(2, dim=-1)[0].sum(dim=-1)
        group_idx = torch.topk(group_scores, k=self.topk_group, dim=-1, sorted=False)[1]
        group_mask = torch.zeros_like(group_scores)
        group_mask.scatter_(1, group_idx, 1)
        score_mask = group_mask.unsqueeze(-1).expand(n, self.n_group, self.n_routed_experts // self.n_group).reshape(n, -1)
        tmp = choice.masked_fill(~score_mask.bool(), float("-inf"))
        _, topk_idx = torch.topk(tmp, k=self.top_k, dim=-1, sorted=False)
        topk_weight = scores.gather(1, topk_idx)
        if self.top_k > 1 and self.norm_topk_prob:
            topk_weight = topk_weight / (topk_weight.sum(dim=-1, keepdim=True) + 1e-20)
        topk_weight = topk_weight * self.routed_scaling_factor
        return topk_idx, topk_weight


class DeepseekV3MoE(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.num_experts_per_tok = config.num_experts_per_tok
        self.ep_size, self.experts_per_rank, self.ep_rank = 1, config.n_routed_experts, 0
        self.experts = nn.ModuleList([DeepseekV3MLP(config, intermediate_size=config.moe_intermediate_size)
                                      for _ in range(config.n_routed_experts)])
        self.gate = MoEGate(config)
        if config.n_shared_experts is not None:
            self.shared_experts = DeepseekV3MLP(config, intermediate_size=config.moe_intermediate_size * config.n_shared_experts)

    def forward(self, hidden_states):
        identity = hidden_states
        orig_shape = hidden_states.shape
        topk_idx, topk_weight = self.gate(hidden_states)
        x = hidden_states.view(-1, hidden_states.shape[-1])
        y = torch.zeros_like(x)
        for t in range(x.shape[0]):
            for j in range(topk_idx.shape[1]):
                y[t] += self.experts[int(topk_idx[t, j])](x[t]) * topk_weight[t, j].to(x.dtype)
        y = y.view(*orig_shape)
        if self.config.n_shared_experts is not None:
            y = y + self.shared_experts(identity)
        return y


class DeepseekV3DecoderLayerMoEOnly(nn.Module):
    """A decoder layer reduced to its MoE block (what the hot path covers)."""

    def __init__(self, config, layer_idx: int):
        super().__init__()
        dense = layer_idx < config.first_k_dense_replace
        self.mlp = DeepseekV3MLP(config) if dense else DeepseekV3MoE(config)

    def forward(self, hidden_states):
        return hidden_states + self.mlp(hidden_states)


class DeepseekV3MoEStack(nn.Module):
    """`model.layers.N.mlp` naming so that GGUF keys (`blk.N.ffn_*`) and rule regexes line up with the
    reference's full model."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.layers = nn.ModuleList([DeepseekV3DecoderLayerMoEOnly(config, i) for i in range(config.num_hidden_layers)])

    def forward(self, hidden_states):
        for layer in self.layers:
            hidden_states = layer(hidden_states)
        return hidden_states


class DeepseekV3MoEOnlyForCausalLM(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.model = DeepseekV3MoEStack(config)
        self.lm_head = nn.Identity()

    def forward(self, hidden_states):
        return self.model(hidden_states)
