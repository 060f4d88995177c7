"""Minimal DeepSeek-V2 MoE block (softmax scoring, greedy / group_limited_greedy routing) — injection
target for V2 / V2-Lite rules.  Restates archive/ktransformers/models/modeling_deepseek.py:386-459."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn

from .modeling_deepseek_v3 import DeepseekV3Config, DeepseekV3MLP


class DeepseekV2Config(DeepseekV3Config):
    model_type = "deepseek_v2"

    def __init__(self, hidden_size=2048, moe_intermediate_size=1408, n_routed_experts=64, n_shared_experts=2,
                 num_experts_per_tok=6, n_group=1, topk_group=1, routed_scaling_factor=1.0, norm_topk_prob=False,
                 scoring_func="softmax", topk_method="greedy", **kwargs):
        super().__init__(hidden_size=hidden_size, moe_intermediate_size=moe_intermediate_size,
                         n_routed_experts=n_routed_experts, n_shared_experts=n_shared_experts,
                         num_experts_per_tok=num_experts_per_tok, n_group=n_group, topk_group=topk_group,
                         routed_scaling_factor=routed_scaling_factor, norm_topk_prob=norm_topk_prob,
                         scoring_func=scoring_func, topk_method=topk_method, **kwargs)


DeepseekV2MLP = DeepseekV3MLP


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
        if self.weight.device.type != "meta":
            nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))

    def forward(self, hidden_states):
        bsz, seq_len, h = hidden_states.shape
        x = hidden_states.view(-1, h)
        logits = F.linear(x.type(torch.float32), self.weight.type(torch.float32), None)
        if self.scoring_func != "softmax":
            raise NotImplementedError(f"insupportable scoring function for MoE gating: {self.scoring_func}")
        scores = logits.softmax(dim=-1, dtype=torch.float32)
        n = bsz * seq_len
        if self.topk_method == "greedy":
            topk_weight, topk_idx = torch.topk(scores, k=self.top_k, dim=-1, sorted=False)
        elif self.topk_method == "group_limited_greedy":
            group_scores = scores.view(n, self.n_group, -1).max(dim=-1).values
            group_idx = torch.topk(group_scores, k=self.topk_group, dim=-1, sorted=False)[1]
            group_mask = torch.zeros_like(group_scores)
            group_mask.scatter_(1, group_idx, 1)
            score_mask = group_mask.unsqueeze(-1).expand(n, self.n_group, self.n_routed_experts // self.n_group).reshape(n, -1)
            tmp = scores.masked_fill(~score_mask.bool(), 0.0)
            topk_weight, topk_idx = torch.topk(tmp, k=self.top_k, dim=-1, sorted=False)
        else:
            raise NotImplementedError(f"insupportable TopK function for MoE gating: {self.topk_method}")
        if self.top_k > 1 and self.norm_topk_prob:
            topk_weight = topk_weight / (topk_weight.sum(dim=-1, keepdim=True) + 1e-20)
        else:
            topk_weight = topk_weight * self.routed_scaling_factor
        return topk_idx, topk_weight


class DeepseekV2MoE(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.num_experts_per_tok = config.num_experts_per_tok
        self.experts = nn.ModuleList([DeepseekV2MLP(config, intermediate_size=config.moe_intermediate_size)
                                      for _ in range(config.n_routed_experts)])
        self.gate = MoEGate(config)
        if config.n_shared_experts is not None:
            self.shared_experts = DeepseekV2MLP(config, intermediate_size=config.moe_intermediate_size * config.n_shared_experts)

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
