"""Expert-parallel dispatch/combine for decode (spec: the HF EP branch of DeepseekV3MoE.moe_infer,
archive/ktransformers/models/modeling_deepseek_v3.py:550-605, minus its host round trips).

GPU g owns experts [g*E/N, (g+1)*E/N).  Every rank holds one (or a few) decode tokens.  Per MoE layer:

  dispatch  all-gather of the ranks' tokens, expert ids and routing weights (N x 14 KB at bs=1 per rank:
            latency-bound, so no count exchange and no variable-size all-to-all — every rank sees all
            (token, expert) pairs and keeps the ones it owns: ids outside its shard are skipped by the kernel,
            like the reference's gpu_experts_mask / should_skip_expert, kt-kernel/operators/common.hpp:255-258)
  compute   local experts on the gathered tokens, fp32 partial sums
  combine   reduce-scatter (sum) of the [N*t, H] fp32 partials: each rank receives its own tokens' totals, which
            are rounded to the hidden dtype once — the same sum the single-GPU path forms.

All buffers are static so the whole layer is CUDA-graph capturable with NCCL.
"""
from __future__ import annotations

from typing import Callable

import torch
import torch.distributed as dist


def shard_range(n_experts: int, rank: int, world: int) -> tuple[int, int]:
    assert n_experts % world == 0, "expert count must divide evenly across the EP group"
    per = n_experts // world
    return rank * per, (rank + 1) * per


class ExpertParallelCombine:
    """local_forward(x_all[T,H] (fp32), ids_all[T,k], w_all[T,k], out_partial[T,H] fp32) computes this rank's share."""

    def __init__(self, hidden_size: int, top_k: int, tokens_per_rank: int, device, group=None, in_dtype=torch.bfloat16):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.t = tokens_per_rank
        T = self.world * tokens_per_rank
        self.x_all = torch.zeros((T, hidden_size), dtype=in_dtype, device=device)
        self.x_all_f32 = torch.zeros((T, hidden_size), dtype=torch.float32, device=device)
        self.ids_all = torch.zeros((T, top_k), dtype=torch.int64, device=device)
        self.w_all = torch.zeros((T, top_k), dtype=torch.float32, device=device)
        self.partial = torch.zeros((T, hidden_size), dtype=torch.float32, device=device)
        self.own = torch.zeros((tokens_per_rank, hidden_size), dtype=torch.float32, device=device)
        self._nccl = dist.is_initialized() and dist.get_backend(group) == "nccl"

    def dispatch(self, x: torch.Tensor, ids: torch.Tensor, w: torch.Tensor):
        if self.world == 1:
            self.x_all.copy_(x); self.ids_all.copy_(ids); self.w_all.copy_(w)
        else:
            dist.all_gather_into_tensor(self.x_all, x.contiguous(), group=self.group)
            dist.all_gather_into_tensor(self.ids_all, ids.contiguous(), group=self.group)
            dist.all_gather_into_tensor(self.w_all, w.contiguous(), group=self.group)
        self.x_all_f32.copy_(self.x_all)
        return self.x_all_f32, self.ids_all, self.w_all

    def combine(self) -> torch.Tensor:
        if self.world == 1:
            self.own.copy_(self.partial)
        elif self._nccl:
            dist.reduce_scatter_tensor(self.own, self.partial, group=self.group)
        else:  # gloo has no reduce_scatter: all-reduce and keep the own rows (CPU tests)
            dist.all_reduce(self.partial, group=self.group)
            self.own.copy_(self.partial[self.rank * self.t:(self.rank + 1) * self.t])
        return self.own

    def forward(self, x, ids, w, local_forward: Callable) -> torch.Tensor:
        xa, ia, wa = self.dispatch(x, ids, w)
        local_forward(xa, ia, wa, self.partial)
        return self.combine()
