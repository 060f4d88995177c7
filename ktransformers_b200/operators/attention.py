"""KDeepseekV2Attention — absorbed multi-head latent attention for decode (V3 MLA is the same as V2).

Mirrors archive/ktransformers/operators/attention.py:49-75 (`get_absorbed`) and :349-478 (`forward_linux_flashinfer`,
decode branch): q projections -> RoPE -> paged latent-cache update -> q_nope . W_UK (batched matmul) -> MLA paged decode
over the 576-wide latents -> . W_UV^T -> o_proj.  The attention itself is `MLAWrapper.run` = ktb200_mla_decode (tcgen05 +
TMEM + TMA, csrc/mla.cu); the cache write is ktb200_mla_kv_write; the projections are whatever modules the rules injected
(KLinearB200 on raw GGUF blocks, or nn.Linear); the two absorb products are ktb200_mla_absorb_q / _o (HBM-bound batched
GEMVs over the bf16 halves of kv_b_proj; torch.matmul for other dtypes).  Prefill (q_len > 1 without absorb) is outside this path and raises."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .. import native
from ..models.modeling_deepseek_v3 import DeepseekV3Attention, apply_rotary_pos_emb
from .base_operator import BaseInjectedModule
from .flashinfer_wrapper import MLAWrapperSingleton


class KDeepseekV2Attention(BaseInjectedModule, DeepseekV3Attention):
    def __init__(self, key, gguf_loader, config, orig_module, prefill_device: str = "cuda", generate_device: str = "cuda",
                 chunck_size: int = 1000, absorb_for_prefill: bool = False, **kwargs):
        BaseInjectedModule.__init__(self, key, gguf_loader, config, orig_module, prefill_device, generate_device, **kwargs)
        self.chunck_size = chunck_size
        self.mla_wrapper = None
        self.absorb_for_prefill = absorb_for_prefill

    def get_absorbed(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """kv_b_proj [heads * (128 + 128), 512] viewed per head: q_absorb = W_UK [h, 128, 512], out_absorb = W_UV [h, 128, 512]
        (attention.py:69-75); kv_b_proj is the one projection the rules keep dense."""
        if not (hasattr(self, "q_absorb") and hasattr(self, "out_absorb")):
            kv_b = self.kv_b_proj.weight.view(self.num_heads, -1, self.kv_lora_rank)
            object.__setattr__(self, "q_absorb", kv_b[:, : self.qk_nope_head_dim, :].contiguous())
            object.__setattr__(self, "out_absorb", kv_b[:, self.qk_nope_head_dim:, :].contiguous())
        return self.q_absorb, self.out_absorb

    def forward(self, hidden_states: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                position_ids: Optional[torch.Tensor] = None, past_key_value=None, output_attentions: bool = False,
                use_cache: bool = False, cache_position: Optional[torch.Tensor] = None, **kwargs):
        bsz, q_len, _ = hidden_states.size()
        if q_len != 1 and not self.absorb_for_prefill:
            raise NotImplementedError("KDeepseekV2Attention: the B200 path covers absorbed decode (q_len == 1)")
        assert past_key_value is not None, "decode needs the paged latent cache (models/custom_cache.StaticCache)"
        q = self.q_proj(hidden_states) if self.q_lora_rank is None else self.q_b_proj(self.q_a_layernorm(self.q_a_proj(hidden_states)))
        q = q.view(bsz, q_len, self.num_heads, self.q_head_dim)
        q_nope, q_pe = torch.split(q, [self.qk_nope_head_dim, self.qk_rope_head_dim], dim=-1)
        compressed_kv = self.kv_a_proj_with_mqa(hidden_states)
        compressed_kv, k_pe = torch.split(compressed_kv, [self.kv_lora_rank, self.qk_rope_head_dim], dim=-1)
        compressed_kv = self.kv_a_layernorm(compressed_kv).view(bsz, q_len, 1, self.kv_lora_rank)
        k_pe = k_pe.view(bsz, q_len, 1, self.qk_rope_head_dim)
        cos, sin = self.rotary_emb(q_pe, position_ids)
        q_pe, k_pe = apply_rotary_pos_emb(q_pe, k_pe, cos, sin, unsqueeze_dim=2)

        cache_kwargs = {"sin": sin, "cos": cos, "cache_position": cache_position}
        kv_with_k_pe, page_table = past_key_value.update(compressed_kv, k_pe, self.layer_idx, cache_kwargs)
        ckv_pages = kv_with_k_pe[:, :, :, : self.kv_lora_rank].view(-1, past_key_value.page_size, self.kv_lora_rank)
        kpe_pages = kv_with_k_pe[:, :, :, self.kv_lora_rank:].view(-1, past_key_value.page_size, self.qk_rope_head_dim)

        q_absorb, out_absorb = self.get_absorbed()
        fused = (q.dtype == torch.bfloat16 and q_absorb.dtype == torch.bfloat16 and q.is_contiguous())
        stream = torch.cuda.current_stream(hidden_states.device).cuda_stream
        if fused:   # q_nope . W_UK straight from the q_b output (no slice copy): ktb200_mla_absorb_q
            q_abs = torch.empty((bsz * q_len, self.num_heads, self.kv_lora_rank), dtype=q.dtype, device=q.device)
            native.check(native.lib().ktb200_mla_absorb_q(q.data_ptr(), self.q_head_dim, self.num_heads * self.q_head_dim, q_absorb.data_ptr(), self.num_heads,
                                                          self.qk_nope_head_dim, self.kv_lora_rank, q_abs.data_ptr(), bsz * q_len, stream))
            q_nope = q_abs
        else:
            q_nope = torch.matmul(q_nope.transpose(1, 2), q_absorb).transpose(1, 2).contiguous().reshape(bsz * q_len, self.num_heads, self.kv_lora_rank)
        q_pe = q_pe.reshape(bsz * q_len, self.num_heads, self.qk_rope_head_dim)

        if self.mla_wrapper is None:
            self.mla_wrapper = MLAWrapperSingleton.get_instance(str(hidden_states.device), bsz, past_key_value.max_pages * bsz, use_cuda_graph=True)
        w = self.mla_wrapper
        if w.need_plan:
            # decode: one query per sequence, kv length = position + 1, identity page table of the static cache
            kv_len = (position_ids.reshape(bsz, -1)[:, -1] + 1).to(torch.int32)
            pages = past_key_value.max_pages
            indptr = torch.arange(0, bsz + 1, dtype=torch.int32, device=hidden_states.device) * pages
            w.plan(None, indptr, page_table.reshape(-1), kv_len, None, self.num_heads, self.kv_lora_rank, self.qk_rope_head_dim,
                   past_key_value.page_size, self.softmax_scale, q_nope.dtype, ckv_pages.dtype)
            w.max_pages_per_seq = pages
        else:   # the plan is static (identity page table); only the lengths move from step to step (a captured device copy)
            w.kv_len_arr_buf[:bsz].copy_((position_ids.reshape(bsz, -1)[:, -1] + 1).to(torch.int32))
        attn = w.run(q_nope, q_pe.contiguous(), ckv_pages, kpe_pages).view(bsz, q_len, self.num_heads, self.kv_lora_rank)
        if fused:
            o = torch.empty((bsz, q_len, self.num_heads, self.v_head_dim), dtype=attn.dtype, device=attn.device)
            native.check(native.lib().ktb200_mla_absorb_o(attn.data_ptr(), out_absorb.data_ptr(), self.num_heads, self.v_head_dim, self.kv_lora_rank, o.data_ptr(),
                                                          bsz * q_len, stream))
            attn = o
        else:
            attn = torch.matmul(attn.transpose(1, 2), out_absorb.mT).transpose(1, 2).contiguous()     # [b, 1, h, 128]
        attn = self.o_proj(attn.reshape(bsz, q_len, self.num_heads * self.v_head_dim))
        return attn, None, past_key_value
