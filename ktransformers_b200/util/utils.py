"""Module surgery and weight-loading recursion.  Mirrors archive/ktransformers/util/utils.py
(get_module :130, set_module :138, set_param :147, get_device :159, load_cur_state_dict :263-308,
load_weights :335-342, InferenceState :809)."""
from __future__ import annotations

import enum
import itertools

import torch
from torch import nn


class InferenceState(enum.Enum):
    UNLOAD = 0
    PREFILL = 1
    GENERATE = 2
    RESTORE = 3


def get_module_from_name(module: nn.Module, name: str) -> nn.Module:
    for part in [p for p in name.split(".") if p]:
        module = getattr(module, part)
    return module


def set_module(model: nn.Module, submodule_key: str, module: nn.Module) -> None:
    tokens = submodule_key.split(".")
    cur = model
    for s in tokens[:-1]:
        cur = getattr(cur, s)
    setattr(cur, tokens[-1], module)


def set_param(module: nn.Module, name: str, weights: torch.Tensor) -> None:
    param = nn.parameter.Parameter(weights, requires_grad=False)
    if isinstance(module, nn.Linear) and len(weights.shape) == 1:
        param.unsqueeze_(0)
    setattr(module, name, param)


def get_device(gguf_module_key: str, device_map: dict) -> str:
    if gguf_module_key in device_map:
        return device_map[gguf_module_key]["generate_device"]
    return "cuda"


def get_all_used_cuda_device(device_map: dict) -> list:
    devs = set()
    for v in device_map.values():
        for k in ("generate_device", "prefill_device"):
            if k in v and "cpu" not in str(v[k]).lower():
                devs.add(v[k])
    return sorted(devs)


def load_cur_state_dict(module: nn.Module, gguf_loader, prefix: str = "", device: str = "cuda") -> None:
    prefix = prefix.replace("orig_module.", "")
    persistent = {k: v for k, v in module._buffers.items() if k not in module._non_persistent_buffers_set}
    for name, param in itertools.chain(module._parameters.items(), persistent.items()):
        if param is None:
            continue
        key = prefix + name
        if gguf_loader.has_tensor(key) or "kv_b_proj" in key:
            target_dtype = torch.get_default_dtype()
            dev = get_device(key[: key.rfind(".")], gguf_loader.tensor_device_map)
            if "kv_b_proj" in key and not gguf_loader.has_tensor(key):
                # newer GGUFs split kv_b into attn_k_b / attn_v_b (utils.py:288-296)
                k_b = gguf_loader.load_gguf_tensor(key.replace("self_attn.kv_b_proj", "attn_k_b"), device=dev).to(target_dtype)
                k_b = k_b.transpose(1, 2).contiguous()
                v_b = gguf_loader.load_gguf_tensor(key.replace("self_attn.kv_b_proj", "attn_v_b"), device=dev).to(target_dtype)
                kv_b = torch.cat((k_b, v_b), dim=1)
                kv_b = kv_b.contiguous() if kv_b.ndim == 2 else kv_b.flatten(0, 1).contiguous()
                set_param(module, name, kv_b)
            else:
                set_param(module, name, gguf_loader.load_gguf_tensor(key, device=dev).to(target_dtype))
        else:
            raise Exception(f"can't find {key} in GGUF file!")


def load_weights(module: nn.Module, gguf_loader, prefix: str = "", device: str = "cuda") -> None:
    from ..operators import base_operator
    if not isinstance(module, base_operator.BaseInjectedModule):
        load_cur_state_dict(module, gguf_loader, prefix, device=device)
        for name, child in module._modules.items():
            if child is not None:
                load_weights(child, gguf_loader, prefix + name + ".", device=device)
    else:
        module.load()
