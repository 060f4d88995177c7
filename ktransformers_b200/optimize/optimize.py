"""YAML-rule driven module injection.  Same grammar and behaviour as
archive/ktransformers/optimize/optimize.py (inject :28-54, gen_optimize_config :67-118,
optimize_and_load_gguf :129-163):

    - match:   {name: <regex on module path>, class: <import path, isinstance>}   (either or both)
      replace: {class: <import path> | default, kwargs: {...}}
      recursive: bool        # False stops descending below a matched module

The first matching rule wins for each module; unmatched modules get a "default" entry carrying the
default devices.  ``inject`` instantiates the replacement as
``cls(key=, gguf_loader=, config=, orig_module=child, **kwargs)`` and records the kwargs in
``gguf_loader.tensor_device_map``.
"""
from __future__ import annotations

import copy
import importlib
import itertools
import re
from typing import List, Mapping

import torch
import yaml
from torch import nn

from ..util import utils
from ..util.custom_loader import ModelLoaderFactory
from ..util.utils import load_weights, set_module


def _import_cls(path: str):
    mod, _, cls = path.rpartition(".")
    return getattr(importlib.import_module(mod), cls)


def inject(module, local_optimization_dict, model_config, gguf_loader, prefix=""):
    for name, child in module._modules.items():
        if child is None:
            continue
        child_prefix = prefix + name
        if child_prefix not in local_optimization_dict:
            continue
        meta = local_optimization_dict[child_prefix]
        gguf_loader.tensor_device_map[meta["key"]] = meta["kwargs"] if "kwargs" in meta else dict()
        if meta["class"] != "default":
            module_cls = _import_cls(meta["class"])
            print(f"Injecting {child_prefix} as", meta["class"])
            new = module_cls(key=meta["key"], gguf_loader=gguf_loader, config=model_config, orig_module=child, **meta["kwargs"])
            set_module(module, name, new)
        child_prefix += "."
        sub = {k: v for k, v in local_optimization_dict.items() if k.startswith(child_prefix)}
        inject(child, sub, model_config, gguf_loader, child_prefix)


def del_meta(module: nn.Module):
    persistent = {k: v for k, v in module._buffers.items() if k not in module._non_persistent_buffers_set}
    for name, p in list(itertools.chain(module._parameters.items(), persistent.items())):
        if p is not None and p.device == torch.device("meta"):
            module.__delattr__(name)
    for child in module._modules.values():
        if child is not None:
            del_meta(child)


def gen_optimize_config(module: nn.Module, out_data: Mapping, rule_list: List, prefix: str = "", default_device: str = "cuda:0"):
    module_name = prefix[:-1]
    recursive = True
    for rule in rule_list:
        match = rule["match"]
        if "class" not in match and "name" not in match:
            raise Exception("match must have at least one of \"class\" and \"name\"")
        if "class" in match and not isinstance(module, _import_cls(match["class"])):
            continue
        if "name" in match and re.search(match["name"], module_name) is None:
            continue
        if "replace" not in rule:
            raise Exception("replace must be in rule")
        rep = rule["replace"]
        kwargs = copy.deepcopy(rep["kwargs"]) if "kwargs" in rep else dict()
        if module_name not in out_data:
            out_data[module_name] = {"key": module_name, "class": rep.get("class", "default"), "kwargs": kwargs}
        else:
            if out_data[module_name]["class"] == "default":
                out_data[module_name]["class"] = rep.get("class", "default")
            out_data[module_name]["kwargs"].update(kwargs)
        if "recursive" in rule:
            recursive = bool(rule["recursive"])
        break
    if module_name not in out_data:
        out_data[module_name] = {"class": "default", "key": module_name,
                                 "kwargs": {"generate_device": default_device, "prefill_device": default_device}}
    if recursive:
        for name, child in module._modules.items():
            if child is not None:
                gen_optimize_config(child, out_data, rule_list, prefix + name + ".", default_device=default_device)


def translate_model_config(model_config):
    if getattr(model_config, "model_type", None) == "mixtral":
        model_config.moe_intermediate_size = model_config.intermediate_size
    return model_config


def optimize_and_load_gguf(module: nn.Module, rule_file: str, gguf_path: str, model_config, default_device: str = "cuda:0"):
    with open(rule_file, "r", encoding="utf-8") as f:
        rule_list = yaml.load(f.read(), Loader=yaml.FullLoader)
    optimize_config = dict()
    gen_optimize_config(module, optimize_config, rule_list, default_device=default_device)
    model_config = translate_model_config(model_config)
    loader = ModelLoaderFactory.create_loader(gguf_path)
    with torch.device("meta"):
        inject(module, optimize_config, model_config, loader)
    if hasattr(module, "lm_head"):
        load_weights(module.lm_head, loader, "lm_head.", device=default_device)   # pre-load (optimize.py:156)
    load_weights(module, loader, device=default_device)
    module.gguf_loader = loader
    del_meta(module)
    if torch.cuda.is_available():
        torch.cuda.empty_cache()
    return optimize_config
