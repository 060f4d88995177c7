"""BaseInjectedModule — attribute-forwarding proxy around the module it replaces.
Same contract as archive/ktransformers/operators/base_operator.py:12-63: the ctor signature
``(key, gguf_loader, config, orig_module, prefill_device, generate_device, **kwargs)``, attributes not
found on the wrapper resolve on ``orig_module``, and ``load()`` recurses into children."""
from __future__ import annotations

from typing import Any

from torch import nn


class BaseInjectedModule(nn.Module):
    def __init__(self, key, gguf_loader, config, orig_module, prefill_device: str = "cuda",
                 generate_device: str = "cuda", **kwargs):
        nn.Module.__init__(self)
        nn.Module.__setattr__(self, "orig_module", orig_module)
        for name, value in (("key", key), ("gguf_loader", gguf_loader), ("config", config),
                            ("prefill_device", prefill_device), ("generate_device", generate_device),
                            ("device", generate_device)):
            object.__setattr__(self, name, value)

    def __getattr__(self, name: str) -> Any:
        try:
            return object.__getattribute__(self, name)
        except AttributeError:
            pass
        orig = nn.Module.__getattr__(self, "orig_module")
        if name == "orig_module":
            return orig
        try:
            return orig.__getattr__(name)                       # parameters / buffers / submodules of orig
        except AttributeError:
            return object.__getattribute__(orig, name)          # plain attributes of orig

    def __setattr__(self, name: str, value) -> None:
        if name == "orig_module":
            return nn.Module.__setattr__(self, "orig_module", value)
        if hasattr(self, name):
            return object.__setattr__(self, name, value)
        return nn.Module.__getattr__(self, "orig_module").__setattr__(name, value)

    def forward(self, *args, **kwargs):
        return self.orig_module.forward(*args, **kwargs)

    def load(self):
        from ..util import utils
        for name, child in self._modules.items():
            utils.load_weights(child, self.gguf_loader, self.key + ".")
