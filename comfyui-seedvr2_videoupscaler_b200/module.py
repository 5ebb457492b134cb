"""Model-slot base class: the engines are ``nn.Module``s because the reference pipeline treats its slots as such.

What the reference does to ``runner.dit`` / ``runner.vae`` (SURVEY.md §8(b) "Lifecycle the slots must survive"):
``next(model.parameters()).device / .dtype`` sniffing (``generation_phases.py:620,708-712``, ``infer.py:125,152``),
``model.to(device)`` + ``zero_grad`` (``manage_model_device``, ``memory_manager.py:670-738``), ``named_modules()`` walks
(``clear_rope_lru_caches`` ``:427-455``; ``apply_model_specific_config`` matching ``'FlashAttentionVarlen'`` by class name,
``model_configuration.py:1190-1213``), ``requires_grad_(False).eval()``, and at the end of the phase
``release_model_memory`` (``:544-581``: ``param.data.set_()`` / ``buffer.data.set_()`` on everything on the GPU).

Engine weights are kept in the kernels' compute layout (bf16 K-major matrices, fp32 modulation vectors), registered as
non-persistent buffers so all of the above reaches them; one frozen bf16 parameter answers the device / dtype sniffing.
Dtype casts are refused (the layout is the kernels' contract); device moves are honoured, and ``forward`` raises when
the weights are not on a CUDA device — there is no CPU path.
"""
from __future__ import annotations

from typing import Dict

import torch
from torch import nn


class BufferView:
    """dict-like read access to a module's buffers under their original (dotted) names; aliases share one buffer."""

    def __init__(self, module: nn.Module, names: Dict[str, str]):
        self._m, self._names = module, names

    def __getitem__(self, key: str) -> torch.Tensor:
        return self._m._buffers[self._names[key]]

    def __contains__(self, key: str) -> bool:
        return key in self._names

    def get(self, key: str, default=None):
        return self[key] if key in self._names else default

    def keys(self):
        return self._names.keys()

    def values(self):
        return (self[k] for k in self._names)


class EngineModule(nn.Module):
    def __init__(self, device="cuda"):
        super().__init__()
        self.probe = nn.Parameter(torch.zeros(1, dtype=torch.bfloat16, device=device), requires_grad=False)
        self._n_buffers = 0

    @property
    def device(self) -> torch.device:
        return self.probe.device

    def _register(self, group: str, tensors: Dict[str, torch.Tensor]) -> BufferView:
        """Register ``tensors`` as non-persistent buffers; entries that are the same tensor object share a buffer."""
        names: Dict[str, str] = {}
        seen: Dict[int, str] = {}
        for key, t in tensors.items():
            if id(t) in seen:
                names[key] = seen[id(t)]
                continue
            name = f"{group}{self._n_buffers}"
            self._n_buffers += 1
            self.register_buffer(name, t, persistent=False)
            names[key] = seen[id(t)] = name
        return BufferView(self, names)

    # ---- dtype casts would break the kernels' layout contract: device moves only
    def to(self, *args, **kwargs):
        device, _dtype, non_blocking, _fmt = torch._C._nn._parse_to(*args, **kwargs)
        if device is None:
            return self
        return super().to(device=device, non_blocking=non_blocking)

    def float(self):
        return self

    def half(self):
        return self

    def bfloat16(self):
        return self

    def double(self):
        return self

    def _apply(self, fn, recurse=True):
        out = super()._apply(fn, recurse)
        self._device_state_moved()
        return out

    def _device_state_moved(self) -> None:
        """Hook: drop cached device-side tables after the weights moved."""

    def _require_cuda(self, what: str) -> None:
        from . import lib
        if self.probe.device.type != "cuda" or self.probe.numel() == 0:
            raise lib.Svr2Error(f"{what}: engine weights are on '{self.probe.device}' (offloaded or released) — move the "
                                "module back to a CUDA device; there is no CPU path")
