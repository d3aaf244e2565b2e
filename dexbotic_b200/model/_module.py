"""nn.Module plumbing: parameters registered under the reference's dotted names as fp32 views of the
ParamStore master buffer, so state_dict()/load_state_dict() interoperate with reference checkpoints."""
from __future__ import annotations

import torch
import torch.nn as nn

from ..params import ParamSpec, ParamStore


class B200Module(nn.Module):
    def _materialize(self, specs: list[ParamSpec], device) -> ParamStore:
        store = ParamStore(specs, device)
        for sp in specs:
            p = nn.Parameter(store.master_view(sp.name), requires_grad=sp.trainable)
            mod = self
            *path, leaf = sp.name.split(".")
            for part in path:
                if part not in mod._modules:
                    mod.add_module(part, nn.Module())
                mod = mod._modules[part]
            mod.register_parameter(leaf, p)
        self.store = store
        return store

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        res = super().load_state_dict(state_dict, strict=strict, assign=False)
        self.store.refresh_shadow()
        self._after_weights_changed()
        return res

    def _after_weights_changed(self) -> None:
        pass

    @torch.no_grad()
    def init_weights_(self, seed: int = 0, std: float = 0.02) -> None:
        """Seeded N(0, std) init on device (norm scales 1, biases 0) — synthetic weights for bench/smoke."""
        g = torch.Generator(device=self.store.device).manual_seed(seed)
        for name in self.store.order:
            v = self.store.master_view(name)
            leaf = name.rsplit(".", 1)[-1]
            if v.dim() == 1 and leaf == "weight":
                v.fill_(1.0)
            elif leaf == "bias":
                v.zero_()
            else:
                v.normal_(0.0, std, generator=g)
        self.store.refresh_shadow()
        self._after_weights_changed()
