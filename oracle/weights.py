"""TEST INFRASTRUCTURE — deterministic synthetic weights shared by the oracle, the golden-vector
generator and the parity tests: every tensor is drawn from its own generator seeded by (seed, name), so
the values do not depend on parameter ordering, and no zero-initialised layer survives (SURVEY.md §7
"zero-init trap": DiT final_layer, LayerNorm biases ...)."""
from __future__ import annotations

import zlib

import torch


def _tensor_seed(seed: int, name: str) -> int:
    return (zlib.crc32(name.encode()) + 1000003 * seed) % (2 ** 31 - 1)


def seeded_tensor(name: str, shape, seed: int, dtype=torch.float32) -> torch.Tensor:
    g = torch.Generator().manual_seed(_tensor_seed(seed, name))
    shape = tuple(shape)
    if name.endswith("position_ids"):
        return torch.arange(shape[-1]).expand(shape).clone()
    if "inv_freq" in name:
        raise KeyError(name)
    t = torch.randn(shape, generator=g, dtype=torch.float32)
    last = name.rsplit(".", 1)[-1]
    is_norm_weight = last == "weight" and len(shape) == 1
    if is_norm_weight:                       # RMSNorm / LayerNorm scale: around 1
        t = 1.0 + 0.1 * t
    elif last == "bias" or len(shape) == 1:  # biases, class_embedding
        t = 0.05 * t
    elif len(shape) >= 2:
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        t = 0.05 * t if "embed" in name else t * (1.0 / max(fan_in, 1)) ** 0.5
    return t.to(dtype)


def seeded_state_dict(shapes: dict, seed: int, dtype=torch.float32) -> dict:
    out = {}
    for name, shape in shapes.items():
        if "inv_freq" in name:
            continue
        out[name] = seeded_tensor(name, shape, seed, dtype if "position_ids" not in name else torch.long)
    return out
