"""nn.Module plumbing: parameters registered under the reference's dotted names as fp32 views of the
ParamStore master buffer, so state_dict()/load_state_dict() interoperate with reference checkpoints, plus the
HF-shaped surface SURVEY.md §8b lists (save_pretrained / from_pretrained, gradient-checkpointing switches, and the
`model.model.<property>` accessors the reference's optimizer grouping reads, dexbotic_arch.py:125-155)."""
from __future__ import annotations

import json
from pathlib import Path

import torch
import torch.nn as nn

from ..params import ParamSpec, ParamStore


class _ParamNamespace(nn.Module):
    """A node of the parameter tree (`model`, `model.llm`, ...).  The top node `model.model` additionally answers the
    reference's DexboticVLMModel properties (backbone, mm_projector_prefix, action_head_module, initialize_model ...)
    by forwarding unknown attributes to the engine object that executes the kernels."""

    def __getattr__(self, name: str):
        try:
            return super().__getattr__(name)
        except AttributeError:
            engine = self.__dict__.get("_engine")
            if engine is not None and not name.startswith("__"):
                return getattr(engine, name)
            raise


def _config_to_dict(cfg) -> dict:
    out = {}
    for k, v in vars(cfg).items():
        if k.startswith("_"):
            continue
        if hasattr(v, "to_dict"):
            v = v.to_dict()
        try:
            json.dumps(v)
        except TypeError:
            continue
        out[k] = v
    out["model_type"] = getattr(cfg, "model_type", None)
    return out


class B200Module(nn.Module):
    supports_gradient_checkpointing = True      # dexbotic_arch.py:40
    base_model_prefix = "model"

    def _materialize(self, specs: list[ParamSpec], device) -> ParamStore:
        store = ParamStore(specs, device)
        for sp in specs:
            p = nn.Parameter(store.master_view(sp.name), requires_grad=sp.trainable)
            mod = self
            *path, leaf = sp.name.split(".")
            for part in path:
                if part not in mod._modules:
                    mod.add_module(part, _ParamNamespace())
                mod = mod._modules[part]
            mod.register_parameter(leaf, p)
        self.store = store
        return store

    def __setattr__(self, name, value):
        # `model_engine` is a plain object (not an nn.Module): once it exists, `model.model` forwards to it
        super().__setattr__(name, value)
        if name == "model_engine" and "model" in self._modules:
            self._modules["model"].__dict__["_engine"] = value

    # ------------------------------------------------------------------ dtype / device (reference: torch_dtype=bf16)
    @property
    def dtype(self) -> torch.dtype:
        """The compute dtype, as the reference's `model.dtype` reads it after `from_pretrained(torch_dtype=bfloat16)`
        (cogact_exp.py:145-177 casts the image tensor with it).  Storage stays fp32 master + bf16 shadow."""
        return torch.bfloat16

    @property
    def device(self) -> torch.device:
        return self.store.device

    def _apply(self, fn, recurse: bool = True):
        """`model.to(torch.bfloat16)`, `.half()`, `.float()`, `.cuda(1)` would rebind every Parameter to a fresh tensor
        that is no longer a view of the ParamStore master buffer: the kernels would keep training the store while
        state_dict() / save_pretrained() serialise stale copies.  Dtype requests are already satisfied (bf16 compute
        copies are maintained by the store), so a conversion that changes nothing is a no-op and anything else is an
        error instead of a silent detach."""
        store = self.__dict__.get("store")
        if store is None:
            return super()._apply(fn, recurse)
        probe = torch.empty(0, device=store.device, dtype=torch.float32)
        res = fn(probe)
        if res.device != probe.device:
            raise RuntimeError(
                f"{type(self).__name__}: moving the model ({probe.device} -> {res.device}) would detach the parameters "
                "from the flat ParamStore buffers; construct it with device=... instead")
        return self          # dtype casts: the master stays fp32, the kernels read the store's bf16 shadow

    def to(self, *args, **kwargs):
        return self._apply(lambda t: t.to(*args, **kwargs))

    def state_dict(self, *args, **kwargs):
        self.store.wait_all_params()           # an overlapped optimizer step may still be writing the master buffer
        if self.store.sharder is not None:     # ZeRO-1: every rank holds the current fp32 master of its pieces only
            self.store.sharder.gather_master()
        return super().state_dict(*args, **kwargs)

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        self.store.wait_all_params()
        res = super().load_state_dict(state_dict, strict=strict, assign=False)
        self.store.refresh_shadow()
        self._after_weights_changed()
        return res

    def _after_weights_changed(self) -> None:
        pass

    # ------------------------------------------------------------------ gradient checkpointing (base_exp.py:245)
    def _decoders(self):
        eng = getattr(self, "model_engine", None)
        return [m for m in (getattr(eng, "llm", None), getattr(eng, "mm_vision_tower", None)) if m is not None
                and hasattr(m, "keep_layers")]

    def gradient_checkpointing_enable(self, gradient_checkpointing_kwargs=None):
        """Every block keeps only its input and is recomputed in backward (the reference's default trade)."""
        for m in self._decoders():
            m.keep_layers = 0

    def gradient_checkpointing_disable(self):
        """Keep intermediates for as many blocks as HBM allows (decided at the next forward)."""
        for m in self._decoders():
            m.keep_layers = None

    # ------------------------------------------------------------------ HF-style checkpoint directory
    def save_pretrained(self, save_directory, safe_serialization: bool = True, dtype: torch.dtype = None,
                        max_shard_size: int = 5 << 30):
        """config.json + model(-0000N-of-0000M).safetensors (+ index) with the reference's parameter names, so the
        directory loads with the reference's `from_pretrained` and vice versa (trainer.py:145-189 saves the same)."""
        from safetensors.torch import save_file
        d = Path(save_directory)
        d.mkdir(parents=True, exist_ok=True)
        (d / "config.json").write_text(json.dumps(_config_to_dict(self.config), indent=2, default=str))
        sd = {k: (v.detach().to(dtype) if dtype is not None else v.detach()).cpu().contiguous()
              for k, v in self.state_dict().items()}
        if not safe_serialization:
            torch.save(sd, d / "pytorch_model.bin")
            return
        shards, cur, size = [], {}, 0
        for k, v in sd.items():
            nb = v.numel() * v.element_size()
            if cur and size + nb > max_shard_size:
                shards.append(cur)
                cur, size = {}, 0
            cur[k] = v
            size += nb
        shards.append(cur)
        if len(shards) == 1:
            save_file(shards[0], str(d / "model.safetensors"), metadata={"format": "pt"})
            return
        weight_map = {}
        for i, sh in enumerate(shards):
            name = f"model-{i + 1:05d}-of-{len(shards):05d}.safetensors"
            save_file(sh, str(d / name), metadata={"format": "pt"})
            weight_map.update({k: name for k in sh})
        total = sum(v.numel() * v.element_size() for v in sd.values())
        (d / "model.safetensors.index.json").write_text(json.dumps({"metadata": {"total_size": total},
                                                                    "weight_map": weight_map}, indent=2))

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, device="cuda", strict: bool = True, **config_overrides):
        """Build from a checkpoint directory written by this class or by the reference (HF layout)."""
        d = Path(pretrained_model_name_or_path)
        cfg_dict = json.loads((d / "config.json").read_text())
        cfg_dict.pop("model_type", None)
        # loader arguments HF accepts but that do not describe the model (the reference passes torch_dtype=bfloat16)
        for k in ("torch_dtype", "dtype", "device_map", "low_cpu_mem_usage", "attn_implementation", "trust_remote_code",
                  "use_safetensors", "cache_dir", "local_files_only", "revision", "token"):
            config_overrides.pop(k, None)
        cfg_dict.pop("torch_dtype", None)
        cfg_dict.update(config_overrides)
        config = cls.config_class(**{k: v for k, v in cfg_dict.items() if k not in ("architectures", "transformers_version")})
        model = cls(config, device=device)
        sd = {}
        index = d / "model.safetensors.index.json"
        if index.exists():
            from safetensors.torch import load_file
            for name in sorted(set(json.loads(index.read_text())["weight_map"].values())):
                sd.update(load_file(str(d / name)))
        elif (d / "model.safetensors").exists():
            from safetensors.torch import load_file
            sd = load_file(str(d / "model.safetensors"))
        else:
            sd = torch.load(d / "pytorch_model.bin", map_location="cpu", weights_only=True)
        own = set(model.state_dict().keys())
        sd = {k: v for k, v in sd.items() if k in own or strict}       # HF buffers (position_ids) are not parameters
        sd = {k: v for k, v in sd.items() if not k.endswith("position_ids")}
        model.load_state_dict(sd, strict=strict)
        return model

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
