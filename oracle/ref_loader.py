"""TEST INFRASTRUCTURE — not part of the product path.

Loads the UNMODIFIED reference (dexmal/dexbotic, read-only at /root/reference) in this container so
that the oracle restatement (oracle/vla_oracle.py) can be validated against it and golden vectors can be
generated (oracle/make_golden.py).  /root/reference does not exist on the GPU box; nothing under
tests/ -m gpu, smoke() or bench.py's own arm imports this module (`bench.py --impl reference_gpu`, the reference's
PyTorch GPU baseline, loads the vendored copy under baseline/_ref/ through it).

Compatibility recipe, all in memory, /root/reference untouched:
  1. import transformers first (it probes find_spec('timm'));
  2. inject stub modules for the un-vendored, un-pinned third-party deps the reference imports:
     timm.models.vision_transformer.{Attention, Mlp} (restated from timm's published definition: fused
     qkv Linear, SDPA, proj; fc1/act/fc2) and diffusers' DDIMScheduler (oracle/ddim_oracle.py, restated);
  3. exec dexbotic_arch.py / pi0_arch.py with the un-defaulted dataclass fields given `= None` (transformers
     >= 5 turns PretrainedConfig subclasses into dataclasses);
  4. vision towers are built from config objects instead of from_pretrained (no network / weights).
"""
from __future__ import annotations

import importlib
import importlib.machinery
import sys
import types
from pathlib import Path

# /root/reference in the build container; on the GPU box the same unmodified tree vendored by
# tools/install_reference.sh into the git-ignored baseline/_ref/ (used by `bench.py --impl reference_gpu` only)
REFERENCE_ROOT = Path("/root/reference")
if not (REFERENCE_ROOT / "dexbotic" / "model" / "dexbotic_arch.py").exists():
    REFERENCE_ROOT = Path(__file__).resolve().parent.parent / "baseline" / "_ref"


def reference_available() -> bool:
    return (REFERENCE_ROOT / "dexbotic" / "model" / "dexbotic_arch.py").exists()


def _stub_module(name: str) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    m.__path__ = []  # behaves like a package
    sys.modules[name] = m
    return m


def _install_timm_stub():
    import torch
    import torch.nn as nn
    import torch.nn.functional as F

    if "timm" in sys.modules:
        return

    class Attention(nn.Module):
        """timm.models.vision_transformer.Attention (fused_attn path), restated."""

        def __init__(self, dim, num_heads=8, qkv_bias=False, qk_norm=False, attn_drop=0.0, proj_drop=0.0,
                     norm_layer=nn.LayerNorm, **_):
            super().__init__()
            assert dim % num_heads == 0 and not qk_norm
            self.num_heads = num_heads
            self.head_dim = dim // num_heads
            self.scale = self.head_dim ** -0.5
            self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
            self.proj = nn.Linear(dim, dim)

        def forward(self, x, attn_mask=None):
            B, N, C = x.shape
            qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
            q, k, v = qkv.unbind(0)
            x = F.scaled_dot_product_attention(q, k, v, attn_mask=attn_mask)
            x = x.transpose(1, 2).reshape(B, N, C)
            return self.proj(x)

    class Mlp(nn.Module):
        """timm.layers.Mlp, restated (drop=0)."""

        def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, norm_layer=None,
                     bias=True, drop=0.0, **_):
            super().__init__()
            out_features = out_features or in_features
            hidden_features = hidden_features or in_features
            self.fc1 = nn.Linear(in_features, hidden_features, bias=bias)
            self.act = act_layer()
            self.fc2 = nn.Linear(hidden_features, out_features, bias=bias)

        def forward(self, x):
            return self.fc2(self.act(self.fc1(x)))

    _stub_module("timm")
    _stub_module("timm.models")
    vt = _stub_module("timm.models.vision_transformer")
    vt.Attention, vt.Mlp = Attention, Mlp
    layers = _stub_module("timm.layers")
    layers.Mlp = Mlp


def _install_diffusers_stub():
    if "diffusers" in sys.modules:
        return

    # diffusers is absent: the reference's DiffusionActionHead runs around the restated scheduler (parity of the
    # scheduler itself is unpinned, see oracle/ddim_oracle.py)
    from oracle.ddim_oracle import DDIMSchedulerOracle as DDIMScheduler

    _stub_module("diffusers")
    _stub_module("diffusers.schedulers")
    m = _stub_module("diffusers.schedulers.scheduling_ddim")
    m.DDIMScheduler = DDIMScheduler
    sys.modules["diffusers"].DDIMScheduler = DDIMScheduler


def _exec_patched(modname: str, relpath: str, replacements: list[tuple[str, str]]):
    src = (REFERENCE_ROOT / relpath).read_text()
    for old, new in replacements:
        assert old in src, f"compat patch target not found in {relpath}: {old!r}"
        src = src.replace(old, new)
    mod = types.ModuleType(modname)
    mod.__file__ = str(REFERENCE_ROOT / relpath)
    mod.__spec__ = importlib.machinery.ModuleSpec(modname, loader=None)
    sys.modules[modname] = mod
    exec(compile(src, mod.__file__, "exec"), mod.__dict__)
    parent, _, child = modname.rpartition(".")
    setattr(importlib.import_module(parent), child, mod)
    return mod


_loaded = False


def load_reference():
    """Make `import dexbotic.model...` work here; returns the patched dexbotic_arch module."""
    global _loaded
    if not reference_available():
        raise RuntimeError("/root/reference is not present (expected on the GPU box): the oracle restatement "
                           "and the committed golden vectors are the checker there")
    import transformers  # noqa: F401  (must come first)
    if _loaded:
        return sys.modules["dexbotic.model.dexbotic_arch"]
    _install_timm_stub()
    _install_diffusers_stub()
    if str(REFERENCE_ROOT) not in sys.path:
        sys.path.insert(0, str(REFERENCE_ROOT))
    import dexbotic.model  # noqa: F401  (namespace/package import only)
    arch = _exec_patched("dexbotic.model.dexbotic_arch", "dexbotic/model/dexbotic_arch.py",
                         [("    llm_config: str | PretrainedConfig\n", "    llm_config: str | PretrainedConfig = None\n")])
    _patch_vision_towers()
    _loaded = True
    return arch


def _patch_vision_towers():
    """Build HF vision models from config objects (random init) instead of from_pretrained."""
    from transformers import CLIPVisionConfig, CLIPVisionModel
    from dexbotic.model.modules.mm_vision.clip import clip_encoder

    def clip_load_model(self):
        if self.is_loaded:
            return
        cfg = self.vision_tower_name
        assert isinstance(cfg, CLIPVisionConfig), "oracle builds CLIP towers from a CLIPVisionConfig"
        self.image_processor = None
        self.vision_tower = CLIPVisionModel(cfg)
        self.vision_tower.requires_grad_(False)
        self.is_loaded = True

    clip_encoder.CLIPVisionTower.load_model = clip_load_model

    from dexbotic.model.modules.mm_vision import builder as vbuilder

    orig_build = vbuilder.build_vision_tower

    def build_vision_tower(mm_vision_tower, **kwargs):
        if isinstance(mm_vision_tower, CLIPVisionConfig):
            return clip_encoder.CLIPVisionTower(mm_vision_tower)
        return orig_build(mm_vision_tower, **kwargs)

    vbuilder.build_vision_tower = build_vision_tower
    sys.modules["dexbotic.model.dexbotic_arch"].build_vision_tower = build_vision_tower


def build_reference_cogact(llm_config, clip_config, action_model_type: str = "DiT-S", action_dim: int = 7,
                           chunk_size: int = 16, mm_projector_type: str = "mlp2x_gelu"):
    """Reference CogACTForCausalLM (cogact_arch.py:47-54) with random-init weights."""
    load_reference()
    from dexbotic.model.cogact.cogact_arch import CogActConfig, CogACTForCausalLM
    cfg = CogActConfig(llm_config=llm_config, mm_projector_type=mm_projector_type, mm_vision_tower=clip_config,
                       action_model_type=action_model_type, action_dim=action_dim, chunk_size=chunk_size)
    return CogACTForCausalLM(cfg)


def build_reference_oft_discrete(llm_config, clip_config, action_dim: int = 7, chunk_size: int = 8, num_bins: int = 256,
                                 mm_projector_type: str = "mlp2x_gelu", use_proprio: bool = False, proprio_dim=None):
    """Reference OFTDiscreteForCausalLM (oft_discrete_arch.py:20-24) with random-init weights."""
    load_reference()
    from dexbotic.model.oft.oft_discrete_arch import OFTDiscreteConfig, OFTDiscreteForCausalLM
    cfg = OFTDiscreteConfig(llm_config=llm_config, mm_projector_type=mm_projector_type, mm_vision_tower=clip_config,
                            action_model_type="Discrete", action_dim=action_dim, chunk_size=chunk_size,
                            use_proprio=use_proprio, proprio_dim=proprio_dim, num_bins=num_bins)
    return OFTDiscreteForCausalLM(cfg)


def build_reference_oft_linear(llm_config, clip_config, action_dim: int = 7, chunk_size: int = 8,
                               use_proprio: bool = False, proprio_dim=None, mm_projector_type: str = "mlp2x_gelu"):
    """Reference OFTForCausalLM (oft_arch.py:50-56) with the `Linear` L1-regression head, random-init weights."""
    load_reference()
    from dexbotic.model.oft.oft_arch import OFTConfig, OFTForCausalLM
    cfg = OFTConfig(llm_config=llm_config, mm_projector_type=mm_projector_type, mm_vision_tower=clip_config,
                    action_model_type="Linear", action_dim=action_dim, chunk_size=chunk_size,
                    use_proprio=use_proprio, proprio_dim=proprio_dim)
    return OFTForCausalLM(cfg)


def build_reference_oft_diffusion(llm_config, clip_config, action_dim: int = 7, chunk_size: int = 8,
                                  use_proprio: bool = False, proprio_dim=None, mm_projector_type: str = "mlp2x_gelu"):
    """Reference OFTForCausalLM with the `DiT` DiffusionActionHead (oft/action_model/model.py:197-271) built around the
    restated DDIM scheduler."""
    load_reference()
    from dexbotic.model.oft.oft_arch import OFTConfig, OFTForCausalLM
    cfg = OFTConfig(llm_config=llm_config, mm_projector_type=mm_projector_type, mm_vision_tower=clip_config,
                    action_model_type="DiT", action_dim=action_dim, chunk_size=chunk_size,
                    use_proprio=use_proprio, proprio_dim=proprio_dim)
    return OFTForCausalLM(cfg)


def build_reference_memvla(llm_config, clip_config, action_model_type: str = "DiT-S", action_dim: int = 7,
                           chunk_size: int = 16, mm_projector_type: str = "mlp2x_gelu", dropout: float = 0.0, **mem):
    """Reference MemVLAForCausalLM (memvla_arch.py:536-544) with random-init weights.  `mem` = the memory-module config
    keys memvla_exp.py:198-236 sets (dataloader_type, group_size, per_token_size, mem_length, retrieval_layers,
    use_timestep_pe, fusion_type, consolidate_type, update_fused).  CrossTransformerBlock hard-codes dropout=0.1 and
    passes it to SDPA unconditionally (memvla_arch.py:83,122-124); parity runs set the instance attributes to
    `dropout` (0 by default) — no reference code is changed."""
    load_reference()
    from dexbotic.model.memvla.memvla_arch import CrossTransformerBlock, MemVLAConfig, MemVLAForCausalLM
    import torch.nn as nn
    cfg = MemVLAConfig(llm_config=llm_config, mm_projector_type=mm_projector_type, mm_vision_tower=clip_config,
                       action_model_type=action_model_type, action_dim=action_dim, chunk_size=chunk_size, **mem)
    model = MemVLAForCausalLM(cfg)
    for m in model.modules():
        if isinstance(m, CrossTransformerBlock):
            m.dropout = dropout
            for sub in m.ffn:
                if isinstance(sub, nn.Dropout):
                    sub.p = dropout
    return model


_pi0_loaded = False


def load_reference_pi0():
    """pi0_arch.py needs the same dataclass-default patch (Pi0Config's annotated fields) and a SigLIP tower
    that is built from a config object (no image processor / weights download)."""
    global _pi0_loaded
    load_reference()
    if _pi0_loaded:
        return sys.modules["dexbotic.model.pi0.pi0_arch"]
    from transformers import SiglipVisionModel
    from dexbotic.model.modules.mm_vision.siglip import siglip_encoder

    def siglip_load_model(self):
        if self.is_loaded:
            return
        self.image_processor = None
        self.vision_tower = SiglipVisionModel(self.vision_tower_config)
        self.is_loaded = True

    siglip_encoder.SiglipVisionTower.load_model = siglip_load_model
    # Version artifact, neutralised: pi0_arch.py:179-182 reads `past_key_values.key_cache[layer]` /
    # `.value_cache[layer]` — list attributes DynamicCache had under the pinned transformers 4.5x; 5.5.0 keeps the
    # same tensors in `cache.layers[layer].keys / .values`.  Expose the old read-only views.
    from transformers import DynamicCache
    if not hasattr(DynamicCache, "key_cache"):
        DynamicCache.key_cache = property(lambda self: [l.keys for l in self.layers])
        DynamicCache.value_cache = property(lambda self: [l.values for l in self.layers])
    import dexbotic.model.pi0  # noqa: F401
    mod = _exec_patched("dexbotic.model.pi0.pi0_arch", "dexbotic/model/pi0/pi0_arch.py",
                        [("    vision_config: dict | str\n", "    vision_config: dict | str = None\n"),
                         ("    processor_config: str\n", "    processor_config: str = None\n"),
                         ("    action_config: dict | str\n", "    action_config: dict | str = None\n")])
    _pi0_loaded = True
    return mod


def build_reference_pi0(llm_config: dict, action_config: dict, vision_config: dict, action_dim: int = 32,
                        chunk_size: int = 50):
    """Reference Pi0ForCausalLM (pi0_arch.py:109-114) with random-init weights; configs are plain dicts with a
    `model_type` key (Pi0Config resolves them through CONFIG_MAPPING, pi0_arch.py:61-83)."""
    mod = load_reference_pi0()
    cfg = mod.Pi0Config(llm_config=llm_config, action_config=action_config, vision_config=vision_config,
                        processor_config="unused", mm_projector_type="linear", action_dim=action_dim,
                        chunk_size=chunk_size)
    model = mod.Pi0ForCausalLM(cfg)
    # Version artifact, neutralised: under the reference's pinned transformers (4.51.0 / 4.54.0, Dockerfile:37,
    # Dockerfile.c130t28:22) GemmaModel.embed_tokens is a plain nn.Embedding and pi0_arch.py:258-261 applies the
    # sqrt(hidden) scale itself; transformers 5.5.0 (this image) moved that scale INTO the embedding module
    # (GemmaTextScaledWordEmbedding), which would apply it twice.  Reset the in-module scale to 1 so the run here
    # computes what the pinned reference computes.
    for m in (model.model.llm, model.model.action_expert):
        if hasattr(m.embed_tokens, "embed_scale"):
            m.embed_tokens.embed_scale.fill_(1.0)
    return model


_pi05_loaded = False


def load_reference_pi05():
    """pi05_arch.py + its vendored AdaRMS Gemma (pi05/transformers_pi05/gemma/modeling_gemma.py).  That file imports
    five names from `transformers.models.gemma.modeling_gemma` which the pinned transformers 4.5x re-exported there and
    5.5.0 keeps in their home modules; they are only referenced by code paths pi05 never calls (AdaRMSGemmaModel.forward,
    the *ForCausalLM heads).  Alias them in memory so the import succeeds — no reference code is changed."""
    global _pi05_loaded
    load_reference_pi0()
    if _pi05_loaded:
        return sys.modules["dexbotic.model.pi05.pi05_arch"]
    from typing import TypedDict
    import transformers.models.gemma.modeling_gemma as mg
    from transformers.cache_utils import StaticCache
    from transformers.modeling_attn_mask_utils import AttentionMaskConverter
    from transformers.modeling_outputs import SequenceClassifierOutputWithPast, TokenClassifierOutput

    class KwargsForCausalLM(TypedDict, total=False):
        pass

    for name, obj in dict(AttentionMaskConverter=AttentionMaskConverter, StaticCache=StaticCache,
                          SequenceClassifierOutputWithPast=SequenceClassifierOutputWithPast,
                          TokenClassifierOutput=TokenClassifierOutput, KwargsForCausalLM=KwargsForCausalLM).items():
        if not hasattr(mg, name):
            setattr(mg, name, obj)
    import dexbotic.model.pi05  # noqa: F401   (registers the adarms_gemma config / model with the Auto classes)
    mod = _exec_patched("dexbotic.model.pi05.pi05_arch", "dexbotic/model/pi05/pi05_arch.py",
                        [("    vision_config: dict | str\n", "    vision_config: dict | str = None\n"),
                         ("    processor_config: str\n", "    processor_config: str = None\n"),
                         ("    action_config: dict | str\n", "    action_config: dict | str = None\n")])
    _pi05_loaded = True
    return mod


def build_reference_pi05(llm_config: dict, action_config: dict, vision_config: dict, action_dim: int = 32,
                         chunk_size: int = 50):
    """Reference Pi05ForCausalLM (pi05_arch.py:110-116), random init.  llm / action configs are `adarms_gemma` dicts;
    `rope_parameters` is added because transformers 5.5's GemmaRotaryEmbedding reads it from the config object
    (the 4.5x one read rope_theta) — a constructor-compat field, same RoPE."""
    mod = load_reference_pi05()
    rp = dict(rope_type="default", rope_theta=float(llm_config.get("rope_theta", 10000.0)))
    llm_config = dict(llm_config, model_type="adarms_gemma", rope_parameters=rp)
    action_config = dict(action_config, model_type="adarms_gemma", rope_parameters=rp)
    cfg = mod.Pi05Config(llm_config=llm_config, action_config=action_config, vision_config=vision_config,
                         processor_config="unused", mm_projector_type="linear", action_dim=action_dim,
                         chunk_size=chunk_size)
    return mod.Pi05ForCausalLM(cfg)


_navila_loaded = False


def load_reference_navila():
    """navila_arch.py: the same un-defaulted dataclass field as DexboticConfig (`llm_config`), and a SigLIP tower built
    from a config object: build_vision_tower() insists on a processor_config for config objects
    (mm_vision/builder.py:26-29) and DexboticVLMModel passes none for NaVILA (dexbotic_arch.py:100-104), so a
    SiglipVisionConfig is routed to SiglipVisionTower directly with its default select_layer=-2 — exactly what the
    string name "google/siglip-so400m-patch14-384" resolves to (builder.py:16-17), minus the download."""
    global _navila_loaded
    load_reference_pi0()          # SigLIP load_model patch
    if _navila_loaded:
        return sys.modules["dexbotic.model.navila.navila_arch"]
    from transformers import SiglipVisionConfig
    from dexbotic.model.modules.mm_vision import builder as vbuilder
    from dexbotic.model.modules.mm_vision.siglip import siglip_encoder
    prev = vbuilder.build_vision_tower

    def build_vision_tower(mm_vision_tower, **kwargs):
        if isinstance(mm_vision_tower, SiglipVisionConfig) and "processor_config" not in kwargs:
            return siglip_encoder.SiglipVisionTower(mm_vision_tower, processor_config="unused")
        return prev(mm_vision_tower, **kwargs)

    vbuilder.build_vision_tower = build_vision_tower
    sys.modules["dexbotic.model.dexbotic_arch"].build_vision_tower = build_vision_tower
    import dexbotic.model.navila  # noqa: F401
    mod = _exec_patched("dexbotic.model.navila.navila_arch", "dexbotic/model/navila/navila_arch.py",
                        [("    llm_config: str | PretrainedConfig\n", "    llm_config: str | PretrainedConfig = None\n")])
    _navila_loaded = True
    return mod


def build_reference_navila(llm_config: dict, vision_config, time_token_ids=None, soft_ce_std: float = 1.0):
    """Reference NaVILAForCausalLM (navila_arch.py:222-231), random init.  llm_config: dict with `model_type`
    (NaVILAModel resolves it through CONFIG_MAPPING, :27-31); vision_config: a SiglipVisionConfig."""
    mod = load_reference_navila()
    cfg = mod.NaVILAConfig(llm_config=dict(llm_config), mm_projector_type="mlp_downsample", mm_vision_tower=vision_config)
    if time_token_ids:
        cfg.time_token_ids, cfg.soft_ce_std = list(time_token_ids), soft_ce_std
    return mod.NaVILAForCausalLM(cfg)


def build_reference_hybrid_cogact(llm_config, clip_config, action_model_type: str = "DiT-S", action_dim: int = 7,
                                  chunk_size: int = 16, mm_projector_type: str = "mlp2x_gelu"):
    """Reference HybridCogACTForCausalLM (cogact/hybrid_cogact_arch.py:51-58): text + action co-training."""
    load_reference()
    from dexbotic.model.cogact.hybrid_cogact_arch import CogActConfig, HybridCogACTForCausalLM
    cfg = CogActConfig(llm_config=llm_config, mm_projector_type=mm_projector_type, mm_vision_tower=clip_config,
                       action_model_type=action_model_type, action_dim=action_dim, chunk_size=chunk_size)
    return HybridCogACTForCausalLM(cfg)
