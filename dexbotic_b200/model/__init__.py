"""Host-side mirror of dexbotic/model (reference: /root/reference/dexbotic/model): same class names,
forward signatures, config fields and state-dict keys; the arithmetic runs in libdexbotic_b200.so."""
from .dexbotic_arch import (CausalLMOutputDexbotic, DexboticConfig, DexboticVLMModel, IGNORE_INDEX,  # noqa: F401
                            IMAGE_TOKEN_INDEX)
from .cogact_arch import CogActConfig, CogActModel, CogACTForCausalLM, HybridCogACTForCausalLM  # noqa: F401
from .oft_arch import OFTConfig, OFTDiscreteConfig, OFTDiscreteForCausalLM, OFTForCausalLM  # noqa: F401
from .pi0_arch import Pi0Config, Pi0ForCausalLM  # noqa: F401
from .memvla_arch import MemVLAConfig, MemVLAForCausalLM, MemVLAModel  # noqa: F401
from .pi05_arch import Pi05Config, Pi05ForCausalLM  # noqa: F401
from .navila_arch import NaVILAConfig, NaVILAForCausalLM, NaVILAModel  # noqa: F401

# model_type string (DexboticConfig subclasses' `model_type`, written to config.json by save_pretrained) -> classes
MODEL_TYPES = {
    CogActConfig.model_type: (CogActConfig, CogACTForCausalLM),
    Pi0Config.model_type: (Pi0Config, Pi0ForCausalLM),
    Pi05Config.model_type: (Pi05Config, Pi05ForCausalLM),
    OFTConfig.model_type: (OFTConfig, OFTForCausalLM),
    OFTDiscreteConfig.model_type: (OFTDiscreteConfig, OFTDiscreteForCausalLM),
    MemVLAConfig.model_type: (MemVLAConfig, MemVLAForCausalLM),
    NaVILAConfig.model_type: (NaVILAConfig, NaVILAForCausalLM),
}


def from_pretrained(path, device="cuda", **config_overrides):
    """Build whichever policy the checkpoint directory's config.json names (HF AutoModel-style dispatch)."""
    import json
    from pathlib import Path
    mt = json.loads((Path(path) / "config.json").read_text()).get("model_type")
    if mt not in MODEL_TYPES:
        raise ValueError(f"unknown model_type {mt!r}; known: {sorted(MODEL_TYPES)}")
    return MODEL_TYPES[mt][1].from_pretrained(path, device=device, **config_overrides)
