"""Host-side mirror of dexbotic/model (reference: /root/reference/dexbotic/model): same class names,
forward signatures, config fields and state-dict keys; the arithmetic runs in libdexbotic_b200.so."""
from .dexbotic_arch import (CausalLMOutputDexbotic, DexboticConfig, DexboticVLMModel, IGNORE_INDEX,  # noqa: F401
                            IMAGE_TOKEN_INDEX)
from .cogact_arch import CogActConfig, CogActModel, CogACTForCausalLM  # noqa: F401
from .oft_arch import OFTConfig, OFTDiscreteConfig, OFTDiscreteForCausalLM, OFTForCausalLM  # noqa: F401
from .pi0_arch import Pi0Config, Pi0ForCausalLM  # noqa: F401
from .memvla_arch import MemVLAConfig, MemVLAForCausalLM, MemVLAModel  # noqa: F401
from .pi05_arch import Pi05Config, Pi05ForCausalLM  # noqa: F401
