"""CPU test: the C-ABI shared library loads here (no GPU needed) and exports every symbol that
include/*.h declares; calling a compute entry point without a device fails loudly, never silently."""
import ctypes

import pytest

from dexbotic_b200 import _lib


def test_library_built_and_loads():
    assert _lib.lib_available(), "run `python -c 'import __graft_entry__ as g; g.build()'` first"
    lib = _lib.load()
    assert lib.b200_version() >= 100


def test_every_declared_symbol_is_exported():
    lib = _lib.load()
    assert len(_lib.EXPORTED) >= 30
    for name in _lib.EXPORTED:
        assert hasattr(lib, name), f"include/*.h declares {name} but the library does not export it"


def test_gemm_args_struct_matches_header():
    # field order/size sanity: the ctypes mirror must have the same size on both sides of the ABI
    assert ctypes.sizeof(_lib.GemmArgs) % 8 == 0
    assert _lib.GemmArgs.a.offset == 0 and _lib.GemmArgs.d.offset == 24 and _lib.GemmArgs.m.offset == 48


def test_no_cpu_fallback():
    import torch
    from dexbotic_b200 import ops
    with pytest.raises(RuntimeError):
        ops.gemm(torch.zeros(8, 8), torch.zeros(8, 8))
    with pytest.raises(RuntimeError):
        ops.rmsnorm_fwd(torch.zeros(4, 8), torch.ones(8), 1e-6)


def test_config_json_round_trip_for_every_policy():
    """save_pretrained's config.json -> from_pretrained's config object (host logic only): every DexboticConfig
    subclass keeps its fields, nested HF configs become plain dicts that the spec builders read the same way."""
    import json
    from transformers import CLIPVisionConfig, Qwen2Config
    import dexbotic_b200.model as m
    from dexbotic_b200.model._module import _config_to_dict
    from dexbotic_b200.model.dexbotic_arch import cfg_get
    llm = Qwen2Config(vocab_size=128, hidden_size=64, intermediate_size=160, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2)
    clip = CLIPVisionConfig(hidden_size=32, intermediate_size=64, num_hidden_layers=3, num_attention_heads=2,
                            image_size=28, patch_size=14)
    gemma = dict(model_type="gemma", vocab_size=128, hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                 num_attention_heads=4, num_key_value_heads=1, head_dim=16)
    siglip = dict(model_type="siglip_vision_model", hidden_size=32, intermediate_size=64, num_hidden_layers=2,
                  num_attention_heads=2, image_size=28, patch_size=14)
    made = {
        "dexbotic_cogact": m.CogActConfig(llm_config=llm, mm_vision_tower=clip, action_model_type="DiT-S", action_dim=7,
                                          chunk_size=16, freeze_mm_vision=True),
        "dexbotic_oft": m.OFTConfig(llm_config=llm, mm_vision_tower=clip, action_model_type="Linear", action_dim=7,
                                    chunk_size=8, use_proprio=True, proprio_dim=9),
        "dexbotic_oft_discrete": m.OFTDiscreteConfig(llm_config=llm, mm_vision_tower=clip, action_model_type="Discrete",
                                                     action_dim=7, chunk_size=8, num_bins=256),
        "dexbotic_memvla": m.MemVLAConfig(llm_config=llm, mm_vision_tower=clip, action_model_type="DiT-L", action_dim=7,
                                          chunk_size=16, group_size=4, per_token_size=32, mem_length=3),
        "dexbotic_pi0": m.Pi0Config(llm_config=gemma, action_config=dict(gemma, hidden_size=32), vision_config=siglip,
                                    action_dim=32, chunk_size=10),
        "dexbotic_pi05": m.Pi05Config(llm_config=gemma, action_config=dict(gemma, hidden_size=32, adarms_cond_dim=32),
                                      vision_config=siglip, action_dim=32, chunk_size=10),
        "dexbotic_navila": m.NaVILAConfig(llm_config=dict(model_type="llama", vocab_size=128, hidden_size=64,
                                                          intermediate_size=160, num_hidden_layers=2,
                                                          num_attention_heads=4, num_key_value_heads=2),
                                          mm_vision_tower=siglip, time_token_ids=[120, 121, 122], soft_ce_std=1.5),
    }
    assert set(made) == set(m.MODEL_TYPES)
    for mt, cfg in made.items():
        d = json.loads(json.dumps(_config_to_dict(cfg), default=str))
        assert d.pop("model_type") == mt
        cls = m.MODEL_TYPES[mt][0]
        back = cls(**d)
        for k, v in vars(cfg).items():
            if k.startswith("_"):
                continue
            w = getattr(back, k)
            if hasattr(v, "to_dict"):          # nested HF config -> dict with the same entries the spec builders read
                for key in ("hidden_size", "num_hidden_layers", "num_attention_heads", "intermediate_size"):
                    assert cfg_get(w, key) == cfg_get(v, key), (mt, k, key)
            else:
                assert w == v, (mt, k, w, v)
