"""Architecture description of the VibeVoice generation path.

Mirrors the reference's HF config composition (`vibevoice/modular/configuration_vibevoice.py:13-241`:
`VibeVoiceConfig` = decoder_config (Qwen2) + acoustic/semantic tokenizer configs + diffusion head
config) and reads the same `config.json` files (`vibevoice/configs/qwen2.5_1.5b_64k.json`,
`qwen2.5_7b_32k.json`).  Plain attribute bags, no transformers dependency: the C-ABI runtime only
needs the numbers.
"""
from __future__ import annotations

import copy
import json
import os
from typing import Any, Dict, List, Optional


class _Bag:
    _defaults: Dict[str, Any] = {}

    def __init__(self, **kw):
        d = copy.deepcopy(self._defaults)
        d.update(kw)
        for k, v in d.items():
            setattr(self, k, v)

    def to_dict(self):
        out = {}
        for k, v in self.__dict__.items():
            out[k] = v.to_dict() if isinstance(v, _Bag) else copy.deepcopy(v)
        return out

    def __repr__(self):
        return "%s(%s)" % (type(self).__name__, json.dumps(self.to_dict(), sort_keys=True))


class VibeVoiceAcousticTokenizerConfig(_Bag):
    """`configuration_vibevoice.py:13-73`."""
    model_type = "vibevoice_acoustic_tokenizer"
    _defaults = dict(
        channels=1, corpus_normalize=0.0, causal=True, vae_dim=64, fix_std=0.5, std_dist_type="gaussian",
        mixer_layer="depthwise_conv", conv_norm="none", pad_mode="constant", disable_last_norm=True,
        layernorm="RMSNorm", layernorm_eps=1e-5, layernorm_elementwise_affine=True, conv_bias=True,
        layer_scale_init_value=1e-6, weight_init_value=1e-2, encoder_n_filters=32,
        encoder_ratios=[8, 5, 5, 4, 2, 2], encoder_depths="3-3-3-3-3-3-8", decoder_n_filters=32,
        decoder_ratios=None, decoder_depths=None,
    )

    def __init__(self, **kw):
        super().__init__(**kw)
        if self.decoder_ratios is None:
            self.decoder_ratios = list(self.encoder_ratios)

    @property
    def encoder_depth_list(self) -> List[int]:
        d = self.encoder_depths
        return [int(x) for x in d.split("-")] if isinstance(d, str) else list(d)

    @property
    def decoder_depth_list(self) -> List[int]:
        # `modular_vibevoice_tokenizer.py:1024-1028`: decoder_depths=None => reversed encoder depths
        d = self.decoder_depths
        if d is None:
            return list(reversed(self.encoder_depth_list))
        return [int(x) for x in d.split("-")] if isinstance(d, str) else list(d)


class VibeVoiceSemanticTokenizerConfig(VibeVoiceAcousticTokenizerConfig):
    """`configuration_vibevoice.py:76-127` (encoder only)."""
    model_type = "vibevoice_semantic_tokenizer"
    _defaults = dict(VibeVoiceAcousticTokenizerConfig._defaults, fix_std=0, std_dist_type="none")


class VibeVoiceDiffusionHeadConfig(_Bag):
    """`configuration_vibevoice.py:130-162`."""
    model_type = "vibevoice_diffusion_head"
    _defaults = dict(
        hidden_size=768, head_layers=4, head_ffn_ratio=3.0, rms_norm_eps=1e-5, latent_size=64,
        speech_vae_dim=None, prediction_type="v_prediction", diffusion_type="ddpm", ddpm_num_steps=1000,
        ddpm_num_inference_steps=20, ddpm_beta_schedule="cosine", ddpm_batch_mul=4,
    )


class Qwen2DecoderConfig(_Bag):
    """The Qwen2Config fields the decode path reads (`vibevoice/configs/qwen2.5_1.5b_64k.json:43-64`)."""
    model_type = "qwen2"
    _defaults = dict(
        hidden_size=1536, intermediate_size=8960, num_hidden_layers=28, num_attention_heads=12,
        num_key_value_heads=2, head_dim=None, max_position_embeddings=65536, rms_norm_eps=1e-6,
        rope_theta=1000000.0, vocab_size=151936, tie_word_embeddings=True, hidden_act="silu",
        attention_dropout=0.0, initializer_range=0.02, use_sliding_window=False, sliding_window=None,
        _attn_implementation="b200_paged_splitkv",
    )

    def __init__(self, **kw):
        super().__init__(**kw)
        if self.head_dim is None:
            self.head_dim = self.hidden_size // self.num_attention_heads


class VibeVoiceConfig(_Bag):
    """`configuration_vibevoice.py:164-241`."""
    model_type = "vibevoice"
    _defaults = dict(torch_dtype="bfloat16")

    def __init__(self, acoustic_tokenizer_config=None, semantic_tokenizer_config=None, decoder_config=None,
                 diffusion_head_config=None, **kw):
        super().__init__(**kw)

        def mk(cls, v):
            if v is None:
                return cls()
            if isinstance(v, dict):
                v = {k: x for k, x in v.items() if k != "model_type"}
                return cls(**v)
            return v

        self.acoustic_tokenizer_config = mk(VibeVoiceAcousticTokenizerConfig, acoustic_tokenizer_config)
        self.semantic_tokenizer_config = mk(VibeVoiceSemanticTokenizerConfig, semantic_tokenizer_config)
        if isinstance(decoder_config, dict) and decoder_config.get("model_type", "qwen2") != "qwen2":
            raise ValueError("Unsupported decoder model type: %s" % decoder_config.get("model_type"))
        self.decoder_config = mk(Qwen2DecoderConfig, decoder_config)
        self.diffusion_head_config = mk(VibeVoiceDiffusionHeadConfig, diffusion_head_config)
        self.acoustic_vae_dim = getattr(self.acoustic_tokenizer_config, "vae_dim", 64)
        self.semantic_vae_dim = getattr(self.semantic_tokenizer_config, "vae_dim", 128)

    @classmethod
    def from_dict(cls, d: Dict[str, Any]) -> "VibeVoiceConfig":
        d = dict(d)
        d.pop("model_type", None)
        d.pop("acoustic_vae_dim", None)
        d.pop("semantic_vae_dim", None)
        return cls(**d)

    @classmethod
    def from_pretrained(cls, path: str) -> "VibeVoiceConfig":
        if os.path.isdir(path):
            path = os.path.join(path, "config.json")
        with open(path) as f:
            return cls.from_dict(json.load(f))


_COMMON_TOK = dict(
    causal=True, channels=1, conv_bias=True, conv_norm="none", corpus_normalize=0.0, disable_last_norm=True,
    encoder_depths="3-3-3-3-3-3-8", encoder_n_filters=32, encoder_ratios=[8, 5, 5, 4, 2, 2],
    layer_scale_init_value=1e-6, layernorm="RMSNorm", layernorm_elementwise_affine=True, layernorm_eps=1e-5,
    mixer_layer="depthwise_conv", pad_mode="constant", weight_init_value=0.01,
)


def _preset(hidden, inter, heads, kv_heads, max_pos, vocab, tie):
    return dict(
        acoustic_tokenizer_config=dict(_COMMON_TOK, decoder_depths=None, decoder_n_filters=32,
                                       decoder_ratios=[8, 5, 5, 4, 2, 2], fix_std=0.5, std_dist_type="gaussian",
                                       vae_dim=64),
        semantic_tokenizer_config=dict(_COMMON_TOK, fix_std=0, std_dist_type="none", vae_dim=128),
        decoder_config=dict(hidden_size=hidden, intermediate_size=inter, num_hidden_layers=28,
                            num_attention_heads=heads, num_key_value_heads=kv_heads,
                            max_position_embeddings=max_pos, rms_norm_eps=1e-6, rope_theta=1000000.0,
                            vocab_size=vocab, tie_word_embeddings=tie, model_type="qwen2"),
        diffusion_head_config=dict(ddpm_batch_mul=4, ddpm_beta_schedule="cosine", ddpm_num_inference_steps=20,
                                   ddpm_num_steps=1000, diffusion_type="ddpm", head_ffn_ratio=3.0, head_layers=4,
                                   hidden_size=hidden, latent_size=64, prediction_type="v_prediction",
                                   rms_norm_eps=1e-5, speech_vae_dim=64),
    )


def preset_config(name: str) -> VibeVoiceConfig:
    """Architecture presets with the numbers of the shipped JSONs (`vibevoice/configs/*.json`) plus two
    reduced shapes used only by parity tests ("tiny", "small")."""
    if name in ("1.5b", "1.5B", "vibevoice-1.5b"):
        return VibeVoiceConfig.from_dict(_preset(1536, 8960, 12, 2, 65536, 151936, True))
    if name in ("7b", "7B", "vibevoice-7b"):
        return VibeVoiceConfig.from_dict(_preset(3584, 18944, 28, 4, 32768, 152064, False))
    if name in ("1.5b-l2", "7b-l2"):
        d = _preset(1536, 8960, 12, 2, 65536, 4096, True) if name == "1.5b-l2" else _preset(3584, 18944, 28, 4, 32768, 4096, False)
        d["decoder_config"]["num_hidden_layers"] = 2
        return VibeVoiceConfig.from_dict(d)
    if name in ("tiny", "small"):
        # same topology (7 codec stages, ratios, GQA, 4 head layers), narrow widths
        big = name == "small"
        d = _preset(256 if big else 128, 768 if big else 384, 4 if big else 2, 2 if big else 1,
                    4096, 2048, True)
        d["decoder_config"]["num_hidden_layers"] = 4 if big else 2
        d["decoder_config"]["head_dim"] = 128
        for k in ("acoustic_tokenizer_config", "semantic_tokenizer_config"):
            d[k]["encoder_n_filters"] = 16 if big else 8
            d[k]["encoder_depths"] = "2-1-1-1-1-1-2" if big else "1-1-1-1-1-1-2"
        d["acoustic_tokenizer_config"]["decoder_n_filters"] = 16 if big else 8
        return VibeVoiceConfig.from_dict(d)
    if name == "tiny64":
        # head_dim 64 (the streaming-0.5B attention geometry: GQA group 2 here) at toy widths
        d = _preset(128, 384, 4, 2, 4096, 2048, True)
        d["decoder_config"]["num_hidden_layers"] = 2
        d["decoder_config"]["head_dim"] = 64
        for k in ("acoustic_tokenizer_config", "semantic_tokenizer_config"):
            d[k]["encoder_n_filters"] = 8
            d[k]["encoder_depths"] = "1-1-1-1-1-1-2"
        d["acoustic_tokenizer_config"]["decoder_n_filters"] = 8
        return VibeVoiceConfig.from_dict(d)
    if name in ("streaming-0.5b", "streaming-0.5b-l4"):
        # VibeVoice-Streaming-0.5B language model: Qwen2.5-0.5B geometry (H = 896, 14 query / 2 kv heads of 64, I = 4864, 24 layers = 4 text
        # + 20 TTS layers; inferred from the cached-prompt tensors demo/voices/streaming_model/*.pt and the public Qwen2.5-0.5B config, the
        # reference ships no JSON for it -- SURVEY 8d-5).  "-l4" = the same layer shapes with 1 + 3 layers for parity tests.
        d = _preset(896, 4864, 14, 2, 8192 if name.endswith("l4") else 32768, 4096 if name.endswith("l4") else 151936, True)
        d["decoder_config"]["num_hidden_layers"] = 4 if name.endswith("l4") else 24
        d["decoder_config"]["head_dim"] = 64
        return VibeVoiceConfig.from_dict(d)
    raise ValueError("unknown preset %r" % name)
