"""Synthetic (random-init) checkpoints with the reference's state-dict layout.

There are no weights on disk and no network, so every parity test and benchmark runs on a
seeded random-init model.  Key names follow the module tree of `VibeVoiceModel.__init__`
(`vibevoice/modular/modeling_vibevoice.py:119-142`) + `lm_head`
(`modeling_vibevoice_inference.py:79`), i.e. exactly what a real HF checkpoint holds, so the
same loader path (`VibeVoiceForConditionalGenerationInference.load_state_dict`) is exercised.

Every parameter is re-randomised explicitly: the reference's zero-initialised
AdaLN/final layers (`modular_vibevoice_diffusion_head.py:240-252`) and 1e-6 layer-scale gammas
(`configs/qwen2.5_1.5b_64k.json:32`) would otherwise hide most of the arithmetic.

Each tensor gets its own CPU generator seeded from (seed, name), so a single tensor can be
regenerated without materialising the checkpoint, and CPU/GPU-box runs agree bit for bit.
"""
from __future__ import annotations

import zlib
from typing import Dict, Iterator, List, Tuple

import torch

from .configuration import VibeVoiceConfig

# init kinds
W, B, G, N, E = "weight", "bias", "gamma", "norm", "embed"


def _tok_encoder_specs(prefix: str, tc, vae_dim: int) -> List[Tuple[str, Tuple[int, ...], str]]:
    """`TokenizerEncoder.__init__` (`modular_vibevoice_tokenizer.py:694-774`)."""
    out = []
    nf = tc.encoder_n_filters
    ratios = list(reversed(tc.encoder_ratios))
    depths = tc.encoder_depth_list
    out.append((f"{prefix}.downsample_layers.0.0.conv.conv.weight", (nf, tc.channels, 7), W))
    out.append((f"{prefix}.downsample_layers.0.0.conv.conv.bias", (nf,), B))
    for i, r in enumerate(ratios):
        cin, cout = nf * 2 ** i, nf * 2 ** (i + 1)
        out.append((f"{prefix}.downsample_layers.{i+1}.0.conv.conv.weight", (cout, cin, 2 * r), W))
        out.append((f"{prefix}.downsample_layers.{i+1}.0.conv.conv.bias", (cout,), B))
    for i, d in enumerate(depths):
        c = nf * 2 ** i
        for j in range(d):
            out += _block_specs(f"{prefix}.stages.{i}.{j}", c)
    c = nf * 2 ** (len(depths) - 1)
    out.append((f"{prefix}.head.conv.conv.weight", (vae_dim, c, 7), W))
    out.append((f"{prefix}.head.conv.conv.bias", (vae_dim,), B))
    return out


def _block_specs(p: str, c: int):
    """`Block1D` (`modular_vibevoice_tokenizer.py:620-663`), depthwise mixer, RMSNorm, FFN x4 with bias."""
    return [
        (f"{p}.norm.weight", (c,), N),
        (f"{p}.mixer.conv.conv.conv.weight", (c, 1, 7), W),
        (f"{p}.mixer.conv.conv.conv.bias", (c,), B),
        (f"{p}.gamma", (c,), G),
        (f"{p}.ffn_norm.weight", (c,), N),
        (f"{p}.ffn.linear1.weight", (4 * c, c), W),
        (f"{p}.ffn.linear1.bias", (4 * c,), B),
        (f"{p}.ffn.linear2.weight", (c, 4 * c), W),
        (f"{p}.ffn.linear2.bias", (c,), B),
        (f"{p}.ffn_gamma", (c,), G),
    ]


def _tok_decoder_specs(prefix: str, tc) -> List[Tuple[str, Tuple[int, ...], str]]:
    """`TokenizerDecoder.__init__` (`modular_vibevoice_tokenizer.py:823-912`)."""
    out = []
    nf = tc.decoder_n_filters
    ratios = list(tc.decoder_ratios)
    depths = tc.decoder_depth_list
    nst = len(depths)
    c0 = nf * 2 ** (nst - 1)
    out.append((f"{prefix}.upsample_layers.0.0.conv.conv.weight", (c0, tc.vae_dim, 7), W))
    out.append((f"{prefix}.upsample_layers.0.0.conv.conv.bias", (c0,), B))
    for i, r in enumerate(ratios):
        cin, cout = nf * 2 ** (nst - 1 - i), nf * 2 ** (nst - 2 - i)
        out.append((f"{prefix}.upsample_layers.{i+1}.0.convtr.convtr.weight", (cin, cout, 2 * r), W))
        out.append((f"{prefix}.upsample_layers.{i+1}.0.convtr.convtr.bias", (cout,), B))
    for i, d in enumerate(depths):
        c = nf * 2 ** (nst - 1 - i)
        for j in range(d):
            out += _block_specs(f"{prefix}.stages.{i}.{j}", c)
    out.append((f"{prefix}.head.conv.conv.weight", (tc.channels, nf, 7), W))
    out.append((f"{prefix}.head.conv.conv.bias", (tc.channels,), B))
    return out


def _connector_specs(p: str, din: int, h: int):
    """`SpeechConnector` (`modeling_vibevoice.py:58-69`)."""
    return [(f"{p}.fc1.weight", (h, din), W), (f"{p}.fc1.bias", (h,), B), (f"{p}.norm.weight", (h,), N),
            (f"{p}.fc2.weight", (h, h), W), (f"{p}.fc2.bias", (h,), B)]


def _head_specs(p: str, hc):
    """`VibeVoiceDiffusionHead.__init__` (`modular_vibevoice_diffusion_head.py:204-236`); no biases."""
    h, lat = hc.hidden_size, hc.latent_size
    f = int(h * hc.head_ffn_ratio)
    out = [(f"{p}.noisy_images_proj.weight", (h, lat), W), (f"{p}.cond_proj.weight", (h, h), W),
           (f"{p}.t_embedder.mlp.0.weight", (h, 256), W), (f"{p}.t_embedder.mlp.2.weight", (h, h), W)]
    for i in range(hc.head_layers):
        out += [(f"{p}.layers.{i}.ffn.gate_proj.weight", (f, h), W), (f"{p}.layers.{i}.ffn.up_proj.weight", (f, h), W),
                (f"{p}.layers.{i}.ffn.down_proj.weight", (h, f), W), (f"{p}.layers.{i}.norm.weight", (h,), N),
                (f"{p}.layers.{i}.adaLN_modulation.1.weight", (3 * h, h), W)]
    out += [(f"{p}.final_layer.linear.weight", (lat, h), W), (f"{p}.final_layer.adaLN_modulation.1.weight", (2 * h, h), W)]
    return out


def _lm_specs(p: str, dc):
    """Qwen2Model parameters (third-party `transformers/models/qwen2/modeling_qwen2.py`; q/k/v bias=True)."""
    h, i_, hd = dc.hidden_size, dc.intermediate_size, dc.head_dim
    nq, nkv = dc.num_attention_heads * hd, dc.num_key_value_heads * hd
    out = [(f"{p}.embed_tokens.weight", (dc.vocab_size, h), E)]
    for l in range(dc.num_hidden_layers):
        q = f"{p}.layers.{l}"
        out += [(f"{q}.self_attn.q_proj.weight", (nq, h), W), (f"{q}.self_attn.q_proj.bias", (nq,), B),
                (f"{q}.self_attn.k_proj.weight", (nkv, h), W), (f"{q}.self_attn.k_proj.bias", (nkv,), B),
                (f"{q}.self_attn.v_proj.weight", (nkv, h), W), (f"{q}.self_attn.v_proj.bias", (nkv,), B),
                (f"{q}.self_attn.o_proj.weight", (h, nq), W),
                (f"{q}.mlp.gate_proj.weight", (i_, h), W), (f"{q}.mlp.up_proj.weight", (i_, h), W),
                (f"{q}.mlp.down_proj.weight", (h, i_), W),
                (f"{q}.input_layernorm.weight", (h,), N), (f"{q}.post_attention_layernorm.weight", (h,), N)]
    out.append((f"{p}.norm.weight", (h,), N))
    return out


def param_specs(cfg: VibeVoiceConfig, parts=("lm", "head", "acoustic_decoder", "acoustic_encoder", "semantic",
                                              "connectors", "lm_head")):
    """(name, shape, init-kind) for every tensor of a VibeVoice inference checkpoint."""
    dc = cfg.decoder_config
    out = []
    if "lm" in parts:
        out += _lm_specs("model.language_model", dc)
    if "acoustic_encoder" in parts:
        out += _tok_encoder_specs("model.acoustic_tokenizer.encoder", cfg.acoustic_tokenizer_config,
                                  cfg.acoustic_tokenizer_config.vae_dim)
    if "acoustic_decoder" in parts:
        out += _tok_decoder_specs("model.acoustic_tokenizer.decoder", cfg.acoustic_tokenizer_config)
    if "semantic" in parts:
        out += _tok_encoder_specs("model.semantic_tokenizer.encoder", cfg.semantic_tokenizer_config,
                                  cfg.semantic_tokenizer_config.vae_dim)
    if "connectors" in parts:
        out += _connector_specs("model.acoustic_connector", cfg.acoustic_vae_dim, dc.hidden_size)
        out += _connector_specs("model.semantic_connector", cfg.semantic_vae_dim, dc.hidden_size)
    if "head" in parts:
        out += _head_specs("model.prediction_head", cfg.diffusion_head_config)
    if "lm_head" in parts and not dc.tie_word_embeddings:
        out.append(("lm_head.weight", (dc.vocab_size, dc.hidden_size), E))
    return out


def synth_tensor(name: str, shape, kind: str, seed: int, dtype=torch.float32, device="cpu") -> torch.Tensor:
    """Deterministic value for one parameter.  Scales are chosen so activations stay O(1) through
    28 LM layers / 26 codec blocks and so every term of every block contributes visibly."""
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    numel = 1
    for s in shape:
        numel *= s
    if kind in (W, E):
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        std = 0.02 if kind == E else min(0.05, 0.7 / max(fan_in, 1) ** 0.5)
        if "convtr" in name:  # ConvTranspose1d weight is [Cin, Cout, k]; every output sees 2*Cin taps
            std = min(0.05, 0.7 / (2.0 * shape[0]) ** 0.5)
        t = torch.empty(shape, dtype=torch.float32)
        # generate in chunks to bound peak memory for the 7B embedding (152064 x 3584)
        flat = t.view(-1)
        step = 1 << 24
        for s in range(0, numel, step):
            flat[s:s + step].normal_(0.0, std, generator=g)
    elif kind == B:
        t = torch.empty(shape, dtype=torch.float32).normal_(0.0, 0.02, generator=g)
    elif kind == G:
        t = torch.empty(shape, dtype=torch.float32).uniform_(0.2, 0.6, generator=g)
    elif kind == N:
        t = torch.empty(shape, dtype=torch.float32).uniform_(0.5, 1.5, generator=g)
    else:  # pragma: no cover
        raise ValueError(kind)
    return t.to(dtype=dtype, device=device)


def iter_synth_state_dict(cfg: VibeVoiceConfig, seed: int = 1234, dtype=torch.bfloat16, parts=None,
                          device="cpu") -> Iterator[Tuple[str, torch.Tensor]]:
    specs = param_specs(cfg) if parts is None else param_specs(cfg, parts)
    for name, shape, kind in specs:
        yield name, synth_tensor(name, shape, kind, seed, dtype=dtype, device=device)
    yield "model.speech_scaling_factor", torch.tensor(SPEECH_SCALING_FACTOR, dtype=torch.float32)
    yield "model.speech_bias_factor", torch.tensor(SPEECH_BIAS_FACTOR, dtype=torch.float32)


def synth_state_dict(cfg: VibeVoiceConfig, seed: int = 1234, dtype=torch.bfloat16, parts=None) -> Dict[str, torch.Tensor]:
    return dict(iter_synth_state_dict(cfg, seed, dtype, parts))


# Random-init models have NaN scaling buffers (`modeling_vibevoice.py:131-132`); BASELINE.md section 3 fixes these.
SPEECH_SCALING_FACTOR = 0.2
SPEECH_BIAS_FACTOR = -0.05


class SynthTokenizer:
    """The six ids `generate()` reads from its `tokenizer` argument
    (`modeling_vibevoice_inference.py:257-279`; `modular_vibevoice_text_tokenizer.py:163-181`).  The numeric values
    are those of the public Qwen2.5 vocabulary; reduced-vocab test presets place them at the top of the vocab."""

    def __init__(self, vocab_size: int = 151936):
        if vocab_size >= 151936:
            self.eos_token_id, self.speech_start_id, self.speech_end_id = 151643, 151652, 151653
            self.speech_diffusion_id, self.pad_token_id = 151654, 151655
        else:
            self.eos_token_id, self.speech_start_id, self.speech_end_id = vocab_size - 13, vocab_size - 4, vocab_size - 3
            self.speech_diffusion_id, self.pad_token_id = vocab_size - 2, vocab_size - 1
        self.bos_token_id = None
        self.pad_id = self.pad_token_id


def iter_synth_state_dict_fast(cfg: VibeVoiceConfig, seed: int = 1234, device="cuda", parts=None):
    """Same layout and scales as `iter_synth_state_dict`, generated directly on the GPU (device RNG, so values differ
    from the CPU-seeded checkpoint).  Used by bench.py only, where weight *values* are irrelevant."""
    specs = param_specs(cfg) if parts is None else param_specs(cfg, parts)
    g = torch.Generator(device=device)
    for name, shape, kind in specs:
        g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
        if kind in (W, E):
            fan_in = 1
            for s_ in shape[1:]:
                fan_in *= s_
            std = 0.02 if kind == E else min(0.05, 0.7 / max(fan_in, 1) ** 0.5)
            if "convtr" in name:
                std = min(0.05, 0.7 / (2.0 * shape[0]) ** 0.5)
            t = torch.empty(shape, dtype=torch.bfloat16, device=device).normal_(0.0, std, generator=g)
        elif kind == B:
            t = torch.empty(shape, dtype=torch.float32, device=device).normal_(0.0, 0.02, generator=g)
        elif kind == G:
            t = torch.empty(shape, dtype=torch.float32, device=device).uniform_(0.2, 0.6, generator=g)
        else:
            t = torch.empty(shape, dtype=torch.float32, device=device).uniform_(0.5, 1.5, generator=g)
        yield name, t
    yield "model.speech_scaling_factor", torch.tensor(SPEECH_SCALING_FACTOR, dtype=torch.float32)
    yield "model.speech_bias_factor", torch.tensor(SPEECH_BIAS_FACTOR, dtype=torch.float32)
