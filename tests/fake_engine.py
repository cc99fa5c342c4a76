"""TEST INFRASTRUCTURE: a CPU stand-in for `vibevoice_b200.engine.Engine` whose arithmetic is the oracle's.

Purpose: run the PRODUCT's host-side state machine (`vibevoice_b200/modeling.py::generate` -- token bookkeeping, which KV entries
each stream keeps, when codec state is zeroed, how noise rows are packed, what is streamed out) on a machine without a GPU and hold
it to the fixtures produced by the reference's own `generate()` (`tests/golden/loop.pt`).  The CUDA library is not involved and
nothing here is importable from the product; the GPU tests hold the real engine to the same oracle.

The engine contract mirrored here (include/vibevoice_b200.h):
  * 2B rows: 0..B-1 positive streams, B..2B-1 negative (CFG) streams; `lm_decode` consumes `embeds[2B,H]`, appends one K/V entry per
    row SPECULATIVELY at kv_len, writes `hidden[2B,H]`, and the constrained logits / argmax for rows 0..B-1;
  * `kv_commit(advance[2B])` keeps (1) or discards (0) the speculative entry; `kv_set_len(seq, n)` truncates;
  * `frame_tail(cfg)`: for active rows  hidden(pos,neg) + noise -> latent -> audio frame -> semantic feature -> connectors ->
    embeds[b] and embeds[B+b]; inactive rows keep their token embeddings (select_embeds_kernel).
"""
from __future__ import annotations

from typing import List

import numpy as np
import torch

from oracle import vv_oracle as O


class FakeEngine:
    def __init__(self, config, valid_ids, max_batch: int = 1, device=0, max_diffusion_steps=64, weights=None):   # Engine's signature + weights
        if weights is None:                                  # constructed by the product code (Engine signature): weights arrive
            self._pending, self.finalized_args = {}, None    # through load_tensor()/finalize() like the C ABI's vv_load_tensor
            self._init_args = (config, valid_ids, max_batch)
            self.config, self.B, self.finalized, self.kv_pages = config, max_batch, False, 0
            self.device, self.stream = torch.device("cpu"), None
            self.valid_ids = sorted(set(int(v) for v in valid_ids))
            self.scheduler = None
            return
        self._setup(config, valid_ids, max_batch, weights)

    def load_tensor(self, name: str, t: torch.Tensor) -> int:
        self._pending[name] = t.detach().clone()
        return 0

    def finalize(self, speech_scaling_factor=None, speech_bias_factor=None):
        w = dict(self._pending)
        w["model.speech_scaling_factor"] = torch.tensor(float(speech_scaling_factor))
        w["model.speech_bias_factor"] = torch.tensor(float(speech_bias_factor))
        config, valid_ids, max_batch = self._init_args
        self._setup(config, valid_ids, max_batch, w)

    def _setup(self, config, valid_ids, max_batch: int, weights):
        self.scheduler = None
        self.config, self.w = config, weights
        self.B = max_batch
        self.device = torch.device("cpu")
        self.stream = None                                   # torch.cuda.stream(None) is a no-op context
        self.valid_ids = sorted(set(int(v) for v in valid_ids))
        self.finalized = True
        self.kv_pages = 0
        self.n_steps = 0
        dc = config.decoder_config
        H, B = dc.hidden_size, max_batch
        self.embeds = torch.zeros(2 * B, H)
        self.hidden = torch.zeros(2 * B, H)
        self.logits = torch.zeros(B, len(self.valid_ids))
        self.tokens = torch.zeros(B, dtype=torch.int32)
        self.noise = torch.zeros(B, 64)
        self.active = torch.zeros(B, dtype=torch.int32)
        self.latent = torch.zeros(B, 64)
        self.audio = torch.zeros(B, 3200)
        self.feat = torch.zeros(B, config.semantic_vae_dim)
        self.kv: List[O.KVCache] = [O.KVCache(dc.num_hidden_layers) for _ in range(2 * B)]
        self.spec = [False] * (2 * B)                        # a speculative (uncommitted) entry sits at the end of the stream
        self.a_state, self.s_state = O.StreamState(B), O.StreamState(B)
        self.embed_w = weights["model.language_model.embed_tokens.weight"].float()
        self.head_rows = O.lm_head_weight(weights, dc)[self.valid_ids].float()
        self.calls = {"lm_decode": 0, "frame_tail": 0, "kv_commit": 0}

    # ---- KV ----
    def kv_init(self, total_tokens: int):
        self.kv_pages = (total_tokens + 63) // 64 + 4 * self.B

    def kv_len(self, seq: int) -> int:
        return len(self.kv[seq]) - (1 if self.spec[seq] else 0)

    def kv_set_len(self, seq: int, n: int):
        self.kv[seq].truncate(n)
        self.spec[seq] = False

    def kv_write(self, seq: int, layer: int, pos0: int, k: torch.Tensor, v: torch.Tensor):
        """prefill hand-off: k, v [n_tokens, kv_heads, head_dim] for positions pos0.. of (seq, layer)"""
        assert pos0 == 0 and k.dim() == 3
        self.kv[seq].k[layer] = k.float().transpose(0, 1).contiguous()
        self.kv[seq].v[layer] = v.float().transpose(0, 1).contiguous()

    def kv_delete(self, seq: int, pos: int):
        assert not self.spec[seq]
        kv = self.kv[seq]
        n = len(kv)
        assert 0 <= pos < n
        for l in range(len(kv.k)):
            if kv.k[l] is not None:
                k, v = kv.k[l].clone(), kv.v[l].clone()
                k[:, pos], v[:, pos] = k[:, n - 1], v[:, n - 1]
                kv.k[l], kv.v[l] = k[:, :n - 1], v[:, :n - 1]

    def kv_commit(self, advance):
        assert len(advance) == 2 * self.B
        self.calls["kv_commit"] += 1
        for s, a in enumerate(advance):
            if self.spec[s] and not a:
                self.kv[s].truncate(max(0 if k is None else k.shape[1] for k in self.kv[s].k) - 1)
            self.spec[s] = False

    # ---- programs ----
    @property
    def sde(self) -> bool:
        return self.scheduler is not None and getattr(self.scheduler.config, "algorithm_type", "dpmsolver++") == "sde-dpmsolver++"

    def set_scheduler(self, scheduler):
        self.scheduler = scheduler
        self.n_steps = 0

    def set_diffusion_steps(self, n_steps: int):
        self.n_steps = int(n_steps)
        self.step_noise = torch.zeros(self.n_steps, self.B, 64)

    def upload_step_noise(self, draw, active_rows):
        n = len(active_rows)
        rows = torch.as_tensor(list(active_rows), dtype=torch.long)
        for i in range(self.n_steps):
            self.step_noise[i].index_copy_(0, rows, draw(i)[:n].float())

    def embed_tokens(self, tokens, out: torch.Tensor):
        out[:len(tokens)] = self.embed_w[torch.as_tensor(list(tokens), dtype=torch.long)]

    def lm_decode(self):
        self.calls["lm_decode"] += 1
        dc = self.config.decoder_config
        for s in range(2 * self.B):
            if self.spec[s]:                                 # a second decode without a commit overwrites the speculative slot
                self.kv[s].truncate(len(self.kv[s]) - 1)
            self.hidden[s] = O.qwen2_forward(self.w, dc, self.embeds[s][None].clone(), self.kv[s], len(self.kv[s]))[-1]
            self.spec[s] = True
        self.lm_head(self.hidden)

    # ---- streaming-0.5B split stack (vv_set_row_mode / vv_lm_decode_range) ----
    def set_row_mode(self, modes):
        self.row_mode = [int(m) for m in modes]

    def lm_decode_range(self, layer_begin: int, layer_end: int, final_norm: bool, out=None):
        out = self.hidden if out is None else out
        dc = self.config.decoder_config
        for s in range(2 * self.B):
            if not getattr(self, "row_mode", [1] * (2 * self.B))[s]:
                out[s] = float("nan")                      # the real kernels leave rows with mode 0 undefined: make misuse visible
                continue
            if self.spec[s]:
                self.kv[s].truncate(self._len(s, layer_begin) - 1)
            pos = self._len(s, layer_begin)
            out[s] = O.qwen2_forward(self.w, dc, self.embeds[s][None].clone(), self.kv[s], pos, n_layers=layer_end - layer_begin,
                                     final_norm=final_norm, layer_begin=layer_begin)[-1]
            self.spec[s] = True

    def _len(self, s: int, layer: int) -> int:
        k = self.kv[s].k[layer]
        return 0 if k is None else k.shape[1]

    def lm_head(self, hidden: torch.Tensor):
        lg = hidden[:self.B] @ self.head_rows.T
        self.logits.copy_(lg)
        self.tokens.copy_(torch.tensor(self.valid_ids)[torch.argmax(lg, dim=-1)].to(torch.int32))   # lowest id among ties

    def read_tokens(self):
        return self.tokens.numpy().copy(), self.logits.numpy().copy()

    def stage_audio(self, rows):
        return self.audio[list(rows)].clone()

    def fetch_audio(self, ticket) -> torch.Tensor:
        return ticket

    def lm_logits_full(self) -> torch.Tensor:
        return self.hidden[:self.B] @ O.lm_head_weight(self.w, self.config.decoder_config).float().T

    def upload_frame_inputs(self, noise_rows: torch.Tensor, active_rows):
        self.noise.zero_()
        self.active.zero_()
        for i, b in enumerate(active_rows):
            self.noise[b] = noise_rows[i]
            self.active[b] = 1

    def frame_tail(self, cfg_scale: float):
        self.calls["frame_tail"] += 1
        cfg, w, B = self.config, self.w, self.B
        hc = cfg.diffusion_head_config
        rows = self.active.nonzero().flatten()
        scale, bias = float(w["model.speech_scaling_factor"]), float(w["model.speech_bias_factor"])
        n = rows.numel()
        noise = torch.cat([self.noise[rows], torch.zeros(n, 64)])            # the sampler uses rows [:n] of the 2n draw (:701-704)
        kw = dict(algorithm_type="sde-dpmsolver++", step_noise=[self.step_noise[i][rows] for i in range(self.n_steps)]) if self.sde else {}
        lat = O.sample_speech_tokens(w, self.hidden[rows], self.hidden[B + rows], cfg_scale, self.n_steps, noise, hc.head_layers, hc.rms_norm_eps, **kw)
        audio = O.decoder_frame(w, cfg.acoustic_tokenizer_config, (lat / scale - bias)[:, None, :], self.a_state, rows)
        sem = O.encoder_frame(w, cfg.semantic_tokenizer_config, audio, self.s_state, rows)
        emb = O.connector(w, "model.acoustic_connector", lat) + O.connector(w, "model.semantic_connector", sem[:, 0])
        self.latent[rows] = lat
        self.audio[rows] = audio[:, 0]
        self.embeds[rows] = emb
        self.embeds[B:] = self.embeds[:B]                                     # select_embeds_kernel: negative rows see the same input

    # individual stages (the streaming loop has no semantic encoder, so it calls them one by one)
    def diffusion_sample(self, cfg_scale: float):
        hc = self.config.diffusion_head_config
        rows = self.active.nonzero().flatten()
        n = rows.numel()
        noise = torch.cat([self.noise[rows], torch.zeros(n, 64)])
        self.latent[rows] = O.sample_speech_tokens(self.w, self.hidden[rows], self.hidden[self.B + rows], cfg_scale, self.n_steps, noise,
                                                   hc.head_layers, hc.rms_norm_eps)

    def codec_decode(self):
        rows = self.active.nonzero().flatten()
        scale, bias = float(self.w["model.speech_scaling_factor"]), float(self.w["model.speech_bias_factor"])
        audio = O.decoder_frame(self.w, self.config.acoustic_tokenizer_config, (self.latent[rows] / scale - bias)[:, None, :], self.a_state, rows)
        self.audio[rows] = audio[:, 0]

    def connect(self):
        rows = self.active.nonzero().flatten()
        emb = O.connector(self.w, "model.acoustic_connector", self.latent[rows]) + O.connector(self.w, "model.semantic_connector", self.feat[rows])
        self.embeds[rows] = emb
        self.embeds[self.B:] = self.embeds[:self.B]

    def codec_state_zero(self, rows):
        if len(rows):
            r = torch.as_tensor(list(rows), dtype=torch.long)
            self.a_state.set_to_zero(r)
            self.s_state.set_to_zero(r)

    def codec_state_reset(self):
        self.a_state, self.s_state = O.StreamState(self.B), O.StreamState(self.B)

    def launch_count(self) -> int:
        return 0

    def sync(self):
        pass

    def close(self):
        pass


def make_model(cfg, tok, weights, max_batch: int):
    """The product's `VibeVoiceForConditionalGenerationInference` with a FakeEngine plugged in (no CUDA library is touched)."""
    from vibevoice_b200.modeling import VibeVoiceForConditionalGenerationInference
    m = VibeVoiceForConditionalGenerationInference(cfg, tok, max_batch=max_batch)
    m.engine = FakeEngine(cfg, m._valid_ids(tok), max_batch, weights=weights)
    m._scale, m._bias = float(weights["model.speech_scaling_factor"]), float(weights["model.speech_bias_factor"])
    return m
