"""`-m gpu` parity tests at the sizes that are actually benchmarked / targeted (VERDICT r01 "parity gaps"):

  * 61 440-token context on the 1.5B layer shapes: every split of the split-KV attention is populated, each split walks several
    64-token tiles through the cp.async double buffer, the page table has ~960 entries, the combine merges 128 real partials;
  * the 7B layer shapes (H=3584, I=18 944, 28/4 heads = GQA group 7, untied lm_head) in a closed loop;
  * four prompts per GPU (M = 8 rows) in a closed loop on the 1.5B layer shapes (BASELINE config #4 runs 4 prompts per GPU).

The oracle is the CPU restatement (`oracle/vv_oracle.py`); tolerances as in test_gpu_parity.py.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from vibevoice_b200.configuration import preset_config
from vibevoice_b200.synth import SynthTokenizer, synth_state_dict

from test_gpu_parity import SEED, _scripted, rel_l2, report

PARTS = ("lm", "head", "acoustic_decoder", "semantic", "connectors", "lm_head")


def _model(preset, max_batch, **kw):
    from vibevoice_b200.modeling import VibeVoiceForConditionalGenerationInference
    cfg = preset_config(preset)
    tok = SynthTokenizer(cfg.decoder_config.vocab_size)
    sd = synth_state_dict(cfg, SEED, torch.bfloat16, parts=PARTS)
    m = VibeVoiceForConditionalGenerationInference(cfg, tok, max_batch=max_batch, **kw)
    m.load_state_dict(sd, tok)
    return m, cfg, tok, sd


def _structured_kv(nkv, L, hd, g):
    """A synthetic prefix whose attention output depends on every region of the context: unit-variance keys (peaky softmax at
    |q| ~ 10) and values that carry a slow position-dependent ramp, so a dropped / duplicated split or tile moves the output by
    far more than the tolerance."""
    k = torch.randn(nkv, L, hd, generator=g)
    t = torch.arange(L, dtype=torch.float32)[None, :, None]
    d = torch.arange(hd, dtype=torch.float32)[None, None, :]
    v = torch.randn(nkv, L, hd, generator=g) + 2.0 * torch.sin(t / 997.0 + 0.37 * d)
    return k.to(torch.bfloat16), v.to(torch.bfloat16)


@pytest.mark.parametrize("ctx_pos,ctx_neg", [(61440, 777), (8191, 64), (12345, 0)])
def test_lm_decode_long_context_vs_oracle(ctx_pos, ctx_neg):
    """Positive row at `ctx_pos` tokens, negative row at `ctx_neg`: 4 decode steps against `qwen2_forward` over the same bf16
    prefix (imported through vv_kv_write on the CUDA side, preloaded into the oracle's cache on the CPU side)."""
    from oracle import vv_oracle as O
    model, cfg, tok, sd = _model("1.5b-l2", 1)
    try:
        eng = model.engine
        dc = cfg.decoder_config
        nl, nkv, hd = dc.num_hidden_layers, dc.num_key_value_heads, dc.head_dim
        eng.kv_init(ctx_pos + ctx_neg + 256)
        g = torch.Generator().manual_seed(ctx_pos)
        caches = [O.KVCache(nl, kv_bf16=True), O.KVCache(nl, kv_bf16=True)]
        for seq, L in ((0, ctx_pos), (1, ctx_neg)):
            eng.kv_set_len(seq, 0)
            if L == 0:
                continue
            for l in range(nl):
                k, v = _structured_kv(nkv, L, hd, g)
                caches[seq].preload(l, k.float(), v.float())
                kd, vd = k.transpose(0, 1).contiguous().cuda(), v.transpose(0, 1).contiguous().cuda()      # [L, nkv, hd]
                with torch.cuda.stream(eng.stream):
                    eng.kv_write(seq, l, 0, kd, vd)
                eng.sync()
            eng.kv_set_len(seq, L)
        assert eng.kv_len(0) == ctx_pos and eng.kv_len(1) == ctx_neg
        errs = []
        for step in range(4):
            x = torch.randn(2, dc.hidden_size, generator=g) * 0.05
            with torch.cuda.stream(eng.stream):
                eng.embeds.copy_(x.cuda())
            eng.lm_decode()
            eng.read_tokens()
            adv = [1, step % 2]
            want = []
            for r in range(2):
                n0 = len(caches[r])
                want.append(O.qwen2_forward(sd, dc, x[r][None], caches[r], n0)[0])
                if not adv[r]:
                    caches[r].truncate(n0)
            eng.kv_commit(adv)
            got = eng.hidden.cpu()
            errs.append([rel_l2(got[r], want[r]) for r in range(2)])
        report("lm_decode_long_context", ctx_pos=ctx_pos, ctx_neg=ctx_neg, rel_l2=errs)
        assert max(max(e) for e in errs) < 2e-3, errs
        assert eng.kv_len(0) == ctx_pos + 4 and eng.kv_len(1) == ctx_neg + 2
    finally:
        model.engine.close()


def test_generate_torch_prefill_multi_thousand_token_prompt():
    """The TorchPrefill -> vv_kv_write hand-off at scale: a 5 000-token prompt prefilled on library kernels (bf16) and imported into the
    paged pool, then the CUDA loop; the oracle prefills in fp32 -> same looser audio tolerance as the voice-prompt test."""
    from oracle import vv_oracle as O
    from vibevoice_b200.modeling import ForcedTokenScript
    model, cfg, tok, sd = _model("1.5b-l2", 1, torch_prefill=True)
    try:
        dc = cfg.decoder_config
        g = torch.Generator().manual_seed(21)
        ids = torch.randint(0, dc.vocab_size - 20, (1, 5000), generator=g)
        ids[:, -1] = tok.speech_start_id
        script = [_scripted(tok, "dddx")]
        model.set_ddpm_inference_steps(5)
        torch.manual_seed(0)
        out = model.generate(input_ids=ids, tokenizer=tok, cfg_scale=1.3, is_prefill=False, logits_processor=[ForcedTokenScript(script)],
                             max_new_tokens=8, show_progress_bar=False)
        torch.manual_seed(0)
        ref = O.generate(sd, cfg, ids, None, tok, cfg_scale=1.3, num_steps=5, max_new_tokens=8, forced_tokens=script, kv_bf16=True)
        assert torch.equal(out.sequences, ref.sequences)
        e = rel_l2(out.speech_outputs[0].cpu(), ref.speech_outputs[0])
        report("generate_torch_prefill_5000", audio_rel_l2=e)
        assert e < 5e-2, e
    finally:
        model.engine.close()


def test_7b_shapes_closed_loop_vs_oracle():
    """VibeVoice-7B layer shapes, 2 LM layers: GQA group 7 in the attention kernel, K = 3584 / 18 944 / 10 752 weight streams,
    untied lm_head, 3584-wide connectors and diffusion head; speaker turn in the script."""
    from oracle import vv_oracle as O
    from vibevoice_b200.modeling import ForcedTokenScript
    model, cfg, tok, sd = _model("7b-l2", 1)
    try:
        dc = cfg.decoder_config
        assert dc.hidden_size == 3584 and dc.num_attention_heads // dc.num_key_value_heads == 7 and not dc.tie_word_embeddings
        g = torch.Generator().manual_seed(6)
        ids = torch.randint(0, dc.vocab_size - 20, (1, 70), generator=g)
        ids[:, -1] = tok.speech_start_id
        script = [_scripted(tok, "ddesdx")]
        model.set_ddpm_inference_steps(10)
        torch.manual_seed(0)
        out = model.generate(input_ids=ids, tokenizer=tok, cfg_scale=1.3, is_prefill=False, logits_processor=[ForcedTokenScript(script)],
                             max_new_tokens=12, show_progress_bar=False)
        torch.manual_seed(0)
        ref = O.generate(sd, cfg, ids, None, tok, cfg_scale=1.3, num_steps=10, max_new_tokens=12, forced_tokens=script, kv_bf16=True)
        assert torch.equal(out.sequences, ref.sequences)
        a, b = out.speech_outputs[0].cpu(), ref.speech_outputs[0]
        assert a.shape == b.shape == (1, 3 * 3200)
        e = rel_l2(a, b)
        report("generate_real_shapes_7b", audio_rel_l2=e)
        assert e < 1e-2, e
        # free-running constrained argmax on the untied head for a few steps (token ids must agree unless the margin is a tie)
        torch.manual_seed(1)
        out = model.generate(input_ids=ids, tokenizer=tok, cfg_scale=1.3, is_prefill=False, max_new_tokens=4, show_progress_bar=False)
        torch.manual_seed(1)
        ref = O.generate(sd, cfg, ids, None, tok, cfg_scale=1.3, num_steps=10, max_new_tokens=4, kv_bf16=True, trace=True)
        margins = [float(l.sort(dim=-1).values[0, -1] - l.sort(dim=-1).values[0, -2]) for l in ref.trace["logits"]]
        first_tie = next((i for i, m in enumerate(margins) if m < 1e-3), len(margins))
        upto = min(out.sequences.shape[1], ref.sequences.shape[1], ids.shape[1] + first_tie)
        assert torch.equal(out.sequences[:, :upto], ref.sequences[:, :upto])
    finally:
        model.engine.close()


def test_batch4_closed_loop_vs_oracle():
    """Four prompts per GPU on the 1.5B layer shapes (M = 8 rows through every LM / sampler stage): ragged left-padded prompts, rows
    that finish at different steps, one speaker turn, one row that never diffuses after the first frame."""
    from oracle import vv_oracle as O
    from vibevoice_b200.modeling import ForcedTokenScript
    model, cfg, tok, sd = _model("1.5b-l2", 4)
    try:
        dc = cfg.decoder_config
        g = torch.Generator().manual_seed(9)
        L0 = 40
        ids = torch.randint(0, dc.vocab_size - 20, (4, L0), generator=g)
        ids[:, -1] = tok.speech_start_id
        mask = torch.ones(4, L0, dtype=torch.long)
        for r, pad in enumerate((0, 7, 0, 19)):
            mask[r, :pad] = 0
            ids[r, :pad] = tok.pad_token_id
        scripts = [_scripted(tok, "dddesddx"), _scripted(tok, "ddddddx"), _scripted(tok, "dx"), _scripted(tok, "dddddddx")]
        model.set_ddpm_inference_steps(10)
        torch.manual_seed(0)
        out = model.generate(input_ids=ids, attention_mask=mask, tokenizer=tok, cfg_scale=1.3, is_prefill=False,
                             logits_processor=[ForcedTokenScript(scripts)], max_new_tokens=20, show_progress_bar=False)
        torch.manual_seed(0)
        ref = O.generate(sd, cfg, ids, mask, tok, cfg_scale=1.3, num_steps=10, max_new_tokens=20, forced_tokens=scripts, kv_bf16=True)
        assert torch.equal(out.sequences, ref.sequences)
        assert torch.equal(out.reach_max_step_sample, ref.reach_max_step_sample)
        for r in range(4):
            a, b = out.speech_outputs[r].cpu(), ref.speech_outputs[r]
            assert a.shape == b.shape == (1, 3200 * scripts[r].count(tok.speech_diffusion_id))
            e = rel_l2(a, b)
            report("generate_batch4_1.5b_shapes", row=r, audio_rel_l2=e)
            assert e < 1e-2, (r, e)
    finally:
        model.engine.close()


def test_head_dim_64_lm_decode_vs_oracle():
    """head_dim 64 (the streaming-0.5B attention geometry) through the weight-stream path: K / V pages are one 64-column TMA box, the
    o-projection merges one head per k-block.  Crosses a page boundary; both rows long enough to have several partials."""
    from test_gpu_parity import _lm_roundtrip
    model, cfg, tok, sd = _model("tiny64", 2)
    try:
        assert cfg.decoder_config.head_dim == 64
        errs = _lm_roundtrip(model, cfg, tok, sd, n_prompt=70, n_steps=6)
        report("lm_decode_head_dim_64", max_rel_l2=max(errs))
        assert max(errs) < 2e-3, errs
    finally:
        model.engine.close()


def test_streaming_variant_real_05b_shapes_vs_oracle():
    """SURVEY 8f-1 at the real VibeVoice-Streaming-0.5B layer shapes (H = 896, 14 / 2 heads of 64, I = 4864; 1 text + 3 TTS layers here):
    the split-stack loop against `oracle/vv_streaming.py` (pinned to the reference's own streaming generate())."""
    from oracle import vv_streaming as VS
    from vibevoice_b200 import streaming as S
    cfg = preset_config("streaming-0.5b-l4")
    base = synth_state_dict(cfg, SEED, torch.bfloat16, parts=("lm", "head", "acoustic_decoder", "connectors", "lm_head"))
    sd = VS.streaming_state_dict(base, cfg, 3, eos_bias=-6.0)
    m = S.VibeVoiceStreamingForConditionalGenerationInference(cfg, tts_backbone_num_hidden_layers=3)
    m.load_state_dict(sd)
    try:
        m.set_ddpm_inference_steps(5)
        g = torch.Generator().manual_seed(7)
        prompt = torch.randint(0, 2000, (40,), generator=g)
        text = torch.randint(0, 2000, (12,), generator=g)
        torch.manual_seed(0)
        out = m.generate(input_ids=prompt[None], tts_text_ids=text[None], neg_text_input_id=2047, cfg_scale=1.5, max_new_tokens=24)
        torch.manual_seed(0)
        ref = VS.generate_streaming(sd, cfg, 3, prompt, text, 2047, cfg_scale=1.5, num_steps=5, max_new_tokens=24, kv_bf16=True)
        assert torch.equal(out.sequences, ref.sequences), (out.sequences, ref.sequences)
        e = rel_l2(out.speech_outputs[0].cpu(), ref.speech_outputs[0])
        report("streaming_generate_0.5b_shapes", audio_rel_l2=e)
        assert e < 1e-2, e
    finally:
        m.engine.close()


def test_batch8_tiny_closed_loop_vs_oracle():
    """The largest batch the engine accepts (8 prompts -> 16 LM rows: MMA N = 32 in the weight-stream kernel, 16 attention rows, 8-row codec
    stages): ragged prompts, rows finishing at different steps."""
    from oracle import vv_oracle as O
    from vibevoice_b200.modeling import ForcedTokenScript
    model, cfg, tok, sd = _model("tiny", 8)
    try:
        dc = cfg.decoder_config
        g = torch.Generator().manual_seed(13)
        L0 = 14
        ids = torch.randint(0, dc.vocab_size - 20, (8, L0), generator=g)
        ids[:, -1] = tok.speech_start_id
        mask = torch.ones(8, L0, dtype=torch.long)
        for r in range(8):
            mask[r, :r] = 0
            ids[r, :r] = tok.pad_token_id
        plans = ["ddddx", "dddesddx", "dx", "ddddddx", "ddx", "dddx", "desdx", "dddddx"]
        scripts = [_scripted(tok, p) for p in plans]
        model.set_ddpm_inference_steps(5)
        torch.manual_seed(0)
        out = model.generate(input_ids=ids, attention_mask=mask, tokenizer=tok, cfg_scale=1.3, is_prefill=False,
                             logits_processor=[ForcedTokenScript(scripts)], max_new_tokens=16, show_progress_bar=False)
        torch.manual_seed(0)
        ref = O.generate(sd, cfg, ids, mask, tok, cfg_scale=1.3, num_steps=5, max_new_tokens=16, forced_tokens=scripts, kv_bf16=True)
        assert torch.equal(out.sequences, ref.sequences)
        for r in range(8):
            e = rel_l2(out.speech_outputs[r].cpu(), ref.speech_outputs[r])
            report("generate_batch8_tiny", row=r, audio_rel_l2=e)
            assert e < 1e-2, (r, e)
    finally:
        model.engine.close()
