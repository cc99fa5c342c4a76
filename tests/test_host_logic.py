"""CPU tests: C-ABI surface, scheduler tables, prompt sharding + gather over gloo (world_size 2), config/synth plumbing."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_abi_exports_every_declared_symbol():
    """include/vibevoice_b200.h <-> libvibevoice_b200.so <-> ctypes table (no compute calls, no GPU needed)."""
    from vibevoice_b200 import _native as N
    hdr = open(os.path.join(ROOT, "include", "vibevoice_b200.h")).read()
    declared = set(re.findall(r"\b(vv_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"vv_ctx", "vv_model_desc", "vv_status", "vv_dtype"}
    assert len(declared) >= 25
    lib = N.load_library()
    for name in declared:
        assert hasattr(lib, name), "missing export %s" % name
    assert declared == set(N.SYMBOLS), (declared ^ set(N.SYMBOLS))
    assert lib.vv_abi_version() == 1


def test_compute_fails_loudly_without_gpu():
    from vibevoice_b200 import _native as N
    from vibevoice_b200.configuration import preset_config
    from vibevoice_b200.engine import Engine
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(N.VVError):
        Engine(preset_config("tiny"), [1, 2, 3, 4])


@pytest.mark.parametrize("n", [5, 10, 20, 30])
def test_product_schedule_equals_oracle_tables(n):
    from oracle import vv_oracle as O
    from vibevoice_b200.schedule import DPMSolverMultistepScheduler
    s = DPMSolverMultistepScheduler(num_train_timesteps=1000, beta_schedule="cosine", prediction_type="v_prediction").set_timesteps(n)
    t = O.dpm_tables(n)
    assert np.array_equal(s.timesteps.numpy(), t.timesteps)
    assert np.array_equal(s.sigmas.numpy(), t.sigmas)
    for j, a in enumerate([t.a0, t.s0, t.ks, t.kx, t.rinv]):
        assert np.array_equal(s.coef[:, j], a)
    assert np.array_equal(s.coef[:, 5].astype(np.int32), t.order)
    # sde-dpmsolver++ (seven columns): the oracle's tables are bit-exact against the reference scheduler (test_oracle_golden.py)
    s = DPMSolverMultistepScheduler(algorithm_type="sde-dpmsolver++", beta_schedule="squaredcos_cap_v2").set_timesteps(n)
    t = O.dpm_tables(n, algorithm_type="sde-dpmsolver++")
    for j, a in enumerate([t.a0, t.s0, t.ks, t.kx, t.rinv, t.order.astype(np.float32), t.kn]):
        assert np.array_equal(s.coef[:, j], a), j


def test_scheduler_rejects_variants_off_the_path():
    from vibevoice_b200.schedule import DPMSolverMultistepScheduler
    with pytest.raises(NotImplementedError):
        DPMSolverMultistepScheduler(algorithm_type="dpmsolver")
    with pytest.raises(NotImplementedError):
        DPMSolverMultistepScheduler(solver_order=3)
    sde = DPMSolverMultistepScheduler(algorithm_type="sde-dpmsolver++").set_timesteps(10)
    assert sde.coef.shape == (10, 7) and sde.coef[-1, 6] == 0.0 and (sde.coef[:-1, 6] > 0).all()
    s = DPMSolverMultistepScheduler()
    s2 = DPMSolverMultistepScheduler.from_config(s.config)
    assert s2.config.solver_order == 2


def test_configs_match_the_shipped_jsons():
    from vibevoice_b200.configuration import VibeVoiceConfig, preset_config
    from vibevoice_b200.synth import param_specs
    import math
    for name, n_params in (("1.5b", 2704021985), ("7b", 9343355361)):
        cfg = preset_config(name)
        assert sum(math.prod(s) for _, s, _ in param_specs(cfg)) == n_params
    c = preset_config("1.5b")
    assert c.acoustic_tokenizer_config.decoder_depth_list == [8, 3, 3, 3, 3, 3, 3]
    ref = "/root/reference/vibevoice/configs/qwen2.5_1.5b_64k.json"
    if os.path.exists(ref):
        r = VibeVoiceConfig.from_pretrained(ref)
        assert r.decoder_config.hidden_size == 1536 and r.decoder_config.num_key_value_heads == 2
        assert r.diffusion_head_config.head_layers == 4 and r.semantic_vae_dim == 128


def test_shard_prompts():
    from vibevoice_b200.distributed import shard_prompts
    for n, w in ((32, 8), (7, 2), (3, 4), (0, 2)):
        seen = []
        for r in range(w):
            seen += shard_prompts(n, r, w)
        assert seen == list(range(n))
    assert shard_prompts(32, 3, 8) == [12, 13, 14, 15]


def _gather_worker(rank, world, port, q):
    import torch.distributed as dist
    from vibevoice_b200.distributed import gather_waveforms, shard_prompts
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    mine = shard_prompts(5, rank, world)
    wavs = [None if i == 3 else torch.full((1, 3200 * (i + 1)), float(i)) for i in mine]
    res = gather_waveforms(wavs)
    if rank == 0:
        flat = [w for row in res for w in row]
        q.put([(None if w is None else (tuple(w.shape), float(w.mean()))) for w in flat])
    dist.destroy_process_group()


def test_gather_waveforms_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    ps = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = q.get(timeout=120)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    assert out == [((1, 3200), 0.0), ((1, 6400), 1.0), ((1, 9600), 2.0), None, ((1, 16000), 4.0)]


def test_streamer_contract():
    from vibevoice_b200.streamer import AudioStreamer
    st = AudioStreamer(batch_size=2, stop_signal=None)
    st.put(torch.ones(2, 1, 4), torch.tensor([0, 1]))
    st.end(torch.tensor([1]))
    assert st.finished_flags == [False, True]
    st.put(torch.ones(1, 1, 4) * 2, torch.tensor([1]))          # ignored: finished
    st.end()
    assert [c.sum().item() for c in st.get_stream(0)] == [4.0]
    assert len(list(st.get_stream(1))) == 1


def test_sample_valid_tokens_distribution():
    """do_sample host logic (reference :493-496): draws follow softmax over the valid ids, ids come back in vocabulary space,
    and the draw does not consume the global CPU RNG that the diffusion noise uses (:701)."""
    from vibevoice_b200.modeling import sample_valid_tokens
    valid = [7, 11, 13, 20]
    logits = np.log(np.array([[0.1, 0.2, 0.3, 0.4]], dtype=np.float32)).repeat(4000, 0)
    torch.manual_seed(5)
    before = torch.get_rng_state().clone()
    toks = sample_valid_tokens(logits, valid, torch.Generator().manual_seed(0))
    assert torch.equal(before, torch.get_rng_state())
    assert set(toks.tolist()) <= set(valid)
    freq = np.array([(toks == v).mean() for v in valid])
    assert np.abs(freq - np.array([0.1, 0.2, 0.3, 0.4])).max() < 0.03
    hot = np.full((3, 4), -np.inf, dtype=np.float32)
    hot[0, 2] = hot[1, 0] = hot[2, 3] = 0.0
    assert sample_valid_tokens(hot, valid, torch.Generator().manual_seed(1)).tolist() == [13, 7, 20]


def _write_adapter(d, tensors, r, alpha, **extra):
    import json
    from safetensors.torch import save_file
    d.mkdir(parents=True, exist_ok=True)
    (d / "adapter_config.json").write_text(json.dumps(dict(peft_type="LORA", r=r, lora_alpha=alpha, **extra)))
    save_file(tensors, str(d / "adapter_model.safetensors"))


def test_lora_assets_fold_into_base_weights(tmp_path):
    """lora.py vs the published PEFT LoRA merge W' = W + (alpha/r) B A, on the directory layout the reference's loader reads
    (`lora_loading.py:46-55, 72-127, 164-170`): LM adapter at the root, head adapter under diffusion_head/ behind the `base.` shim,
    connector state dicts replaced wholesale; unknown targets raise instead of being dropped."""
    from vibevoice_b200 import lora as L
    g = torch.Generator().manual_seed(0)
    rn = lambda *s: torch.randn(*s, generator=g)
    base = {
        "model.language_model.layers.0.self_attn.q_proj.weight": rn(16, 8).bfloat16(),
        "model.language_model.layers.0.self_attn.q_proj.bias": rn(16).bfloat16(),
        "model.language_model.layers.0.mlp.down_proj.weight": rn(8, 24).bfloat16(),
        "model.prediction_head.layers.0.ffn.gate_proj.weight": rn(24, 8).bfloat16(),
        "model.acoustic_connector.fc1.weight": rn(8, 4).bfloat16(),
        "model.acoustic_connector.fc1.bias": rn(8).bfloat16(),
    }
    root = tmp_path / "ckpt" / "lora"
    Aq, Bq, Ad, Bd = rn(4, 8), rn(16, 4), rn(2, 24), rn(8, 2)
    _write_adapter(root, {"base_model.model.layers.0.self_attn.q_proj.lora_A.weight": Aq,
                          "base_model.model.layers.0.self_attn.q_proj.lora_B.weight": Bq,
                          "base_model.model.layers.0.mlp.down_proj.lora_A.default.weight": Ad,
                          "base_model.model.layers.0.mlp.down_proj.lora_B.default.weight": Bd}, r=4, alpha=8,
                   rank_pattern={"down_proj": 2}, alpha_pattern={"down_proj": 6})
    Ah, Bh = rn(4, 8), rn(24, 4)
    _write_adapter(root / "diffusion_head", {"base_model.model.base.layers.0.ffn.gate_proj.lora_A.weight": Ah,
                                             "base_model.model.base.layers.0.ffn.gate_proj.lora_B.weight": Bh}, r=4, alpha=4, use_rslora=True)
    (root / "acoustic_connector").mkdir()
    new_fc1 = {"fc1.weight": rn(8, 4), "fc1.bias": rn(8)}
    torch.save(new_fc1, root / "acoustic_connector" / "pytorch_model.bin")
    deltas, repl, rep = L.collect_overrides(tmp_path / "ckpt")
    assert rep.language_model and rep.diffusion_head_lora and rep.acoustic_connector and not rep.semantic_connector and not rep.diffusion_head_full
    assert rep.adapter_root == root
    out = dict(L.merged_state_dict(base.items(), deltas, repl))
    want = {
        "model.language_model.layers.0.self_attn.q_proj.weight": (base["model.language_model.layers.0.self_attn.q_proj.weight"].float() + (8 / 4) * Bq @ Aq).bfloat16(),
        "model.language_model.layers.0.mlp.down_proj.weight": (base["model.language_model.layers.0.mlp.down_proj.weight"].float() + (6 / 2) * Bd @ Ad).bfloat16(),
        "model.prediction_head.layers.0.ffn.gate_proj.weight": (base["model.prediction_head.layers.0.ffn.gate_proj.weight"].float() + (4 / 2.0) * Bh @ Ah).bfloat16(),
        "model.acoustic_connector.fc1.weight": new_fc1["fc1.weight"].bfloat16(),
        "model.acoustic_connector.fc1.bias": new_fc1["fc1.bias"].bfloat16(),
        "model.language_model.layers.0.self_attn.q_proj.bias": base["model.language_model.layers.0.self_attn.q_proj.bias"],
    }
    assert set(out) == set(base)
    for k, v in want.items():
        assert out[k].dtype == torch.bfloat16 and torch.equal(out[k], v), k
    # an adapter that targets a tensor the checkpoint lacks must not be dropped silently
    small = {k: v for k, v in base.items() if "down_proj" not in k}
    with pytest.raises(KeyError):
        list(L.merged_state_dict(small.items(), deltas, repl))
    # full-head fallback is only read when no head adapter is present (:97-112)
    import shutil
    shutil.rmtree(root / "diffusion_head")
    torch.save({"layers.0.ffn.gate_proj.weight": rn(24, 8)}, root / "diffusion_head_full.bin")
    _, repl2, rep2 = L.collect_overrides(tmp_path / "ckpt")
    assert rep2.diffusion_head_full and not rep2.diffusion_head_lora and "model.prediction_head.layers.0.ffn.gate_proj.weight" in repl2
    with pytest.raises(FileNotFoundError):
        L.collect_overrides(tmp_path / "nope" / "lora")
    from vibevoice.modular.lora_loading import load_lora_assets  # noqa: F401  (drop-in import path)


@pytest.mark.parametrize("case", ["scripted", "free", "maxlen", "quirk", "norefresh1", "norefresh", "sde"])
def test_product_generate_host_logic_against_reference_generate_fixture(golden, case):
    """`modeling.generate` (the product's host state machine, a-1/a-2/a-8: token bookkeeping, which KV entries the negative stream
    keeps, restart on <speech_start>, codec-state zeroing, per-row finishing, noise-row packing) driven through a CPU stand-in of the
    engine (`tests/fake_engine.py`, oracle arithmetic) and held to what the REFERENCE's own generate() produced on the same
    checkpoint (`tests/golden/loop.pt`).  Sequences and flags exact; audio 1e-5."""
    from fake_engine import make_model
    from vibevoice_b200.configuration import preset_config
    from vibevoice_b200.modeling import ForcedTokenScript
    from vibevoice_b200.synth import SynthTokenizer, synth_state_dict
    g = golden("loop")
    c = g[case]
    cfg = preset_config(g["preset"])
    tok = SynthTokenizer(cfg.decoder_config.vocab_size)
    sd = synth_state_dict(cfg, 1234, torch.float32)
    model = make_model(cfg, tok, sd, max_batch=c["ids"].shape[0])
    model.set_ddpm_inference_steps(g["num_steps"])
    if c.get("algorithm_type") == "sde-dpmsolver++":            # the way demo/gradio_demo.py:141-146 switches solvers
        from vibevoice_b200.schedule import DPMSolverMultistepScheduler
        base = DPMSolverMultistepScheduler()
        model.model.noise_scheduler = base.from_config(base.config, algorithm_type="sde-dpmsolver++", beta_schedule="squaredcos_cap_v2")
    torch.manual_seed(c["seed"])
    out = model.generate(input_ids=c["ids"], attention_mask=c["mask"], tokenizer=tok, cfg_scale=g["cfg_scale"], is_prefill=False,
                         max_new_tokens=c["max_new_tokens"], max_length_times=c["max_length_times"], show_progress_bar=False,
                         logits_processor=[ForcedTokenScript(c["scripts"])] if c["scripts"] else None,
                         refresh_negative=c["refresh_negative"])
    assert torch.equal(out.sequences, c["sequences"])
    assert torch.equal(out.reach_max_step_sample, c["reach_max"])
    for r, (a, b) in enumerate(zip(out.speech_outputs, c["audio"])):
        assert (a is None) == (b is None)
        if a is not None:
            assert a.shape == b.shape
            rel = float((a.double() - b.double()).norm() / b.double().norm())
            if case == "quirk" and r == 0:
                # stated deviation: the ill-formed d,e,d row keeps a different negative context in the reference
                # (guard off-by-one at :603/:613); the product drops the newest entry.  Same tokens, same length, different audio.
                assert 1e-3 < rel < 0.2, rel
            else:
                assert rel < 1e-5, rel
    assert model.engine.calls["frame_tail"] > 0 or case == "free" and all(a is None for a in c["audio"])


@pytest.mark.parametrize("case", ["streamed", "stopped"])
def test_product_streaming_and_stop_hooks_against_reference_generate_fixture(golden, case):
    """Boundary behaviour (b): `audio_streamer` hand-off and `stop_check_fn` through the product's generate() + AudioStreamer vs what the
    reference's generate() + its own AudioStreamer did (fixture): same chunks in the same order in every per-row queue, the stop
    signal where the reference put it, the loop ending as soon as any row's stream is finished (:443-447), same sequences / audio."""
    from fake_engine import make_model
    from vibevoice_b200.configuration import preset_config
    from vibevoice_b200.modeling import ForcedTokenScript
    from vibevoice_b200.streamer import AudioStreamer
    from vibevoice_b200.synth import SynthTokenizer, synth_state_dict
    g = golden("loop")
    c = g[case]
    cfg = preset_config(g["preset"])
    tok = SynthTokenizer(cfg.decoder_config.vocab_size)
    model = make_model(cfg, tok, synth_state_dict(cfg, 1234, torch.float32), max_batch=2)
    model.set_ddpm_inference_steps(g["num_steps"])
    st = AudioStreamer(batch_size=2)
    extra = {}
    if c["stop_after_calls"] is not None:
        calls = {"n": 0}

        def stop_fn():
            calls["n"] += 1
            return calls["n"] > c["stop_after_calls"]
        extra["stop_check_fn"] = stop_fn
    torch.manual_seed(c["seed"])
    out = model.generate(input_ids=c["ids"], attention_mask=c["mask"], tokenizer=tok, cfg_scale=g["cfg_scale"], is_prefill=False,
                         max_new_tokens=c["max_new_tokens"], show_progress_bar=False, audio_streamer=st,
                         logits_processor=[ForcedTokenScript(c["scripts"])], **extra)
    assert torch.equal(out.sequences, c["sequences"])
    assert torch.equal(out.reach_max_step_sample, c["reach_max"])
    for a, b in zip(out.speech_outputs, c["audio"]):
        assert a.shape == b.shape and float((a.double() - b.double()).norm() / b.double().norm()) < 1e-5
    for r in range(2):
        got = []
        while not st.audio_queues[r].empty():
            got.append(st.audio_queues[r].get())
        want = c["streamed"][r]
        assert len(got) == len(want), (r, len(got), len(want))
        for x, y in zip(got, want):
            assert (x is None) == (y is None)                      # the stop signal sits where the reference put it
            if x is not None:
                assert tuple(x.shape) == tuple(y.shape)
                assert float((x.double() - y.double()).norm() / y.double().norm()) < 1e-5


def test_from_pretrained_reads_an_hf_checkpoint_directory(tmp_path, monkeypatch, golden):
    """Boundary (b): `from_pretrained(dir)` on the checkpoint layout the reference loads (`demo/inference_from_file.py:295-332`):
    `config.json` in the shipped format + sharded `*.safetensors` with the reference's key names -- tied `lm_head.weight` absent,
    acoustic *encoder* and `fix_std` tensors present but off the path, scaling/bias factors as 0-d buffers.  The engine is replaced by
    the CPU stand-in, so this exercises the real file reading / key routing code; the loaded model must generate exactly what a model
    given the same state dict in memory generates."""
    import json
    from safetensors.torch import save_file
    import fake_engine
    from vibevoice_b200 import modeling
    from vibevoice_b200.configuration import preset_config
    from vibevoice_b200.modeling import ForcedTokenScript, VibeVoiceForConditionalGenerationInference
    from vibevoice_b200.synth import SynthTokenizer, synth_state_dict
    cfg = preset_config("tiny")
    tok = SynthTokenizer(cfg.decoder_config.vocab_size)
    sd = synth_state_dict(cfg, 1234, torch.bfloat16)
    assert "lm_head.weight" not in sd or cfg.decoder_config.tie_word_embeddings is False
    ck = tmp_path / "ckpt"
    ck.mkdir()
    d = cfg.to_dict()
    d["model_type"] = "vibepod"                              # the shipped JSONs carry this (configs/qwen2.5_1.5b_64k.json:37)
    (ck / "config.json").write_text(json.dumps(d))
    keys = sorted(sd)
    half = len(keys) // 2
    for i, part in enumerate((keys[:half], keys[half:])):
        save_file({k: sd[k].contiguous() for k in part}, str(ck / ("model-%05d-of-00002.safetensors" % (i + 1))))
    monkeypatch.setattr(modeling, "Engine", fake_engine.FakeEngine)
    m = VibeVoiceForConditionalGenerationInference.from_pretrained(str(ck), torch_dtype=torch.bfloat16, device_map="cuda:0", tokenizer=tok,
                                                                   torch_prefill=False)   # exact comparison: same prompt ingestion on both sides
    assert m.engine.finalized and m.config.decoder_config.hidden_size == cfg.decoder_config.hidden_size
    assert abs(float(m.model.speech_scaling_factor) - float(sd["model.speech_scaling_factor"])) < 1e-6
    ref = fake_engine.make_model(cfg, tok, sd, max_batch=1)
    c = golden("loop")["free"]
    outs = []
    for model in (m, ref):
        model.set_ddpm_inference_steps(5)
        torch.manual_seed(1)
        outs.append(model.generate(input_ids=c["ids"], tokenizer=tok, cfg_scale=1.3, is_prefill=False, max_new_tokens=6, show_progress_bar=False))
    assert torch.equal(outs[0].sequences, outs[1].sequences)
    assert torch.equal(outs[0].speech_outputs[0], outs[1].speech_outputs[0])


@pytest.mark.parametrize("case", ["eos", "windows", "short"])
def test_streaming_product_host_logic_against_reference_fixture(golden, monkeypatch, case):
    """`vibevoice_b200/streaming.py` (split stack over `lm_decode_range`, type embeddings, EOS classifier, text/speech
    windows, zero-semantic connector) through the engine stand-in against the fixtures of the reference's own streaming generate()."""
    import fake_engine
    from oracle import vv_streaming as VS
    from vibevoice_b200 import streaming as S
    from vibevoice_b200.configuration import preset_config
    from vibevoice_b200.synth import synth_state_dict
    g = golden("streaming")
    c = g[case]
    cfg = preset_config(g["preset"])
    sd = VS.streaming_state_dict(synth_state_dict(cfg, 1234, torch.float32), cfg, g["tts_layers"], eos_bias=c["eos_bias"])
    m = S.VibeVoiceStreamingForConditionalGenerationInference(cfg, tts_backbone_num_hidden_layers=g["tts_layers"])
    monkeypatch.setattr(m, "_new_engine", lambda: fake_engine.FakeEngine(cfg, [0, 1], 2))
    m.load_state_dict(sd)
    m.set_ddpm_inference_steps(g["num_steps"])
    from types import SimpleNamespace
    prefilled = {k: SimpleNamespace(past_key_values=tuple(v["kv"]), last_hidden_state=v["hidden"]) for k, v in c["prefilled"].items()}
    # prompt state from ids (fp32 throughout: 1e-5) / imported from the reference's cached-prompt format (K/V stored as bf16, like the pool)
    for extra, tol in (({}, 1e-5), ({"all_prefilled_outputs": prefilled, "tts_lm_input_ids": c["prompt"][None]}, 5e-3)):
        torch.manual_seed(c["seed"])
        out = m.generate(input_ids=c["prompt"][None], tts_text_ids=c["text"][None], neg_text_input_id=g["neg_id"], cfg_scale=c["cfg_scale"],
                         max_new_tokens=c["max_new_tokens"], **extra)
        assert torch.equal(out.sequences, c["sequences"])
        assert torch.equal(out.reach_max_step_sample, c["reach_max"])
        a, b = out.speech_outputs[0], c["audio"]
        assert (a is None) == (b is None)
        if a is not None:
            assert a.shape == b.shape and float((a.double() - b.double()).norm() / b.double().norm()) < tol


def test_streaming_from_pretrained_refuses_non_cuda_devices(tmp_path):
    """`demo/streaming_inference_from_file.py:259-262` passes device_map="cpu" on CPU hosts: there is no CPU path, and the refusal must come
    before any weights are touched."""
    import json
    from vibevoice.modular.modeling_vibevoice_streaming_inference import VibeVoiceStreamingForConditionalGenerationInference as M
    from vibevoice_b200 import _native as N
    from vibevoice_b200.configuration import preset_config
    d = preset_config("tiny").to_dict()
    d["tts_backbone_num_hidden_layers"] = 1
    (tmp_path / "config.json").write_text(json.dumps(d))
    with pytest.raises(N.VVError):
        M.from_pretrained(str(tmp_path), device_map="cpu")


@pytest.mark.parametrize("seed", list(range(16)))
def test_product_negative_stream_bookkeeping_on_random_scripts(seed):
    """The product's host state machine (fake engine, oracle arithmetic) against `oracle.generate` -- which reproduces the reference's own
    generate() on the committed fixtures -- on RANDOM forced-token scripts: batches of 2-3 ragged rows, speaker turns, early EOS, and with
    `refresh_negative=False` also ill-formed orders (diffusion tokens right after <speech_end>), which is where the reference's mask / cache
    shifting and its guard off-by-one (modeling_vibevoice_inference.py:599-624) decide which negative KV entries stay visible.
    Sequences and flags exact, audio 1e-5."""
    import random
    from fake_engine import make_model
    from oracle import vv_oracle as O
    from vibevoice_b200.configuration import preset_config
    from vibevoice_b200.modeling import ForcedTokenScript
    from vibevoice_b200.synth import SynthTokenizer, synth_state_dict
    rnd = random.Random(1000 + seed)
    cfg = preset_config("tiny")
    tok = SynthTokenizer(cfg.decoder_config.vocab_size)
    sd = synth_state_dict(cfg, 1234, torch.float32)
    s, e, d, x = tok.speech_start_id, tok.speech_end_id, tok.speech_diffusion_id, tok.eos_token_id
    refresh = seed % 2 == 0
    B = 2 + seed % 2
    scripts = []
    for _ in range(B):
        row, prev = [], s                                   # the prompt ends with <speech_start>
        for _ in range(rnd.randint(3, 9)):
            if refresh:                                     # well-formed: turns are s d+ e, in between only s or EOS
                nxt = rnd.choice([d, d, d, e]) if prev in (s, d) else s
            else:                                           # any order the constraint processor could emit
                nxt = rnd.choice([d, d, e, s])
            row.append(nxt)
            prev = nxt
        row.append(x)
        scripts.append(row)
    g = torch.Generator().manual_seed(seed)
    lens = [rnd.randint(4, 9) for _ in range(B)]
    L = max(lens)
    ids = torch.randint(0, 1000, (B, L), generator=g)
    mask = torch.zeros(B, L, dtype=torch.long)
    for r, n in enumerate(lens):                           # left-padded ragged prompts, as the processor builds them
        mask[r, L - n:] = 1
        ids[r, :L - n] = tok.pad_id
    ids[:, -1] = s
    n_new = max(len(r) for r in scripts) + 1
    model = make_model(cfg, tok, sd, max_batch=B)
    model.set_ddpm_inference_steps(5)
    torch.manual_seed(77)
    out = model.generate(input_ids=ids, attention_mask=mask, tokenizer=tok, cfg_scale=1.3, is_prefill=False, max_new_tokens=n_new,
                         max_length_times=1e9, show_progress_bar=False, logits_processor=[ForcedTokenScript(scripts)],
                         refresh_negative=refresh)
    torch.manual_seed(77)
    ref = O.generate(sd, cfg, ids, mask, tok, cfg_scale=1.3, num_steps=5, max_new_tokens=n_new, max_length_times=1e9,
                     forced_tokens=scripts, refresh_negative=refresh)
    assert torch.equal(out.sequences, ref.sequences), (scripts, out.sequences, ref.sequences)
    assert torch.equal(out.reach_max_step_sample, ref.reach_max_step_sample)
    for a, b in zip(out.speech_outputs, ref.speech_outputs):
        assert (a is None) == (b is None)
        if a is not None:
            assert a.shape == b.shape, scripts
            rel = float((a.double() - b.double()).norm() / b.double().norm())
            assert rel < 1e-5, (rel, scripts, refresh)


@pytest.mark.parametrize("seed", list(range(8)))
def test_streaming_product_host_logic_on_random_inputs(monkeypatch, seed):
    """The streaming variant's host loop (engine stand-in) against `oracle/vv_streaming.generate_streaming` -- which reproduces the reference's
    own streaming generate() on the committed fixtures -- for random prompt / text lengths (0..3 text windows, partial last window), EOS
    classifier biases (early stop, never stop) and step limits (`reach_max_step_sample`)."""
    import random
    import fake_engine
    from oracle import vv_streaming as VS
    from vibevoice_b200 import streaming as S
    from vibevoice_b200.configuration import preset_config
    from vibevoice_b200.synth import synth_state_dict
    rnd = random.Random(500 + seed)
    cfg = preset_config("tiny")
    tts_layers = 1
    eos_bias = rnd.choice([-6.0, -0.3, 0.5])
    sd = VS.streaming_state_dict(synth_state_dict(cfg, 1234, torch.float32), cfg, tts_layers, eos_bias=eos_bias)
    m = S.VibeVoiceStreamingForConditionalGenerationInference(cfg, tts_backbone_num_hidden_layers=tts_layers)
    monkeypatch.setattr(m, "_new_engine", lambda: fake_engine.FakeEngine(cfg, [0, 1], 2))
    m.load_state_dict(sd)
    m.set_ddpm_inference_steps(5)
    g = torch.Generator().manual_seed(seed)
    prompt = torch.randint(0, 2000, (rnd.randint(2, 9),), generator=g)
    text = torch.randint(0, 2000, (rnd.randint(1, 14),), generator=g)
    max_new = rnd.randint(3, 30)
    cfg_scale = rnd.choice([1.0, 1.5, 3.0])
    torch.manual_seed(9)
    out = m.generate(input_ids=prompt[None], tts_text_ids=text[None], neg_text_input_id=2047, cfg_scale=cfg_scale, max_new_tokens=max_new)
    torch.manual_seed(9)
    ref = VS.generate_streaming(sd, cfg, tts_layers, prompt, text, 2047, cfg_scale=cfg_scale, num_steps=5, max_new_tokens=max_new)
    assert torch.equal(out.sequences, ref.sequences), (out.sequences, ref.sequences)
    assert torch.equal(out.reach_max_step_sample, ref.reach_max_step_sample)
    a, b = out.speech_outputs[0], ref.speech_outputs[0]
    assert (a is None) == (b is None)
    if a is not None:
        assert a.shape == b.shape and float((a.double() - b.double()).norm() / b.double().norm()) < 1e-5
