"""Per-segment device time of one speech frame (CUDA events on the engine stream, graphs on): LM decode step, N-step diffusion
sampler, acoustic decoder frame, semantic encoder frame, connectors -- each through its own C-ABI entry point, steady state.
    python tools/profile_segments.py [--model 1.5b --ctx 61440 --iters 20 --steps 30 --batch 1]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vibevoice_b200 import _native as N
from vibevoice_b200.configuration import preset_config
from vibevoice_b200.modeling import VibeVoiceForConditionalGenerationInference
from vibevoice_b200.synth import SynthTokenizer, iter_synth_state_dict_fast

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="1.5b")
ap.add_argument("--ctx", type=int, default=61440)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--batch", type=int, default=1)
a = ap.parse_args()
cfg = preset_config(a.model)
tok = SynthTokenizer(cfg.decoder_config.vocab_size)
B = a.batch
m = VibeVoiceForConditionalGenerationInference(cfg, tok, max_batch=B)
m.load_state_dict(iter_synth_state_dict_fast(cfg, 1234, device="cuda", parts=("lm", "head", "acoustic_decoder", "semantic", "connectors", "lm_head")), tok)
eng = m.engine
eng.kv_init(B * (a.ctx + 8 * a.iters + 8) + B * (8 * a.iters + 8))
eng.set_diffusion_steps(a.steps)
for r in range(B):
    N.check(eng.lib.vv_kv_reserve(eng.h, r, a.ctx + 8 * a.iters + 1, eng.s))
    eng.kv_set_len(r, a.ctx)
    eng.kv_set_len(B + r, 0)
eng.embed_tokens([tok.speech_start_id] * (2 * B), eng.embeds)
with torch.cuda.stream(eng.stream):
    eng.active.fill_(1)
    eng.noise.normal_()
eng.sync()


def lm():
    eng.lm_decode()
    eng.kv_commit([1] * (2 * B))


segs = [("lm_decode", lm), ("diffusion_sample", lambda: eng.diffusion_sample(1.3)), ("codec_decode", eng.codec_decode),
        ("semantic_encode", eng.semantic_encode), ("connect", eng.connect), ("frame_tail(all four fused in one graph)", lambda: eng.frame_tail(1.3))]
out = {}
for name, fn in segs:
    for _ in range(3):
        fn()
    eng.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = eng.launch_count()
    e0.record(eng.stream)
    for _ in range(a.iters):
        fn()
    e1.record(eng.stream)
    eng.sync()
    out[name] = {"us": 1000.0 * e0.elapsed_time(e1) / a.iters, "launches": (eng.launch_count() - l0) // a.iters}
    print("%-42s %9.1f us   %4d launches" % (name, out[name]["us"], out[name]["launches"]), flush=True)
wb = eng.weight_bytes()
print(json.dumps({"segments": out, "weight_bytes": wb, "model": a.model, "ctx": a.ctx, "steps": a.steps, "batch": B}))
