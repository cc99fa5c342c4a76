"""Runs a few steady-state frames of the hot path for ncu (launch list / --set full captures).

    VV_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -s S -c N --csv --log-file out.csv \
        python tools/profile_frame.py --model 1.5b --ctx 61440 --frames 4

The KV prefix is NOT computed here (kv_len is simply set to --ctx over zero-initialised pages): attention cost does not
depend on the values, and this script only exists to expose kernels to the profiler; bench.py does the real prefill."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vibevoice_b200.configuration import preset_config
from vibevoice_b200.modeling import VibeVoiceForConditionalGenerationInference
from vibevoice_b200.synth import SynthTokenizer, iter_synth_state_dict_fast

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="1.5b")
ap.add_argument("--ctx", type=int, default=61440)
ap.add_argument("--frames", type=int, default=4)
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--batch", type=int, default=1)
a = ap.parse_args()
cfg = preset_config(a.model)
tok = SynthTokenizer(cfg.decoder_config.vocab_size)
B = a.batch
m = VibeVoiceForConditionalGenerationInference(cfg, tok, max_batch=B)
parts = ("lm", "head", "acoustic_decoder", "semantic", "connectors", "lm_head")
m.load_state_dict(iter_synth_state_dict_fast(cfg, 1234, device="cuda", parts=parts), tok)
eng = m.engine
eng.kv_init(B * (a.ctx + a.frames + 8) + B * (a.frames + 8))
eng.set_diffusion_steps(a.steps)
for r in range(B):
    N = __import__("vibevoice_b200._native", fromlist=["check"])
    N.check(eng.lib.vv_kv_reserve(eng.h, r, a.ctx + a.frames + 1, eng.s))
    eng.kv_set_len(r, a.ctx)
    eng.kv_set_len(B + r, 0)
eng.embed_tokens([tok.speech_start_id] * (2 * B), eng.embeds)
with torch.cuda.stream(eng.stream):
    eng.active.fill_(1)
    eng.noise.normal_()
eng.sync()
l0 = eng.launch_count()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for f in range(a.frames):
    if f == a.frames - 1:
        lf = eng.launch_count()
        eng.sync()
        torch.cuda.profiler.start()
        ev0.record(eng.stream)
    eng.lm_decode()
    eng.kv_commit([1] * (2 * B))
    eng.frame_tail(1.3)
ev1.record(eng.stream)
eng.sync()
torch.cuda.profiler.stop()
print("launches per frame:", eng.launch_count() - lf, " last frame ms:", ev0.elapsed_time(ev1), " launches before last frame:", lf - l0, flush=True)
