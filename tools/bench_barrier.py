import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vibevoice_b200 import _native as NV
from vibevoice_b200.configuration import preset_config
from vibevoice_b200.synth import SynthTokenizer, synth_state_dict
from vibevoice_b200.modeling import VibeVoiceForConditionalGenerationInference
cfg = preset_config("tiny"); tok = SynthTokenizer(cfg.decoder_config.vocab_size)
m = VibeVoiceForConditionalGenerationInference(cfg, tok, max_batch=1); m.load_state_dict(synth_state_dict(cfg, 1, torch.bfloat16), tok)
eng = m.engine
for per_sm in (1, 2):
    ms = C.c_float()
    NV.check(eng.lib.vv_debug_barrier_bench(eng.h, 2000, per_sm, C.byref(ms)))
    print("ctas/SM", per_sm, "barrier us:", ms.value * 1e3 / 2000)
