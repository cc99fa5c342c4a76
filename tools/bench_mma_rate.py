"""tcgen05.mma issue / execution rate for the skinny shapes of the weight-stream kernel (128 x nB x 16, operands in shared memory)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vibevoice_b200 import _native as N
from vibevoice_b200.configuration import preset_config
from vibevoice_b200.modeling import VibeVoiceForConditionalGenerationInference
from vibevoice_b200.synth import SynthTokenizer, synth_state_dict

cfg = preset_config("tiny")
tok = SynthTokenizer(cfg.decoder_config.vocab_size)
m = VibeVoiceForConditionalGenerationInference(cfg, tok, max_batch=1)
m.load_state_dict(synth_state_dict(cfg, 1234, torch.bfloat16), tok)
eng = m.engine
names = {0: "tid 0 issues, one commit", 1: "commit + wait per 4", 3: "converged warp, elect.sync", 4: "loop-invariant operands"}
for ctas in (1, 148):
    for nB in (16, 64, 256):
        for nacc in (1, 2):
            if nacc * nB > 512:
                continue
            for mode in (0, 1, 3, 4):
                cyc = (C.c_longlong * 2)(0, 0)
                n = 2000
                N.check(eng.lib.vv_debug_mma_rate(eng.h, n, nB, mode, nacc, ctas, cyc))
                print("ctas %3d  nB %3d  accumulators %2d  mode %d (%-28s): %7.1f cycles / MMA   (issue loop alone %7.1f)" % (
                    ctas, nB, nacc, mode, names[mode], cyc[0] / n, cyc[1] / n), flush=True)
