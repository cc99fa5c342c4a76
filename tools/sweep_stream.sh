#!/bin/bash
# A/B of the weight-stream kernel's knobs through tools/profile_segments.py (sampler + LM segment times)
for inf in 12 6 4 3 2; do
  echo "== VV_STREAM_INFLIGHT=$inf"
  VV_STREAM_INFLIGHT=$inf timeout 300 python tools/profile_segments.py 2>&1 | grep -E "^lm_decode|^diffusion_sample"
done
