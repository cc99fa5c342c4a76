"""Drop-in import path for the reference's `vibevoice/modular/modeling_vibevoice_streaming_inference.py` (streaming-0.5B variant)."""
from vibevoice_b200.modeling import VibeVoiceGenerationOutput  # noqa: F401
from vibevoice_b200.streaming import (TTS_SPEECH_WINDOW_SIZE, TTS_TEXT_WINDOW_SIZE,  # noqa: F401
                                      VibeVoiceStreamingForConditionalGenerationInference)
