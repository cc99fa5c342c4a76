from vibevoice_b200.configuration import (VibeVoiceAcousticTokenizerConfig, VibeVoiceConfig, VibeVoiceDiffusionHeadConfig,  # noqa: F401
                                          VibeVoiceSemanticTokenizerConfig)
from vibevoice_b200.modeling import VibeVoiceForConditionalGenerationInference, VibeVoiceGenerationOutput  # noqa: F401
