from vibevoice_b200.configuration import (VibeVoiceAcousticTokenizerConfig, VibeVoiceConfig, VibeVoiceDiffusionHeadConfig,  # noqa: F401
                                          VibeVoiceSemanticTokenizerConfig)
