from vibevoice_b200.processor import AudioNormalizer, VibeVoiceProcessor  # noqa: F401

__all__ = ["VibeVoiceProcessor"]
