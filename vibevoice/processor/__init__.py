from vibevoice_b200.processor import VibeVoiceProcessor  # noqa: F401
