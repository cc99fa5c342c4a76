from vibevoice_b200.streamer import AsyncAudioStreamer, AudioStreamer  # noqa: F401
