"""Per-sample audio chunk queues with the reference's contract (`vibevoice/modular/streamer.py:13-86, 150-203`):
`put(audio_chunks, sample_indices)`, `end(sample_indices=None)`, `finished_flags`, `get_stream(i)`, iteration.
`generate()` calls exactly these (`modeling_vibevoice_inference.py:443-447, 527-528, 653-655, 677-678`)."""
from __future__ import annotations

import asyncio
from queue import Queue
from typing import Optional

import torch


class AudioStreamer:
    def __init__(self, batch_size: int, stop_signal=None, timeout: Optional[float] = None):
        self.batch_size, self.stop_signal, self.timeout = batch_size, stop_signal, timeout
        self.audio_queues = [Queue() for _ in range(batch_size)]
        self.finished_flags = [False] * batch_size
        self.sample_indices_map = {}

    def _emit(self, idx: int, item):
        self.audio_queues[idx].put(item, timeout=self.timeout)

    def put(self, audio_chunks: torch.Tensor, sample_indices: torch.Tensor):
        for i, s in enumerate(sample_indices):
            idx = int(s)
            if idx < self.batch_size and not self.finished_flags[idx]:
                self._emit(idx, audio_chunks[i].detach().cpu())

    def end(self, sample_indices=None):
        rows = range(self.batch_size) if sample_indices is None else [int(s) for s in sample_indices]
        for idx in rows:
            if idx < self.batch_size and not self.finished_flags[idx]:
                self._emit(idx, self.stop_signal)
                self.finished_flags[idx] = True

    def get_stream(self, sample_idx: int):
        if sample_idx >= self.batch_size:
            raise ValueError(f"Sample index {sample_idx} exceeds batch size {self.batch_size}")
        return _SampleIter(self, sample_idx)

    def __iter__(self):
        return _BatchIter(self)


class _SampleIter:
    def __init__(self, st, idx):
        self.st, self.idx = st, idx

    def __iter__(self):
        return self

    def __next__(self):
        v = self.st.audio_queues[self.idx].get(timeout=self.st.timeout)
        if v is self.st.stop_signal or (not torch.is_tensor(v) and v == self.st.stop_signal):
            raise StopIteration
        return v


class _BatchIter:
    def __init__(self, st):
        self.st, self.active = st, set(range(st.batch_size))

    def __iter__(self):
        return self

    def __next__(self):
        if not self.active:
            raise StopIteration
        out = {}
        for idx in list(self.active):
            q = self.st.audio_queues[idx]
            if q.empty():
                continue
            v = q.get(block=False)
            if not torch.is_tensor(v) and v == self.st.stop_signal:
                self.active.discard(idx)
            else:
                out[idx] = v
        if out or self.active:
            return out
        raise StopIteration


class AsyncAudioStreamer(AudioStreamer):
    """asyncio flavour: queues live on the consumer's loop; `put`/`end` may be called from the generation thread."""

    def __init__(self, batch_size: int, stop_signal=None, timeout: Optional[float] = None):
        super().__init__(batch_size, stop_signal, timeout)
        self.audio_queues = [asyncio.Queue() for _ in range(batch_size)]
        self.loop = asyncio.get_running_loop()

    def _emit(self, idx: int, item):
        self.loop.call_soon_threadsafe(self.audio_queues[idx].put_nowait, item)

    async def get_stream(self, sample_idx: int):
        if sample_idx >= self.batch_size:
            raise ValueError(f"Sample index {sample_idx} exceeds batch size {self.batch_size}")
        while True:
            v = await self.audio_queues[sample_idx].get()
            if not torch.is_tensor(v) and v == self.stop_signal:
                break
            yield v
