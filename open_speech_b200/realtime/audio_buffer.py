"""Drop-in for src/realtime/audio_buffer.py (reference lines 20-166), GPU-backed.

Same names, argument meaning and error behaviour; the codec and the np.interp resampler run
in libosb200 (osb_resample_linear_host: G.711 expand/compress fused with the f64 linear
interpolation, bit-exact).  The speech start/stop gate stays host-side Python, exactly the
reference's integer state machine (audio_buffer.py:125-156).
"""
from __future__ import annotations

import logging
from typing import Any

import numpy as np

from .. import _native as N
from ..vad.silero import VAD_SAMPLE_RATE, SileroVAD

logger = logging.getLogger(__name__)

_FMT = {"pcm16": N.FMT_PCM16, "g711_ulaw": N.FMT_ULAW, "g711_alaw": N.FMT_ALAW}


def _linear(data: bytes, in_fmt: int, out_fmt: int, from_rate: int, to_rate: int) -> bytes:
    n_in = len(data) // (2 if in_fmt == N.FMT_PCM16 else 1)
    if n_in == 0:
        return data if in_fmt == out_fmt else b""
    n_out = int(n_in * (to_rate / from_rate)) if from_rate != to_rate else n_in
    if n_out == 0:
        return b""
    out = np.empty(n_out, dtype=np.int16 if out_fmt == N.FMT_PCM16 else np.uint8)
    if from_rate == to_rate:
        # same-rate: pure codec (the reference skips np.interp here, audio_buffer.py:22-23)
        if in_fmt == out_fmt:
            return data
        if in_fmt == N.FMT_PCM16:
            N.call("osb_g711_encode_host", data, N.ptr(out), n_in, out_fmt)
        else:
            N.call("osb_g711_decode_host", data, N.ptr(out), n_in, in_fmt)
        return out.tobytes()
    N.call("osb_resample_linear_host", data, in_fmt, N.ptr(out), out_fmt, n_in, n_out, 1, n_in, n_out)
    return out.tobytes()


def _resample_linear(pcm_bytes: bytes, from_rate: int, to_rate: int) -> bytes:
    """Simple linear interpolation resample for PCM16 mono (reference :20-34)."""
    if from_rate == to_rate:
        return pcm_bytes
    if len(pcm_bytes) // 2 == 0:
        return pcm_bytes
    return _linear(pcm_bytes, N.FMT_PCM16, N.FMT_PCM16, from_rate, to_rate)


def decode_audio_to_pcm16(data: bytes, fmt: str, target_rate: int = 16000) -> bytes:
    """Decode 'pcm16' (24 kHz) / 'g711_ulaw' / 'g711_alaw' (8 kHz) to PCM16 mono at target_rate."""
    if fmt == "pcm16":
        return _resample_linear(data, 24000, target_rate)
    if fmt in ("g711_ulaw", "g711_alaw"):
        return _linear(data, _FMT[fmt], N.FMT_PCM16, 8000, target_rate)
    raise ValueError(f"Unsupported audio format: {fmt}")


def encode_pcm16_to_format(pcm16_data: bytes, from_rate: int, fmt: str) -> bytes:
    """Encode PCM16 mono at from_rate to 'pcm16' (24 kHz) / 'g711_ulaw' / 'g711_alaw' (8 kHz)."""
    if fmt == "pcm16":
        return _resample_linear(pcm16_data, from_rate, 24000)
    if fmt in ("g711_ulaw", "g711_alaw"):
        if len(pcm16_data) // 2 == 0:
            return b""
        return _linear(pcm16_data, N.FMT_PCM16, _FMT[fmt], from_rate, 8000)
    raise ValueError(f"Unsupported audio format: {fmt}")


class InputAudioBuffer:
    """Input audio buffer (PCM16 16 kHz mono) with optional VAD gating (reference :84-166)."""

    def __init__(self, vad: SileroVAD | None = None, threshold: float = 0.5,
                 silence_duration_ms: int = 500, max_buffer_bytes: int = 50 * 1024 * 1024):
        self._buffer = bytearray()
        self._vad = vad
        self._threshold = threshold
        self._silence_duration_ms = silence_duration_ms
        self._in_speech = False
        self._silence_samples = 0
        self._speech_start_ms = 0
        self._total_samples = 0
        self._max_buffer_bytes = max_buffer_bytes

    @property
    def in_speech(self) -> bool:
        return self._in_speech

    def clear(self) -> None:
        self._buffer.clear()
        self._silence_samples = 0

    def append(self, pcm16_16khz: bytes) -> list[dict[str, Any]]:
        events: list[dict[str, Any]] = []
        frame_size = len(pcm16_16khz)
        if frame_size > self._max_buffer_bytes:
            self.clear()
            raise BufferError(f"Audio frame exceeds max buffer size ({self._max_buffer_bytes} bytes)")
        if len(self._buffer) + frame_size > self._max_buffer_bytes:
            raise BufferError(f"Input audio buffer exceeded max size ({self._max_buffer_bytes} bytes)")
        self._buffer.extend(pcm16_16khz)

        num_samples = frame_size // 2
        current_ms = (self._total_samples * 1000) // VAD_SAMPLE_RATE
        self._total_samples += num_samples
        if self._vad is None or num_samples == 0:
            return events

        # the reference converts to float32/32768 and calls vad(audio); a SileroVAD from this
        # package scores the int16 bytes on the GPU directly (same arithmetic, fused convert)
        score = getattr(self._vad, "score_pcm16", None)
        if score is not None:
            prob = score(pcm16_16khz)
        else:  # any callable with the reference's __call__(float32 ndarray) contract (e.g. a mock)
            prob = self._vad(np.frombuffer(pcm16_16khz, dtype=np.int16).astype(np.float32) / 32768.0)
        if prob >= self._threshold:
            self._silence_samples = 0
            if not self._in_speech:
                self._in_speech = True
                self._speech_start_ms = current_ms
                events.append({"type": "speech_started", "audio_start_ms": current_ms})
        elif self._in_speech:
            self._silence_samples += num_samples
            if (self._silence_samples * 1000) // VAD_SAMPLE_RATE >= self._silence_duration_ms:
                self._in_speech = False
                self._silence_samples = 0
                events.append({"type": "speech_stopped", "audio_end_ms": current_ms})
        return events

    def commit(self) -> bytes:
        data = bytes(self._buffer)
        self.clear()
        return data

    def get_audio(self) -> bytes:
        return bytes(self._buffer)
