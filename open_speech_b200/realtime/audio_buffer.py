"""Drop-in for src/realtime/audio_buffer.py (reference lines 20-166), GPU-backed.

Same names, argument meaning and error behaviour; the codec and the np.interp resampler run
in libosb200 (osb_resample_linear_host: G.711 expand/compress fused with the f64 linear
interpolation, bit-exact).  The speech start/stop gate runs in libosb200 as well
(csrc/realtime_gate.cu): InputAudioBuffer is a per-connection view of it.
"""
from __future__ import annotations

import logging
from typing import Any

import numpy as np

from .. import _native as N
from ..vad.silero import VAD_SAMPLE_RATE, SileroVAD, VadSession

N_EVT_STARTED, N_EVT_STOPPED, N_EVT_FRAME_TOO_LARGE, N_EVT_BUFFER_FULL = 1, 2, 3, 4  # OSB_EVT_* (include/osb200.h)

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


GATE_STATE = np.dtype([("total_samples", "<i8"), ("silence_samples", "<i8"), ("buffered_samples", "<i8"),
                       ("in_speech", "<i4"), ("speech_start_ms", "<i4")])  # osb_gate_state (include/osb200.h)
_EVENT = {N_EVT_STARTED: ("speech_started", "audio_start_ms"), N_EVT_STOPPED: ("speech_stopped", "audio_end_ms")}


class InputAudioBuffer:
    """Per-connection view of the device gate (reference class: src/realtime/audio_buffer.py:84-166).

    The object holds what the drop-in boundary has to hand back as host ``bytes`` (the committed audio goes to the
    transcriber as a WAV) and one ``osb_gate_state`` record; every gated ``append`` is one ``osb_gate_append_host``
    call: VAD scoring of the chunk and the start / stop machine run in libosb200, there is no host copy of that logic.
    S concurrent connections are better served by :class:`open_speech_b200.realtime.gate.RealtimeGate`, which keeps
    the same records, the LSTM states and the audio arena resident on the GPU and advances all of them per tick.
    """

    def __init__(self, vad: SileroVAD | None = None, threshold: float = 0.5,
                 silence_duration_ms: int = 500, max_buffer_bytes: int = 50 * 1024 * 1024):
        if vad is not None and not isinstance(getattr(vad, "session", None), VadSession):
            raise TypeError("InputAudioBuffer needs an open_speech_b200 SileroVAD over a VadSession (GPU-resident weights); "
                            "there is no host implementation of the gate")
        self._buffer = bytearray()
        self._vad = vad
        self._threshold = threshold
        self._silence_duration_ms = silence_duration_ms
        self._max_buffer_bytes = max_buffer_bytes
        self._gate = np.zeros(1, dtype=GATE_STATE)
        self._event = np.zeros(2, dtype=np.int32)

    # the reference's private counters, read-only views of the record
    @property
    def in_speech(self) -> bool:
        return bool(self._gate["in_speech"][0])

    _in_speech = in_speech

    @property
    def _silence_samples(self) -> int:
        return int(self._gate["silence_samples"][0])

    @property
    def _total_samples(self) -> int:
        return int(self._gate["total_samples"][0])

    @property
    def _speech_start_ms(self) -> int:
        return int(self._gate["speech_start_ms"][0])

    def clear(self) -> None:
        self._buffer.clear()
        self._gate["silence_samples"] = 0

    def append(self, pcm16_16khz: bytes) -> list[dict[str, Any]]:
        size = len(pcm16_16khz)
        if size > self._max_buffer_bytes:
            self.clear()
            raise BufferError(f"Audio frame exceeds max buffer size ({self._max_buffer_bytes} bytes)")
        if len(self._buffer) + size > self._max_buffer_bytes:
            raise BufferError(f"Input audio buffer exceeded max size ({self._max_buffer_bytes} bytes)")
        self._buffer += pcm16_16khz
        n = size // 2
        if self._vad is None or n == 0:
            self._gate["total_samples"] += n  # no gate configured: the clock is all there is (reference :125-129)
            return []
        if size % 2:
            raise ValueError("buffer size must be a multiple of element size")  # np.frombuffer(..., int16) in the reference (:133)
        st = np.ascontiguousarray(self._vad._state, dtype=np.float32)
        N.call("osb_gate_append_host", self._vad.session.handle, pcm16_16khz, n, N.ptr(self._gate), N.ptr(st), 1,
               float(self._threshold), int(self._silence_duration_ms), N.ptr(self._event))
        self._vad._state = st
        kind = _EVENT.get(int(self._event[0]))
        return [{"type": kind[0], kind[1]: int(self._event[1])}] if kind else []

    def commit(self) -> bytes:
        data = bytes(self._buffer)
        self.clear()
        return data

    def get_audio(self) -> bytes:
        return bytes(self._buffer)
