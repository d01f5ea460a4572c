"""Drop-in for the resampler of src/streaming.py (reference lines 55-91), GPU-backed."""
from __future__ import annotations

from math import gcd

import numpy as np

from . import _native as N

INTERNAL_SAMPLE_RATE = 16000
MAX_UTTERANCE_SECONDS = 30
MAX_UTTERANCE_BYTES = MAX_UTTERANCE_SECONDS * INTERNAL_SAMPLE_RATE * 2  # reference :42-43
STREAM_STATE = np.dtype([("silence_samples", "<i8"), ("utterance_bytes", "<i8"), ("speech_active", "<i4"), ("reserved", "<i4")])
ACT_SPEECH_START, ACT_UTTERANCE_RESET, ACT_APPEND, ACT_TRANSCRIBE, ACT_FINALIZE, ACT_SPEECH_END = 1, 2, 4, 8, 16, 32  # OSB_ACT_*


def resample_pcm16(pcm_bytes: bytes, from_rate: int, to_rate: int) -> bytes:
    """Resample PCM16 LE mono: polyphase FIR, scipy.signal.resample_poly(padtype='line')
    arithmetic (bit-exact), clip, truncate -- computed by osb_resample_poly_host."""
    if from_rate == to_rate:
        return pcm_bytes
    n = len(pcm_bytes) // 2
    if n == 0:
        return pcm_bytes
    if n == 1:
        out_len = int(n * (to_rate / from_rate))
        if out_len <= 0:
            return b""
        return np.full(out_len, np.frombuffer(pcm_bytes, dtype=np.int16)[0], dtype=np.int16).tobytes()
    g = gcd(to_rate, from_rate)
    up, down = to_rate // g, from_rate // g
    n_out = (n * up + down - 1) // down
    out = np.empty(n_out, dtype=np.int16)
    N.call("osb_resample_poly_host", pcm_bytes, N.ptr(out), n, 1, n, n_out, up, down)
    return out.tobytes()


class SessionGate:
    """The VAD / endpointing half of one ``StreamingSession`` (reference src/streaming.py:186-197, :290-355, :429-498).

    ``process_chunk(chunk)`` is ``_process_chunk`` up to the point where the reference awaits the transcriber: the
    client-rate chunk is resampled (resample_pcm16 arithmetic), scored and the utterance machine advances, all in one
    ``osb_stream_chunk_host`` call.  It returns the 16 kHz chunk and the OSB_ACT_* bits that say what the session does
    next (send speech_start, reset / extend ``utterance_audio``, transcribe, finalize, send speech_end).  The state
    record lives in this object, like the reference keeps it on the session.  S sessions at once:
    :class:`open_speech_b200.realtime.gate.StreamGate`.
    """

    def __init__(self, sample_rate: int, endpointing_ms: int, vad=None, vad_enabled: bool = True, threshold: float = 0.5):
        self.client_sample_rate = sample_rate
        self.needs_resample = sample_rate != INTERNAL_SAMPLE_RATE
        self.endpointing_samples = int(INTERNAL_SAMPLE_RATE * endpointing_ms / 1000)
        self.vad_state = vad
        self.vad_enabled = bool(vad_enabled) and vad is not None
        self.threshold = threshold
        self._rec = np.zeros(1, dtype=STREAM_STATE)
        self._act = np.zeros(1, dtype=np.int32)

    @property
    def speech_active(self) -> bool:
        return bool(self._rec["speech_active"][0])

    @property
    def silence_samples(self) -> int:
        return int(self._rec["silence_samples"][0])

    @property
    def utterance_bytes(self) -> int:
        return int(self._rec["utterance_bytes"][0])

    def process_chunk(self, chunk: bytes) -> tuple[bytes, int]:
        n_in = len(chunk) // 2
        if self.needs_resample and n_in >= 2:
            g = gcd(INTERNAL_SAMPLE_RATE, self.client_sample_rate)
            up, down = INTERNAL_SAMPLE_RATE // g, self.client_sample_rate // g
            n_out = (n_in * up + down - 1) // down
            rate = self.client_sample_rate
        elif self.needs_resample and n_in == 1:  # single-sample special case of resample_pcm16 (:69-73): nothing to filter
            chunk = resample_pcm16(chunk, self.client_sample_rate, INTERNAL_SAMPLE_RATE)
            n_in = n_out = len(chunk) // 2
            rate = INTERNAL_SAMPLE_RATE
        else:
            n_out, rate = n_in, INTERNAL_SAMPLE_RATE
        out = np.empty(n_out, dtype=np.int16)
        if self.vad_enabled:
            st = np.ascontiguousarray(self.vad_state._state, dtype=np.float32)
            N.call("osb_stream_chunk_host", self.vad_state.session.handle, chunk, n_in, rate, N.ptr(out), n_out, N.ptr(self._rec), N.ptr(st), 1,
                   float(self.threshold), self.endpointing_samples, MAX_UTTERANCE_BYTES, N.ptr(self._act))
            self.vad_state._state = st
        else:
            N.call("osb_stream_chunk_host", None, chunk, n_in, rate, N.ptr(out), n_out, N.ptr(self._rec), None, 0,
                   float(self.threshold), self.endpointing_samples, MAX_UTTERANCE_BYTES, N.ptr(self._act))
        return out.tobytes(), int(self._act[0])
