"""Oracle for the SURVEY 8(f) "next" rows.  TEST ONLY.

Follows: MultiTrackComposer._resample / _mix_prepared / _float_to_int16 (src/composer.py:167-189, :254-257) and
_resample_to_16k (src/wyoming/tts_handler.py:37-44); the realtime response framing (src/realtime/server.py:238-277).  Pinned by tests/golden/reference_vectors_next.npz, produced by
the reference's own code (oracle/make_golden_next.py; the wyoming function is executed from its source text because the
`wyoming` package it imports at module level is not installed).
"""
from __future__ import annotations

import math

import numpy as np


def composer_resample(samples: np.ndarray, src_rate: int, dst_rate: int) -> np.ndarray:
    from scipy.signal import resample_poly

    if src_rate == dst_rate:
        return samples.astype(np.float32, copy=False)
    g = math.gcd(src_rate, dst_rate)
    return resample_poly(samples, dst_rate // g, src_rate // g).astype(np.float32, copy=False)


def mix_prepared(prepared, sample_rate: int) -> np.ndarray:
    total = 0
    for t in prepared:
        start = int(round(max(0.0, float(t.get("offset_s", 0.0))) * sample_rate))
        total = max(total, start + len(t["samples"]))
    if total <= 0:
        return np.zeros(0, dtype=np.float32)
    mixed = np.zeros(total, dtype=np.float32)
    for t in prepared:
        start = int(round(max(0.0, float(t.get("offset_s", 0.0))) * sample_rate))
        s = np.asarray(t["samples"], dtype=np.float32)
        mixed[start:start + len(s)] += s
    return np.clip(mixed, -1.0, 1.0)


def resample_to_16k(audio: np.ndarray, source_rate: int = 24000) -> np.ndarray:
    if source_rate == 16000:
        return audio
    m = int(len(audio) * (16000 / source_rate))
    return np.interp(np.linspace(0, len(audio) - 1, m), np.arange(len(audio)), audio).astype(audio.dtype)


def realtime_response_payload(chunks, output_format: str) -> bytes:
    """src/realtime/server.py:238-251: concatenate, quantise (multiply, clip, truncate), encode to the output format."""
    from . import codec

    parts = [c if isinstance(c, np.ndarray) else np.array(c, dtype=np.float32) for c in chunks]
    if not parts:
        return b""
    combined = np.concatenate(parts)
    pcm16 = (combined * 32767).clip(-32768, 32767).astype(np.int16).tobytes()
    return codec.encode_pcm16_to_format(pcm16, 24000, output_format)


def realtime_deltas(chunks, output_format: str) -> list[str]:
    """src/realtime/server.py:268-277: base64 of consecutive 3000-byte pieces."""
    import base64

    data = realtime_response_payload(chunks, output_format)
    return [base64.b64encode(data[i:i + 3000]).decode("ascii") for i in range(0, len(data), 3000)]
