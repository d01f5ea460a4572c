"""Drop-in for the resampler of src/streaming.py (reference lines 55-91), GPU-backed."""
from __future__ import annotations

from math import gcd

import numpy as np

from . import _native as N

INTERNAL_SAMPLE_RATE = 16000


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
