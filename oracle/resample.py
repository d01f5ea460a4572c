"""Oracle: polyphase PCM16 resampler.  TEST ONLY.

Follows src/streaming.py:55-91 (``resample_pcm16``), which calls
``scipy.signal.resample_poly(x_f32, up, down, padtype="line")`` (scipy pinned
1.17.0 in requirements.lock:13; 1.18.1 installed here, same algorithm:
scipy/signal/_signaltools.py ``resample_poly`` and scipy/signal/_upfirdn_apply.pyx
``_apply_impl`` with MODE_LINE).

``resample_pcm16``          calls scipy exactly like the reference does.
``resample_poly_restated``  the same arithmetic written out tap by tap in
                            float32 (this is what the CUDA kernel implements);
                            tests prove it is bit-identical to scipy.
"""
from __future__ import annotations

from math import gcd

import numpy as np


def design(up: int, down: int) -> tuple[np.ndarray, int, int]:
    """Filter exactly as resample_poly builds it (f64 firwin -> f32, *up, pre-pad).

    Returns (h_padded_f32, n_pre_remove, half_len).
    """
    from scipy.signal import firwin

    max_rate = max(up, down)
    half_len = 10 * max_rate
    h = firwin(2 * half_len + 1, 1.0 / max_rate, window=("kaiser", 5.0)).astype(np.float32)
    h *= up
    n_pre_pad = down - half_len % down
    n_pre_remove = (half_len + n_pre_pad) // down
    h = np.concatenate([np.zeros(n_pre_pad, np.float32), h])
    return h, n_pre_remove, half_len


def _output_len(len_h: int, in_len: int, up: int, down: int) -> int:
    return (((in_len - 1) * up + len_h) - 1) // down + 1


def resample_poly_restated(x: np.ndarray, up: int, down: int) -> np.ndarray:
    """float32 in -> float32 out; bit-identical to scipy resample_poly(padtype='line')."""
    g = gcd(up, down)
    up //= g
    down //= g
    x = np.asarray(x, dtype=np.float32)
    if up == down == 1:
        return x.copy()
    n_in = len(x)
    n_out = (n_in * up + down - 1) // down
    h, n_pre_remove, _ = design(up, down)
    n_post = 0
    while _output_len(len(h) + n_post, n_in, up, down) < n_out + n_pre_remove:
        n_post += 1
    if n_post:
        h = np.concatenate([h, np.zeros(n_post, np.float32)])
    # _pad_h: per-phase, flipped
    hp = len(h) + (-len(h) % up)
    hf = np.zeros(hp, np.float32)
    hf[: len(h)] = h
    per_phase = hp // up
    h_tf = hf.reshape(-1, up).T[:, ::-1]  # [phase][k], k ascending = oldest sample first

    y_idx = np.arange(n_pre_remove, n_pre_remove + n_out, dtype=np.int64)
    x_idx = (y_idx * down) // up
    phase = (y_idx * down) % up
    slope = np.float32((x[-1] - x[0]) / np.float32(n_in - 1))
    acc = np.zeros(n_out, np.float32)
    for k in range(per_phase):
        xi = x_idx - per_phase + 1 + k
        inside = (xi >= 0) & (xi < n_in)
        xv = np.where(
            inside,
            x[np.clip(xi, 0, n_in - 1)],
            np.where(
                xi < 0,
                x[0] + xi.astype(np.float32) * slope,
                x[-1] + (xi - n_in + 1).astype(np.float32) * slope,
            ),
        ).astype(np.float32)
        acc = acc + xv * h_tf[phase, k]
    return acc


def resample_pcm16(pcm_bytes: bytes, from_rate: int, to_rate: int) -> bytes:
    """src/streaming.py:55-91."""
    if from_rate == to_rate:
        return pcm_bytes
    s = np.frombuffer(pcm_bytes, dtype=np.int16).astype(np.float32)
    if len(s) == 0:
        return pcm_bytes
    if len(s) == 1:
        m = int(len(s) * (to_rate / from_rate))
        if m <= 0:
            return b""
        return np.full(m, s[0], dtype=np.int16).tobytes()
    from scipy.signal import resample_poly

    g = gcd(to_rate, from_rate)
    y = resample_poly(s, to_rate // g, from_rate // g, padtype="line")
    y = np.clip(y, -32768, 32767)
    return y.astype(np.int16).tobytes()


def resample_pcm16_restated(pcm_bytes: bytes, from_rate: int, to_rate: int) -> bytes:
    if from_rate == to_rate:
        return pcm_bytes
    s = np.frombuffer(pcm_bytes, dtype=np.int16).astype(np.float32)
    if len(s) == 0:
        return pcm_bytes
    if len(s) == 1:
        m = int(len(s) * (to_rate / from_rate))
        if m <= 0:
            return b""
        return np.full(m, s[0], dtype=np.int16).tobytes()
    g = gcd(to_rate, from_rate)
    y = resample_poly_restated(s, to_rate // g, from_rate // g)
    return np.clip(y, -32768, 32767).astype(np.int16).tobytes()
