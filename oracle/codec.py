"""Oracle: G.711 codecs and the realtime path's linear resampler.  TEST ONLY.

Follows (reference file:line):
  * audioop.ulaw2lin / alaw2lin / lin2ulaw / lin2alaw as called from
    src/realtime/audio_buffer.py:52,55,76,79 (CPython 3.12 Modules/audioop.c,
    ITU-T G.711; closed forms in SURVEY.md Appendix B).
  * _resample_linear            src/realtime/audio_buffer.py:20-34
  * decode_audio_to_pcm16       src/realtime/audio_buffer.py:37-58
  * encode_pcm16_to_format      src/realtime/audio_buffer.py:61-81

Pinned: tables hash to the sha256 values of SURVEY.md Appendix B (taken from
this container's CPython audioop) -- see tests/test_oracle_codec.py.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------- G.711 expand


def ulaw2lin_table() -> np.ndarray:
    """256-entry mu-law -> int16 table (audioop.c st_ulaw2linear16)."""
    b = np.arange(256, dtype=np.int32)
    u = (~b) & 0xFF
    t = ((u & 0x0F) << 3) + 0x84
    t = t << ((u & 0x70) >> 4)
    y = np.where(u & 0x80, 0x84 - t, t - 0x84)
    return y.astype(np.int16)


def alaw2lin_table() -> np.ndarray:
    """256-entry A-law -> int16 table (audioop.c st_alaw2linear16)."""
    b = np.arange(256, dtype=np.int32)
    a = b ^ 0x55
    t = (a & 0x0F) << 4
    seg = (a & 0x70) >> 4
    t = np.where(seg == 0, t + 8, np.where(seg == 1, t + 0x108, (t + 0x108) << np.maximum(seg - 1, 0)))
    y = np.where(a & 0x80, t, -t)
    return y.astype(np.int16)


_ULAW = ulaw2lin_table()
_ALAW = alaw2lin_table()


def ulaw2lin(data: bytes) -> bytes:
    return _ULAW[np.frombuffer(data, dtype=np.uint8)].tobytes()


def alaw2lin(data: bytes) -> bytes:
    return _ALAW[np.frombuffer(data, dtype=np.uint8)].tobytes()


# --------------------------------------------------------------------------- G.711 compress

_SEG_UEND = np.array([0x3F, 0x7F, 0xFF, 0x1FF, 0x3FF, 0x7FF, 0xFFF, 0x1FFF], dtype=np.int32)
_SEG_AEND = np.array([0x1F, 0x3F, 0x7F, 0xFF, 0x1FF, 0x3FF, 0x7FF, 0xFFF], dtype=np.int32)


def lin2ulaw_array(pcm: np.ndarray) -> np.ndarray:
    """int16 -> mu-law byte (audioop.lin2ulaw width 2: st_14linear2ulaw(s >> 2))."""
    v = pcm.astype(np.int32) >> 2
    neg = v < 0
    mag = np.where(neg, -v, v)
    mask = np.where(neg, 0x7F, 0xFF)
    mag = np.minimum(mag, 8159) + 0x21
    seg = np.searchsorted(_SEG_UEND, mag, side="left")  # first i with mag <= end[i]
    uval = (seg << 4) | ((mag >> (np.minimum(seg, 7) + 1)) & 0xF)
    out = np.where(seg >= 8, 0x7F ^ mask, uval ^ mask)
    return out.astype(np.uint8)


def lin2alaw_array(pcm: np.ndarray) -> np.ndarray:
    """int16 -> A-law byte (audioop.lin2alaw width 2: st_linear2alaw(s >> 3))."""
    v = pcm.astype(np.int32) >> 3
    neg = v < 0
    mask = np.where(neg, 0x55, 0xD5)
    mag = np.where(neg, -v - 1, v)
    seg = np.searchsorted(_SEG_AEND, mag, side="left")
    segc = np.minimum(seg, 7)
    aval = (segc << 4) | np.where(segc < 2, (mag >> 1) & 0xF, (mag >> segc) & 0xF)
    out = np.where(seg >= 8, 0x7F ^ mask, aval ^ mask)
    return out.astype(np.uint8)


def lin2ulaw(pcm16: bytes) -> bytes:
    return lin2ulaw_array(np.frombuffer(pcm16, dtype=np.int16)).tobytes()


def lin2alaw(pcm16: bytes) -> bytes:
    return lin2alaw_array(np.frombuffer(pcm16, dtype=np.int16)).tobytes()


# --------------------------------------------------------------------------- linear resample


def resample_linear(pcm_bytes: bytes, from_rate: int, to_rate: int) -> bytes:
    """src/realtime/audio_buffer.py:20-34 -- np.interp on [0,1] grids, f64, trunc."""
    if from_rate == to_rate:
        return pcm_bytes
    x = np.frombuffer(pcm_bytes, dtype=np.int16).astype(np.float32)
    if len(x) == 0:
        return pcm_bytes
    m = int(len(x) * (to_rate / from_rate))
    if m == 0:
        return b""
    y = np.interp(np.linspace(0, 1, m), np.linspace(0, 1, len(x)), x)
    return y.astype(np.int16).tobytes()


def linear_out_len(n: int, from_rate: int, to_rate: int) -> int:
    return int(n * (to_rate / from_rate))


def interp_explicit(x_i16: np.ndarray, m: int) -> np.ndarray:
    """The closed form the CUDA kernel implements, written out without np.interp.

    numpy's arr_interp (numpy/_core/src/multiarray/compiled_base.c): for each
    x_new find i with xp[i] <= x_new < xp[i+1]; i == n-1 -> fp[n-1];
    xp[i] == x_new -> fp[i]; else slope=(fp[i+1]-fp[i])/(xp[i+1]-xp[i]),
    y = slope*(x_new-xp[i]) + fp[i]  (separate f64 mul and add).
    np.linspace(0,1,k)[i] = i*(1.0/(k-1)), last element forced to 1.0.
    Used by tests to prove the closed form == np.interp bit for bit.
    """
    n = len(x_i16)
    fp = x_i16.astype(np.float64)
    if n == 1:
        return np.full(m, fp[0]).astype(np.int16)
    step_o = np.float64(1.0) / np.float64(n - 1)
    xo = np.arange(n, dtype=np.float64) * step_o
    xo[-1] = 1.0
    if m == 1:
        xn = np.array([0.0])
    else:
        xn = np.arange(m, dtype=np.float64) * (np.float64(1.0) / np.float64(m - 1))
        xn[-1] = 1.0
    i = np.minimum((xn * (n - 1)).astype(np.int64), n - 1)
    for _ in range(2):  # fix-up: candidate may be off by one
        i = np.where(xo[i] > xn, i - 1, i)
        ip = np.minimum(i + 1, n - 1)
        i = np.where((i < n - 1) & (xo[ip] <= xn), i + 1, i)
    ip = np.minimum(i + 1, n - 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        slope = (fp[ip] - fp[i]) / (xo[ip] - xo[i])
        y = slope * (xn - xo[i]) + fp[i]
    y = np.where((i == n - 1) | (xo[i] == xn), fp[i], y)
    return y.astype(np.int16)


def decode_audio_to_pcm16(data: bytes, fmt: str, target_rate: int = 16000) -> bytes:
    """src/realtime/audio_buffer.py:37-58."""
    if fmt == "pcm16":
        return resample_linear(data, 24000, target_rate)
    if fmt == "g711_ulaw":
        return resample_linear(ulaw2lin(data), 8000, target_rate)
    if fmt == "g711_alaw":
        return resample_linear(alaw2lin(data), 8000, target_rate)
    raise ValueError(f"Unsupported audio format: {fmt}")


def encode_pcm16_to_format(pcm16_data: bytes, from_rate: int, fmt: str) -> bytes:
    """src/realtime/audio_buffer.py:61-81."""
    if fmt == "pcm16":
        return resample_linear(pcm16_data, from_rate, 24000)
    if fmt == "g711_ulaw":
        return lin2ulaw(resample_linear(pcm16_data, from_rate, 8000))
    if fmt == "g711_alaw":
        return lin2alaw(resample_linear(pcm16_data, from_rate, 8000))
    raise ValueError(f"Unsupported audio format: {fmt}")
