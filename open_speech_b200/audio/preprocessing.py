"""Drop-in for src/audio/preprocessing.py (reference lines 9-63), GPU-backed.

Header parsing stays on the host (stdlib ``wave``, as in the reference); every sample-level
operation (int16->f32, channel mean, RMS reduce, gain, clip, requantise, spectral gating) runs
in libosb200.
"""
from __future__ import annotations

import ctypes
import io
import wave

import numpy as np

from .. import _native as N


def _parse_wav(wav_bytes: bytes):
    with wave.open(io.BytesIO(wav_bytes), "rb") as wf:
        sr, ch, width = wf.getframerate(), wf.getnchannels(), wf.getsampwidth()
        raw = wf.readframes(wf.getnframes())
    if width != 2:
        raise ValueError("Only 16-bit WAV is supported for preprocessing")
    return raw, sr, ch


def _wav_header(n_samples: int, sample_rate: int) -> bytes:
    """44-byte mono/16-bit header, byte-identical to what ``wave`` writes (reference :26-32)."""
    buf = io.BytesIO()
    with wave.open(buf, "wb") as wf:
        wf.setnchannels(1)
        wf.setsampwidth(2)
        wf.setframerate(sample_rate)
        wf.setnframes(n_samples)
        wf.writeframes(b"")
    hdr = bytearray(buf.getvalue()[:44])
    hdr[4:8] = (36 + 2 * n_samples).to_bytes(4, "little")
    hdr[40:44] = (2 * n_samples).to_bytes(4, "little")
    return bytes(hdr)


def wav_bytes_to_float32_mono(wav_bytes: bytes) -> tuple[np.ndarray, int]:
    raw, sr, ch = _parse_wav(wav_bytes)
    n = len(raw) // 2
    out = np.empty(n // ch, dtype=np.float32)
    if out.size:
        N.call("osb_pcm16_to_f32_host", raw, N.ptr(out), n, ch)
    return out, sr


def float32_mono_to_wav_bytes(audio: np.ndarray, sample_rate: int) -> bytes:
    a = np.ascontiguousarray(audio, dtype=np.float32)
    pcm = np.empty(a.size, dtype=np.int16)
    if a.size:
        N.call("osb_f32_to_pcm16_host", N.ptr(a), N.ptr(pcm), a.size)
    return _wav_header(a.size, sample_rate) + pcm.tobytes()


def normalize_gain(audio: np.ndarray, target_dbfs: float = -18.0) -> np.ndarray:
    a = np.ascontiguousarray(audio, dtype=np.float32)
    if a.size == 0:
        return audio  # np.mean of an empty array is nan; nan <= 1e-8 is False -> clip(empty) == empty
    out = np.empty_like(a)
    unchanged = ctypes.c_int(0)
    N.call("osb_normalize_gain_f32_host", N.ptr(a), N.ptr(out), 0, a.size, 1, float(target_dbfs), ctypes.byref(unchanged))
    if unchanged.value:
        return audio
    if isinstance(audio, np.ndarray) and audio.dtype != np.float32 and np.issubdtype(audio.dtype, np.floating):
        out = out.astype(audio.dtype)  # float32 arithmetic, the caller's dtype (deviations)
    return out


def reduce_noise(audio: np.ndarray, sample_rate: int) -> np.ndarray:
    """noisereduce.reduce_noise(y=audio, sr=sample_rate) defaults: non-stationary spectral gating."""
    a = np.ascontiguousarray(audio, dtype=np.float32)
    out = np.empty_like(a)
    if a.size:
        N.call("osb_spectral_gate_host", N.ptr(a), N.FMT_F32, N.ptr(out), a.size, int(sample_rate))
    return out


def preprocess_stt_audio(wav_bytes: bytes, *, noise_reduce: bool, normalize: bool) -> bytes:
    try:
        raw, sr, ch = _parse_wav(wav_bytes)
    except Exception:
        # Keep backward compatibility for tests/inputs that provide non-WAV bytes (reference :56-58)
        return wav_bytes
    n = len(raw) // 2
    frames = n // ch
    pcm = np.empty(frames, dtype=np.int16)
    if frames:
        if ch == 1 and not noise_reduce:
            # fused int16 -> (/32768) -> RMS -> gain -> clip -> *32767 -> int16, one H2D + one D2H
            N.call("osb_normalize_gain_pcm16_host", raw, N.ptr(pcm), n, int(bool(normalize)), -18.0)
        else:
            N.call("osb_preprocess_stt_host", raw, n, ch, int(sr), int(bool(noise_reduce)), int(bool(normalize)), -18.0, N.ptr(pcm))
    return _wav_header(frames, sr) + pcm.tobytes()
