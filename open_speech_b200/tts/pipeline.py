"""Drop-in for the PCM edge of src/tts/pipeline.py (reference lines 32-66), GPU-backed.

The ffmpeg encoders of the reference (pipeline.py:69-264) are an external codec process and out of
scope.
"""
from __future__ import annotations

import struct

import numpy as np

from .. import _native as N


def float32_to_int16(audio) -> np.ndarray:
    """Convert float32 [-1, 1] to int16: clip, *32767, truncate toward zero."""
    if hasattr(audio, "numpy"):
        audio = audio.numpy()
    a = np.ascontiguousarray(audio, dtype=np.float32)
    out = np.empty(a.shape, dtype=np.int16)
    if a.size:
        N.call("osb_f32_to_pcm16_host", N.ptr(a), N.ptr(out), a.size)
    return out


def wav_header(num_samples: int, sample_rate: int) -> bytes:
    data_size = num_samples * 2
    return (b"RIFF" + struct.pack("<I", 36 + data_size) + b"WAVE" + b"fmt " + struct.pack("<I", 16)
            + struct.pack("<H", 1) + struct.pack("<H", 1) + struct.pack("<I", sample_rate)
            + struct.pack("<I", sample_rate * 2) + struct.pack("<H", 2) + struct.pack("<H", 16)
            + b"data" + struct.pack("<I", data_size))


def encode_wav(audio: np.ndarray, sample_rate: int = 24000) -> bytes:
    pcm = float32_to_int16(audio)
    return wav_header(len(pcm), sample_rate) + pcm.tobytes()


def encode_pcm(audio: np.ndarray) -> bytes:
    return float32_to_int16(audio).tobytes()
