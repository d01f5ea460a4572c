"""src/wyoming/tts_handler.py::_resample_to_16k (lines 37-44), GPU-backed."""
from __future__ import annotations

import numpy as np

from . import _native as N

WYOMING_RATE = 16000
TTS_SAMPLE_RATE = 24000


def _resample_to_16k(audio: np.ndarray, source_rate: int = TTS_SAMPLE_RATE) -> np.ndarray:
    """Linear interpolation on the index grid linspace(0, len-1, new_len) (float64), cast back to audio.dtype."""
    if source_rate == WYOMING_RATE:
        return audio
    new_length = int(len(audio) * (WYOMING_RATE / source_rate))
    a = np.ascontiguousarray(audio, dtype=np.float32)
    out = np.empty(new_length, dtype=np.float32)
    if new_length:
        N.call("osb_interp_index_f32_host", N.ptr(a), a.size, N.ptr(out), new_length)
    return out.astype(audio.dtype, copy=False)
