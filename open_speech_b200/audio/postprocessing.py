"""Drop-in for src/audio/postprocessing.py (reference lines 8-40), GPU-backed (osb_tts_post_host)."""
from __future__ import annotations

import ctypes
from typing import Iterator

import numpy as np

from .. import _native as N


def _post(audio: np.ndarray, trim: bool, normalize: bool, threshold: float, peak: float) -> np.ndarray:
    a = np.ascontiguousarray(audio, dtype=np.float32)
    out = np.empty_like(a)
    m = ctypes.c_int64(0)
    N.call("osb_tts_post_host", N.ptr(a), a.size, int(trim), int(normalize), float(threshold), float(peak), N.ptr(out), ctypes.byref(m))
    out = out[: m.value]
    # the kernels compute in float32; a float64 caller gets its dtype back like the reference's numpy expressions would give
    # (values carry float32 precision: INTEGRATION.md, deviations)
    if isinstance(audio, np.ndarray) and audio.dtype != np.float32 and np.issubdtype(audio.dtype, np.floating):
        out = out.astype(audio.dtype)
    return out


def trim_silence(audio: np.ndarray, threshold: float = 0.01) -> np.ndarray:
    if len(audio) == 0:
        return audio
    out = _post(audio, True, False, threshold, 0.95)
    # nothing above the threshold: the reference returns its input object unchanged (:12-13)
    return audio if len(out) == len(audio) else out


def normalize_output(audio: np.ndarray, peak: float = 0.95) -> np.ndarray:
    if len(audio) == 0:
        return audio
    return _post(audio, False, True, 0.01, peak)


def process_tts_chunks(chunks: Iterator[np.ndarray], *, trim: bool = True, normalize: bool = True) -> Iterator[np.ndarray]:
    all_chunks = list(chunks)
    if not all_chunks:
        return iter(())
    audio = np.concatenate(all_chunks)
    if len(audio) and (trim or normalize):
        audio = _post(audio, trim, normalize, 0.01, 0.95)  # one H2D, trim + peak-normalise fused, one D2H
    return iter([audio.astype(np.float32)])
