"""the sample-level pieces of src/composer.py::MultiTrackComposer, GPU-backed.

  resample(samples, src_rate, dst_rate)  == MultiTrackComposer._resample        (composer.py:167-173)
  mix_prepared(prepared, sample_rate)    == MultiTrackComposer._mix_prepared    (composer.py:175-189)
  float_to_int16(samples)                == MultiTrackComposer._float_to_int16  (composer.py:254-257)
Bind as methods with ``MultiTrackComposer._resample = staticmethod(lambda s, a, b: composer.resample(s, a, b))`` etc.
"""
from __future__ import annotations

import math

import numpy as np

from . import _native as N
from .tts.pipeline import float32_to_int16 as float_to_int16  # same arithmetic: clip, *32767, truncate


def resample(samples: np.ndarray, src_rate: int, dst_rate: int) -> np.ndarray:
    if src_rate == dst_rate:
        return samples.astype(np.float32, copy=False)
    g = math.gcd(src_rate, dst_rate)
    up, down = dst_rate // g, src_rate // g
    a = np.ascontiguousarray(samples, dtype=np.float32)
    n_out = (a.size * up + down - 1) // down
    out = np.empty(n_out, dtype=np.float32)
    if a.size:
        N.call("osb_resample_poly_f32_host", N.ptr(a), N.ptr(out), a.size, up, down)
    return out


def mix_prepared(prepared: list[dict], sample_rate: int) -> np.ndarray:
    starts, arrays, total = [], [], 0
    for track in prepared:
        start = int(round(max(0.0, float(track.get("offset_s", 0.0))) * sample_rate))
        arr = np.ascontiguousarray(track["samples"], dtype=np.float32)
        starts.append(start)
        arrays.append(arr)
        total = max(total, start + len(arr))
    if total <= 0:
        return np.zeros(0, dtype=np.float32)
    lens = np.array([len(a) for a in arrays], dtype=np.int64)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    flat = np.concatenate(arrays) if arrays else np.zeros(0, np.float32)
    st = np.array(starts, dtype=np.int64)
    out = np.empty(total, dtype=np.float32)
    N.call("osb_mix_tracks_host", N.ptr(flat), N.ptr(offs), N.ptr(lens), N.ptr(st), len(arrays), flat.size, total, N.ptr(out))
    return out
