"""Drop-in for src/effects/chain.py (reference lines 15-74), GPU-backed (osb_fx_chain_host).

Same effect dictionaries, same order semantics, unknown types skipped, result float32.  The whole chain is
one C call: the samples cross PCIe once in each direction regardless of the number of effects.
``pitch`` follows librosa.effects.pitch_shift's stretch-then-resample (phase vocoder 2048/512 on the GPU); librosa and
soxr are absent here, so that effect is parity-unpinned and uses a Kaiser-sinc resampler (csrc/pitch.cu).
"""
from __future__ import annotations

import numpy as np

from .. import _native as N

SUPPORTED_EFFECTS = {"normalize", "pitch", "reverb", "podcast_eq", "robot"}
_FX = {"normalize": 1, "reverb": 2, "podcast_eq": 3, "robot": 4, "pitch": 5}
_ROOM_MS = {"small": 50, "medium": 120, "large": 300}
_ROOM_MIX = {"small": 0.25, "medium": 0.4, "large": 0.55}


def encode_effects(effects: list[dict] | None):
    """[{type: ...}, ...] -> (types int32[], p0 f64[], p1 f64[]) as osb_fx_chain expects."""
    types, p0, p1 = [], [], []
    for fx in effects or []:
        t = fx.get("type")
        if t not in _FX:
            continue
        a = b = 0.0
        if t == "normalize":
            a = float(fx.get("target_lufs", -16))
        elif t == "pitch":
            a = float(fx.get("semitones", 0))
        elif t == "reverb":
            room = fx.get("room", "small")
            a = float(_ROOM_MS.get(room, 50))
            b = float(fx.get("mix", _ROOM_MIX.get(room, 0.3)))
        types.append(_FX[t]); p0.append(a); p1.append(b)
    return np.asarray(types, dtype=np.int32), np.asarray(p0, dtype=np.float64), np.asarray(p1, dtype=np.float64)


def _run(samples: np.ndarray, sample_rate: int, effects: list[dict]) -> np.ndarray:
    types, p0, p1 = encode_effects(effects)
    a = np.ascontiguousarray(samples, dtype=np.float32)
    if len(types) == 0 or a.size == 0:
        return samples.astype(np.float32, copy=False)
    out = np.empty_like(a)
    N.call("osb_fx_chain_host", N.ptr(a), a.size, int(sample_rate), N.ptr(types), N.ptr(p0), N.ptr(p1), len(types), N.ptr(out), 0)
    return out


def apply_chain(samples: np.ndarray, sample_rate: int, effects: list[dict] | None) -> np.ndarray:
    """Apply ordered list of effects. Each dict: {type: str, ...params}"""
    return _run(samples, sample_rate, list(effects or []))


# single-effect helpers with the reference's names.  NOTE: inside apply_chain intermediate results stay float64
# after the first float64 effect, exactly like the reference; these helpers return the float32 cast of one effect.
def _normalize(samples, target_lufs: float = -16):
    return _run(samples, 24000, [{"type": "normalize", "target_lufs": target_lufs}])


def _pitch_shift(samples, sample_rate: int, semitones: float = 0):
    return _run(samples, sample_rate, [{"type": "pitch", "semitones": semitones}])


def _reverb(samples, sample_rate: int, room: str = "small", mix: float = 0.2):
    return _run(samples, sample_rate, [{"type": "reverb", "room": room, "mix": mix}])


def _podcast_eq(samples, sample_rate: int):
    return _run(samples, sample_rate, [{"type": "podcast_eq"}])


def _robot(samples, sample_rate: int):
    return _run(samples, sample_rate, [{"type": "robot"}])
