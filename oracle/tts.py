"""Oracle: TTS output post-processing, effects chain, voice blending, PCM edge.  TEST ONLY.

Follows (reference file:line):
  trim_silence / normalize_output / process_tts_chunks  src/audio/postprocessing.py:8-40
  apply_chain, _normalize, _reverb, _podcast_eq, _robot src/effects/chain.py:15-74
  parse_voice_spec / normalized_weights                 src/tts/voices.py:29-82
  KokoroBackend._blend_voices                           src/tts/backends/kokoro.py:289-308
  float32_to_int16 / encode_pcm / encode_wav            src/tts/pipeline.py:32-66

Pinned against the reference modules imported in the build container
(oracle/make_golden.py; src/effects/chain.py needs a stub ``librosa`` module
because librosa is not installed).  ``_pitch_shift`` (librosa + soxr) is
PARITY UNPINNED and not restated.
"""
from __future__ import annotations

import re
import struct

import numpy as np


def trim_silence(audio: np.ndarray, threshold: float = 0.01) -> np.ndarray:
    if len(audio) == 0:
        return audio
    idx = np.where(np.abs(audio) > threshold)[0]
    if len(idx) == 0:
        return audio
    return audio[idx[0] : idx[-1] + 1]


def normalize_output(audio: np.ndarray, peak: float = 0.95) -> np.ndarray:
    if len(audio) == 0:
        return audio
    m = float(np.max(np.abs(audio)))
    if m <= 1e-8:
        return audio
    return np.clip(audio * (peak / m), -1.0, 1.0)


def process_tts_chunks(chunks, *, trim: bool = True, normalize: bool = True):
    allc = list(chunks)
    if not allc:
        return iter(())
    a = np.concatenate(allc)
    if trim:
        a = trim_silence(a)
    if normalize:
        a = normalize_output(a)
    return iter([a.astype(np.float32)])


# --------------------------------------------------------------------------- effects


def fx_normalize(x: np.ndarray, target_lufs: float = -16) -> np.ndarray:
    rms = np.sqrt(np.mean(x**2)) if len(x) > 0 else 1.0
    if rms < 1e-8:
        return x
    return x * (10 ** (target_lufs / 20) / rms)


def reverb_ir(sample_rate: int, room: str = "small") -> np.ndarray:
    room_ms = {"small": 50, "medium": 120, "large": 300}.get(room, 50)
    n = max(1, int(sample_rate * room_ms / 1000))
    ir = np.exp(-np.linspace(0, 6, n))
    return ir / ir.sum()


def fx_reverb(x: np.ndarray, sample_rate: int, room: str = "small", mix: float = 0.2) -> np.ndarray:
    from scipy.signal import fftconvolve

    wet = fftconvolve(x, reverb_ir(sample_rate, room), mode="full")[: len(x)]
    return (1 - mix) * x + mix * wet


def podcast_eq_coeffs(sample_rate: int):
    from scipy import signal

    nyq = sample_rate / 2
    b_hp, a_hp = signal.butter(2, 80 / nyq, btype="high")
    b_pk, a_pk = signal.iirpeak(3000 / nyq, Q=2)
    return (b_hp, a_hp), (b_pk, a_pk)


def fx_podcast_eq(x: np.ndarray, sample_rate: int) -> np.ndarray:
    from scipy.signal import lfilter

    (b1, a1), (b2, a2) = podcast_eq_coeffs(sample_rate)
    return lfilter(b2, a2, lfilter(b1, a1, x))


def fx_robot(x: np.ndarray, sample_rate: int) -> np.ndarray:
    t = np.arange(len(x)) / sample_rate
    return x * np.sin(2 * np.pi * 100 * t)


def apply_chain(samples: np.ndarray, sample_rate: int, effects) -> np.ndarray:
    for fx in effects or []:
        t = fx.get("type")
        if t == "normalize":
            samples = fx_normalize(samples, fx.get("target_lufs", -16))
        elif t == "pitch":
            if fx.get("semitones", 0) != 0:
                raise NotImplementedError("pitch shift: librosa absent, parity unpinned (SURVEY 8(f))")
        elif t == "reverb":
            room = fx.get("room", "small")
            mix = fx.get("mix", {"small": 0.25, "medium": 0.4, "large": 0.55}.get(room, 0.3))
            samples = fx_reverb(samples, sample_rate, room, mix)
        elif t == "podcast_eq":
            samples = fx_podcast_eq(samples, sample_rate)
        elif t == "robot":
            samples = fx_robot(samples, sample_rate)
    return samples.astype(np.float32, copy=False)


# --------------------------------------------------------------------------- voices

_COMPONENT = re.compile(r"([a-zA-Z0-9_]+)(?:\((\d+(?:\.\d+)?)\))?")
_ALIASES = {"alloy": "af_heart", "echo": "am_adam", "fable": "bf_emma", "onyx": "am_michael",
            "nova": "af_nova", "shimmer": "af_bella"}


def parse_voice_spec(voice: str) -> list[tuple[str, float]]:
    if "+" not in voice and "(" not in voice:
        voice = _ALIASES.get(voice, voice)
    out = []
    for part in voice.split("+"):
        part = part.strip()
        m = _COMPONENT.fullmatch(part)
        if not m:
            raise ValueError(f"Invalid voice spec component: {part!r}")
        out.append((m.group(1), float(m.group(2)) if m.group(2) else 1.0))
    return out


def normalized_weights(components) -> list[float]:
    total = sum(w for _, w in components)
    if total == 0:
        return [1.0 / len(components)] * len(components)
    return [w / total for _, w in components]


def blend_voices(packs: list[np.ndarray], weights: list[float]) -> np.ndarray:
    """result = zeros; result += w_i * t_i in order, float32 (torch semantics)."""
    r = np.zeros_like(packs[0], dtype=np.float32)
    for w, t in zip(weights, packs):
        r += np.float32(w) * t.astype(np.float32)
    return r


# --------------------------------------------------------------------------- PCM edge


def float32_to_int16(audio: np.ndarray) -> np.ndarray:
    return (np.clip(audio, -1.0, 1.0) * 32767).astype(np.int16)


def wav_header(n_samples: int, sample_rate: int) -> bytes:
    d = n_samples * 2
    return (b"RIFF" + struct.pack("<I", 36 + d) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, 1, sample_rate,
            sample_rate * 2, 2, 16) + b"data" + struct.pack("<I", d))


def encode_wav(audio: np.ndarray, sample_rate: int = 24000) -> bytes:
    pcm = float32_to_int16(audio)
    return wav_header(len(pcm), sample_rate) + pcm.tobytes()


def tts_chain(chunks, effects, sample_rate: int = 24000) -> np.ndarray:
    """BASELINE config 5 per utterance: trim + peak normalise -> effects -> int16."""
    out = list(process_tts_chunks(iter(chunks), trim=True, normalize=True))
    a = out[0] if out else np.zeros(0, np.float32)
    return float32_to_int16(apply_chain(a, sample_rate, effects))
