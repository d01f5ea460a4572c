"""Seeded synthetic inputs for the BASELINE configs.

No datasets or checkpoints are reachable (no network), so every benchmark and
parity test runs on these "speech-like" signals: voiced bursts (8 harmonics of
a random f0 with a 4 Hz envelope) separated by gaps, over a white-noise floor.
Pure numpy; used by tests, bench.py and __graft_entry__.smoke().
"""
from __future__ import annotations

import numpy as np

SEED_C1, SEED_C2, SEED_C3, SEED_C4, SEED_C5 = 1001, 1002, 1003, 1004, 1005


def speech_like(n_samples: int, sr: int = 16000, seed: int = 0, noise_rms: float = 0.002,
                extra_noise_rms: float = 0.0) -> np.ndarray:
    """float32[n_samples] in [-1, 1]."""
    rng = np.random.default_rng(seed)
    x = np.zeros(n_samples, dtype=np.float32)
    pos = int(rng.uniform(0.0, 0.5) * sr)
    while pos < n_samples:
        dur = int(rng.uniform(0.4, 3.0) * sr)
        end = min(n_samples, pos + dur)
        t = np.arange(end - pos, dtype=np.float64) / sr
        f0 = rng.uniform(90.0, 250.0)
        burst = np.zeros(end - pos, dtype=np.float64)
        for k in range(1, 9):
            burst += np.sin(2 * np.pi * f0 * k * t + rng.uniform(0, 2 * np.pi)) / k
        burst *= 0.6 + 0.4 * np.sin(2 * np.pi * 4.0 * t)
        rms = np.sqrt(np.mean(burst**2)) + 1e-12
        burst *= rng.uniform(0.03, 0.25) / rms
        x[pos:end] = burst.astype(np.float32)
        pos = end + int(rng.uniform(0.2, 1.0) * sr)
    x += (rng.standard_normal(n_samples) * noise_rms).astype(np.float32)
    if extra_noise_rms > 0:
        x += (rng.standard_normal(n_samples) * extra_noise_rms).astype(np.float32)
    return np.clip(x, -1.0, 1.0)


def to_pcm16(x: np.ndarray) -> np.ndarray:
    """The reference's own quantisation (src/audio/preprocessing.py:24-25)."""
    return (np.clip(x, -1.0, 1.0) * 32767.0).astype(np.int16)


def clip_pcm16(seconds: float, sr: int = 16000, seed: int = SEED_C1, extra_noise_rms: float = 0.0) -> np.ndarray:
    return to_pcm16(speech_like(int(seconds * sr), sr, seed, extra_noise_rms=extra_noise_rms))


def clip_batch_pcm16(n_clips: int, seconds: float, sr: int = 16000, seed: int = SEED_C4,
                     extra_noise_rms: float = 0.01, distinct: int = 8) -> np.ndarray:
    """int16[n_clips, n]; ``distinct`` different clips tiled (generation cost stays bounded)."""
    n = int(seconds * sr)
    base = [clip_pcm16(seconds, sr, seed + i, extra_noise_rms) for i in range(min(distinct, n_clips))]
    out = np.empty((n_clips, n), dtype=np.int16)
    for i in range(n_clips):
        out[i] = base[i % len(base)]
    return out


def tts_utterance(seconds: float, seed: int, sr: int = 24000) -> np.ndarray:
    """Kokoro-shaped f32 utterance: near-silent lead/tail (|x|<0.005) around a speech-like body."""
    rng = np.random.default_rng(seed)
    n = int(seconds * sr)
    lead = int(rng.uniform(0.05, 0.3) * sr)
    tail = int(rng.uniform(0.05, 0.3) * sr)
    body = speech_like(max(n - lead - tail, sr // 10), sr, seed + 7)
    body *= rng.uniform(0.2, 0.8) / (np.max(np.abs(body)) + 1e-9)
    quiet = lambda k: (rng.uniform(-0.004, 0.004, size=k)).astype(np.float32)
    return np.concatenate([quiet(lead), body.astype(np.float32), quiet(tail)])


def tts_batch(n_utts: int, seed: int = SEED_C5, sr: int = 24000, distinct: int = 16,
              min_s: float = 2.0, max_s: float = 12.0) -> list[np.ndarray]:
    rng = np.random.default_rng(seed)
    base = [tts_utterance(rng.uniform(min_s, max_s), seed + 100 + i, sr) for i in range(min(distinct, n_utts))]
    return [base[i % len(base)] for i in range(n_utts)]


def voice_packs(n: int = 3, seed: int = SEED_C5) -> list[np.ndarray]:
    """Synthetic Kokoro-82M-shaped voice packs f32[510,1,256] ~ N(0, 0.1)."""
    rng = np.random.default_rng(seed)
    return [(rng.standard_normal((510, 1, 256)) * 0.1).astype(np.float32) for _ in range(n)]


def ulaw_streams(n_streams: int, n_ticks: int, chunk: int = 160, seed: int = SEED_C3) -> np.ndarray:
    """uint8[n_ticks, n_streams, chunk] mu-law bytes of speech-like 8 kHz audio."""
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore", DeprecationWarning)
        import audioop  # stdlib (<=3.12); only used to make synthetic input bytes

    def lin2ulaw_np(p):
        return np.frombuffer(audioop.lin2ulaw(p.tobytes(), 2), dtype=np.uint8)

    distinct = min(n_streams, 16)
    total = n_ticks * chunk
    base = np.stack([lin2ulaw_np(clip_pcm16(total / 8000.0, 8000, seed + i)[:total]) for i in range(distinct)])
    idx = np.arange(n_streams) % distinct
    return np.ascontiguousarray(base[idx].reshape(n_streams, n_ticks, chunk).transpose(1, 0, 2))
