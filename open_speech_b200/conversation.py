"""Audio half of src/conversation.py::ConversationManager.render (reference lines 96-158), GPU-backed.

The reference's render() interleaves database rows, file writes and audio; only the audio is on the hot path:
per turn ``apply_chain(samples, sample_rate, effects)`` (:125-126) and ``encode_wav`` (:131), then the turns joined with
``SILENCE_MS`` of zeros between them (:113, :138-142) and ``duration_ms = int(1000 * len / sample_rate)`` (:128, :156).
Storage, profiles and synthesis stay with the caller.
"""
from __future__ import annotations

import numpy as np

from .effects.chain import apply_chain
from .tts.pipeline import encode_wav

SILENCE_MS = 500


def render_turns(turn_samples, turn_effects=None, sample_rate: int = 24000, save_turn_audio: bool = True) -> dict:
    """turn_samples: list of float32 arrays (what ``_synthesize_turn`` returned); turn_effects: list of effect lists.

    Returns {"merged": float32 array, "duration_ms": int, "turn_wavs": [bytes | None], "turn_duration_ms": [int]} with the
    reference's arithmetic: effects only when the list is non-empty, silence between turns but not after the last.
    """
    turn_effects = turn_effects or [None] * len(turn_samples)
    if len(turn_effects) != len(turn_samples):
        raise ValueError("turn_effects must have one entry per turn")
    silence = np.zeros(int(sample_rate * SILENCE_MS / 1000), dtype=np.float32)
    parts, wavs, durs = [], [], []
    for n, (samples, effects) in enumerate(zip(turn_samples, turn_effects), start=1):
        samples = np.asarray(samples, dtype=np.float32)
        if effects:
            samples = apply_chain(samples, sample_rate, effects)
        durs.append(int(1000 * len(samples) / sample_rate) if len(samples) else 0)
        wavs.append(encode_wav(samples, sample_rate=sample_rate) if save_turn_audio else None)
        parts.append(samples)
        if n < len(turn_samples):
            parts.append(silence)
    merged = np.concatenate(parts) if parts else np.zeros(0, dtype=np.float32)
    return {"merged": merged, "duration_ms": int(1000 * len(merged) / sample_rate) if len(merged) else 0,
            "turn_wavs": wavs, "turn_duration_ms": durs}
