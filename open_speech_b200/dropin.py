"""Alias the reference's module paths onto this package."""
from __future__ import annotations

import importlib
import sys

_MAP = {
    "src.audio.preprocessing": "open_speech_b200.audio.preprocessing",
    "src.audio.postprocessing": "open_speech_b200.audio.postprocessing",
    "src.effects.chain": "open_speech_b200.effects.chain",
    "src.tts.voices": "open_speech_b200.tts.voices",
}
# modules the reference keeps MORE than the hot path in: patch the hot-path names only
_PATCH = {
    "src.realtime.audio_buffer": ("open_speech_b200.realtime.audio_buffer",
                                  ["_resample_linear", "decode_audio_to_pcm16", "encode_pcm16_to_format", "InputAudioBuffer"]),
    "src.streaming": ("open_speech_b200.streaming", ["resample_pcm16"]),
    "src.vad.silero": ("open_speech_b200.vad.silero", ["SileroVAD", "Segment", "get_vad_model"]),
    "src.tts.pipeline": ("open_speech_b200.tts.pipeline", ["float32_to_int16", "encode_wav", "encode_pcm"]),
}


def install() -> None:
    """Call after the reference's ``src`` package is importable and before the server starts."""
    for ref, ours in _MAP.items():
        sys.modules[ref] = importlib.import_module(ours)
    for ref, (ours, names) in _PATCH.items():
        try:
            target = importlib.import_module(ref)
        except Exception:
            continue
        mod = importlib.import_module(ours)
        for n in names:
            setattr(target, n, getattr(mod, n))
