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


def _pcm_to_wav(audio_bytes: bytes, rate: int, width: int, channels: int) -> bytes:
    """Raw PCM -> WAV container (src/wyoming/stt_handler.py:22-40; host-side header only)."""
    import struct

    n = len(audio_bytes)
    return (b"RIFF" + struct.pack("<I", 36 + n) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, channels, rate, rate * channels * width,
            channels * width, width * 8) + b"data" + struct.pack("<I", n) + audio_bytes)


def _extract_speech_segments(pcm_data: bytes, rate: int, width: int, channels: int, *, session=None, vad_enabled: bool = True,
                             threshold: float = 0.5, min_speech_ms: int = 250, silence_ms: int = 800) -> bytes:
    """VAD-gated utterance assembly (src/wyoming/stt_handler.py:43-115), one device-resident pipeline:
    [polyphase resample to 16 kHz] -> VAD -> segments -> gather of the speech spans at the original rate.
    The reference reads threshold / min_speech / silence from settings (defaults 0.5 / 250 / 800) and skips
    filtering when no VAD model is loaded; here they are keyword arguments and `session` is a VadSession."""
    import ctypes

    if not vad_enabled or not pcm_data or width != 2 or channels != 1 or session is None:
        return pcm_data
    n = len(pcm_data) // 2
    out = np.empty(n, dtype=np.int16)
    out_n, n_seg = ctypes.c_int64(0), ctypes.c_int(0)
    try:
        N.call("osb_vad_extract_speech_host", session.handle, pcm_data, n, int(rate), float(threshold), int(min_speech_ms), int(silence_ms),
               N.ptr(out), ctypes.byref(out_n), ctypes.byref(n_seg))
    except Exception:  # "VAD filtering failed, using original audio" (stt_handler.py:112-115)
        return pcm_data
    if out_n.value == 0:
        return pcm_data
    return out[: out_n.value].tobytes()
