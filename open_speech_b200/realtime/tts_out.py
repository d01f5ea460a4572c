"""Realtime TTS output framing, GPU-backed (reference src/realtime/server.py:231-277).

The reference's response handler concatenates the synthesised float32 24 kHz chunks, quantises
them with ``(x * 32767).clip(-32768, 32767).astype(int16)``, encodes to the session's output
format (``encode_pcm16_to_format``) and sends the payload as base64 deltas of 3000 bytes.
``encode_response_audio`` is that tail of ``_synthesize`` and ``audio_deltas`` the delta loop:
one host->device copy, quantise / resample+G.711 / base64 kernels, the text comes back once.
"""
from __future__ import annotations

from typing import Iterable

import numpy as np

from .. import _native as N
from .audio_buffer import _FMT

CHUNK_SIZE = 3000  # payload bytes per delta (server.py:268); 4000 base64 characters
_B64_PER_DELTA = CHUNK_SIZE // 3 * 4


def _collect(chunks: Iterable) -> np.ndarray | None:
    # server.py:238-246: ndarray chunks are taken as they are, anything else through np.array(dtype=float32)
    parts = [c if isinstance(c, np.ndarray) else np.array(c, dtype=np.float32) for c in chunks]
    if not parts:
        return None
    combined = np.concatenate(parts)
    # the reference multiplies in the chunks' own dtype; backends deliver float32, anything else is cast (deviations)
    return np.ascontiguousarray(combined, dtype=np.float32)


def _encode(chunks: Iterable, output_format: str, want_payload: bool, want_text: bool) -> tuple[bytes, str]:
    if output_format not in _FMT:
        raise ValueError(f"Unsupported audio format: {output_format}")
    combined = _collect(chunks)
    if combined is None or combined.size == 0:
        return b"", ""
    fmt, n = _FMT[output_format], int(combined.size)
    # encode_pcm16_to_format(pcm16, 24000, fmt): pcm16 stays at 24 kHz, G.711 goes to 8 kHz (audio_buffer.py:20-34, 65-81)
    n_out = n if fmt == N.FMT_PCM16 else int(n * (8000 / 24000))
    if n_out == 0:
        return b"", ""
    n_bytes = 2 * n if fmt == N.FMT_PCM16 else n_out
    payload = np.empty(n_bytes, dtype=np.uint8) if want_payload else None
    text = np.empty((n_bytes + 2) // 3 * 4, dtype=np.uint8) if want_text else None
    N.call("osb_realtime_tts_encode_host", N.ptr(combined), n, fmt, n_out,
           N.ptr(payload) if want_payload else None, N.ptr(text) if want_text else None)
    return (payload.tobytes() if want_payload else b""), (text.tobytes().decode("ascii") if want_text else "")


def encode_response_audio(chunks: Iterable, output_format: str) -> bytes:
    """Payload bytes of a response: what ``_synthesize`` returns (server.py:232-251)."""
    return _encode(chunks, output_format, True, False)[0]


def audio_deltas(chunks: Iterable, output_format: str) -> list[str]:
    """The ``delta`` strings of the response.audio.delta events, in order (server.py:268-277)."""
    text = _encode(chunks, output_format, False, True)[1]
    return [text[i:i + _B64_PER_DELTA] for i in range(0, len(text), _B64_PER_DELTA)]
