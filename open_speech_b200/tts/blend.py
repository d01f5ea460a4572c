"""Kokoro voice-style blending on the GPU (drop-in for KokoroBackend._blend_voices, kokoro.py:289-308)."""
from __future__ import annotations

import ctypes

import numpy as np

from .. import _native as N
from .voices import VoiceSpec


def blend_voice_arrays(packs: list[np.ndarray], weights: list[float]) -> np.ndarray:
    """result = zeros_like(packs[0]); result += w_i * pack_i  (float32, in order) via osb_voice_blend_host."""
    arrs = [np.ascontiguousarray(p, dtype=np.float32) for p in packs]
    shape, n = arrs[0].shape, arrs[0].size
    if any(a.size != n for a in arrs):
        raise ValueError("voice packs must have the same shape")
    ptrs = (ctypes.c_void_p * len(arrs))(*[N.ptr(a) for a in arrs])
    w = np.asarray(weights, dtype=np.float32)
    out = np.empty(n, dtype=np.float32)
    N.call("osb_voice_blend_host", ctypes.cast(ptrs, ctypes.c_void_p), N.ptr(w), len(arrs), n, N.ptr(out))
    return out.reshape(shape)


def blend_voices(pipeline, spec: VoiceSpec):
    """Same contract as KokoroBackend._blend_voices(spec): loads each component with pipeline.load_voice()
    and returns the weighted sum as a torch.FloatTensor (bind as a method: ``KokoroBackend._blend_voices =
    lambda self, spec: blend_voices(self._pipeline, spec)``)."""
    import torch

    tensors = [pipeline.load_voice(c.voice_id) for c in spec.components]
    out = blend_voice_arrays([t.detach().cpu().numpy() for t in tensors], spec.normalized_weights())
    return torch.from_numpy(out).to(tensors[0].device)
