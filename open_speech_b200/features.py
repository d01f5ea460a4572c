"""B200 drop-in for faster_whisper.FeatureExtractor (duck-type).

Usage with the reference backend (src/backends/faster_whisper.py:40-45 creates the WhisperModel):
    model.feature_extractor = B200FeatureExtractor(feature_size=model.feature_extractor.mel_filters.shape[0])
"""
from __future__ import annotations

import numpy as np

from . import _native as N


class B200FeatureExtractor:
    def __init__(self, feature_size: int = 80, sampling_rate: int = 16000, hop_length: int = 160,
                 chunk_length: int = 30, n_fft: int = 400):
        if (sampling_rate, hop_length, n_fft) != (16000, 160, 400):
            raise ValueError("B200FeatureExtractor supports Whisper's 16 kHz / n_fft 400 / hop 160 front-end only")
        if feature_size not in (80, 128):
            raise ValueError("feature_size must be 80 or 128")
        self.n_fft = n_fft
        self.hop_length = hop_length
        self.chunk_length = chunk_length
        self.n_samples = chunk_length * sampling_rate
        self.nb_max_frames = self.n_samples // hop_length
        self.time_per_frame = hop_length / sampling_rate
        self.sampling_rate = sampling_rate
        self.feature_size = feature_size
        mf = np.empty((feature_size, 201), dtype=np.float32)
        N.call("osb_mel_filters", feature_size, N.ptr(mf), mf.size)
        self.mel_filters = mf

    def __call__(self, waveform: np.ndarray, padding: int = 160, chunk_length: int | None = None) -> np.ndarray:
        """float32[N] (or int16[N], taken as /32768) -> float32[n_mels, (N+160)//160]."""
        if chunk_length is not None:
            self.n_samples = chunk_length * self.sampling_rate
            self.nb_max_frames = self.n_samples // self.hop_length
        if padding != 160:
            raise ValueError("only faster-whisper's default padding=160 is implemented")
        w = np.asarray(waveform)
        if w.dtype == np.int16:
            w, fmt = np.ascontiguousarray(w), N.FMT_PCM16
        else:
            w, fmt = np.ascontiguousarray(w, dtype=np.float32), N.FMT_F32
        n_frames = N.lib().osb_logmel_frames(w.size)
        out = np.empty((self.feature_size, n_frames), dtype=np.float32)
        N.call("osb_logmel_host", N.ptr(w), fmt, w.size, self.feature_size, N.ptr(out), 0, -18.0)
        return out
