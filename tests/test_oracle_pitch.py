"""The pitch-shift oracle (oracle/tts.py) is PARITY UNPINNED: librosa and soxr are not installed and the reference holds no
value for this effect.  What can be checked here is that the restated pieces are the algorithms they claim to be, against
independent implementations that ARE installed: torch.stft / torch.istft and torchaudio.functional.phase_vocoder (the same
phase-vocoder recurrence as librosa's, accumulated in float64) and torchaudio's Kaiser-windowed sinc resampler."""
import math

import numpy as np
import pytest

from oracle import tts

torch = pytest.importorskip("torch")


def _signal(n=36000, seed=1):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 24000
    return (0.3 * np.sin(2 * np.pi * 220 * t) + 0.1 * np.sin(2 * np.pi * 1500 * t) + 0.02 * rng.standard_normal(n)).astype(np.float32)


def _hann64():
    return torch.hann_window(2048, periodic=True, dtype=torch.float64)


def test_stft_istft_match_torch():
    x = _signal()
    D = tts.ps_stft(x)
    T = torch.stft(torch.from_numpy(x).double(), n_fft=2048, hop_length=512, window=_hann64(), center=True, pad_mode="constant",
                   return_complex=True).numpy()
    assert D.shape == T.shape and np.abs(D - T).max() <= 1e-5 * np.abs(T).max()
    y = torch.istft(torch.from_numpy(D.astype(np.complex128)), n_fft=2048, hop_length=512, window=_hann64(), center=True, length=len(x)).numpy()
    assert np.abs(tts.ps_istft(D, len(x)) - y).max() <= 1e-6
    assert np.abs(tts.ps_istft(D, len(x)) - x).max() <= 1e-6  # and the pair is an identity


@pytest.mark.parametrize("semitones", [4, -3])
def test_phase_vocoder_matches_torchaudio(semitones):
    ta = pytest.importorskip("torchaudio")
    x = _signal()
    rate = 2.0 ** (-semitones / 12)
    D = tts.ps_stft(x)
    mine = tts.ps_phase_vocoder(D, rate)
    adv = torch.linspace(0, math.pi * 512, 1025, dtype=torch.float64)[..., None]
    ref = ta.functional.phase_vocoder(torch.from_numpy(D.astype(np.complex128)), rate, adv).numpy()
    assert mine.shape == ref.shape
    assert np.abs(np.abs(mine) - np.abs(ref)).max() <= 1e-4  # magnitudes: plain interpolation
    # phases: librosa accumulates in float32 (restated so), torchaudio here in float64
    assert np.sqrt(np.mean(np.abs(mine - ref) ** 2) / np.mean(np.abs(ref) ** 2)) <= 2e-3
    n2 = int(round(len(x) / rate))
    y_ref = torch.istft(torch.from_numpy(ref), n_fft=2048, hop_length=512, window=_hann64(), center=True, length=n2).numpy()
    assert np.abs(tts.ps_istft(mine, n2) - y_ref).max() <= 5e-3 * np.abs(y_ref).max()


def test_sinc_resampler_matches_torchaudio_kaiser():
    ta = pytest.importorskip("torchaudio")
    x = _signal(24000, seed=2)
    orig, new = 20181, 24000  # the ratio of a -3 semitone shift
    ref = ta.functional.resample(torch.from_numpy(x.astype(np.float64)), orig, new, lowpass_filter_width=64, rolloff=tts.PS_ROLLOFF,
                                 resampling_method="sinc_interp_kaiser", beta=tts.PS_BETA).numpy()
    mine = tts.ps_resample(x, new / orig)
    m = min(len(ref), len(mine))
    assert abs(len(ref) - len(mine)) <= 1
    assert np.abs(ref[:m] - mine[:m]).max() <= 1e-3  # same design family, different tabulation


@pytest.mark.parametrize("semitones", [5, -7, 12])
def test_pitch_moves_by_the_semitone_ratio(semitones):
    tone = (0.5 * np.sin(2 * np.pi * 440 * np.arange(48000) / 24000)).astype(np.float32)
    y = tts.pitch_shift(tone, 24000, semitones)
    assert y.shape == tone.shape and y.dtype == np.float32
    f = np.fft.rfftfreq(16384, 1 / 24000)[np.argmax(np.abs(np.fft.rfft(y[8000 : 8000 + 16384] * np.hanning(16384))))]
    assert abs(f - 440 * 2 ** (semitones / 12)) <= 3.0
    assert tts.pitch_shift(tone, 24000, 0) is tone  # identity returns its input (chain.py:46-47)
