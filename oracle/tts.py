"""Oracle: TTS output post-processing, effects chain, voice blending, PCM edge.  TEST ONLY.

Follows (reference file:line):
  trim_silence / normalize_output / process_tts_chunks  src/audio/postprocessing.py:8-40
  apply_chain, _normalize, _reverb, _podcast_eq, _robot src/effects/chain.py:15-74
  parse_voice_spec / normalized_weights                 src/tts/voices.py:29-82
  KokoroBackend._blend_voices                           src/tts/backends/kokoro.py:289-308
  float32_to_int16 / encode_pcm / encode_wav            src/tts/pipeline.py:32-66

Pinned against the reference modules imported in the build container
(oracle/make_golden.py; src/effects/chain.py needs a stub ``librosa`` module
because librosa is not installed).

``pitch_shift`` is PARITY UNPINNED: librosa (>=0.10, requirements.lock:14, unpinned) and its default
resampler soxr are absent and no reference test holds a value (tests/test_effects_chain.py:14-18 checks
length and "changed" only).  The time-stretch half is librosa's published algorithm restated from memory
(stft 2048/512 Hann centred, phase_vocoder, istft with window-sum-square normalisation); the resampling
half substitutes a Kaiser-windowed sinc interpolator (the "kaiser_best" design librosa used before soxr:
64 zero crossings, 512 table entries per crossing, roll-off 0.9476, beta 14.77) for soxr_hq.
tests/test_oracle_pitch.py checks the restated pieces against torch.stft / torch.istft, torchaudio's phase_vocoder and
torchaudio's Kaiser-sinc resampler (independent implementations that are installed).
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


# ---- pitch shift (src/effects/chain.py:44-48 -> librosa.effects.pitch_shift), PARITY UNPINNED, see module docstring

PS_NFFT, PS_HOP = 2048, 512
PS_ZEROS, PS_PREC, PS_ROLLOFF, PS_BETA = 64, 9, 0.9475937167399596, 14.769656459379492


def _ps_window() -> np.ndarray:
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(PS_NFFT) / PS_NFFT)  # get_window("hann", 2048, fftbins=True)


def ps_stft(y: np.ndarray) -> np.ndarray:
    """librosa.stft(y) defaults: centre zero padding, complex64 [1025, 1 + n // 512]."""
    yp = np.pad(y.astype(np.float32), PS_NFFT // 2)
    n_frames = 1 + (len(yp) - PS_NFFT) // PS_HOP
    idx = np.arange(PS_NFFT)[None, :] + PS_HOP * np.arange(n_frames)[:, None]
    return np.fft.rfft(_ps_window()[None, :] * yp[idx], axis=1).T.astype(np.complex64)


def ps_phase_vocoder(D: np.ndarray, rate: float) -> np.ndarray:
    """librosa.phase_vocoder(D, rate=rate, hop_length=512, n_fft=2048): f32 phase accumulator, f64 increments."""
    n_bins, n_frames = D.shape
    steps = np.arange(0, n_frames, rate, dtype=np.float64)
    out = np.zeros((n_bins, len(steps)), np.complex64)
    phi = np.linspace(0, np.pi * PS_HOP, n_bins)
    acc = np.angle(D[:, 0]).astype(np.float32)
    Dp = np.pad(D, ((0, 0), (0, 2)))
    mag_all, ang_all = np.abs(Dp), np.angle(Dp)  # float32
    for t, step in enumerate(steps):
        i = int(step)
        alpha = float(np.mod(step, 1.0))
        mag = (1.0 - alpha) * mag_all[:, i].astype(np.float64) + alpha * mag_all[:, i + 1].astype(np.float64)
        out[:, t] = ((np.cos(acc) + 1j * np.sin(acc)) * mag).astype(np.complex64)
        d = ang_all[:, i + 1].astype(np.float64) - ang_all[:, i].astype(np.float64) - phi
        d = d - 2.0 * np.pi * np.round(d / (2.0 * np.pi))
        acc = (acc.astype(np.float64) + (phi + d)).astype(np.float32)
    return out


def ps_istft(D: np.ndarray, length: int) -> np.ndarray:
    """librosa.istft(D, length=length) defaults (centred): overlap-add / window sum-square, float32."""
    n_frames = min(D.shape[1], int(np.ceil((length + PS_NFFT) / PS_HOP)))
    w = _ps_window()
    total = PS_NFFT + PS_HOP * (n_frames - 1)
    y = np.zeros(total)
    wss = np.zeros(total)
    fr = np.fft.irfft(D[:, :n_frames].astype(np.complex128), n=PS_NFFT, axis=0) * w[:, None]
    for f in range(n_frames):
        y[f * PS_HOP : f * PS_HOP + PS_NFFT] += fr[:, f]
        wss[f * PS_HOP : f * PS_HOP + PS_NFFT] += w * w
    nz = wss > np.finfo(np.float32).tiny
    y[nz] /= wss[nz]
    y = y[PS_NFFT // 2 :]
    out = np.zeros(length, np.float32)
    m = min(length, len(y))
    out[:m] = y[:m]
    return out


def ps_sinc_table():
    n = (2**PS_PREC) * PS_ZEROS
    sinc_win = PS_ROLLOFF * np.sinc(PS_ROLLOFF * np.linspace(0, PS_ZEROS, num=n + 1, endpoint=True))
    return np.kaiser(2 * n + 1, PS_BETA)[n:] * sinc_win


def ps_resample(x: np.ndarray, ratio: float) -> np.ndarray:
    """Band-limited interpolation with the table above (resampy's scheme), n_out = int(len(x) * ratio); f32 out."""
    win = ps_sinc_table()
    nb = 2**PS_PREC
    if ratio < 1:
        win = ratio * win
    delta = np.diff(win, append=win[-1])
    scale = min(1.0, ratio)
    step = int(scale * nb)
    n_orig, n_out = len(x), int(len(x) * ratio)
    t = np.arange(n_out) * (1.0 / ratio)
    n = t.astype(np.int64)
    y = np.zeros(n_out)
    xd = x.astype(np.float64)
    for sign in (0, 1):
        frac = scale * (t - n)
        if sign:
            frac = scale - frac
        fi = frac * nb
        off = fi.astype(np.int64)
        eta = fi - off
        kmax = (len(win) - off) // step
        kmax = np.minimum(kmax, n + 1) if not sign else np.minimum(kmax, n_orig - n - 1)
        for k in range(int(kmax.max()) if n_out else 0):
            m = k < kmax
            idx = np.where(m, off + k * step, 0)
            src = np.where(m, (n - k) if not sign else (n + k + 1), 0)
            y += np.where(m, (win[idx] + eta * delta[idx]) * xd[src], 0.0)
    return y.astype(np.float32)


def pitch_shift(x: np.ndarray, sample_rate: int, semitones: float) -> np.ndarray:
    """_pitch_shift (src/effects/chain.py:44-48): identity for 0, else stretch by 2^(-n/12) and resample back."""
    if semitones == 0:
        return x
    x = x.astype(np.float32)
    rate = 2.0 ** (-float(semitones) / 12)
    stretched = ps_istft(ps_phase_vocoder(ps_stft(x), rate), int(round(len(x) / rate)))
    ratio = float(sample_rate) / (float(sample_rate) / rate)
    y = ps_resample(stretched, ratio)
    out = np.zeros(len(x), np.float32)  # fix_length(ceil(len*ratio)) then fix_length(len(x)): zero-pad or crop
    m = min(len(x), len(y), int(np.ceil(len(stretched) * ratio)))
    out[:m] = y[:m]
    return out


def apply_chain(samples: np.ndarray, sample_rate: int, effects) -> np.ndarray:
    for fx in effects or []:
        t = fx.get("type")
        if t == "normalize":
            samples = fx_normalize(samples, fx.get("target_lufs", -16))
        elif t == "pitch":
            samples = pitch_shift(samples, sample_rate, fx.get("semitones", 0))
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
