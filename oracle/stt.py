"""Oracle: STT pre-processing, spectral-gating denoise and Whisper log-mel.  TEST ONLY.

Follows (reference file:line):
  wav_bytes_to_float32_mono   src/audio/preprocessing.py:9-20
  float32_mono_to_wav_bytes   src/audio/preprocessing.py:23-32
  normalize_gain              src/audio/preprocessing.py:35-42
  reduce_noise                src/audio/preprocessing.py:45-50  -> noisereduce (3P, absent)
  preprocess_stt_audio        src/audio/preprocessing.py:53-63
  log-mel                     call site src/backends/faster_whisper.py:245
                              -> faster_whisper.FeatureExtractor (3P, absent)

PARITY UNPINNED for ``spectral_gate`` (noisereduce>=3.0, pyproject.toml:38) and
``logmel`` (faster-whisper==1.2.1, requirements.lock:7): neither package is in
/root/reference nor installed; both are restated from their published
algorithm.  logmel is cross-checked against the
installed transformers.WhisperFeatureExtractor in tests/test_oracle_stt.py.
"""
from __future__ import annotations

import io
import wave

import numpy as np

# --------------------------------------------------------------------------- WAV edge


def wav_bytes_to_float32_mono(wav_bytes: bytes) -> tuple[np.ndarray, int]:
    with wave.open(io.BytesIO(wav_bytes), "rb") as wf:
        sr, ch, width = wf.getframerate(), wf.getnchannels(), wf.getsampwidth()
        raw = wf.readframes(wf.getnframes())
    if width != 2:
        raise ValueError("Only 16-bit WAV is supported for preprocessing")
    a = np.frombuffer(raw, dtype=np.int16).astype(np.float32) / 32768.0
    if ch > 1:
        a = a.reshape(-1, ch).mean(axis=1)
    return a, sr


def quantise_pcm16(audio: np.ndarray) -> np.ndarray:
    """clip -> *32767.0 -> astype(int16) (truncation toward zero), preprocessing.py:24-25."""
    return (np.clip(audio, -1.0, 1.0) * 32767.0).astype(np.int16)


def float32_mono_to_wav_bytes(audio: np.ndarray, sample_rate: int) -> bytes:
    pcm = quantise_pcm16(audio)
    buf = io.BytesIO()
    with wave.open(buf, "wb") as wf:
        wf.setnchannels(1)
        wf.setsampwidth(2)
        wf.setframerate(sample_rate)
        wf.writeframes(pcm.tobytes())
    return buf.getvalue()


def normalize_gain(audio: np.ndarray, target_dbfs: float = -18.0) -> np.ndarray:
    rms = np.sqrt(np.mean(np.square(audio)))
    if rms <= 1e-8:
        return audio
    gain = 10 ** ((target_dbfs - 20 * np.log10(rms)) / 20)
    return np.clip(audio * gain, -1.0, 1.0)


def preprocess_stt_audio(wav_bytes: bytes, *, noise_reduce: bool, normalize: bool) -> bytes:
    try:
        audio, sr = wav_bytes_to_float32_mono(wav_bytes)
    except Exception:
        return wav_bytes
    if noise_reduce:
        audio = spectral_gate(audio, sr)
    if normalize:
        audio = normalize_gain(audio)
    return float32_mono_to_wav_bytes(audio, sr)


# --------------------------------------------------------------------------- spectral gating
# noisereduce.reduce_noise(y, sr) with every default:
# stationary=False, prop_decrease=1.0, time_constant_s=2.0, freq_mask_smooth_hz=500,
# time_mask_smooth_ms=50, thresh_n_mult_nonstationary=2, sigmoid_slope_nonstationary=10,
# chunk_size=600000, padding=30000, n_fft=1024, win_length=1024, hop_length=256.

NR_CHUNK = 600000
NR_PAD = 30000
NR_NFFT = 1024
NR_HOP = 256


def _nr_smoothing_filter(n_f: int, n_t: int) -> np.ndarray:
    def tri(n):
        return np.concatenate([np.linspace(0, 1, n + 1, endpoint=False), np.linspace(1, 0, n + 2)])[1:-1]

    f = np.outer(tri(n_f), tri(n_t))
    return f / np.sum(f)


def nr_params(sr: int) -> dict:
    t_frames = 2.0 * sr / float(NR_HOP)
    b = (np.sqrt(1 + 4 * t_frames**2) - 1) / (2 * t_frames**2)
    n_f = int(500 / (sr / (NR_NFFT / 2)))
    n_t = int(50 / ((NR_HOP / sr) * 1000))
    return {"b": float(b), "n_f": n_f, "n_t": n_t}


def _gate_chunk(chunk: np.ndarray, sr: int) -> np.ndarray:
    """One padded chunk -> filtered padded chunk, in the chunk's dtype: float64 is what noisereduce does; float32 is the
    PRECISION CONTROL of tests/test_gpu_denoise.py (same recipe, complex64 FFTs, float32 mask arithmetic)."""
    from scipy.signal import fftconvolve, filtfilt, istft, stft

    p = nr_params(sr)
    dt = chunk.dtype
    _, _, S = stft(chunk, nfft=NR_NFFT, noverlap=NR_NFFT - NR_HOP, nperseg=NR_NFFT, padded=False)
    A = np.abs(S)
    A_s = filtfilt([p["b"]], [1, p["b"] - 1], A, axis=-1, padtype=None).astype(dt)
    with np.errstate(divide="ignore", invalid="ignore"):
        M = (dt.type(1.0) / (dt.type(1.0) + np.exp(-((A - A_s) / A_s - dt.type(2.0)) * dt.type(10.0)))).astype(dt)
    if not (p["n_f"] == 1 and p["n_t"] == 1):
        M = fftconvolve(M, _nr_smoothing_filter(p["n_f"], p["n_t"]).astype(dt), mode="same")
    M = M * dt.type(1.0) + np.ones(M.shape, dt) * dt.type(0.0)
    _, xr = istft(S * M, nfft=NR_NFFT, noverlap=NR_NFFT - NR_HOP, nperseg=NR_NFFT)
    out = np.zeros(chunk.shape, dt)
    out[: len(xr)] = xr
    return out


def spectral_gate(y: np.ndarray, sr: int, work_dtype=np.float64) -> np.ndarray:
    """noisereduce SpectralGateNonStationary.get_traces() for a 1-D signal (work_dtype=float64, as upstream reads every chunk
    into a float64 buffer).  work_dtype=float32 is the precision control: the identical recipe carried out in single precision."""
    y = np.asarray(y)
    n = len(y)

    def read(i1, i2):
        c = np.zeros(i2 - i1, dtype=work_dtype)
        a, b = max(i1, 0), min(i2, n)
        c[a - i1 : b - i1] = y[a:b]
        return c

    def filt(start, end):
        i1, i2 = start - NR_PAD, end + NR_PAD
        return _gate_chunk(read(i1, i2), sr)[start - i1 : end - i1]

    if n > NR_CHUNK:
        out = np.zeros(n, dtype=y.dtype)
        pos = 0
        for k in range(0, (n - 1) // NR_CHUNK + 1):
            end0 = min(n - k * NR_CHUNK, NR_CHUNK)
            out[pos : pos + end0] = filt(k * NR_CHUNK, (k + 1) * NR_CHUNK)[:end0]
            pos += end0
        return out
    return filt(0, n).astype(y.dtype)


# --------------------------------------------------------------------------- Whisper log-mel


def mel_filters(sr: int = 16000, n_fft: int = 400, n_mels: int = 128) -> np.ndarray:
    """faster_whisper.FeatureExtractor.get_mel_filters (Slaney scale + area norm), f32."""
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    min_mel, max_mel = 0.0, 45.245640471924965
    mels = np.linspace(min_mel, max_mel, n_mels + 2)
    f_sp = 200.0 / 3
    freqs = f_sp * mels
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    log_t = mels >= min_log_mel
    freqs[log_t] = min_log_hz * np.exp(logstep * (mels[log_t] - min_log_mel))
    fdiff = np.diff(freqs)
    ramps = freqs.reshape(-1, 1) - fftfreqs.reshape(1, -1)
    lower = -ramps[:-2] / np.expand_dims(fdiff[:-1], axis=1)
    upper = ramps[2:] / np.expand_dims(fdiff[1:], axis=1)
    weights = np.maximum(np.zeros_like(lower), np.minimum(lower, upper))
    enorm = 2.0 / (freqs[2 : n_mels + 2] - freqs[:n_mels])
    weights *= np.expand_dims(enorm, axis=1)
    return weights.astype(np.float32)


def _stft_400(x: np.ndarray, window: np.ndarray, n_fft: int = 400, hop: int = 160) -> np.ndarray:
    """faster-whisper's numpy stft: reflect pad n_fft//2, strided frames, rfft (c64)."""
    pad = n_fft // 2
    xp = np.pad(x, (pad, pad), mode="reflect")
    n_frames = 1 + (len(xp) - n_fft) // hop
    frames = np.lib.stride_tricks.as_strided(
        xp, shape=(n_frames, n_fft), strides=(xp.strides[0] * hop, xp.strides[0])
    )
    return np.fft.rfft(frames * window, n=n_fft).T  # [201, n_frames]


def logmel(waveform: np.ndarray, n_mels: int = 128, padding: int = 160) -> np.ndarray:
    """FeatureExtractor.__call__(waveform, padding=160) -> f32[n_mels, frames]."""
    w = np.asarray(waveform, dtype=np.float32)
    if padding:
        w = np.pad(w, (0, padding))
    window = np.hanning(400 + 1)[:-1].astype(np.float32)
    stft = _stft_400(w, window)
    mag = np.abs(stft[:, :-1]) ** 2
    mel = mel_filters(16000, 400, n_mels) @ mag
    log_spec = np.log10(np.clip(mel, a_min=1e-10, a_max=None))
    log_spec = np.maximum(log_spec, log_spec.max() - 8.0)
    return ((log_spec + 4.0) / 4.0).astype(np.float32)


def logmel_n_frames(n_samples: int, padding: int = 160) -> int:
    return (n_samples + padding) // 160


def stt_frontend(pcm16: np.ndarray, *, noise_reduce: bool, normalize: bool = True, n_mels: int = 128,
                 sr: int = 16000, gate_dtype=np.float64) -> np.ndarray:
    """BASELINE configs 1 / 4 chain: int16 clip -> (denoise) -> normalise -> int16 -> /32768 -> log-mel.
    gate_dtype=float32: the single-precision control of the denoise stage (see spectral_gate)."""
    a = pcm16.astype(np.float32) / 32768.0
    if noise_reduce:
        a = spectral_gate(a, sr, gate_dtype)
    if normalize:
        a = normalize_gain(a)
    q = quantise_pcm16(a)
    return logmel(q.astype(np.float32) / 32768.0, n_mels)


def stt_full(wire: bytes, fmt: str, from_rate: int, *, linear_chunk: int = 0, noise_reduce: bool = True, normalize: bool = True,
             n_mels: int = 128, net=None, threshold: float = 0.5, min_speech_ms: int = 250, silence_ms: int = 800):
    """The north_star chain for ONE recording, stage by stage as the reference runs it:

    wire bytes -> audioop expand (realtime/audio_buffer.py:52-56) -> 16 kHz, either per ``linear_chunk`` input samples with
    ``_resample_linear`` (one ``decode_audio_to_pcm16`` per append, realtime/server.py:137) or whole-buffer
    ``resample_pcm16`` (streaming.py:55-91, as wyoming/stt_handler.py:75-79 does) -> [``SileroVAD.get_speech_segments`` on
    the pcm16 (vad/silero.py:109-177)] and [``preprocess_stt_audio`` -> /32768 -> FeatureExtractor (main.py:295-296,
    backends/faster_whisper.py:245)].  Returns (pcm16k int16, probs, segments [(start_ms, end_ms)], mel).
    """
    from . import codec, resample
    from . import vad as ovad

    if fmt == "g711_ulaw":
        lin = codec.ulaw2lin(wire)
    elif fmt == "g711_alaw":
        lin = codec.alaw2lin(wire)
    elif fmt == "pcm16":
        lin = wire
    else:
        raise ValueError(f"Unsupported audio format: {fmt}")
    if from_rate == 16000:
        pcm = lin
    elif linear_chunk:
        step = 2 * linear_chunk
        pcm = b"".join(codec.resample_linear(lin[i:i + step], from_rate, 16000) for i in range(0, len(lin) - step + 1, step))
    else:
        pcm = resample.resample_pcm16(lin, from_rate, 16000)
    x = np.frombuffer(pcm, dtype=np.int16)
    probs, segs = np.zeros(0, np.float32), []
    if net is not None:
        probs, _ = net.score_stream(x.astype(np.float32) / 32768.0)
        segs = [(s.start_ms, s.end_ms) for s in ovad.segments_from_probs(probs, len(x), threshold, min_speech_ms, silence_ms)]
    mel = stt_frontend(x, noise_reduce=noise_reduce, normalize=normalize, n_mels=n_mels)
    return x, probs, segs, mel
