"""Oracle: Silero VAD framing, network arithmetic, segmenter, input-buffer gate.  TEST ONLY.

Follows (reference file:line):
  SileroVAD.__call__            src/vad/silero.py:63-91   (512-sample windows, NO 64-sample
                                                           context, max prob, state carried)
  SileroVAD.is_speech           src/vad/silero.py:93-107
  SileroVAD.get_speech_segments src/vad/silero.py:109-177
  InputAudioBuffer.append       src/realtime/audio_buffer.py:111-156

PARITY UNPINNED for the network arithmetic: the reference downloads an unpinned
ONNX file at run time (src/vad/silero.py:28,196-206) and runs it with
onnxruntime==1.24.1 (requirements.lock:11); neither is available offline and
every reference VAD test mocks the session (tests/test_vad.py:19-43).
``SileroNet`` restates the published Silero VAD v5 16 kHz graph (SURVEY.md
App. A.6) with seeded random-init weights, which BASELINE.json config 2 allows.
The state machines ARE pinned: tests drive the reference's own classes with a
scripted session and compare (tests/golden/vad_segments.json).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

VAD_SAMPLE_RATE = 16000
WINDOW = 512
HIDDEN = 128

# (name, out_ch, in_ch, kernel, stride)
ENCODER = (("enc1", 128, 129, 3, 1), ("enc2", 64, 128, 3, 2), ("enc3", 64, 64, 3, 2), ("enc4", 128, 64, 3, 1))


def stft_basis() -> np.ndarray:
    """[258,256] f32: Hann(256, periodic) * {cos, -sin}(2 pi k n / 256), k=0..128."""
    n = np.arange(256)
    win = 0.5 - 0.5 * np.cos(2 * np.pi * n / 256)
    fb = np.fft.fft(np.eye(256))
    basis = np.vstack([np.real(fb[:129]), np.imag(fb[:129])]) * win[None, :]
    return basis.astype(np.float32)


def make_weights(seed: int = 1002) -> dict[str, np.ndarray]:
    """PyTorch-default-style uniform init, numpy RNG (independent of torch version)."""
    rng = np.random.default_rng(seed)
    w: dict[str, np.ndarray] = {"stft_basis": stft_basis()}

    def u(shape, bound):
        return rng.uniform(-bound, bound, size=shape).astype(np.float32)

    for name, oc, ic, k, _s in ENCODER:
        bound = 1.0 / np.sqrt(ic * k)
        w[f"{name}.weight"] = u((oc, ic, k), bound)
        if name == "enc1":  # lift |STFT| of ~0.1-amplitude audio to O(1) features
            w[f"{name}.weight"] = (w[f"{name}.weight"] * 16.0).astype(np.float32)
        w[f"{name}.bias"] = u((oc,), bound)
    b = 1.0 / np.sqrt(HIDDEN)
    w["lstm.weight_ih"] = u((4 * HIDDEN, HIDDEN), b)
    w["lstm.weight_hh"] = u((4 * HIDDEN, HIDDEN), b)
    w["lstm.bias_ih"] = u((4 * HIDDEN,), b)
    w["lstm.bias_hh"] = u((4 * HIDDEN,), b)
    # head scaled/biased so that the seeded random network gives probabilities that straddle
    # 0.5 and follow frame energy on the section-8(d) speech-like inputs (corr ~ +0.83)
    w["dec.weight"] = (u((HIDDEN,), b) * -120.0).astype(np.float32)
    w["dec.bias"] = np.array([-2.9], dtype=np.float32)
    return w


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _conv1d(x: np.ndarray, w: np.ndarray, b: np.ndarray, stride: int) -> np.ndarray:
    """x [B, Cin, T], w [Cout, Cin, 3], pad=1 -> [B, Cout, Tout] (f32)."""
    B, C, T = x.shape
    xp = np.zeros((B, C, T + 2), np.float32)
    xp[:, :, 1:-1] = x
    t_out = (T + 2 - 3) // stride + 1
    cols = np.stack([xp[:, :, t * stride : t * stride + 3] for t in range(t_out)], axis=1)  # [B,Tout,C,3]
    y = cols.reshape(B * t_out, C * 3) @ w.reshape(w.shape[0], -1).T + b
    return y.reshape(B, t_out, -1).transpose(0, 2, 1).astype(np.float32)


class SileroNet:
    """Numpy restatement of Silero VAD v5 (16 kHz branch) with an ORT-like ``run``."""

    def __init__(self, weights: dict[str, np.ndarray] | None = None):
        self.w = weights if weights is not None else make_weights()

    # -- stateless front: windows [B,512] f32 -> LSTM input-gate pre-activations [B,512]
    def front(self, windows: np.ndarray) -> np.ndarray:
        w = self.w
        x = np.asarray(windows, np.float32)
        xp = np.concatenate([x, x[:, -2:-66:-1]], axis=1)  # right reflect pad 64 -> 576
        frames = np.stack([xp[:, 128 * f : 128 * f + 256] for f in range(3)], axis=1)  # [B,3,256]
        spec = frames @ w["stft_basis"].T  # [B,3,258]
        mag = np.sqrt(spec[..., :129] ** 2 + spec[..., 129:] ** 2).transpose(0, 2, 1)  # [B,129,3]
        h = mag.astype(np.float32)
        for name, _oc, _ic, _k, s in ENCODER:
            h = np.maximum(_conv1d(h, w[f"{name}.weight"], w[f"{name}.bias"], s), 0.0)
        feat = h[:, :, 0]  # [B,128]
        return (feat @ w["lstm.weight_ih"].T + w["lstm.bias_ih"] + w["lstm.bias_hh"]).astype(np.float32)

    def step(self, pre: np.ndarray, h: np.ndarray, c: np.ndarray):
        """One LSTMCell step + head.  pre [512], h/c [128] -> (prob, h', c')."""
        w = self.w
        g = pre + w["lstm.weight_hh"] @ h
        i, f, gg, o = g[:128], g[128:256], g[256:384], g[384:]
        c2 = (_sigmoid(f) * c + _sigmoid(i) * np.tanh(gg)).astype(np.float32)
        h2 = (_sigmoid(o) * np.tanh(c2)).astype(np.float32)
        logit = np.float32(np.dot(np.maximum(h2, 0.0), w["dec.weight"]) + w["dec.bias"][0])
        return np.float32(_sigmoid(logit)), h2, c2

    def run(self, _names, inputs):
        """onnxruntime.InferenceSession.run look-alike (src/vad/silero.py:86)."""
        x = np.asarray(inputs["input"], np.float32)
        st = np.asarray(inputs["state"], np.float32)
        pre = self.front(x)[0]
        p, h2, c2 = self.step(pre, st[0, 0], st[1, 0])
        return [np.array([[p]], np.float32), np.stack([h2, c2])[:, None, :].astype(np.float32)]

    def score_stream(self, audio_f32: np.ndarray, state: np.ndarray | None = None):
        """All full windows of one stream -> (probs [W] f32, final state [2,1,128])."""
        n_win = len(audio_f32) // WINDOW
        st = np.zeros((2, 1, HIDDEN), np.float32) if state is None else state.copy()
        if n_win == 0:
            return np.zeros(0, np.float32), st
        pre = self.front(np.asarray(audio_f32[: n_win * WINDOW], np.float32).reshape(n_win, WINDOW))
        h, c = st[0, 0].copy(), st[1, 0].copy()
        probs = np.zeros(n_win, np.float32)
        for t in range(n_win):
            probs[t], h, c = self.step(pre[t], h, c)
        return probs, np.stack([h, c])[:, None, :]


@dataclass
class Segment:
    start_ms: int
    end_ms: int


def segments_from_probs(probs, n_samples: int, threshold: float = 0.5, min_speech_ms: int = 250,
                        silence_ms: int = 800) -> list[Segment]:
    """The integer state machine of src/vad/silero.py:133-177 on per-window probabilities."""
    window_ms = WINDOW * 1000 // VAD_SAMPLE_RATE
    silence_windows = max(1, silence_ms // window_ms)
    min_speech_windows = max(1, min_speech_ms // window_ms)
    segs: list[Segment] = []
    in_speech, speech_start, silence_count, speech_windows = False, 0, 0, 0
    for k, prob in enumerate(probs):
        cur = (k * WINDOW) * 1000 // VAD_SAMPLE_RATE
        if float(prob) >= threshold:
            silence_count = 0
            if not in_speech:
                in_speech, speech_start, speech_windows = True, cur, 0
            speech_windows += 1
        elif in_speech:
            silence_count += 1
            if silence_count >= silence_windows:
                if speech_windows >= min_speech_windows:
                    segs.append(Segment(speech_start, cur))
                in_speech, silence_count, speech_windows = False, 0, 0
    if in_speech and speech_windows >= min_speech_windows:
        segs.append(Segment(speech_start, n_samples * 1000 // VAD_SAMPLE_RATE))
    return segs


def input_buffer_events(chunk_probs, chunk_samples, threshold: float = 0.5, silence_duration_ms: int = 500):
    """InputAudioBuffer.append gate (src/realtime/audio_buffer.py:125-156) over a chunk sequence.

    chunk_probs[i] is the VAD's return value for chunk i (max over its windows, 0.0 if
    none); chunk_samples[i] its length in 16 kHz samples.  Returns [(chunk_idx, type, ms)].
    """
    ev = []
    total, in_speech, sil = 0, False, 0
    for i, (p, n) in enumerate(zip(chunk_probs, chunk_samples)):
        cur = (total * 1000) // VAD_SAMPLE_RATE
        total += n
        if n == 0:
            continue
        if float(p) >= threshold:
            sil = 0
            if not in_speech:
                in_speech = True
                ev.append((i, "speech_started", cur))
        elif in_speech:
            sil += n
            if (sil * 1000) // VAD_SAMPLE_RATE >= silence_duration_ms:
                in_speech, sil = False, 0
                ev.append((i, "speech_stopped", cur))
    return ev


ACT_SPEECH_START, ACT_UTTERANCE_RESET, ACT_APPEND, ACT_TRANSCRIBE, ACT_FINALIZE, ACT_SPEECH_END = 1, 2, 4, 8, 16, 32


def stream_gate_steps(chunk_probs, chunk_samples_16k: int, *, vad_enabled: bool = True, threshold: float = 0.5,
                      endpointing_samples: int = 4800, max_utterance_bytes: int = 960000):
    """StreamingSession._process_chunk + the state half of _transcribe_utterance / _finalize_utterance
    (src/streaming.py:290-355, :357-360, :429-436, :493-498) over a chunk sequence.

    Returns one (actions, speech_active, silence_samples, utterance_bytes) tuple per chunk; ``actions`` uses the
    OSB_ACT_* bits of include/osb200.h.  Pinned by tests/golden/stream_gate.json (traces of the reference's own class).
    """
    out = []
    active, silence, utt = False, 0, 0
    cb = 2 * chunk_samples_16k
    for p in chunk_probs:
        act, work, fin = 0, False, False
        if not vad_enabled:
            if not active:
                active, utt = True, 0
                act |= ACT_UTTERANCE_RESET
            utt += cb
            work, fin = True, utt >= max_utterance_bytes
        elif float(np.float32(p)) >= float(np.float32(threshold)):
            silence = 0
            if not active:
                active, utt = True, 0
                act |= ACT_UTTERANCE_RESET | ACT_SPEECH_START
            utt += cb
            work, fin = True, utt >= max_utterance_bytes
        elif active:
            silence += chunk_samples_16k
            utt += cb
            work, fin = True, silence >= endpointing_samples
        if work:
            act |= ACT_APPEND
            if fin:
                was = active
                active, silence = False, 0
                if utt < 3200:
                    if was and vad_enabled:
                        act |= ACT_SPEECH_END
                else:
                    act |= ACT_FINALIZE | (ACT_SPEECH_END if vad_enabled else 0)
                    utt = 0
            elif utt >= 3200:
                act |= ACT_TRANSCRIBE
        out.append((act, active, silence, utt))
    return out
